import sys, ctypes as C, torch, numpy as np
sys.path.insert(0,'.')
from sparsernns_b200 import _abi
from sparsernns_b200.model import scan_c64
B,L,P = 32,3751,64
Bu = torch.randn(B,L,P,dtype=torch.complex64,device='cuda')
lam = torch.from_numpy(np.exp(-np.random.uniform(0.01,0.1,P)+1j*np.random.uniform(-1,1,P)).astype(np.complex64)).cuda()
flush = torch.empty(256*1024*1024,dtype=torch.uint8,device='cuda')
for i in range(2):
    flush.zero_(); torch.cuda.synchronize()
    print('--- launch', i, flush=True)
    scan_c64(Bu, lam); torch.cuda.synchronize()
