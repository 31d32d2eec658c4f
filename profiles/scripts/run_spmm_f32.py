import ctypes as C, sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from sparsernns_b200 import _abi
lib=_abi.lib(); dev=torch.device("cuda:0")
rows,K,N=120032,96,128
rng=np.random.default_rng(0)
W=(rng.standard_normal((K,N))).astype(np.float32)
h=C.c_void_p(); _abi.check(lib.s5b200_sparse_create(W.ctypes.data,K,N,0,0,C.byref(h)))
X=torch.randn(rows,K,device=dev); Y=torch.empty(rows,N,device=dev)
st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    _abi.check(lib.s5b200_spmm_f32(h,X.data_ptr(),None,Y.data_ptr(),rows,0,st))
torch.cuda.synchronize()
