import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from sparsernns_b200 import _abi
rows, K, N = map(int, sys.argv[1:4])
lib = _abi.lib(); rng = np.random.default_rng(0)
W = rng.standard_normal((K, N)).astype(np.float32)
h = C.c_void_p(); _abi.check(lib.s5b200_sparse_create(W.ctypes.data, K, N, _abi.F32, 0, C.byref(h)))
dev = torch.device('cuda:0')
dX = torch.randn(rows, K, device=dev); dY = torch.empty(rows, N, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2): _abi.check(lib.s5b200_spmm_f32(h, dX.data_ptr(), None, dY.data_ptr(), rows, 0, st))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5): _abi.check(lib.s5b200_spmm_f32(h, dX.data_ptr(), None, dY.data_ptr(), rows, 0, st))
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 5 * 1e3
print(rows, K, N, f"{us:.1f} us  {rows * (K + N) * 4 / us / 1e3:.0f} GB/s")
