import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from sparsernns_b200 import _abi
rows, K, N = map(int, sys.argv[1:4])
lib = _abi.lib(); rng = np.random.default_rng(0)
W = rng.standard_normal((K, N)).astype(np.float32); X = rng.standard_normal((rows, K)).astype(np.float32)
bias = rng.standard_normal(N).astype(np.float32)
h = C.c_void_p(); _abi.check(lib.s5b200_sparse_create(W.ctypes.data, K, N, _abi.F32, 0, C.byref(h)))
dev = torch.device('cuda:0')
dX, db = torch.from_numpy(X).to(dev), torch.from_numpy(bias).to(dev); dY = torch.empty(rows, N, device=dev)
_abi.check(lib.s5b200_spmm_f32(h, dX.data_ptr(), db.data_ptr(), dY.data_ptr(), rows, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
try:
    torch.cuda.synchronize()
    got = dY.cpu().numpy(); ref = X.astype(np.float64) @ W.astype(np.float64) + bias
    err = np.abs(got - ref).max(1) / np.abs(ref).max(1)
    print(rows, K, N, 'max err', err.max(), 'bad rows', np.nonzero(err > 1e-5)[0][:20])
except Exception as e:
    print(rows, K, N, 'FAILED', str(e)[:100])
