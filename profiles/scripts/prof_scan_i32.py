import ctypes as C, sys, numpy as np, torch
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from sparsernns_b200 import _abi
lib = _abi.lib(); dev = torch.device("cuda:0")
B, L, P = 32, 3751, 64
mk = lambda: torch.randint(-(1 << 14), 1 << 14, (B, L, P), dtype=torch.int32, device=dev)
bre, bim = mk(), mk(); xre, xim = torch.empty_like(bre), torch.empty_like(bim)
rng = np.random.default_rng(0)
lam = np.exp(-rng.uniform(0.01, 0.1, P) + 1j * rng.uniform(-1, 1, P))
are = torch.from_numpy(np.round(lam.real * (1 << 14)).astype(np.int32)).to(dev)
aim = torch.from_numpy(np.round(lam.imag * (1 << 14)).astype(np.int32)).to(dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    _abi.check(lib.s5b200_scan_i32(bre.data_ptr(), bim.data_ptr(), are.data_ptr(), aim.data_ptr(), xre.data_ptr(), xim.data_ptr(), B, L, P, 14, 14, 10, 10, 10, 10, 1, st))
torch.cuda.synchronize()
for relu in (0, 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _abi.check(lib.s5b200_scan_i32(bre.data_ptr(), bim.data_ptr(), are.data_ptr(), aim.data_ptr(), xre.data_ptr(), xim.data_ptr(), B, L, P, 14, 14, 10, 10, 10, 10, relu, st))
    e1.record(); torch.cuda.synchronize()
    print("relu", relu, "us per launch", e0.elapsed_time(e1) * 100)
