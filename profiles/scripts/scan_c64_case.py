"""One stand-alone scan_c64 at a given shape (default: the bench shape, 32 clips x 3751 frames, P = 64), timed with CUDA events over
rotating buffers; with a -DS5_SCAN_TRACE build the kernel prints its phase times."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from sparsernns_b200 import _abi
B, L, P = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 3751, 64)))
lib = _abi.lib(); dev = torch.device('cuda:0')
lib.s5b200_scan_c64_workspace_bytes.restype = C.c_size_t
need = lib.s5b200_scan_c64_workspace_bytes(B, L, P)
ws = torch.zeros(max(need, 16), dtype=torch.uint8, device=dev)
nb = 4
Bu = [torch.randn(B, L, P, 2, device=dev) for _ in range(nb)]
xs = [torch.empty(B, L, P, 2, device=dev) for _ in range(nb)]
rng = np.random.default_rng(0)
lam = np.exp((-rng.uniform(1e-4, 1, P) + 1j * rng.uniform(-3, 3, P)) * 0.01).astype(np.complex64)
dlam = torch.from_numpy(np.stack([lam.real, lam.imag], -1).copy()).to(dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(i):
    _abi.check(lib.s5b200_scan_c64(Bu[i % nb].data_ptr(), dlam.data_ptr(), xs[i % nb].data_ptr(), B, L, P, 0, ws.data_ptr(), C.c_size_t(need), st))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
for i in range(3): run(i)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(reps): run(i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"scan_c64 B={B} L={L} P={P}: {us:.1f} us, {B * L * P * 16 / us / 1e3:.0f} GB/s algorithmic")
