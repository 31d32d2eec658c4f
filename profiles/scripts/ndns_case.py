"""One stft -> istft round trip (32 clips x 480 000 samples, (B, L, 257) layout) and one few-clip scan_c64 - the launches profiled in
profiles/r02_ndns_audio.md / r02_scan_resident.md (run under ncu with -k regex:stft_kernel|istft_kernel|scan_c64_res)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from sparsernns_b200 import ndns
from sparsernns_b200.model import scan_c64

dev = torch.device("cuda:0")
a = (0.05 * torch.randn(32, 480000, device=dev))
for _ in range(2):
    mag, ph, xin = ndns.stft_splitter(a, blf=True, model_input=True)
    mask = 0.1 * torch.randn_like(mag)
    out = ndns.stft_mixer(mag, ph, blf=True, mask=mask)
Bu = torch.randn(32, 3751, 64, dtype=torch.complex64, device=dev)
lam = torch.exp(torch.complex(-torch.rand(64, device=dev) * 0.05, torch.rand(64, device=dev)))
for _ in range(2):
    xs = scan_c64(Bu, lam)
torch.cuda.synchronize()
print("ok", float(out.abs().max()), float(xs.abs().max()))
