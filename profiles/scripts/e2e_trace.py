"""Device timeline of one s5b200_model_forward_host call (S5B200_E2E_TRACE=1), default chunk schedule and copy-only."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from sparsernns_b200 import synth
from sparsernns_b200.model import BoundModel, S5Config

dims = synth.ndns_dims(0.5)
params, batch_stats, _ = bench.load_fixture(0.5)
cfg = S5Config(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size, relufication=True, quantization="w8a16")
model = BoundModel(cfg, {"params": params, "batch_stats": batch_stats})
x = torch.randn(32, 3751, 257).abs().mul_(1e-3).pin_memory()
y = torch.empty_like(x).pin_memory()
for _ in range(3):
    model.forward_host(x, out=y)
os.environ["S5B200_E2E_TRACE"] = "1"
model.forward_host(x, out=y)
