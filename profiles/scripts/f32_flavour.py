"""Dense fp32 (GELU) flavour at the bench shape: ms per step and per kernel class (3 x TF32 dense-tile kernels, tc_f32.cu).
usage: python profiles/scripts/f32_flavour.py [dim_scale]    (S5B200_F32_V2=0 selects the first form of the kernel)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from sparsernns_b200.model import BoundModel, S5Config

ds = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
dev = torch.device("cuda:0")
params, bs, cal = bench.load_fixture(ds)
B, L = 32, 3751
xs = bench.synth_inputs(torch, 4, B, L, 0, dev)
from sparsernns_b200 import synth
dims = synth.ndns_dims(ds)
cfg = S5Config(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size, relufication=False, quantization="none", use_sparse=False)
fm = BoundModel(cfg, {"params": params, "batch_stats": bs})
for i in range(3):
    fm(xs[i % 4])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
n = 10
for i in range(n):
    fm(xs[i % 4])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"dim_scale": ds, "v2": os.environ.get("S5B200_F32_V2", "1"), "ms_per_step": ms, "frames_per_s": B * L / ms * 1e3,
                  "kernels_ms_per_step": bench.kernel_breakdown(lambda: fm(xs[0]))}))
