"""Fixed-point flavour at the bench shape (32 clips x 3751 frames): tensor-core (fxp_tc.cu) against CUDA-core kernels
(fxp_model.cu).  Prints ms per forward, the per-kernel-class breakdown, and whether the two paths agree bit for bit."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench
from sparsernns_b200 import fxputils, synth
from sparsernns_b200.fxpmodel import FxpRegressionModel

ds = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
B, L = 32, 3751
dev = torch.device("cuda:0")
params, batch_stats, cal_stats = bench.load_fixture(ds)
dims = synth.ndns_dims(ds)
xs = bench.synth_inputs(torch, 2, B, L, 0, dev)
md = fxputils.load_modeldict(params, cal_stats)
q = fxputils.create_fxp_qconfig(md, "w8a16", shared_exp=True)
res = {}
outs = {}
for name, env in (("cuda_cores", {"S5B200_FXP_TC": "0"}), ("dense_tiles", {"S5B200_FXP_TC": "1"})):
    os.environ.update(env)
    fx = FxpRegressionModel(md, q, dims.n_layers)
    for _ in range(3):
        y, _m = fx.forward_float(xs[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(10):
        fx.forward_float(xs[i % 2])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    outs[name] = fx.forward_float(xs[1])[0].data.clone()
    res[name] = {"ms": ms, "frames_per_s": B * L / ms * 1e3, "kernels_ms": bench.kernel_breakdown(lambda: fx.forward_float(xs[0]))}
res["bit_identical"] = bool(torch.equal(outs["cuda_cores"], outs["dense_tiles"]))
print(json.dumps(res))
