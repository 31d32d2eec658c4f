"""A few fixed-point forwards at the bench shape (32 clips x 3751 frames, dim_scale 0.5) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from sparsernns_b200 import fxputils
from sparsernns_b200.fxpmodel import FxpRegressionModel
dev = torch.device("cuda:0")
params, bs, cal = bench.load_fixture(0.5)
xs = bench.synth_inputs(torch, 2, 32, 3751, 0, dev)
md = fxputils.load_modeldict(params, cal)
fx = FxpRegressionModel(md, fxputils.create_fxp_qconfig(md, "w8a16", shared_exp=True), 3)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    fx.forward_float(xs[i % 2])
torch.cuda.synchronize()
