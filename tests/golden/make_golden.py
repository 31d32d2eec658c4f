"""Generates the committed fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference cannot be imported in the build container (no JAX), so the vectors come from the oracle
(oracle/, the NumPy restatement of the reference).  Two kinds of files:

  ndns_ds{0.5,1.0}_sparse_w8a16.npz   recipe-shaped random-init model, ERK magnitude-pruned to 90 %, BatchNorm
                                      statistics fitted and static-quant scales calibrated on one synthetic batch
                                      (SURVEY.md Appendix C): params (with scales), batch_stats, calibration stats.
                                      bench.py and the tests load these; the CUDA path never imports oracle/.
  golden_io_ds0.5.npz                 a small input and the oracle's outputs for every numerics flavour.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import modelzoo as Z  # noqa: E402
from oracle import fxpmodel as OFM  # noqa: E402
from oracle import model as OM  # noqa: E402
from sparsernns_b200 import synth  # noqa: E402
from sparsernns_b200.checkpoint import save_npz  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def model_fixture(dim_scale, n_clips, n_frames):
    zoo = Z.build_sparse_calibrated(dim_scale, n_clips=n_clips, n_frames=n_frames, seed=1919)
    tree = {"params": zoo["cal_params"], "batch_stats": zoo["variables"]["batch_stats"], "cal_stats": zoo["cal_stats"]}
    save_npz(os.path.join(HERE, f"ndns_ds{dim_scale}_sparse_w8a16.npz"), tree)
    return zoo


def main():
    zoo = model_fixture(0.5, 4, 1024)
    model_fixture(1.0, 2, 512)
    dims = zoo["dims"]
    x = synth.synthetic_stft_batch(2, 96, seed=123)
    var_f = zoo["variables"]
    var_s = {"params": zoo["cal_params"], "batch_stats": zoo["variables"]["batch_stats"]}
    out = {"x": x}
    out["y_float_gelu"] = OM.model_forward(x, var_f, Z.oracle_cfg(dims, relufication=False, static_quant=False, associative_scan=True))
    out["y_float_relu"] = OM.model_forward(x, var_f, Z.oracle_cfg(dims, relufication=True, static_quant=False, associative_scan=True))
    out["y_static_w8a16"] = OM.model_forward(x, var_s, Z.oracle_cfg(dims, relufication=True, static_quant=True))
    md, q = Z.fxp_setup(zoo)
    m = OFM.FxpRegressionModel(md, q, dims.n_layers)
    taps = {}
    y = m(m.quantize_input(x), taps)
    out["y_fxp_int"] = y.data
    out["y_fxp_exp"] = np.int32(y.exp)
    out["fxp_layer_out_exp"] = np.array([taps[f"layers_{i}"]["output"].exp for i in range(dims.n_layers)], np.int32)
    for i in range(dims.n_layers):
        out[f"fxp_layer{i}_output"] = taps[f"layers_{i}"]["output"].data
    np.savez_compressed(os.path.join(HERE, "golden_io_ds0.5.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
