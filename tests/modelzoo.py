"""Test helper: build a synthetic, pruned, calibrated NDNS model.

Composes the product-side generator (``sparsernns_b200.synth``: shapes, init, masks, inputs) with
the oracle's float forward / calibration.  Used by the golden-fixture script and by tests.
"""

from __future__ import annotations

import numpy as np

from oracle import model as OM
from oracle.fxpconfig import build_fxp_qconfig
from sparsernns_b200 import synth


def oracle_cfg(dims, *, relufication, static_quant, associative_scan=False):
    return OM.ModelConfig(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size,
                          relufication=relufication, static_quant=static_quant,
                          associative_scan=associative_scan)


def fit_batchnorm(variables, x, cfg):
    """Fill BN running stats layer by layer from the float forward on ``x`` so activations are O(1)."""
    P, BS = variables["params"], variables["batch_stats"]
    h = OM.dense_float(np.asarray(x, np.float32), P["encoder"]["encoder"])
    if cfg.relufication:
        h = np.maximum(h, 0)
    for i in range(cfg.n_layers):
        flat = h.reshape(-1, h.shape[-1]).astype(np.float64)
        BS["encoder"][f"layers_{i}"]["norm"]["mean"] = flat.mean(0).astype(np.float32)
        BS["encoder"][f"layers_{i}"]["norm"]["var"] = flat.var(0).astype(np.float32)
        h = OM.layer_forward(h, P["encoder"][f"layers_{i}"], BS["encoder"][f"layers_{i}"], cfg)


def build_sparse_calibrated(dim_scale=0.5, n_clips=2, n_frames=256, seed=1919, sparsity=0.9,
                            bn_affine=True, exact_stft=False):
    """Returns dict(dims, variables (float, pruned), cal_params, cal_stats, x)."""
    dims = synth.ndns_dims(dim_scale)
    var = synth.init_variables(dims, seed=seed, bn_affine=bn_affine)
    if sparsity > 0:
        var = synth.prune_magnitude(var, sparsity)
    x = synth.synthetic_stft_batch(n_clips, n_frames, seed=0, exact_stft=exact_stft)
    cfg = oracle_cfg(dims, relufication=True, static_quant=False)
    fit_batchnorm(var, x, cfg)
    cal_params, cal_stats = OM.calibrate(x, var, oracle_cfg(dims, relufication=True, static_quant=True))
    return dict(dims=dims, variables=var, cal_params=cal_params, cal_stats=cal_stats, x=x)


def fxp_setup(zoo, shared_exp=True):
    md, q = build_fxp_qconfig(zoo["cal_params"], zoo["cal_stats"], "w8a16", shared_exp=shared_exp)
    return md, q
