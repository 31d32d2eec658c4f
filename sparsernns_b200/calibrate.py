"""Static-quant calibration on the GPU.

The reference turns a float checkpoint into the static w8a16 model by running the ``calibrating=True`` graph over a
calibration batch (convert.py:881-971): every ``FakeQuant`` / ``QuantizedDense`` site passes its value through unchanged and
records min / max (quantization.py:291-349, 428-436, 730-745, 793-811), and each site's scale is
``2 ** round(log2(max(absmax / (2**(b-1) - 1), 1e-6)))`` (quantization.py:352-370).  Here the forward and the activation observers
run on the device (``s5b200_model_calibrate``); the weight observers and the scale formula are host NumPy.  The result has
the layout of ``sc_calibrated_params.pkl`` / ``sc_cal_stats.pkl`` (convert.py:963-971), i.e. what ``BoundModel`` (static
flavour) and ``fxputils.load_modeldict`` take.
"""

from __future__ import annotations

import copy
import ctypes as C

import numpy as np

from . import _abi
from .model import QUANT_RECIPES, BoundModel, S5Config, ssm_discrete

F32 = np.float32


def calculate_qparams(minval, maxval, precision: int, eps: float = 1e-6):
    """quantization.py:352-370, per-tensor symmetric, power-of-two scale; float32 throughout, round half to even."""
    max_abs = np.maximum(np.abs(F32(minval)), np.abs(F32(maxval)))
    scale = np.maximum(F32(max_abs / F32(2 ** (precision - 1) - 1)), F32(eps))
    return F32(F32(2.0) ** np.round(np.log2(F32(scale))))


def _obs(lo, hi, prefix=""):
    return {f"{prefix}observer_min": F32(lo), f"{prefix}observer_max": F32(hi)}


def calibrate(x, variables: dict, cfg: S5Config, recipe: str = "w8a16"):
    """One-batch calibration: ``x`` (B, L, d_input) float32 CUDA tensor, ``variables`` the float (pruned) model.

    Returns ``(params_with_scales, stats, y)``; ``y`` is the float model's output on ``x``.
    """
    import torch

    q = QUANT_RECIPES[recipe]
    fcfg = S5Config(n_layers=cfg.n_layers, d_model=cfg.d_model, ssm_size=cfg.ssm_size, d_input=cfg.d_input, d_output=cfg.d_output,
                    relufication=cfg.relufication, quantization="none", clip_eigs=cfg.clip_eigs, conj_sym=cfg.conj_sym,
                    discretization=cfg.discretization, step_rescale=cfg.step_rescale, bn_eps=cfg.bn_eps, use_sparse=cfg.use_sparse)
    model = BoundModel(fcfg, variables)
    lib = _abi.lib()
    x = model._check_in(x, cfg.d_input)
    B, L, _ = x.shape
    sites = lib.s5b200_calibrate_sites(model._handle)
    mm = torch.empty((sites, 2), dtype=torch.float32, device=x.device)
    y = torch.empty((B, L, cfg.d_output), dtype=torch.float32, device=x.device)
    ws, need = model._workspace(B, L)
    _abi.check(lib.s5b200_model_calibrate(model._handle, x.data_ptr(), y.data_ptr(), mm.data_ptr(), B, L, ws.data_ptr(), need,
                                          model._stream()), "s5b200_model_calibrate")
    mm = mm.cpu().numpy()

    P = copy.deepcopy(variables["params"])
    BS = variables["batch_stats"]
    stats = {"encoder": {}, "decoder": {}}
    a_bits, w_bits = q["non_ssm_act_precision"], q["non_ssm_precision"]

    def fq(sp: dict, ss: dict, lo, hi, bits: int):
        s = calculate_qparams(lo, hi, bits)
        sp["scale"] = s
        ss["scale"] = s
        ss["observer"] = _obs(lo, hi)

    def fqc_weights(sp: dict, ss: dict, val, bits: int):       # FakeQuantComplex on a weight tensor, quantization.py:451-474
        for part, name in ((val.real, "quant_real"), (val.imag, "quant_imag")):
            fq(sp.setdefault(name, {}), ss.setdefault(name, {}), part.min(), part.max(), bits)

    def fqc_sites(sp: dict, ss: dict, site_re: int, bits: int):
        for off, name in ((0, "quant_real"), (1, "quant_imag")):
            fq(sp.setdefault(name, {}), ss.setdefault(name, {}), mm[site_re + off, 0], mm[site_re + off, 1], bits)

    def qdense(p: dict, st: dict, site_in: int, site_out: int):
        k = np.asarray(p["kernel"], dtype=F32)
        for name, val in (("act_scale", calculate_qparams(mm[site_in, 0], mm[site_in, 1], a_bits)),
                          ("weight_scale", calculate_qparams(k.min(), k.max(), w_bits)),
                          ("out_scale", calculate_qparams(mm[site_out, 0], mm[site_out, 1], a_bits))):
            p[name] = val
            st[name] = val
        st["input_observer"] = _obs(mm[site_in, 0], mm[site_in, 1], "input_")
        st["output_observer"] = _obs(mm[site_out, 0], mm[site_out, 1], "output_")

    enc = P["encoder"]
    qdense(enc["encoder"], stats["encoder"].setdefault("encoder", {}), 0, 1)
    for i in range(cfg.n_layers):
        lp = enc[f"layers_{i}"]
        st = stats["encoder"].setdefault(f"layers_{i}", {})
        st["norm"] = {k: v for k, v in BS["encoder"][f"layers_{i}"]["norm"].items()}
        mp, ms = lp["mixer"], st.setdefault("mixer", {})
        s0 = 2 + 9 * i
        Lb, Bb, Ct, D = ssm_discrete(mp, fcfg)
        fq(mp.setdefault("quant_ut", {}), ms.setdefault("quant_ut", {}), mm[s0, 0], mm[s0, 1], q["ssm_act_precision"])
        fqc_weights(mp.setdefault("quant_B", {}), ms.setdefault("quant_B", {}), Bb, q["b_precision"])
        fqc_weights(mp.setdefault("quant_A", {}), ms.setdefault("quant_A", {}), Lb, q["a_precision"])
        fqc_weights(mp.setdefault("quant_C", {}), ms.setdefault("quant_C", {}), Ct, q["c_precision"])
        fqc_sites(mp.setdefault("quant_But", {}), ms.setdefault("quant_But", {}), s0 + 1, q["ssm_act_precision"])
        fqc_sites(mp.setdefault("quant_xt", {}), ms.setdefault("quant_xt", {}), s0 + 3, q["ssm_act_precision"])
        fq(mp.setdefault("quant_D", {}), ms.setdefault("quant_D", {}), D.min(), D.max(), q["d_precision"])
        fq(mp.setdefault("quant_yt", {}), ms.setdefault("quant_yt", {}), mm[s0 + 5, 0], mm[s0 + 5, 1], q["ssm_act_precision"])
        qdense(lp["out2"], st.setdefault("out2", {}), s0 + 6, s0 + 7)
        mg_p, mg_s = lp.setdefault("mult_gate", {}), st.setdefault("mult_gate", {})
        fq(mg_p.setdefault("quant_left", {}), mg_s.setdefault("quant_left", {}), mm[s0 + 6, 0], mm[s0 + 6, 1], a_bits)
        fq(mg_p.setdefault("quant_right", {}), mg_s.setdefault("quant_right", {}), mm[s0 + 8, 0], mm[s0 + 8, 1], a_bits)
    qdense(P["decoder"], stats["decoder"], sites - 2, sites - 1)
    model.close()
    return P, stats, y
