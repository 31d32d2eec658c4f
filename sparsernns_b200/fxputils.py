"""Host-side fixed-point configuration (mirror of sparseRNNs/fxputils.py and fxprun.py:302-360).

Turns the calibrated parameter tree + calibration statistics into the ``modeldict`` and the
``fxp_qconfig`` (bits and exponents of every tensor) that ``FxpRegressionModel`` consumes.  Pure Python /
NumPy: the reference does this once per weight set on the host, and so do we.

Rules restated (with the reference lines they come from):
  * every leaf whose key contains "scale" is replaced by log2 of itself - the BatchNorm gamma included
    (fxputils.py:128-131); NaNs in norm/scale and norm/bias become 1.0 (fxputils.py:711-739);
  * observers gain absmax (fxputils.py:67-80);
  * required integer bits: ceil(log2(absmax)) floored at 0, +1 at exact powers of two (fxputils.py:137-142);
  * fraction bits: -ceil(log2(scale)) (fxputils.py:177-179, :249); dense biases use act_scale (:266);
  * exponent = min(fraction bits, bits - 1 - integer bits) (fxputils.py:404-450, 552-586);
  * shared exponents across blocks: per-key maximum (fxputils.py:288-348, agg="max");
  * BatchNorm parameter exponents from the parameter ranges themselves (fxputils.py:635-752).
"""

from __future__ import annotations

import copy
import math

import numpy as np

F32 = np.float32

PRECISIONS = {  # fxprun.py:302-308
    "w8a16": dict(non_ssm_w=8, non_ssm_b=16, non_ssm_act=16, ssm_w=8, ssm_act=16),
    "w8a8": dict(non_ssm_w=8, non_ssm_b=8, non_ssm_act=8, ssm_w=8, ssm_act=8),
}

_SSM_WEIGHTS = {"A_re": ("quant_A", "quant_real"), "A_im": ("quant_A", "quant_imag"), "B_re": ("quant_B", "quant_real"),
                "B_im": ("quant_B", "quant_imag"), "C_re": ("quant_C", "quant_real"), "C_im": ("quant_C", "quant_imag"),
                "D": ("quant_D",)}
_SSM_ACTS = {"u": ("quant_ut",), "Bu_re": ("quant_But", "quant_real"), "Bu_im": ("quant_But", "quant_imag"),
             "x_re": ("quant_xt", "quant_real"), "x_im": ("quant_xt", "quant_imag"), "y": ("quant_yt",)}


def _log2_f32(v) -> np.float32:
    """float32 log2, correctly rounded (evaluated in float64, rounded once)."""
    return F32(math.log2(float(F32(v)))) if float(F32(v)) > 0 else F32(-np.inf)


def _ceil_log2(v) -> int:
    l = _log2_f32(v)
    if not np.isfinite(l):
        raise OverflowError("log2 of a non-positive range: the reference fails on all-zero tensors as well")
    return int(math.ceil(float(l)))


def required_intbits(absmax) -> int:
    l = _log2_f32(absmax)
    if not np.isfinite(l):
        raise OverflowError("log2 of a non-positive range: the reference fails on all-zero tensors as well")
    bits = max(0, int(math.ceil(float(l))))
    return bits + (1 if float(np.round(l)) == float(l) else 0)


def _merge(a, b):
    if isinstance(a, dict) and isinstance(b, dict):
        return {k: (_merge(a[k], b[k]) if k in a and k in b else copy.deepcopy(a.get(k, b.get(k)))) for k in {**a, **b}}
    return copy.deepcopy(a)


def _walk(tree, fn, key=""):
    if isinstance(tree, dict):
        return {k: _walk(v, fn, k) for k, v in tree.items()}
    return fn(key, tree)


def _absmax_of_observer(obs: dict) -> np.float32:
    lo = next(v for k, v in obs.items() if k.endswith("observer_min"))
    hi = next(v for k, v in obs.items() if k.endswith("observer_max"))
    return F32(max(abs(float(F32(lo))), abs(float(F32(hi)))))


def _annotate_observers(tree, key=""):
    if isinstance(tree, dict):
        if key.endswith("observer"):
            out = dict(tree)
            out["absmax"] = _absmax_of_observer(tree)
            return out
        return {k: _annotate_observers(v, k) for k, v in tree.items()}
    return tree


def load_modeldict(params: dict, stats: dict) -> dict:
    """fxputils.py:121-134 on in-memory trees (use checkpoint.load_reference_pickles for the files)."""
    md = _merge(params, stats)
    with np.errstate(invalid="ignore", divide="ignore"):
        md = _walk(md, lambda k, v: np.log2(np.asarray(v, dtype=F32)) if "scale" in k else v)
    return _annotate_observers(md)


def _at(tree, path):
    for k in path:
        tree = tree[k]
    return tree


def _frac(log2_scale) -> int:
    return -int(math.ceil(float(F32(log2_scale))))


class _Entry(dict):
    """One tensor's requirements: absmax, intbits, fracbits (+ bits/exp once a precision is assigned)."""

    def assign(self, bits: int) -> "_Entry":
        self["bits"] = bits
        self["exp"] = min(self["fracbits"], bits - 1 - self["intbits"])
        return self


def _observed(node: dict) -> _Entry:
    am = float(node["observer"]["absmax"])
    return _Entry(absmax=am, intbits=required_intbits(am), fracbits=_frac(node["scale"]))


def _ssm_requirements(mixer: dict) -> dict:
    return {"weights": {k: _observed(_at(mixer, p)) for k, p in _SSM_WEIGHTS.items()},
            "activations": {k: _observed(_at(mixer, p)) for k, p in _SSM_ACTS.items()}}


def _dense_requirements(d: dict) -> dict:
    w = np.asarray(d["kernel"], dtype=F32)
    b = np.asarray(d["bias"], dtype=F32)
    req = {}
    for name, absmax, log2s in (("b", np.abs(b).max(), d["act_scale"]), ("w", np.abs(w).max(), d["weight_scale"]),
                                ("inp", d["input_observer"]["absmax"], d["act_scale"]),
                                ("out", d["output_observer"]["absmax"], d["out_scale"])):
        req[f"{name}_absmax"] = float(absmax)
        req[f"{name}_intbits"] = required_intbits(absmax)
        req[f"{name}_fracbits"] = _frac(log2s)
    return req


def _gate_requirements(g: dict) -> dict:
    req = {}
    for side, name in (("l", "quant_left"), ("r", "quant_right")):
        req[f"{side}_absmax"] = float(g[name]["observer"]["absmax"])
        req[f"{side}_intbits"] = required_intbits(g[name]["observer"]["absmax"])
        req[f"{side}_fracbits"] = _frac(g[name]["scale"])
    return req


def _max_over_layers(per_layer: list) -> dict:
    first = per_layer[0]
    if all(not isinstance(v, dict) for v in first.values()):
        return type(first)({k: max(d[k] for d in per_layer) for k in first})
    return {k: _max_over_layers([d[k] for d in per_layer]) for k in first}


def _assign_dense(req: dict, prec: dict) -> dict:
    for name, bits in (("w", prec["non_ssm_w"]), ("b", prec["non_ssm_b"]), ("inp", prec["non_ssm_act"]),
                       ("out", prec["non_ssm_act"])):
        req[f"{name}_bits"] = bits
        req[f"{name}_exp"] = min(req[f"{name}_fracbits"], bits - 1 - req[f"{name}_intbits"])
    return req


def _assign_gate(req: dict, prec: dict) -> dict:
    bits = prec["non_ssm_act"]
    for side in ("l", "r"):
        req[f"{side}_bits"] = bits
        req[f"{side}_exp"] = min(req[f"{side}_fracbits"], bits - 1 - req[f"{side}_intbits"])
    req["res_bits"] = bits
    req["res_exp"] = bits - 1 - (req["l_intbits"] + req["r_intbits"])  # fxputils.py:529-537
    return req


def _assign_ssm(req: dict, prec: dict) -> dict:
    for k, e in req["weights"].items():
        _Entry.assign(e, prec["ssm_act"] if k in ("A_re", "A_im") else prec["ssm_w"])
    for e in req["activations"].values():
        _Entry.assign(e, prec["ssm_act"])
    return req


def _norm_config(norms: list, prec: dict) -> dict:
    bits = prec["non_ssm_act"]

    def entry(arrays):
        absmax = max(float(np.abs(np.asarray(a, dtype=F32)).max()) for a in arrays)
        ib = max(0, _ceil_log2(absmax))
        return {"intbits": ib, "exp": bits - 1 - ib, "bits": bits}

    cfg = {"mean": entry([n["mean"] for n in norms]), "var": entry([n["var"] for n in norms]),
           "invsq_var": entry([F32(1.0) / np.sqrt(np.asarray(n["var"], dtype=F32) + F32(1e-5)) for n in norms])}
    for name in ("scale", "bias"):
        if name in norms[0]:
            for n in norms:  # in-place NaN -> 1.0, as the reference mutates the modeldict (fxputils.py:711-739)
                a = np.asarray(n[name], dtype=F32)
                n[name] = np.where(np.isnan(a), F32(1.0), a).astype(F32)
            cfg[name] = entry([n[name] for n in norms])
    return cfg


def create_fxp_qconfig(modeldict: dict, quantization: str = "w8a16", shared_exp: bool = True) -> dict:
    """create_fxp_qconfig + add_target_bits_exp (fxputils.py:351-401, 453-786) in one pass.

    NOTE: mutates ``modeldict`` exactly where the reference does (NaN replacement in the BatchNorm leaves).
    """
    prec = PRECISIONS[quantization]
    enc = modeldict["encoder"]
    layer_keys = sorted((k for k in enc if k.startswith("layers_")), key=lambda s: int(s.split("_")[1]))
    q = {"encoder": _assign_dense(_dense_requirements(enc["encoder"]), prec),
         "decoder": _assign_dense(_dense_requirements(modeldict["decoder"]), prec)}
    per_layer = {k: {"ssm": _ssm_requirements(enc[k]["mixer"]), "multgate": _gate_requirements(enc[k]["mult_gate"]),
                     "out2": _dense_requirements(enc[k]["out2"])} for k in layer_keys}
    if shared_exp:
        blocks = {part: _max_over_layers([per_layer[k][part] for k in layer_keys]) for part in ("ssm", "multgate", "out2")}
        blocks["ssm"] = _assign_ssm(blocks["ssm"], prec)
        blocks["multgate"] = _assign_gate(blocks["multgate"], prec)
        blocks["out2"] = _assign_dense(blocks["out2"], prec)
        blocks["norm"] = _norm_config([enc[k]["norm"] for k in layer_keys], prec)
    else:
        blocks = {}
        for k in layer_keys:
            b = per_layer[k]
            blocks[k] = {"ssm": _assign_ssm(b["ssm"], prec), "multgate": _assign_gate(b["multgate"], prec),
                         "out2": _assign_dense(b["out2"], prec), "norm": _norm_config([enc[k]["norm"]], prec)}
    q["blocks"] = blocks
    return q
