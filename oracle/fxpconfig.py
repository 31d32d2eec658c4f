"""Oracle: fixed-point quantisation config (NumPy restatement).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/sparseRNNs/fxputils.py`` (``fxputils.py:LINE``) and the ``precisions`` table of
``fxprun.py:302-308``.  Only the blockwise-shared-exponent configuration (``agg="max"``,
``shared_exp=True``; fxprun.py:314, 350-360 default) and the per-layer one (``agg=None``) are restated.
"""

from __future__ import annotations

import copy

import numpy as np

from .fxp import ceil_log2_f32

F32 = np.float32


def precisions_for(quantization: str = "w8a16") -> dict:
    """fxprun.py:302-308."""
    a16 = "a16" in quantization
    return dict(non_ssm_w=8, non_ssm_b=16 if a16 else 8, non_ssm_act=16 if a16 else 8,
                ssm_w=8, ssm_act=16 if a16 else 8)


def merge_trees(p, s):
    """fxputils.py:12-64: union of two nested dicts, params win on conflicting leaves."""
    if isinstance(p, dict) and isinstance(s, dict):
        out = {}
        for k in set(p) | set(s):
            pv, sv = p.get(k), s.get(k)
            if pv is not None and sv is not None:
                out[k] = merge_trees(pv, sv)
            else:
                out[k] = pv if pv is not None else sv
        return out
    return p if p is not None else s


def _map_with_key(tree, fn, key=""):
    if isinstance(tree, dict):
        return {k: _map_with_key(v, fn, k) for k, v in tree.items()}
    return fn(key, tree)


def _add_integer_bits(tree, key=""):
    """fxputils.py:67-80: every ``*observer`` dict gains ``absmax`` and ``intbits``."""
    if isinstance(tree, dict) and key.endswith("observer"):
        first = list(tree.keys())[0].split("_")[0]
        prefix = "" if first == "observer" else f"{first}_"
        absmax = F32(max(abs(F32(tree[f"{prefix}observer_min"])), abs(F32(tree[f"{prefix}observer_max"]))))
        tree = dict(tree)
        tree["absmax"] = absmax
        tree["intbits"] = ceil_log2_f32(absmax)
        return tree
    if isinstance(tree, dict):
        return {k: _add_integer_bits(v, k) for k, v in tree.items()}
    return tree


def make_modeldict(params: dict, stats: dict) -> dict:
    """``load_modeldict`` minus the file I/O (fxputils.py:121-134).

    Every leaf whose key contains ``"scale"`` is replaced by ``log2`` of itself - including the
    BatchNorm ``norm/scale`` (gamma) leaf, which the reference does not exempt (fxputils.py:128-131);
    negative gammas become NaN and are later replaced by 1.0 (:711-718).
    """
    md = merge_trees(copy.deepcopy(params), copy.deepcopy(stats))
    with np.errstate(invalid="ignore", divide="ignore"):
        md = _map_with_key(md, lambda k, x: np.log2(np.asarray(x, dtype=F32)) if "scale" in k else x)
    return _add_integer_bits(md)


def get_intbits(absmax) -> int:
    """fxputils.py:137-142: ceil(log2(absmax)), floored at 0, +1 at exact powers of two."""
    v = F32(absmax)
    l = F32(np.log2(np.float64(v)))
    bits = max(0, int(np.ceil(l)))
    bits += 1 if np.round(l) == l else 0
    return bits


def get_sign(xmin, xmax):
    """fxputils.py:145-152."""
    xmin, xmax = float(xmin), float(xmax)
    if xmin * xmax < 0:
        return 1
    return 0.75 if [xmin, xmax] == [abs(xmin), abs(xmax)] else 0.25


def _idx(tree, path: str):
    for k in path.split("/"):
        tree = tree[k]
    return tree


def _fracbits(log2scale) -> int:
    return int(-1 * np.ceil(F32(log2scale)))


def _ssm_cfg(md, key):
    """fxputils.py:155-203."""
    d = _idx(md, key)
    out = {"weights": {}, "activations": {}}
    wk = dict(A_re="quant_A/quant_real", A_im="quant_A/quant_imag", B_re="quant_B/quant_real",
              B_im="quant_B/quant_imag", C_re="quant_C/quant_real", C_im="quant_C/quant_imag", D="quant_D")
    ak = dict(u="quant_ut", Bu_re="quant_But/quant_real", Bu_im="quant_But/quant_imag",
              x_re="quant_xt/quant_real", x_im="quant_xt/quant_imag", y="quant_yt")
    for grp, keys in (("weights", wk), ("activations", ak)):
        for ks, kl in keys.items():
            obs = _idx(d, f"{kl}/observer")
            absmax = float(obs["absmax"])
            out[grp][ks] = dict(absmax=absmax, intbits=get_intbits(absmax),
                                signbits=get_sign(obs["observer_min"], obs["observer_max"]),
                                fracbits=_fracbits(_idx(d, f"{kl}/scale")))
    return out


def _multgate_cfg(md, key):
    """fxputils.py:206-236."""
    d = _idx(md, key)
    out = {}
    for side, name in (("l", "quant_left"), ("r", "quant_right")):
        obs = _idx(d, f"{name}/observer")
        out[f"{side}_absmax"] = float(obs["absmax"])
        out[f"{side}_signbits"] = get_sign(obs["observer_min"], obs["observer_max"])
        out[f"{side}_intbits"] = get_intbits(obs["absmax"])
        out[f"{side}_fracbits"] = _fracbits(_idx(d, f"{name}/scale"))
    return out


def _dense_cfg(md, key):
    """fxputils.py:239-285.  The bias fraction bits come from ``act_scale`` (:266)."""
    d = _idx(md, key)
    w = np.asarray(d["kernel"], dtype=F32)
    b = np.asarray(d["bias"], dtype=F32)
    io, oo = d["input_observer"], d["output_observer"]
    out = dict(
        b_absmax=float(np.abs(b).max()), b_signbit=get_sign(b.min(), b.max()),
        b_intbits=get_intbits(np.abs(b).max()), b_fracbits=_fracbits(d["act_scale"]),
        w_absmax=float(np.abs(w).max()), w_signbit=get_sign(w.min(), w.max()),
        w_intbits=get_intbits(np.abs(w).max()), w_fracbits=_fracbits(d["weight_scale"]),
        inp_absmax=float(io["absmax"]), inp_signbit=get_sign(io["input_observer_min"], io["input_observer_max"]),
        inp_intbits=get_intbits(io["absmax"]), inp_fracbits=_fracbits(d["act_scale"]),
        out_absmax=float(oo["absmax"]), out_signbit=get_sign(oo["output_observer_min"], oo["output_observer_max"]),
        out_intbits=get_intbits(oo["absmax"]), out_fracbits=_fracbits(d["out_scale"]),
    )
    return out


def _custom_max(k, v, start=0):
    """fxputils.py:288-293."""
    if k[start:] in ["signbits"]:
        return 1 if len(set(v)) >= 2 else max(v)
    return max(v)


def _join_flat(d, start):
    """fxputils.py:296-307 with agg='max'."""
    keys = d["layers_0"].keys()
    joint = {k: list(set(d[l][k] for l in d)) for k in keys}
    return {k: v[0] if len(v) == 1 else _custom_max(k, v, start) for k, v in joint.items()}


def _join_ssm(d):
    """fxputils.py:310-348 with agg='max'."""
    out = {"weights": {}, "activations": {}}
    for grp in out:
        for name in d["layers_0"][grp]:
            joint = {k: list(set(d[l][grp][name][k] for l in d)) for k in d["layers_0"][grp][name]}
            out[grp][name] = {k: v[0] if len(v) == 1 else _custom_max(k, v) for k, v in joint.items()}
    return out


def create_fxp_qconfig(md: dict, agg="max") -> dict:
    """fxputils.py:351-401 (returns the aggregated config only)."""
    data = {"encoder": _dense_cfg(md, "encoder/encoder"), "blocks": dict(ssm={}, multgate={}, out2={})}
    for key in md["encoder"]:
        if "layers_" in key:
            data["blocks"]["ssm"][key] = _ssm_cfg(md, f"encoder/{key}/mixer")
            data["blocks"]["multgate"][key] = _multgate_cfg(md, f"encoder/{key}/mult_gate")
            data["blocks"]["out2"][key] = _dense_cfg(md, f"encoder/{key}/out2")
    data["decoder"] = _dense_cfg(md, "decoder")
    if agg == "max":
        data["blocks"]["out2"] = _join_flat(data["blocks"]["out2"], 2)
        data["blocks"]["multgate"] = _join_flat(data["blocks"]["multgate"], 2)
        data["blocks"]["ssm"] = _join_ssm(data["blocks"]["ssm"])
        return data
    if agg is None:
        for key in list(data["blocks"]["ssm"].keys()):
            data["blocks"][key] = dict(ssm=data["blocks"]["ssm"][key], multgate=data["blocks"]["multgate"][key],
                                       out2=data["blocks"]["out2"][key])
        for k in ("ssm", "multgate", "out2"):
            del data["blocks"][k]
        return data
    raise ValueError("agg must be 'max' or None")


def _fit(exp, bits, intbits):
    return min(exp, bits - 1 - intbits)


def _dense_exp(c: dict, prec: dict) -> dict:
    """fxputils.py:404-450."""
    c["w_bits"] = prec["non_ssm_w"]
    c["w_exp"] = _fit(c["w_fracbits"], c["w_bits"], c["w_intbits"])
    c["b_bits"] = prec["non_ssm_b"]
    c["b_exp"] = _fit(c["b_fracbits"], c["b_bits"], c["b_intbits"])
    c["inp_bits"] = prec["non_ssm_act"]
    c["inp_exp"] = _fit(c["inp_fracbits"], c["inp_bits"], c["inp_intbits"])
    c["out_bits"] = prec["non_ssm_act"]
    c["out_exp"] = _fit(c["out_fracbits"], c["out_bits"], c["out_intbits"])
    return c


def _multgate_exp(c: dict, prec: dict) -> dict:
    """fxputils.py:478-550."""
    c["l_bits"] = c["r_bits"] = prec["non_ssm_act"]
    c["l_exp"] = _fit(c["l_fracbits"], c["l_bits"], c["l_intbits"])
    c["r_exp"] = _fit(c["r_fracbits"], c["r_bits"], c["r_intbits"])
    c["res_bits"] = prec["non_ssm_act"]
    c["res_exp"] = c["res_bits"] - 1 - (c["l_intbits"] + c["r_intbits"])
    return c


def _ssm_exp(c: dict, prec: dict) -> dict:
    """fxputils.py:552-633."""
    for key, v in c["weights"].items():
        v["bits"] = prec["ssm_act"] if key in ("A_re", "A_im") else prec["ssm_w"]
        v["exp"] = _fit(v["fracbits"], v["bits"], v["intbits"])
    for key, v in c["activations"].items():
        v["bits"] = prec["ssm_act"]
        v["exp"] = _fit(v["fracbits"], v["bits"], v["intbits"])
    return c


def _norm_entry(arrs, bits):
    absmax = max(F32(np.abs(np.asarray(e, dtype=F32)).max()) for e in arrs)
    intbits = max(0, ceil_log2_f32(absmax))
    return {"intbits": intbits, "exp": bits - 1 - intbits, "bits": bits}


def _norm_cfg(md: dict, layers, prec: dict) -> dict:
    """fxputils.py:635-784.  Mutates ``md``: NaNs in norm/scale and norm/bias become 1.0 (:711-739)."""
    bits = prec["non_ssm_act"]
    norms = [md["encoder"][l]["norm"] for l in layers]
    out = {"mean": _norm_entry([n["mean"] for n in norms], bits),
           "var": _norm_entry([n["var"] for n in norms], bits)}
    invsq = [F32(1.0) / np.sqrt(np.asarray(n["var"], dtype=F32) + F32(1e-5)) for n in norms]
    out["invsq_var"] = _norm_entry(invsq, bits)
    for name in ("scale", "bias"):
        if name in norms[0]:
            for n in norms:
                a = np.asarray(n[name], dtype=F32)
                n[name] = np.where(np.isnan(a), F32(1.0), a).astype(F32)
            out[name] = _norm_entry([n[name] for n in norms], bits)
    return out


def add_target_bits_exp(md: dict, q: dict, prec: dict, shared_exp=True) -> dict:
    """fxputils.py:453-786."""
    q["encoder"] = _dense_exp(q["encoder"], prec)
    q["decoder"] = _dense_exp(q["decoder"], prec)
    layers = [l for l in md["encoder"] if "layers_" in l]
    if shared_exp:
        b = q["blocks"]
        b["out2"] = _dense_exp(b["out2"], prec)
        b["multgate"] = _multgate_exp(b["multgate"], prec)
        b["ssm"] = _ssm_exp(b["ssm"], prec)
        b["norm"] = _norm_cfg(md, layers, prec)
    else:
        for key in list(q["blocks"].keys()):
            b = q["blocks"][key]
            b["out2"] = _dense_exp(b["out2"], prec)
            b["multgate"] = _multgate_exp(b["multgate"], prec)
            b["ssm"] = _ssm_exp(b["ssm"], prec)
            b["norm"] = _norm_cfg(md, [key], prec)
    return q


def build_fxp_qconfig(params: dict, stats: dict, quantization="w8a16", shared_exp=True):
    """fxprun.py:346-360: modeldict -> qconfig -> bits/exps.  Returns (modeldict, fxp_qconfig)."""
    md = make_modeldict(params, stats)
    q = create_fxp_qconfig(md, agg="max" if shared_exp else None)
    q = add_target_bits_exp(md, q, precisions_for(quantization), shared_exp=shared_exp)
    return md, q
