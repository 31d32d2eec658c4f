"""Oracle: S5 block and NDNS regression model, float and static-quant (NumPy restatement).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``model/layers.py:180-241`` (QSequenceLayer.__call__), ``model/ssm.py:525-622`` (qS5SSM.__call__),
``model/seq_model.py:100-115, 439-456`` (encoder stack / regression head) and
``utils/quantization.py:684-816`` (QuantizedDense), ``:481-517`` (QuantizedMultiply).

``variables`` is the Flax pytree of SURVEY.md section 5:
``{"params": {"encoder": {"encoder": ..., "layers_i": {"mixer","norm","out2","mult_gate"}}, "decoder": ...},
  "batch_stats": {"encoder": {"layers_i": {"norm": {"mean","var"}}}}}``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import ssm as S
from .quant import MinMax, W8A16, quantdequant

F32 = np.float32


@dataclass
class ModelConfig:
    n_layers: int = 3
    d_model: int = 96          # H
    ssm_size: int = 64         # P (complex states, after the conj_sym halving of train.py:99-101)
    d_input: int = 257
    d_output: int = 257
    relufication: bool = False
    static_quant: bool = False
    associative_scan: bool = True
    clip_eigs: bool = True     # recipes/ndns.json
    conj_sym: bool = True
    glu_variant: str = "half1"
    prenorm: bool = True
    batchnorm: bool = True
    discretization: str = "zoh"
    step_rescale: float = 1.0
    bn_eps: float = 1e-5       # flax nn.BatchNorm default
    quant_input: Optional[int] = None
    q: dict = None             # precision recipe (quantization.py:96-177)

    def __post_init__(self):
        if self.q is None:
            self.q = dict(W8A16)
        if self.glu_variant != "half1":
            raise NotImplementedError("only glu_variant='half1' (recipes/ndns.json) is restated")
        if not (self.prenorm and self.batchnorm):
            raise NotImplementedError("only prenorm + batchnorm (recipes/ndns.json) is restated")


def gelu_tanh(x):
    """jax.nn.gelu(approximate=True) (layers.py:124)."""
    dt = x.dtype.type
    c = dt(np.sqrt(2.0 / np.pi))
    return (x * (dt(0.5) * (dt(1.0) + np.tanh(c * (x + dt(0.044715) * (x ** 3)))))).astype(x.dtype)


def sigmoid(x):
    """jax.nn.sigmoid (layers.py:224)."""
    dt = x.dtype.type
    with np.errstate(over="ignore"):  # exp(-x) -> inf for very negative x gives exactly 0, as in XLA
        return (dt(1.0) / (dt(1.0) + np.exp(-x))).astype(x.dtype)


def batchnorm_eval(x, norm_params: dict, norm_stats: dict, eps: float):
    """flax nn.BatchNorm(use_running_average=True): y = (x - mean) * (rsqrt(var+eps) * scale) + bias."""
    dt = x.dtype.type
    mean = np.asarray(norm_stats["mean"], dtype=x.dtype)
    var = np.asarray(norm_stats["var"], dtype=x.dtype)
    mul = (dt(1.0) / np.sqrt(var + dt(eps))).astype(x.dtype)
    if "scale" in norm_params:
        mul = mul * np.asarray(norm_params["scale"], dtype=x.dtype)
    y = (x - mean) * mul
    if "bias" in norm_params:
        y = y + np.asarray(norm_params["bias"], dtype=x.dtype)
    return y.astype(x.dtype)


def dense_float(x, p: dict):
    return (x @ np.asarray(p["kernel"], dtype=x.dtype) + np.asarray(p["bias"], dtype=x.dtype)).astype(x.dtype)


def dense_static(x, p: dict, a_bits: int, w_bits: int):
    """QuantizedDense frozen branch (quantization.py:742-814).  Bias uses act_scale (:788)."""
    dt = x.dtype.type
    qx = quantdequant(x, p["act_scale"], a_bits, dt)
    qk = quantdequant(p["kernel"], p["weight_scale"], w_bits, dt)
    y = qx @ qk
    y = y + quantdequant(p["bias"], p["act_scale"], a_bits, dt)
    return quantdequant(y, p["out_scale"], a_bits, dt)


def _ssm_scales(mixer: dict):
    g = lambda *k: _get(mixer, k)["scale"]
    return dict(ut=g("quant_ut"), yt=g("quant_yt"), D=g("quant_D"),
                A_re=g("quant_A", "quant_real"), A_im=g("quant_A", "quant_imag"),
                B_re=g("quant_B", "quant_real"), B_im=g("quant_B", "quant_imag"),
                C_re=g("quant_C", "quant_real"), C_im=g("quant_C", "quant_imag"),
                But_re=g("quant_But", "quant_real"), But_im=g("quant_But", "quant_imag"))


def _get(d, keys):
    for k in keys:
        d = d[k]
    return d


def layer_forward(x, lp: dict, ls: dict, cfg: ModelConfig, dtype=F32, taps: Optional[dict] = None):
    """One S5 block on (..., L, H) (layers.py:180-241, non-fused BatchNorm branch :210-213)."""
    skip = x
    u = batchnorm_eval(x, lp.get("norm", {}), ls["norm"], cfg.bn_eps)
    Lb, Bb, Ct, D = S.ssm_discrete_params(lp["mixer"], clip_eigs=cfg.clip_eigs, step_rescale=cfg.step_rescale,
                                          discretization=cfg.discretization, dtype=dtype)
    if cfg.static_quant:
        sc = _ssm_scales(lp["mixer"])
        bits = dict(a=cfg.q["a_precision"], b=cfg.q["b_precision"], c=cfg.q["c_precision"],
                    ssm_act=cfg.q["ssm_act_precision"])
        ys, xs, u_q = S.apply_ssm_static(Lb, Bb, Ct, u, sc, bits, conj_sym=cfg.conj_sym,
                                         relufication=cfg.relufication, dtype=dtype)
        Dq = quantdequant(D, sc["D"], cfg.q["d_precision"], dtype)  # ssm.py:595
        y = quantdequant(ys + Dq * u_q, sc["yt"], cfg.q["ssm_act_precision"], dtype)  # ssm.py:596-598
    else:
        ys, xs = S.apply_ssm_float(Lb, Bb, Ct, u, conj_sym=cfg.conj_sym, relufication=cfg.relufication,
                                   associative_scan=cfg.associative_scan)
        y = (ys + D * u).astype(dtype)  # ssm.py:618-620
    x1 = np.maximum(y, 0) if cfg.relufication else gelu_tanh(y)  # layers.py:116-124, :217
    if cfg.static_quant:
        a_bits, w_bits = cfg.q["non_ssm_act_precision"], cfg.q["non_ssm_precision"]
        o = dense_static(x1, lp["out2"], a_bits, w_bits)
        g = (quantdequant(x1, lp["mult_gate"]["quant_left"]["scale"], a_bits, dtype)
             * quantdequant(sigmoid(o), lp["mult_gate"]["quant_right"]["scale"], a_bits, dtype))
    else:
        o = dense_float(x1, lp["out2"])
        g = x1 * sigmoid(o)
    out = (g + skip).astype(dtype)
    if cfg.relufication:
        out = np.maximum(out, 0)
    if taps is not None:
        taps.update(pre_s5=u, pre_C=xs, pre_GLU=y, out2=o, post_GLU=g, output=out)
    return out


def model_forward(x, variables: dict, cfg: ModelConfig, dtype=F32, taps: Optional[dict] = None):
    """QRegressionModel.__call__ on (..., L, d_input) -> (..., L, d_output) (seq_model.py:439-456)."""
    P = variables["params"]
    BS = variables["batch_stats"]
    x = np.asarray(x, dtype=dtype)
    if cfg.quant_input is not None:  # seq_model.py:14-20
        x = (np.round(x * dtype(2 ** cfg.quant_input)) / dtype(2 ** cfg.quant_input)).astype(dtype)
    enc = P["encoder"]
    if cfg.static_quant:
        a_bits, w_bits = cfg.q["non_ssm_act_precision"], cfg.q["non_ssm_precision"]
        h = dense_static(x, enc["encoder"], a_bits, w_bits)
    else:
        h = dense_float(x, enc["encoder"])
    if cfg.relufication:  # seq_model.py:73-76, :111
        h = np.maximum(h, 0)
    if taps is not None:
        taps["encoder_output"] = h
    for i in range(cfg.n_layers):
        lt = {} if taps is not None else None
        h = layer_forward(h, enc[f"layers_{i}"], BS["encoder"][f"layers_{i}"], cfg, dtype, lt)
        if taps is not None:
            taps[f"layers_{i}"] = lt
    if cfg.static_quant:
        y = dense_static(h, P["decoder"], a_bits, w_bits)
    else:
        y = dense_float(h, P["decoder"])
    if taps is not None:
        taps["pre_decoder"] = h
        taps["output"] = y
    return y


# --------------------------------------------------------------------------- calibration
def calibrate(x, variables: dict, cfg: ModelConfig, dtype=F32):
    """One-batch static-quant calibration.

    Runs the ``static_quant=True, calibrating=True`` graph: every FakeQuant / QuantizedDense passes
    values through unchanged and records min/max (quantization.py:428-436, 730-745, 793-811); the
    scan is sequential; no in-loop state quantisation (first calibration call: the ``quant_xt``
    observer does not exist yet when the loop is built, ssm.py:254, :314).

    Returns ``(params_with_scales, stats)`` where ``stats`` holds the observers in the layout of
    ``sc_cal_stats.pkl`` (batch_stats collection) and ``params_with_scales`` the layout of
    ``sc_calibrated_params.pkl`` (convert.py:963-971).
    """
    import copy

    P = copy.deepcopy(variables["params"])
    BS = variables["batch_stats"]
    stats = {"encoder": {}, "decoder": {}}
    q = cfg.q
    x = np.asarray(x, dtype=dtype)

    def fq(site_params: dict, site_stats: dict, val, bits: int):
        mm = MinMax()
        mm.observe(val)
        s = mm.scale(bits)
        site_params["scale"] = s
        site_stats["scale"] = s
        site_stats["observer"] = mm.as_dict()

    def fqc(site_params: dict, site_stats: dict, val, bits: int):
        for part, name in ((val.real, "quant_real"), (val.imag, "quant_imag")):
            fq(site_params.setdefault(name, {}), site_stats.setdefault(name, {}), part, bits)

    def qdense(p: dict, st: dict, inp):
        a_bits, w_bits = q["non_ssm_act_precision"], q["non_ssm_precision"]
        mi, mo, mw = MinMax(), MinMax(), MinMax()
        mi.observe(inp)
        k = np.asarray(p["kernel"], dtype=dtype)
        mw.observe(k)
        y = (inp @ k + np.asarray(p["bias"], dtype=dtype)).astype(dtype)
        mo.observe(y)
        for name, val in (("act_scale", mi.scale(a_bits)), ("weight_scale", mw.scale(w_bits)),
                          ("out_scale", mo.scale(a_bits))):
            p[name] = val
            st[name] = val
        st["input_observer"] = mi.as_dict("input_")
        st["output_observer"] = mo.as_dict("output_")
        return y

    enc = P["encoder"]
    st_enc = stats["encoder"].setdefault("encoder", {})
    h = qdense(enc["encoder"], st_enc, x)
    if cfg.relufication:
        h = np.maximum(h, 0)
    for i in range(cfg.n_layers):
        lp = enc[f"layers_{i}"]
        ls = BS["encoder"][f"layers_{i}"]
        st = stats["encoder"].setdefault(f"layers_{i}", {})
        st["norm"] = {k: v for k, v in ls["norm"].items()}
        mp = lp["mixer"]
        ms = st.setdefault("mixer", {})
        skip = h
        u = batchnorm_eval(h, lp.get("norm", {}), ls["norm"], cfg.bn_eps)
        Lb, Bb, Ct, D = S.ssm_discrete_params(mp, clip_eigs=cfg.clip_eigs, step_rescale=cfg.step_rescale,
                                              discretization=cfg.discretization, dtype=dtype)
        fq(mp.setdefault("quant_ut", {}), ms.setdefault("quant_ut", {}), u, q["ssm_act_precision"])
        fqc(mp.setdefault("quant_B", {}), ms.setdefault("quant_B", {}), Bb, q["b_precision"])
        fqc(mp.setdefault("quant_A", {}), ms.setdefault("quant_A", {}), Lb, q["a_precision"])
        fqc(mp.setdefault("quant_C", {}), ms.setdefault("quant_C", {}), Ct, q["c_precision"])
        cd = Lb.dtype
        Bu = (u @ Bb.real.T + 1j * (u @ Bb.imag.T)).astype(cd)
        fqc(mp.setdefault("quant_But", {}), ms.setdefault("quant_But", {}), Bu, q["ssm_act_precision"])
        xs = S.scan_sequential(Lb, Bu)
        fqc(mp.setdefault("quant_xt", {}), ms.setdefault("quant_xt", {}), xs, q["ssm_act_precision"])
        if cfg.relufication:
            xs = S.relu_complex(xs)
        ys = xs.real @ Ct.real.T - xs.imag @ Ct.imag.T
        if cfg.conj_sym:
            ys = 2 * ys
        fq(mp.setdefault("quant_D", {}), ms.setdefault("quant_D", {}), D, q["d_precision"])
        y = (ys + D * u).astype(dtype)
        fq(mp.setdefault("quant_yt", {}), ms.setdefault("quant_yt", {}), y, q["ssm_act_precision"])
        x1 = np.maximum(y, 0) if cfg.relufication else gelu_tanh(y)
        o = qdense(lp["out2"], st.setdefault("out2", {}), x1)
        sg = sigmoid(o)
        mg_p = lp.setdefault("mult_gate", {})
        mg_s = st.setdefault("mult_gate", {})
        fq(mg_p.setdefault("quant_left", {}), mg_s.setdefault("quant_left", {}), x1, q["non_ssm_act_precision"])
        fq(mg_p.setdefault("quant_right", {}), mg_s.setdefault("quant_right", {}), sg, q["non_ssm_act_precision"])
        h = (x1 * sg + skip).astype(dtype)
        if cfg.relufication:
            h = np.maximum(h, 0)
    qdense(P["decoder"], stats["decoder"], h)
    return P, stats
