"""Oracle: integer-only S5 NDNS model (NumPy restatement).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/sparseRNNs/fxpmodel.py`` (``fxpmodel.py:LINE``).  Only the configuration that is
reachable in the reference is restated: prenorm BatchNorm (not fused),
relufication, glu_variant "half1", ``use_lax_scan=True``, conj_sym, ZOH.
"""

from __future__ import annotations

from typing import Optional

import numpy as np

from . import ssm as S
from .fxp import (ComplexFxpArray, FxpArray, I32, RoundingMode, fxp_add, fxp_from_fp, fxp_matmul, fxp_mul,
                  fxp_relu, fxp_sub, sar, shl, wrap_add, wrap_mul)

F32 = np.float32


# --------------------------------------------------------------------------- sigmoid LUT
class FxpSigmoid:
    """fxpmodel.py:70-144: 8-entry LUT with linear interpolation on |x|, mirrored for x <= 0."""

    def __init__(self, x_exp: int = 6, y_exp: int = 8, x_extra: int = 3, n_exp: int = 3):
        self.x_exp, self.y_exp, self.x_extra, self.n_exp = x_exp, y_exp, x_extra, n_exp
        x = np.linspace(0, 1 << (x_exp + x_extra), (1 << n_exp) + 1, dtype=F32)[:-1].astype(np.int64)
        xf = (x.astype(F32) / F32(1 << x_exp)).astype(F32)
        sig = (F32(1.0) / (F32(1.0) + np.exp(-xf))).astype(F32)
        self.lut = (np.round(sig * F32(1 << y_exp)) - F32(1 << (y_exp - 1))).astype(np.int64).astype(I32)

    def half(self, xx):
        """fxpmodel.py:97-119 on non-negative ``xx`` (int32, wrapping products)."""
        delta = 1 << self.x_exp
        n_sub_2 = (1 << self.n_exp) - 2
        ind = np.minimum(sar(xx, self.x_exp), n_sub_2)
        mu = np.bitwise_and(xx, I32(delta - 1))
        a = sar(wrap_mul(wrap_add(-mu, delta), self.lut[ind]), self.x_exp)
        b = sar(wrap_mul(mu, self.lut[ind + 1]), self.x_exp)
        return wrap_add(a, b)

    def apply(self, x: FxpArray) -> FxpArray:
        """fxpmodel.py:121-144 for FxpArray input / fxp output."""
        xx = x.change_exp(self.x_exp).data
        sign = (2 * (xx > 0).astype(np.int64) - 1).astype(I32)
        ax = np.abs(xx.astype(np.int64)).astype(I32)  # |INT_MIN| wraps like jnp.abs
        yy = wrap_add(1 << (self.y_exp - 1), wrap_mul(sign, self.half(ax)))
        return FxpArray(yy, x.bits, self.y_exp, True)


# --------------------------------------------------------------------------- recurrence
def recurrent_loop(Bu_re, Bu_im, A_re, A_im, eBu_re, eBu_im, eA_re, eA_im, ex_re, ex_im):
    """fxpmodel.py:147-208.  Bu_*: (..., L, P) int32; A_*: (P,) int32.  Returns (xs_re, xs_im).

    Each of the four products is shifted on its own before the add/sub (floor per product), the
    state is a wrapping int32 with NO saturation, and x_0 = 0.
    """
    Bu_re = np.asarray(Bu_re, dtype=I32)
    Bu_im = np.asarray(Bu_im, dtype=I32)
    bre = sar(Bu_re, eBu_re - ex_re) if eBu_re > ex_re else shl(Bu_re, ex_re - eBu_re)
    bim = sar(Bu_im, eBu_im - ex_im) if eBu_im > ex_im else shl(Bu_im, ex_im - eBu_im)
    xr = np.zeros(Bu_re.shape[:-2] + Bu_re.shape[-1:], dtype=I32)
    xi = np.zeros_like(xr)
    out_r = np.empty_like(Bu_re)
    out_i = np.empty_like(Bu_im)
    for t in range(Bu_re.shape[-2]):
        ar = wrap_add(sar(wrap_mul(A_re, xr), eA_re), -sar(wrap_mul(A_im, xi), eA_re).astype(np.int64))
        ai = wrap_add(sar(wrap_mul(A_re, xi), eA_im), sar(wrap_mul(A_im, xr), eA_im))
        xr = wrap_add(ar, bre[..., t, :])
        xi = wrap_add(ai, bim[..., t, :])
        out_r[..., t, :] = xr
        out_i[..., t, :] = xi
    return out_r, out_i


# --------------------------------------------------------------------------- modules
class FxpDense:
    """fxpmodel.py:291-366."""

    def __init__(self, md: dict, q: dict):
        self.q = q
        self.weight = fxp_from_fp(md["kernel"], bits=q["w_bits"], exp=q["w_exp"], round_mode=RoundingMode.ROUND)
        self.bias = fxp_from_fp(md["bias"], bits=q["b_bits"], exp=q["b_exp"], round_mode=RoundingMode.ROUND)

    def __call__(self, x: FxpArray) -> FxpArray:
        q = self.q
        if x.bits > q["inp_bits"] or x.exp > q["inp_exp"]:  # fxpmodel.py:335-347
            x = x.change_cfg(q["inp_bits"], q["inp_exp"], True)
        wx = fxp_matmul(x, self.weight, result_bits=q["out_bits"], result_exp=q["out_exp"])
        return fxp_add(wx, self.bias, result_bits=q["out_bits"], result_exp=q["out_exp"])


class FxpSSM:
    """fxpmodel.py:396-794 (non-fused-BatchNorm branch; the fused one is unreachable, :537-549)."""

    def __init__(self, md: dict, q: dict, step_rescale: float = 1.0, discretization: str = "zoh"):
        self.q = q
        # clip_eigs is asserted False for the fxp model (fxpmodel.py:434, :440-441)
        Lb, Bb, Ct, D = S.ssm_discrete_params(md, clip_eigs=False, step_rescale=step_rescale,
                                              discretization=discretization, dtype=F32)
        w = q["weights"]
        ff = lambda v, k: fxp_from_fp(v, bits=w[k]["bits"], exp=w[k]["exp"], round_mode=RoundingMode.ROUND)
        self.A = ComplexFxpArray(ff(Lb.real, "A_re"), ff(Lb.imag, "A_im"))
        self.B = ComplexFxpArray(ff(Bb.real, "B_re"), ff(Bb.imag, "B_im"))
        self.C = ComplexFxpArray(ff(Ct.real, "C_re"), ff(Ct.imag, "C_im"))
        self.D = ff(D, "D")

    def __call__(self, u: FxpArray, taps: Optional[dict] = None):
        a = self.q["activations"]
        u = u.change_cfg(a["u"]["bits"], a["u"]["exp"], True)  # :620-624
        Bu = ComplexFxpArray(
            fxp_matmul(u, self.B.real.transpose(), result_exp=a["Bu_re"]["exp"], result_bits=a["Bu_re"]["bits"]),
            fxp_matmul(u, self.B.imag.transpose(), result_exp=a["Bu_im"]["exp"], result_bits=a["Bu_im"]["bits"]))
        xr, xi = recurrent_loop(Bu.real.data, Bu.imag.data, self.A.real.data, self.A.imag.data,
                                Bu.real.exp, Bu.imag.exp, self.A.real.exp, self.A.imag.exp,
                                a["x_re"]["exp"], a["x_im"]["exp"])
        xs = ComplexFxpArray(FxpArray(xr, a["x_re"]["bits"], a["x_re"]["exp"], True),
                             FxpArray(xi, a["x_im"]["bits"], a["x_im"]["exp"], True))
        xs_raw = xs
        xs = fxp_relu(xs)  # :740-742
        ye, yb = a["y"]["exp"], a["y"]["bits"]
        ys = fxp_sub(fxp_matmul(xs.real, self.C.real.transpose(), result_exp=ye, result_bits=yb),
                     fxp_matmul(xs.imag, self.C.imag.transpose(), result_exp=ye, result_bits=yb),
                     result_exp=ye, result_bits=yb)  # :746-763
        ys = FxpArray(wrap_mul(ys.data, 2), ys.bits, ys.exp, ys.signed)  # :765-767, after the clip
        Du = fxp_mul(self.D, u, result_exp=ye, result_bits=yb)  # :771-776
        y = fxp_add(ys, Du, result_exp=ye, result_bits=yb)  # :778-783
        if taps is not None:
            taps.update(u=u, Bu=Bu, xs=xs_raw, xs_relu=xs, two_Cxs=ys, Du=Du, ys=y)
        return y, xs


class FxpBatchNorm:
    """fxpmodel.py:850-944.  Four ``compute_best`` ops; parameters rounded half-even."""

    def __init__(self, md: dict, q: dict, bn_eps: float = 1e-5):
        ff = lambda v, k: fxp_from_fp(v, bits=q[k]["bits"], exp=q[k]["exp"], round_mode=RoundingMode.ROUND)
        self.minus_mean = ff(F32(-1) * np.asarray(md["mean"], dtype=F32), "mean")
        self.invsq_var = ff(F32(1.0) / np.sqrt(np.asarray(md["var"], dtype=F32) + F32(bn_eps)), "invsq_var")
        self.bias = ff(md["bias"], "bias") if "bias" in md else None
        self.scale = ff(md["scale"], "scale") if "scale" in md else None

    def __call__(self, x: FxpArray, taps: Optional[dict] = None) -> FxpArray:
        x1 = fxp_add(x, self.minus_mean, result_exp="compute_best")
        x2 = fxp_mul(x1, self.invsq_var, result_exp="compute_best")
        x3 = fxp_mul(x2, self.scale, result_exp="compute_best") if self.scale is not None else x2
        x4 = fxp_add(x3, self.bias, result_exp="compute_best") if self.bias is not None else x3
        if taps is not None:
            taps.update(norm_input_minus_mean=x1, norm_output_raw=x2, norm_output_scaled=x3, norm_output=x4)
        return x4


class FxpSequenceLayer:
    """fxpmodel.py:971-1161."""

    def __init__(self, md: dict, q: dict, layer_idx: int, step_rescale: float = 1.0):
        if any(k.startswith("layers_") for k in q):  # :988-993 (per-layer exponents)
            q = q[f"layers_{layer_idx}"]
        self.q = q
        self.norm = FxpBatchNorm(md["norm"], q["norm"])
        self.mixer = FxpSSM(md["mixer"], q["ssm"], step_rescale=step_rescale)
        self.out2 = FxpDense(md["out2"], q["out2"])
        sx = min(q["out2"]["out_exp"], 6)  # :1097-1103
        self.sigmoid = FxpSigmoid(x_exp=sx, y_exp=q["out2"]["out_bits"] - 2)

    def mult_gate(self, x: FxpArray, y: FxpArray) -> FxpArray:  # :1075-1093
        m = self.q["multgate"]
        x = x.change_cfg(m["l_bits"], m["l_exp"], True)
        y = y.change_cfg(m["r_bits"], m["r_exp"], True)
        return fxp_mul(x, y, result_exp=m["res_exp"], result_bits=m["res_bits"], round_mode=RoundingMode.FLOOR)

    def __call__(self, x: FxpArray, taps: Optional[dict] = None) -> FxpArray:
        skip = x
        t_norm = {} if taps is not None else None
        t_ssm = {} if taps is not None else None
        u = self.norm(x, t_norm)
        y, xs = self.mixer(u, t_ssm)
        x1 = fxp_relu(y)
        o = self.out2(x1)
        r = self.sigmoid.apply(o)
        g = self.mult_gate(x1, r)
        res = fxp_add(g, skip, result_exp="compute_best", result_bits=self.q["multgate"]["res_bits"])  # :1147-1152
        out = fxp_relu(res)
        if taps is not None:
            taps.update(norm=t_norm, mixer=t_ssm, pre_s5=u, pre_C=xs, pre_GLU=y, out2=o, out2_sigmoid=r,
                        post_GLU=g, residadd=res, output=out)
        return out


class FxpRegressionModel:
    """fxpmodel.py:1210-1271, 1380-1439."""

    def __init__(self, modeldict: dict, fxp_qconfig: dict, n_layers: int):
        self.q = fxp_qconfig
        enc = modeldict["encoder"]
        self.encoder = FxpDense(enc["encoder"], fxp_qconfig["encoder"])
        self.layers = [FxpSequenceLayer(enc[f"layers_{i}"], fxp_qconfig["blocks"], i) for i in range(n_layers)]
        self.decoder = FxpDense(modeldict["decoder"], fxp_qconfig["decoder"])

    def quantize_input(self, x) -> FxpArray:
        """fxprun.py:69-75."""
        e = self.q["encoder"]
        return fxp_from_fp(x, bits=e["inp_bits"], exp=e["inp_exp"], signed=True, round_mode=RoundingMode.FLOOR)

    def __call__(self, x: FxpArray, taps: Optional[dict] = None) -> FxpArray:
        h = fxp_relu(self.encoder(x))  # :1263-1266
        if taps is not None:
            taps["encoder_output_relu"] = h
        for i, layer in enumerate(self.layers):
            lt = {} if taps is not None else None
            h = layer(h, lt)
            if taps is not None:
                taps[f"layers_{i}"] = lt
        y = self.decoder(h)
        if taps is not None:
            taps["output"] = y
        return y
