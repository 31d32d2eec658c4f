"""Oracle: fixed-point tensor type and ops (NumPy restatement).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/sparseRNNs/fxparray.py`` (cited per function as ``fxparray.py:LINE``).

Pinned JAX semantics (JAX default 32-bit mode):
  * payload dtype int32; ``+ - *`` and ``@`` wrap modulo 2**32;
  * ``>>`` is an arithmetic shift, ``<<`` drops bits;
  * ``jnp.round`` is round-half-to-even;
  * ``to_float`` is ``int32 -> float32`` (round-to-nearest-even) divided by ``2**exp`` (exact);
  * data-dependent exponents use float32 ``log2`` then ``ceil``; here ``log2`` is the
    correctly rounded float32 value (computed in float64 and rounded once), which is the
    definition the CUDA path implements as well ("compute_best").
Only ``signed=True`` arrays are reachable on the NDNS path; unsigned ones are supported in
``minval/maxval`` for completeness.
"""

from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional, Union

import numpy as np

I32 = np.int32
F32 = np.float32


class RoundingMode(Enum):  # fxparray.py:13-17
    FLOOR = 0
    CEIL = 1
    ROUND = 2
    STOCHASTIC = 3


# --------------------------------------------------------------------------- helpers
def _i32(x) -> np.ndarray:
    """Wrap anything integer-like to int32 with two's-complement wrap-around."""
    a = np.asarray(x)
    if a.dtype == I32:
        return a
    return a.astype(np.int64).astype(I32)  # int64 -> int32 keeps the low 32 bits


def wrap_mul(a, b) -> np.ndarray:
    """int32 * int32 -> int32 with wrap (what jnp gives in 32-bit mode)."""
    return (np.asarray(a, dtype=np.int64) * np.asarray(b, dtype=np.int64)).astype(I32)


def wrap_add(a, b) -> np.ndarray:
    return (np.asarray(a, dtype=np.int64) + np.asarray(b, dtype=np.int64)).astype(I32)


def shl(a, n: int) -> np.ndarray:
    """int32 << n with bits shifted out dropped (n in [0, 31])."""
    if n < 0 or n > 31:
        raise ValueError(f"left shift by {n} is outside the defined range")
    return (np.asarray(a, dtype=np.int64) << n).astype(I32)


def sar(a, n: int) -> np.ndarray:
    """arithmetic int32 >> n (n in [0, 31])."""
    if n < 0 or n > 31:
        raise ValueError(f"right shift by {n} is outside the defined range")
    return np.asarray(a, dtype=I32) >> I32(n)


def ceil_log2_f32(v) -> int:
    """``int(ceil(log2(v)))`` with v and log2 in float32 (correctly rounded log2).

    Mirrors ``int(np.ceil(np.log2(x)))`` on a float32 scalar (fxparray.py:424, :428, :606;
    fxputils.py:139).  v == 0 gives -inf in the reference (and OverflowError in ``int``);
    callers on the NDNS path always add an epsilon first.
    """
    v32 = F32(v)
    if not v32 > 0:
        raise OverflowError("ceil(log2(v)) of a non-positive value")
    l = F32(np.log2(np.float64(v32)))
    return int(np.ceil(l))


def matmul_i32(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    """int32 @ int32 -> int32 with wrap-around (fxparray.py:662).

    Uses an exact float64 BLAS product when every partial sum provably fits in 2**53,
    otherwise an int64 product; both are then wrapped to int32.
    """
    a = np.asarray(a)
    w = np.asarray(w)
    k = a.shape[-1]
    amax = int(np.abs(a.astype(np.int64)).max(initial=0))
    wmax = int(np.abs(w.astype(np.int64)).max(initial=0))
    if amax * wmax * max(k, 1) < (1 << 52):
        r = a.astype(np.float64) @ w.astype(np.float64)
        return r.astype(np.int64).astype(I32)
    return (a.astype(np.int64) @ w.astype(np.int64)).astype(I32)


def round_array(x, round_mode: RoundingMode = RoundingMode.FLOOR) -> np.ndarray:
    """float -> int32 with the given rounding (fxparray.py:20-30).

    Out-of-range float->int conversion is implementation-defined in XLA; the oracle (and the
    CUDA path) saturate, which only matters for inputs that the following ``fxp_clip`` would
    saturate anyway.
    """
    x = np.asarray(x, dtype=F32)
    if round_mode == RoundingMode.ROUND:
        r = np.round(x)  # half to even
    elif round_mode == RoundingMode.CEIL:
        r = np.ceil(x)
    elif round_mode == RoundingMode.FLOOR:
        r = np.floor(x)
    else:
        raise NotImplementedError(f"rounding mode '{round_mode}' not implemented")
    r = np.clip(r.astype(np.float64), -2147483648.0, 2147483647.0)
    return r.astype(np.int64).astype(I32)


# --------------------------------------------------------------------------- types
@dataclass
class FxpArray:  # fxparray.py:33-167
    data: Optional[np.ndarray] = None
    bits: int = 16
    exp: int = 8
    signed: bool = True

    def __post_init__(self):
        if self.data is not None:
            self.data = _i32(self.data)

    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    def copy(self) -> "FxpArray":
        return FxpArray(self.data.copy(), self.bits, self.exp, self.signed)

    def minval(self) -> int:
        return fxp_minval(self)

    def maxval(self) -> int:
        return fxp_maxval(self)

    def clip(self) -> "FxpArray":
        return fxp_clip(self)

    def to_float(self) -> np.ndarray:  # fxparray.py:72-73
        return self.data.astype(F32) / F32(1 << self.exp)

    def change_exp(self, new_exp, round_mode=RoundingMode.FLOOR, **_) -> "FxpArray":
        return fxp_change_exp(self, new_exp, round_mode)

    def change_cfg(self, new_bits, new_exp, new_signed=True, round_mode=RoundingMode.FLOOR, **_):
        return fxp_change_cfg(self, new_bits, new_exp, new_signed, round_mode)

    def transpose(self) -> "FxpArray":
        return FxpArray(self.data.T, self.bits, self.exp, self.signed)

    def __getitem__(self, key) -> "FxpArray":
        return FxpArray(self.data[key], self.bits, self.exp, self.signed)

    def same_as(self, other: "FxpArray") -> bool:  # fxparray.py:100-109 (__eq__)
        return (
            self.bits == other.bits
            and self.exp == other.exp
            and self.signed == other.signed
            and np.array_equal(self.data, other.data)
        )

    def __add__(self, other):
        return fxp_add(self, other)

    def __sub__(self, other):  # fxparray.py:114-121
        return fxp_add(self, FxpArray(wrap_mul(other.data, -1), other.bits, other.exp, other.signed))

    def __mul__(self, other):
        return fxp_mul(self, other)

    def __matmul__(self, other):
        return fxp_matmul(self, other)

    @staticmethod
    def Zero(shape=(1,), bits=16, exp=8, signed=True):
        return FxpArray(np.zeros(shape, dtype=I32), bits, exp, signed)


@dataclass
class ComplexFxpArray:  # fxparray.py:170-229
    real: FxpArray
    imag: FxpArray

    @property
    def shape(self):
        return self.real.shape

    @property
    def ndim(self):
        return self.real.ndim

    def copy(self):
        return ComplexFxpArray(self.real.copy(), self.imag.copy())

    def to_float(self):
        return self.real.to_float() + 1j * self.imag.to_float()

    def __getitem__(self, key):
        return ComplexFxpArray(self.real[key], self.imag[key])


# --------------------------------------------------------------------------- ops
def fxp_minval(arr: FxpArray) -> int:  # fxparray.py:329-330
    return -(1 << (arr.bits - 1)) if arr.signed else 0


def fxp_maxval(arr: FxpArray) -> int:  # fxparray.py:333-334
    return (1 << (arr.bits - 1)) - 1 if arr.signed else (1 << arr.bits) - 1


def fxp_clip(arr: FxpArray) -> FxpArray:  # fxparray.py:346-357 (saturate to own bit width)
    lo, hi = fxp_minval(arr), fxp_maxval(arr)
    if arr.bits > 32 or (not arr.signed and arr.bits > 31):
        raise OverflowError("clip bounds do not fit int32 (the reference raises here as well)")
    lo = max(lo, -(1 << 31))
    hi = min(hi, (1 << 31) - 1)
    return FxpArray(np.clip(arr.data, lo, hi).astype(I32), arr.bits, arr.exp, arr.signed)


def fxp_rshift_round(x, rshift: int, round_mode: RoundingMode = RoundingMode.FLOOR) -> np.ndarray:
    """fxparray.py:274-284.  The bias add wraps in int32 like the reference."""
    x = _i32(x)
    if round_mode == RoundingMode.FLOOR:
        return sar(x, rshift)
    if round_mode == RoundingMode.CEIL:
        return sar(wrap_add(x, (1 << rshift) - 1), rshift)
    if round_mode == RoundingMode.ROUND:
        if rshift < 1:
            raise ValueError("ROUND needs rshift >= 1 (1 << (rshift-1) is undefined otherwise)")
        return sar(wrap_add(x, 1 << (rshift - 1)), rshift)
    raise NotImplementedError("STOCHASTIC rounding mode not implemented")


def fxp_from_fp(x, bits=16, exp=8, signed=True, round_mode=RoundingMode.FLOOR, **_) -> FxpArray:
    """fxparray.py:287-307: ``x * 2**exp`` in float32, round, saturate."""
    xint = np.asarray(x, dtype=F32) * F32(1 << exp)
    if not signed and np.any(xint < 0):
        xint = np.abs(xint)
    return fxp_clip(FxpArray(round_array(xint, round_mode), bits, exp, signed))


def fxp_change_exp(arr: FxpArray, new_exp: int, round_mode=RoundingMode.FLOOR, **_) -> FxpArray:
    """fxparray.py:310-326.  Saturates to the array's *current* bits after the shift."""
    if new_exp == arr.exp:
        return arr.copy()
    if new_exp > arr.exp:
        data = shl(arr.data, new_exp - arr.exp)
    else:
        data = fxp_rshift_round(arr.data, arr.exp - new_exp, round_mode)
    return fxp_clip(FxpArray(data, arr.bits, new_exp, arr.signed))


def fxp_change_cfg(x: FxpArray, new_bits, new_exp, new_signed=True, round_mode=RoundingMode.FLOOR, **_):
    """fxparray.py:232-271.  change_exp first (saturating at the OLD width), then narrow."""
    if x.bits == new_bits and x.exp == new_exp and x.signed == new_signed:
        return x
    new = fxp_change_exp(x, new_exp, round_mode)
    if new.bits > new_bits:
        new.bits = new_bits
        new = fxp_clip(new)
    else:
        new.bits = new_bits
    if (not new.signed) and new_signed:
        new.signed = new_signed
        new = fxp_clip(new)
    else:
        new.signed = new_signed
    return new


def _neg(op: FxpArray) -> FxpArray:
    return FxpArray(wrap_mul(op.data, -1), op.bits, op.exp, op.signed)


def compute_best_add_plan(op1: FxpArray, op2: FxpArray, result_bits: int, result_signed: bool = True):
    """The data-dependent integers of ``fxp_add(..., result_exp="compute_best")``.

    fxparray.py:420-448.  Returns (result_exp, agg_bits, agg_exp).
    """
    f1 = op1.to_float()
    f2 = op2.to_float()
    m_sum = F32(np.abs(f1 + f2).max())
    intbits = max(0, ceil_log2_f32(F32(m_sum + F32(1e-6))))
    sgn = 1 if result_signed else 0
    result_exp = result_bits - intbits - sgn
    m1 = F32(np.abs(f1).max())
    m2 = F32(np.abs(f2).max())
    agg_intbits = max(
        max(0, ceil_log2_f32(F32(m1 + F32(1e-8)))),
        max(0, ceil_log2_f32(F32(m2 + F32(1e-8)))),
    )
    agg_exp = max(op1.exp, op2.exp)
    agg_bits = agg_intbits + agg_exp + sgn
    return result_exp, agg_bits, agg_exp


def fxp_add(
    op1: FxpArray,
    op2: FxpArray,
    result_bits: Optional[int] = None,
    result_bits_fn=max,
    result_bits_add: int = 0,
    result_exp: Union[None, int, str] = None,
    round_mode: RoundingMode = RoundingMode.FLOOR,
    **_,
) -> FxpArray:
    """fxparray.py:386-466, including the operator-precedence quirk of the default branch."""
    result_signed = op1.signed or op2.signed
    if result_bits is None:
        result_bits = result_bits_fn(op1.bits, op2.bits) + result_bits_add

    if result_exp is None:
        if op1.exp == op2.exp:
            result_exp = op1.exp
            data = wrap_add(op1.data, op2.data)
        elif op1.exp > op2.exp:
            # fxparray.py:416  `a + b << k` parses as `(a + b) << k`
            result_exp = op1.exp
            data = shl(wrap_add(op1.data, op2.data), op1.exp - op2.exp)
        else:
            # fxparray.py:419  `a << (k + b)`: element-wise shift count k + b.data.
            # XLA shift_left yields 0 for counts outside [0, 31].
            result_exp = op2.exp
            cnt = (op2.exp - op1.exp) + op2.data.astype(np.int64)
            a64 = np.broadcast_to(op1.data.astype(np.int64), np.broadcast(op1.data, cnt).shape)
            ok = (cnt >= 0) & (cnt <= 31)
            data = np.where(ok, a64 << np.where(ok, cnt, 0), 0).astype(np.int64).astype(I32)
    elif result_exp == "compute_best":
        result_exp, agg_bits, agg_exp = compute_best_add_plan(op1, op2, result_bits, result_signed)
        op1c = fxp_change_cfg(op1, max(agg_bits, op1.bits), agg_exp, result_signed)
        op2c = fxp_change_cfg(op2, max(agg_bits, op2.bits), agg_exp, result_signed)
        data = wrap_add(op1c.data, op2c.data)
        exp_change = result_exp - agg_exp
        if exp_change > 0:
            data = shl(data, exp_change)
        elif exp_change < 0:
            data = sar(data, -exp_change)
    else:
        data = wrap_add(
            fxp_change_exp(op1, result_exp, round_mode).data,
            fxp_change_exp(op2, result_exp, round_mode).data,
        )
    return fxp_clip(FxpArray(data, result_bits, int(result_exp), result_signed))


def fxp_sub(op1, op2, **kw) -> FxpArray:  # fxparray.py:360-383
    return fxp_add(op1, _neg(op2), **kw)


def compute_best_mul_exp(op1: FxpArray, op2: FxpArray, result_bits: int, result_signed: bool = True) -> int:
    """fxparray.py:601-608."""
    m = F32(np.abs(op1.to_float() * op2.to_float()).max())
    intbits = max(0, ceil_log2_f32(F32(F32(1e-6) + m)))
    return result_bits - intbits - (1 if result_signed else 0)


def fxp_mul(
    op1: FxpArray,
    op2: FxpArray,
    result_exp: Union[None, int, str] = None,
    result_exp_fn=max,
    result_bits: Optional[int] = None,
    result_bits_fn=max,
    round_mode: RoundingMode = RoundingMode.FLOOR,
    **_,
) -> FxpArray:
    """fxparray.py:573-637.  int32 product (wraps), rshift-round, saturate.

    The int64 promotion at :613-616 is a no-op in 32-bit JAX and is not restated.
    """
    result_signed = op1.signed or op2.signed
    if result_bits is None:
        result_bits = result_bits_fn(op1.bits, op2.bits)
    if result_exp is None:
        result_exp = result_exp_fn(op1.exp, op2.exp)
    elif result_exp == "compute_best":
        result_exp = compute_best_mul_exp(op1, op2, result_bits, result_signed)
    rshift = op1.exp + op2.exp - result_exp
    if rshift < 0:
        raise ValueError(f"invalid result_exp: {result_exp}")
    raw = wrap_mul(op1.data, op2.data)
    data = fxp_rshift_round(raw, rshift, round_mode)
    return fxp_clip(FxpArray(data, result_bits, int(result_exp), result_signed))


def fxp_matmul(
    op1: FxpArray,
    op2: FxpArray,
    result_bits: Optional[int] = None,
    result_bits_fn=max,
    result_exp: Optional[int] = None,
    result_exp_fn=max,
    round_mode: RoundingMode = RoundingMode.FLOOR,
) -> FxpArray:
    """fxparray.py:640-678."""
    result_signed = op1.signed or op2.signed
    if result_bits is None:
        result_bits = result_bits_fn(op1.bits, op2.bits)
    if result_exp is None:
        result_exp = result_exp_fn(op1.exp, op2.exp)
    raw = matmul_i32(op1.data, op2.data)
    rshift = op1.exp + op2.exp - result_exp
    data = fxp_rshift_round(raw, rshift, round_mode)
    return fxp_clip(FxpArray(data, result_bits, int(result_exp), result_signed))


def fxp_complex_add(op1: ComplexFxpArray, op2: ComplexFxpArray, result_exp, result_bits, **kw):
    """fxparray.py:469-501."""
    return ComplexFxpArray(
        real=fxp_add(op1.real, op2.real, result_bits=result_bits[0], result_exp=result_exp[0], **kw),
        imag=fxp_add(op1.imag, op2.imag, result_bits=result_bits[1], result_exp=result_exp[1], **kw),
    )


def fxp_complex_mul(op1: ComplexFxpArray, op2: ComplexFxpArray, result_exp=None, result_bits=None,
                    round_mode=RoundingMode.FLOOR, **_):
    """fxparray.py:504-570 (used only by the non-lax-scan branch of FxpSSM)."""
    result_exp = result_exp or (None, None)
    result_bits = result_bits or (None, None)
    kw = dict(round_mode=round_mode)
    arbr = fxp_mul(op1.real, op2.real, result_exp=result_exp[0], result_bits=result_bits[0], **kw)
    aibr = fxp_mul(op1.imag, op2.real, result_exp=result_exp[1], result_bits=result_bits[1], **kw)
    arbi = fxp_mul(op1.real, op2.imag, result_exp=result_exp[1], result_bits=result_bits[1], **kw)
    aibi = fxp_mul(op1.imag, op2.imag, result_exp=result_exp[0], result_bits=result_bits[0], **kw)
    re = fxp_sub(arbr, aibi, result_exp=result_exp[0], result_bits=result_bits[0], round_mode=round_mode)
    im = fxp_add(arbi, aibr, result_exp=result_exp[1], result_bits=result_bits[1], round_mode=round_mode)
    return ComplexFxpArray(real=re, imag=im)


def fxp_relu(x):
    """fxpmodel.py:27-63.

    Real: ``max(x, 0)``.  Complex: ``jax.nn.relu`` on ``re + 1j*im`` is a lexicographic
    ``maximum(z, 0)``: keep z iff re > 0 or (re == 0 and im > 0).  The reference builds the
    complex value as complex64, i.e. both parts pass through float32 and back to int32; that
    round trip (round-to-nearest-even, then truncating/saturating convert) is restated because
    the recurrent state is an unsaturated int32 and may exceed 2**24.
    """
    if isinstance(x, ComplexFxpArray):
        re_f = x.real.data.astype(F32)
        im_f = x.imag.data.astype(F32)
        keep = (re_f > 0) | ((re_f == 0) & (im_f > 0))

        def back(v):
            v = np.clip(np.where(keep, v, F32(0)).astype(np.float64), -2147483648.0, 2147483647.0)
            return v.astype(np.int64).astype(I32)

        return ComplexFxpArray(
            real=FxpArray(back(re_f), x.real.bits, x.real.exp, x.real.signed),
            imag=FxpArray(back(im_f), x.imag.bits, x.imag.exp, x.imag.signed),
        )
    return FxpArray(np.maximum(x.data, 0).astype(I32), x.bits, x.exp, x.signed)
