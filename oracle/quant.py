"""Oracle: static fake-quant primitives (NumPy restatement).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/sparseRNNs/utils/quantization.py`` (``quantization.py:LINE``).
Only the per-tensor-symmetric scheme with power-of-two scales is restated: it is the one the
``w8a16`` static recipe uses (quantization.py:127-136, ssm.py:408-421, layers.py:98-103, 168-174).
"""

from __future__ import annotations

import numpy as np

F32 = np.float32

# quantization.py:127-136  ("w8a16")
W8A16 = dict(a_precision=16, b_precision=8, c_precision=8, d_precision=8,
             non_ssm_precision=8, ssm_act_precision=16, non_ssm_act_precision=16)


def calculate_qparams(minval, maxval, precision: int, pow2scale: bool = True, eps: float = 1e-6):
    """quantization.py:352-370 (symmetric).  float32 throughout, ``round`` half-to-even."""
    max_abs = np.maximum(np.abs(F32(minval)), np.abs(F32(maxval)))
    quant_max = F32(2 ** (precision - 1) - 1)
    scale = F32(max_abs / quant_max)
    scale = np.maximum(scale, F32(eps))
    if pow2scale:
        scale = F32(2.0) ** np.round(np.log2(F32(scale)))
    return F32(scale)


def quantdequant(x, scale, precision: int, dtype=F32):
    """quantization.py:386-400 with zero_point = 0: xdq = clip(round(x/s), -2^(b-1), 2^(b-1)-1) * s, returned through the
    straight-through form ``x + stop_gradient(xdq - x)`` (:398-399).  In floating point that sum IS xdq whenever
    xdq / 2 <= x <= 2 xdq (Sterbenz: the difference is exact) and, beyond that, as long as ulp(x) <= scale; only an
    outlier more than 2^23 scales away from zero comes back within one ulp(x) of xdq instead of on the grid (pinned by
    tests/test_oracle_vs_reference.py::test_quantdequant_and_qparams_bit_exact)."""
    x = np.asarray(x, dtype=dtype)
    s = dtype(scale)
    qmin = -(2 ** (precision - 1))
    qmax = 2 ** (precision - 1) - 1
    xq = np.round(x / s)
    xq = np.clip(xq, qmin, qmax)
    xdq = (xq * s).astype(dtype)
    return (x + (xdq - x).astype(dtype)).astype(dtype)


def quant_int(x, scale, precision: int):
    """The integer code of ``quantdequant`` (used to build bit-exact sparse index sets)."""
    x = np.asarray(x, dtype=F32)
    qmin = -(2 ** (precision - 1))
    qmax = 2 ** (precision - 1) - 1
    return np.clip(np.round(x / F32(scale)), qmin, qmax).astype(np.int32)


class MinMax:
    """Running per-tensor min/max (quantization.py:291-349, per_tensor_symmetric)."""

    def __init__(self):
        self.min = F32(np.inf)
        self.max = F32(-np.inf)

    def observe(self, x):
        x = np.asarray(x)
        self.min = F32(min(self.min, x.min()))
        self.max = F32(max(self.max, x.max()))

    def scale(self, bits: int):
        return calculate_qparams(self.min, self.max, bits, pow2scale=True)

    def as_dict(self, prefix: str = ""):
        return {f"{prefix}observer_min": self.min, f"{prefix}observer_max": self.max}
