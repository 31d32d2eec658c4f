"""Op-level case table shared by the fixture generator (reference sources over the NumPy jax shim) and by the tests
(oracle / CUDA).  A case names a function of ``fxparray.py`` / ``fxpmodel.py``, its FxpArray operands and keyword
arguments; ``run_case`` evaluates it against any module pair that offers the reference's names (the reference itself or
``oracle.fxp`` / ``oracle.fxpmodel``).  Deterministic: operands come from a seeded generator."""

from __future__ import annotations

import numpy as np


def _arr(rng, shape, bits, exp, signed=True, fill=1.0, hot=()):
    """int32 payload that uses ``fill`` of the bits-wide range, plus a few hand-placed values (saturation, ties)."""
    lo, hi = (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1) if signed else (0, 2 ** bits - 1)
    d = rng.integers(int(lo * fill), int(hi * fill) + 1, size=shape).astype(np.int32)
    flat = d.reshape(-1)
    for i, v in enumerate(hot):
        flat[i % flat.size] = v
    return dict(data=d, bits=bits, exp=exp, signed=signed)


def op_cases():
    rng = np.random.default_rng(20261019)
    C = []

    def add(name, fn, ops, **kw):
        C.append(dict(name=name, fn=fn, ops=ops, kw=kw))

    # fxp_from_fp: all rounding modes, half-way ties, saturation at both ends (fxparray.py:287-307)
    xf = np.concatenate([rng.standard_normal(200).astype(np.float32) * 3,
                         np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.125, -0.125, 1e6, -1e6, 127.5, -128.5, 0.0],
                                  np.float32) / 4]).astype(np.float32)
    for mode in ("FLOOR", "ROUND", "CEIL"):
        for bits, exp in ((8, 2), (16, 10), (16, 0)):
            add(f"from_fp_{mode}_{bits}_{exp}", "fxp_from_fp", [], x=xf, bits=bits, exp=exp, signed=True, round_mode=mode)
    # change_exp / change_cfg: up and down shifts, negative odd payloads, saturation at the old / new width
    a16 = _arr(rng, (7, 33), 16, 9, hot=(-32768, 32767, -1, 1, -3, 3, -5, 5, 0))
    for mode in ("FLOOR", "ROUND", "CEIL"):
        for new_exp in (4, 8, 9, 12, 15):
            add(f"change_exp_{mode}_{new_exp}", "fxp_change_exp", [a16], new_exp=new_exp, round_mode=mode)
        for new_bits, new_exp in ((8, 3), (8, 9), (16, 13), (24, 12), (12, 6)):
            add(f"change_cfg_{mode}_{new_bits}_{new_exp}", "fxp_change_cfg", [a16], new_bits=new_bits, new_exp=new_exp,
                new_signed=True, round_mode=mode)
    # clip
    wide = _arr(rng, (50,), 24, 8, hot=(2 ** 23 - 1, -(2 ** 23), 40000, -40000))
    wide["bits"] = 16
    add("clip_16", "fxp_clip", [wide])
    # add / sub: explicit result exponent, compute_best (data-dependent exponent), the default branch with equal and
    # unequal exponents (operator-precedence quirk of fxparray.py:416-419)
    b16 = _arr(rng, (7, 33), 16, 12, hot=(32767, -32768, 7, -7))
    c16 = _arr(rng, (7, 33), 16, 9, fill=0.02)
    small = _arr(rng, (7, 33), 16, 9, fill=0.001)
    for mode in ("FLOOR", "ROUND"):
        for rexp in (6, 9, 11):
            add(f"add_exp{rexp}_{mode}", "fxp_add", [a16, b16], result_exp=rexp, result_bits=16, round_mode=mode)
            add(f"sub_exp{rexp}_{mode}", "fxp_sub", [a16, b16], result_exp=rexp, result_bits=16, round_mode=mode)
    for x, y, tag in ((a16, b16, "big"), (c16, b16, "mixed"), (small, small, "small"), (c16, small, "c_small")):
        for rb in (16, 24):
            add(f"add_best_{tag}_{rb}", "fxp_add", [x, y], result_exp="compute_best", result_bits=rb)
            add(f"sub_best_{tag}_{rb}", "fxp_sub", [x, y], result_exp="compute_best", result_bits=rb)
    add("add_default_equal_exp", "fxp_add", [a16, c16])
    add("add_default_op1_larger", "fxp_add", [_arr(rng, (40,), 8, 5, fill=0.2), _arr(rng, (40,), 8, 3, fill=0.2)])
    add("add_default_op2_larger", "fxp_add", [_arr(rng, (40,), 8, 2, fill=0.2), _arr(rng, (40,), 8, 4, fill=0.2)])
    # mul: explicit exponent in every rounding mode, compute_best, the int32 wrap of 16 x 16 bit products
    for mode in ("FLOOR", "ROUND", "CEIL"):
        add(f"mul_exp_{mode}", "fxp_mul", [a16, b16], result_exp=10, result_bits=16, round_mode=mode)
        add(f"mul_w8_{mode}", "fxp_mul", [_arr(rng, (7, 33), 8, 5), c16], result_exp=9, result_bits=16, round_mode=mode)
    add("mul_best_16", "fxp_mul", [c16, b16], result_exp="compute_best", result_bits=16)
    add("mul_best_24", "fxp_mul", [a16, b16], result_exp="compute_best", result_bits=24)
    add("mul_default", "fxp_mul", [c16, _arr(rng, (7, 33), 8, 9, fill=0.5)])
    # matmul: w8 x a16 with int32 accumulation (incl. wrap for large K), rounding modes
    xa = _arr(rng, (5, 9, 96), 16, 9, hot=(32767, -32768))
    w8 = _arr(rng, (96, 64), 8, 6)
    for mode in ("FLOOR", "ROUND"):
        add(f"matmul_{mode}", "fxp_matmul", [xa, w8], result_exp=7, result_bits=16, round_mode=mode)
    add("matmul_default", "fxp_matmul", [xa, w8])
    big = _arr(rng, (3, 600), 16, 0, fill=1.0)
    bigw = _arr(rng, (600, 4), 16, 0, fill=1.0)
    add("matmul_wraps_int32", "fxp_matmul", [big, bigw], result_exp=0, result_bits=32)
    return C


def to_fxp(mod, spec):
    return mod.FxpArray(data=mod_array(mod, spec["data"]), bits=int(spec["bits"]), exp=int(spec["exp"]),
                        signed=bool(spec["signed"]))


def mod_array(mod, a):
    conv = getattr(mod, "_asarray", None)          # the generator installs jnp.asarray here for the reference
    return conv(a) if conv is not None else np.array(a)


def run_case(mod, case):
    """Evaluate one case with module ``mod`` (reference ``fxparray`` or ``oracle.fxp``); returns (data, bits, exp, signed)."""
    kw = dict(case["kw"])
    if "round_mode" in kw:
        kw["round_mode"] = getattr(mod.RoundingMode, kw["round_mode"])
    if "x" in kw:
        kw["x"] = mod_array(mod, kw["x"])
    ops = [to_fxp(mod, s) for s in case["ops"]]
    out = getattr(mod, case["fn"])(*ops, **kw)
    return np.asarray(out.data).astype(np.int32), int(out.bits), int(out.exp), bool(out.signed)
