"""Oracle: S5 SSM forward, float and static-quant flavours (NumPy restatement).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/sparseRNNs/model/ssm.py`` (``ssm.py:LINE``).  All functions accept a leading
batch axis (the reference vmaps over clips, seq_model.py:459-473).
"""

from __future__ import annotations

import numpy as np

from .quant import quantdequant

F32 = np.float32


def _cdtype(dtype):
    return np.complex64 if dtype == np.float32 else np.complex128


def discretize_zoh(Lambda, B_tilde, Delta):
    """ssm.py:37-50."""
    Lambda_bar = np.exp(Lambda * Delta)
    B_bar = (1 / Lambda * (Lambda_bar - 1))[..., None] * B_tilde
    return Lambda_bar, B_bar


def discretize_bilinear(Lambda, B_tilde, Delta):
    """ssm.py:19-34."""
    BL = 1 / (1 - (Delta / 2.0) * Lambda)
    Lambda_bar = BL * (1 + (Delta / 2.0) * Lambda)
    B_bar = (BL * Delta)[..., None] * B_tilde
    return Lambda_bar, B_bar


def ssm_discrete_params(mixer: dict, *, clip_eigs: bool, step_rescale: float = 1.0,
                        discretization: str = "zoh", dtype=F32):
    """qS5SSM.setup, ssm.py:436-523: (Lambda_bar (P,), B_bar (P,H), C_tilde (H,P), D (H,))."""
    cd = _cdtype(dtype)
    lre = np.asarray(mixer["Lambda_re"], dtype=dtype)
    lim = np.asarray(mixer["Lambda_im"], dtype=dtype)
    if clip_eigs:
        lre = np.minimum(lre, dtype(-1e-4))  # np.clip(x, None, -1e-4), ssm.py:443
    Lambda = (lre + 1j * lim).astype(cd)
    B = np.asarray(mixer["B"], dtype=dtype)
    C = np.asarray(mixer["C"], dtype=dtype)
    B_tilde = (B[..., 0] + 1j * B[..., 1]).astype(cd)
    C_tilde = (C[..., 0] + 1j * C[..., 1]).astype(cd)
    step = (dtype(step_rescale) * np.exp(np.asarray(mixer["log_step"], dtype=dtype)[:, 0])).astype(dtype)
    if discretization == "zoh":
        Lambda_bar, B_bar = discretize_zoh(Lambda, B_tilde, step)
    elif discretization == "bilinear":
        Lambda_bar, B_bar = discretize_bilinear(Lambda, B_tilde, step)
    else:
        raise NotImplementedError(discretization)
    return Lambda_bar.astype(cd), B_bar.astype(cd), C_tilde, np.asarray(mixer["D"], dtype=dtype)


def relu_complex(x):
    """``jax.nn.relu`` on complex = lexicographic ``maximum(x, 0)`` (ssm.py:159, :325)."""
    keep = (x.real > 0) | ((x.real == 0) & (x.imag > 0))
    return np.where(keep, x, 0).astype(x.dtype)


def scan_sequential(lam, Bu):
    """``x_t = lam * x_{t-1} + Bu_t`` with x_{-1} = 0 (ssm.py:130-150, :287-313).

    lam: (P,) complex; Bu: (..., L, P) complex.  The complex product is written out in real
    arithmetic as ssm.py:141-142 does.
    """
    out = np.empty_like(Bu)
    x = np.zeros(Bu.shape[:-2] + Bu.shape[-1:], dtype=Bu.dtype)
    ar, ai = lam.real, lam.imag
    for t in range(Bu.shape[-2]):
        xr, xi = x.real, x.imag
        x = ((ar * xr - ai * xi) + 1j * (ar * xi + ai * xr)).astype(Bu.dtype) + Bu[..., t, :]
        out[..., t, :] = x
    return out


def _binop(qi, qj):
    """ssm.py:54-77 with np.multiply for qhad (the float configuration)."""
    Ai, bi = qi
    Aj, bj = qj
    A_re = Aj.real * Ai.real - Aj.imag * Ai.imag
    A_im = Aj.real * Ai.imag + Aj.imag * Ai.real
    b_re = Aj.real * bi.real - Aj.imag * bi.imag
    b_im = Aj.real * bi.imag + Aj.imag * bi.real
    return (A_re + 1j * A_im).astype(Ai.dtype), ((b_re + 1j * b_im).astype(bi.dtype) + bj)


def scan_associative(lam, Bu):
    """``jax.lax.associative_scan(q_bin_op, (Lambda_elements, Bu))`` along axis -2 (ssm.py:112-127).

    Restates the odd/even recursion of jax.lax.associative_scan so that the combine order (and
    therefore the float32 rounding) follows the reference's parallel path.
    """
    L = Bu.shape[-2]
    A = np.broadcast_to(lam, Bu.shape).astype(Bu.dtype)

    def sl(e, s):
        return (e[0][..., s, :], e[1][..., s, :])

    def rec(e):
        n = e[0].shape[-2]
        if n < 2:
            return e
        red = _binop(sl(e, slice(0, n - 1, 2)), sl(e, slice(1, n, 2)))
        odd = rec(red)
        if n % 2 == 0:
            even = _binop(sl(odd, slice(0, -1)), sl(e, slice(2, n, 2)))
        else:
            even = _binop(odd, sl(e, slice(2, n, 2)))
        outA = np.empty_like(e[0])
        outb = np.empty_like(e[1])
        outA[..., 0, :] = e[0][..., 0, :]
        outb[..., 0, :] = e[1][..., 0, :]
        outA[..., 2::2, :] = even[0]
        outb[..., 2::2, :] = even[1]
        outA[..., 1::2, :] = odd[0]
        outb[..., 1::2, :] = odd[1]
        return outA, outb

    return rec((A, Bu))[1]


def apply_ssm_float(Lambda_bar, B_bar, C_tilde, u, *, conj_sym=True, relufication=False,
                    associative_scan=True, B_bias=None):
    """build_apply_ssm._apply_ssm, ssm.py:84-185 (float configuration, unidirectional).

    u: (..., L, H) real.  Returns (ys (..., L, H), xs (..., L, P)).
    """
    cd = Lambda_bar.dtype
    Bu = (u @ B_bar.real.T + 1j * (u @ B_bar.imag.T)).astype(cd)  # u.imag == 0, ssm.py:116-122
    if B_bias is not None:
        Bu = Bu + B_bias.astype(cd)
    xs = scan_associative(Lambda_bar, Bu) if associative_scan else scan_sequential(Lambda_bar, Bu)
    if relufication:
        xs = relu_complex(xs)
    ys = xs.real @ C_tilde.real.T - xs.imag @ C_tilde.imag.T  # ssm.py:179-180
    if conj_sym:
        ys = 2 * ys
    return ys.astype(u.dtype), xs


def apply_ssm_static(Lambda_bar, B_bar, C_tilde, u, scales: dict, bits: dict, *, conj_sym=True,
                     relufication=True, B_bias=None, dtype=F32):
    """build_apply_ssm_static_quant._apply_ssm, ssm.py:194-343, frozen scales (no observers).

    ``scales`` keys: ut, A_re, A_im, B_re, B_im, C_re, C_im, But_re, But_im.  In frozen inference the
    state is NOT re-quantised inside the loop (ssm.py:254-284: no observer in batch_stats) and the
    result of ``quant_xt(xs)`` is discarded (ssm.py:314).  Returns (ys, xs, u_q).
    """
    cd = _cdtype(dtype)
    a_bits, b_bits, c_bits, act_bits = bits["a"], bits["b"], bits["c"], bits["ssm_act"]
    u_q = quantdequant(u, scales["ut"], act_bits, dtype)
    Bq = (quantdequant(B_bar.real, scales["B_re"], b_bits, dtype)
          + 1j * quantdequant(B_bar.imag, scales["B_im"], b_bits, dtype)).astype(cd)
    Aq = (quantdequant(Lambda_bar.real, scales["A_re"], a_bits, dtype)
          + 1j * quantdequant(Lambda_bar.imag, scales["A_im"], a_bits, dtype)).astype(cd)
    Cq = (quantdequant(C_tilde.real, scales["C_re"], c_bits, dtype)
          + 1j * quantdequant(C_tilde.imag, scales["C_im"], c_bits, dtype)).astype(cd)
    Bu = (u_q @ Bq.real.T + 1j * (u_q @ Bq.imag.T)).astype(cd)
    if B_bias is not None:
        Bu = Bu + B_bias.astype(cd)
    Bu = (quantdequant(Bu.real, scales["But_re"], act_bits, dtype)
          + 1j * quantdequant(Bu.imag, scales["But_im"], act_bits, dtype)).astype(cd)
    xs = scan_sequential(Aq, Bu)
    if relufication:
        xs = relu_complex(xs)
    ys = xs.real @ Cq.real.T - xs.imag @ Cq.imag.T  # real part of the complex dot, ssm.py:334
    if conj_sym:
        ys = 2 * ys
    return ys.astype(dtype), xs, u_q
