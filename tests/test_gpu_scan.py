"""GPU parity: scan kernels through the C ABI vs the oracle (bit-exact for int, 1e-5 rel for complex64)."""
import numpy as np
import pytest

from oracle import fxpmodel as OFM
from oracle import ssm as OS
from oracle.fxp import ComplexFxpArray, FxpArray, fxp_relu

pytestmark = pytest.mark.gpu

TOL_C64 = 1e-5  # fp32 within 1e-5 relative (max-norm) of the float64 evaluation of the same recurrence


def _lam(rng, P):
    dt = rng.uniform(1e-3, 1e-1, P)
    return np.exp((-rng.uniform(1e-4, 1.0, P) + 1j * rng.uniform(-np.pi, np.pi, P)) * dt).astype(np.complex64)


@pytest.mark.parametrize("B,L,P,relu", [(1, 1, 16, 0), (1, 5, 16, 0), (2, 100, 64, 0), (3, 3751, 64, 1), (2, 3751, 128, 0),
                                        (1, 777, 256, 0), (2, 1000, 16, 1), (2, 513, 96, 0), (1, 16, 2, 0),
                                        (1, 65536, 64, 0),
                                        # throughput batches (B >= 96): one CTA per clip, pure stream, ragged last tile
                                        (128, 50, 64, 1), (300, 77, 16, 0), (97, 1000, 128, 1), (160, 16, 256, 0),
                                        # few clips cut into segments: two-pass inner segments, ReLU, odd state counts
                                        (5, 4000, 64, 1), (40, 333, 6, 0), (7, 2049, 48, 1), (33, 31, 64, 0)])
def test_scan_c64_matches_oracle(cuda_device, B, L, P, relu):
    import torch
    from sparsernns_b200.model import scan_c64

    rng = np.random.default_rng(B * 1000 + L + P)
    lam = _lam(rng, P)
    Bu = (rng.standard_normal((B, L, P)) + 1j * rng.standard_normal((B, L, P))).astype(np.complex64)
    got = scan_c64(torch.from_numpy(Bu).to(cuda_device), torch.from_numpy(lam).to(cuda_device), relu=bool(relu)).cpu().numpy()
    ref = OS.scan_sequential(lam.astype(np.complex128), Bu.astype(np.complex128))
    if relu:
        ref = OS.relu_complex(ref)
        # entries whose real part is within rounding of zero may legitimately flip; compare where |re| is resolved
        mask = np.abs(ref.real) > 1e-4 * np.abs(ref).max()
        mask |= (got == 0) == (ref == 0)
        assert np.abs(got - ref)[mask].max() <= TOL_C64 * np.abs(ref).max()
    else:
        assert np.abs(got - ref).max() <= TOL_C64 * np.abs(ref).max()


def test_scan_c64_in_place_and_linearity(cuda_device):
    import torch
    from sparsernns_b200.model import scan_c64

    rng = np.random.default_rng(7)
    B, L, P = 4, 3751, 64
    lam = torch.from_numpy(_lam(rng, P)).to(cuda_device)
    a = torch.randn(B, L, P, dtype=torch.complex64, device=cuda_device)
    b = torch.randn(B, L, P, dtype=torch.complex64, device=cuda_device)
    ya, yb, yab = scan_c64(a, lam), scan_c64(b, lam), scan_c64(a + 2 * b, lam)
    assert (yab - (ya + 2 * yb)).abs().max() <= 1e-4 * yab.abs().max()   # linearity of the recurrence
    buf = a.clone()
    scan_c64(buf, lam, out=buf)                                          # xs may alias Bu
    assert torch.equal(buf, ya)
    big = torch.randn(130, 200, P, dtype=torch.complex64, device=cuda_device)   # the one-CTA-per-clip path, in place too
    yb2 = scan_c64(big, lam)
    scan_c64(big, lam, out=big)
    assert torch.equal(big, yb2)
    # a clip's bits do not depend on the batch it is in, nor on how the grid is cut (the numerics are fixed by L alone):
    # 4 clips (cut into segments, two-pass inner segments), 1 clip (finer cut), 100 / 200 clips (one CTA per clip, pure stream)
    assert torch.equal(scan_c64(a[:1], lam)[0], ya[0])
    assert torch.equal(scan_c64(torch.cat([a[:1]] * 100), lam)[0], ya[0])
    assert torch.equal(scan_c64(torch.cat([a[:1]] * 200), lam)[7], ya[0])
    assert torch.equal(scan_c64(torch.cat([a[1:3]] * 20), lam)[1], ya[2])
    # recurrence identity x_t - lam x_{t-1} == Bu_t at full NDNS size
    resid = ya[:, 1:] - lam * ya[:, :-1] - a[:, 1:]
    assert resid.abs().max() <= 1e-4 * ya.abs().max()


@pytest.mark.parametrize("B,L,P,relu", [(32, 3751, 64, 0), (3, 1000, 100, 1), (2, 3751, 128, 1), (5, 300, 16, 0), (1, 20000, 64, 1),
                                        (4, 3751, 256, 0), (7, 257, 2, 0)])
def test_scan_c64_resident_equals_ring(cuda_device, monkeypatch, B, L, P, relu):
    """Few clips: the resident form (one 128-step block per CTA kept in shared memory, scan_c64_res_kernel) and the segmented ring
    kernel evaluate the same expression - equal bits, also in place, for state groups (P > 64), ragged last blocks and narrow states."""
    import torch
    from sparsernns_b200.model import scan_c64

    rng = np.random.default_rng(B + L + P)
    lam = torch.from_numpy(_lam(rng, P)).to(cuda_device)
    Bu = torch.randn(B, L, P, dtype=torch.complex64, device=cuda_device)
    monkeypatch.setenv("S5B200_SCAN_RESIDENT", "0")
    ring = scan_c64(Bu, lam, relu=bool(relu))
    monkeypatch.setenv("S5B200_SCAN_RESIDENT", "1")
    res = scan_c64(Bu, lam, relu=bool(relu))
    assert torch.equal(ring, res)
    buf = Bu.clone()
    scan_c64(buf, lam, out=buf, relu=bool(relu))
    assert torch.equal(buf, ring)
    ref = OS.scan_sequential(lam.cpu().numpy().astype(np.complex128), Bu[:1].cpu().numpy().astype(np.complex128))
    if not relu:
        assert np.abs(res[:1].cpu().numpy() - ref).max() <= TOL_C64 * np.abs(ref).max()



def test_scan_c64_rejects_bad_shapes(cuda_device):
    import torch
    from sparsernns_b200 import _abi
    from sparsernns_b200.model import scan_c64

    with pytest.raises(_abi.S5Error):
        scan_c64(torch.zeros(1, 8, 257, dtype=torch.complex64, device=cuda_device),
                 torch.zeros(257, dtype=torch.complex64, device=cuda_device))


@pytest.mark.parametrize("B,L,P,relu,exps,big", [
    (1, 7, 16, 0, (15, 15, 15, 14, 13, 15), False),
    (2, 300, 64, 1, (15, 15, 15, 14, 13, 15), False),
    (3, 1000, 128, 1, (15, 14, 3, 5, 4, 4), True),
    (2, 3751, 64, 0, (15, 14, 3, 5, 4, 4), True),
    (1, 50, 33, 1, (14, 15, 10, 10, 12, 11), False),
])
def test_scan_i32_bit_exact(cuda_device, B, L, P, relu, exps, big):
    import ctypes as C
    import torch
    from sparsernns_b200 import _abi

    rng = np.random.default_rng(L + P)
    eA_re, eA_im, eBu_re, eBu_im, ex_re, ex_im = exps
    A_re = rng.integers(-32768, 32767, P).astype(np.int32)
    A_im = rng.integers(-32768, 32767, P).astype(np.int32)
    lim = 2 ** 30 if big else 32767
    Bu_re = rng.integers(-lim, lim, (B, L, P)).astype(np.int32)
    Bu_im = rng.integers(-lim, lim, (B, L, P)).astype(np.int32)
    xr, xi = OFM.recurrent_loop(Bu_re, Bu_im, A_re, A_im, eBu_re, eBu_im, eA_re, eA_im, ex_re, ex_im)
    if relu:
        r = fxp_relu(ComplexFxpArray(FxpArray(xr, 16, ex_re), FxpArray(xi, 16, ex_im)))
        xr, xi = r.real.data, r.imag.data
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    dBr, dBi, dAr, dAi = t(Bu_re), t(Bu_im), t(A_re), t(A_im)
    oR, oI = torch.empty_like(dBr), torch.empty_like(dBi)
    lib = _abi.lib()
    _abi.check(lib.s5b200_scan_i32(dBr.data_ptr(), dBi.data_ptr(), dAr.data_ptr(), dAi.data_ptr(), oR.data_ptr(),
                                   oI.data_ptr(), B, L, P, eA_re, eA_im, eBu_re, eBu_im, ex_re, ex_im, relu,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert np.array_equal(oR.cpu().numpy(), xr) and np.array_equal(oI.cpu().numpy(), xi)
