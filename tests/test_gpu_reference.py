"""GPU parity against vectors produced by the REFERENCE'S OWN SOURCES (tests/golden/reference_*.npz, written by
tests/golden/make_reference_golden.py: the reference's fxparray.py / fxputils.py / fxpmodel.py / ssm.py executed over a
NumPy stand-in for jax.numpy).  The oracle is not involved: CUDA path vs the reference's returned values directly.
Fixed point: bit-exact.  Float / static w8a16 SSM: the tolerances of BASELINE.md."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_F32, TOL_W8A16 = 1e-5, 1e-2


def _load(name):
    with np.load(os.path.join(GOLD, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def ref_model():
    return _load("reference_fxp_model.npz")


@pytest.fixture(scope="module")
def ref_ssm():
    return _load("reference_ssm.npz")


def _product_model(dim_scale, shared_exp=True, **kw):
    from sparsernns_b200 import fxpmodel as PF
    from sparsernns_b200 import fxputils as PU
    from sparsernns_b200.checkpoint import load_npz

    tree = load_npz(os.path.join(GOLD, f"ndns_ds{dim_scale}_sparse_w8a16.npz"))
    md = PU.load_modeldict(tree["params"], tree["cal_stats"])
    q = PU.create_fxp_qconfig(md, "w8a16", shared_exp=shared_exp)
    return PF.FxpRegressionModel(md, q, 3, **kw), PF


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("case,dim_scale,shared", [("A", 0.5, True), ("B", 0.5, True), ("C", 0.5, False), ("E", 1.0, True)],
                         ids=["golden_input", "saturating_and_silent_input", "per_layer_exponents", "dim_scale_1.0"])
def test_fxp_model_bit_exact_vs_reference(cuda_device, ref_model, case, dim_scale, shared, fxp_path):
    """FxpRegressionModel.forward: decoder output and every block output equal the reference's integers and exponents."""
    import torch

    pm, PF = _product_model(dim_scale, shared)
    x = ref_model[f"{case}/x"] if f"{case}/x" in ref_model else ref_model["A/x"]
    xq = pm.quantize_input(torch.from_numpy(x).to(cuda_device))
    y = pm(xq)
    cfg = ref_model[f"{case}/y/cfg"]
    assert (y.bits, y.exp) == (int(cfg[0]), int(cfg[1]))
    assert np.array_equal(y.data.cpu().numpy(), ref_model[f"{case}/y/data"])
    for i in range(3):
        key = f"A/intermediates/encoder/layer_{i}_output" if case == "A" else f"{case}/layer_{i}_output"
        t = pm.tap(xq, i, PF.TAP_OUTPUT)
        assert t.exp == int(ref_model[f"{key}/cfg"][1]), (case, i)
        assert np.array_equal(t.data.cpu().numpy(), ref_model[f"{key}/data"]), (case, i)


def test_fxp_model_taps_bit_exact_vs_reference(cuda_device, ref_model, fxp_path):
    """The sow()-named intermediates the library can tap, per block: ssm_input, pre_s5, pre_C (states after the complex
    ReLU), post_GLU, output; plus the quantised input and the un-batched (L, 257) call."""
    import torch

    pm, PF = _product_model(0.5)
    x = torch.from_numpy(ref_model["A/x"]).to(cuda_device)
    xq = pm.quantize_input(x)
    pre = "A/intermediates/encoder"
    assert np.array_equal(xq.data.cpu().numpy(), ref_model[f"{pre}/pre_encoder/data"])
    for i in range(3):
        L = f"{pre}/layers_{i}"
        for tap, key in ((PF.TAP_LAYER_INPUT, "ssm_input"), (PF.TAP_POST_GLU, "post_GLU"), (PF.TAP_OUTPUT, "output")):
            t = pm.tap(xq, i, tap)
            assert np.array_equal(t.data.cpu().numpy(), ref_model[f"{L}/{key}/data"]), (i, key)
            if tap in (PF.TAP_LAYER_INPUT, PF.TAP_OUTPUT):
                assert t.exp == int(ref_model[f"{L}/{key}/cfg"][1]), (i, key)
        # the library taps the mixer input AFTER its change_cfg to the u format (fxpmodel.py:620-624); the reference sows
        # the value before it ("pre_s5", 16 bits at the BatchNorm's compute_best exponent): floor-shift + saturate by hand
        t = pm.tap(xq, i, PF.TAP_PRE_S5)
        e_ref = int(ref_model[f"{L}/pre_s5/cfg"][1])
        assert e_ref >= t.exp
        u_ref = np.clip(ref_model[f"{L}/pre_s5/data"] >> (e_ref - t.exp), -32768, 32767)
        assert np.array_equal(t.data.cpu().numpy(), u_ref), (i, "u = change_cfg(pre_s5)")
        xs = pm.tap(xq, i, PF.TAP_XS).data.cpu().numpy()
        assert np.array_equal(xs[..., 0], ref_model[f"{L}/pre_C.re/data"]) and np.array_equal(xs[..., 1], ref_model[f"{L}/pre_C.im/data"]), i
    y1 = pm(pm.quantize_input(torch.from_numpy(ref_model["D/x"][None]).to(cuda_device)))
    assert np.array_equal(y1.data.cpu().numpy()[0], ref_model["D/y/data"])


def test_scan_i32_bit_exact_vs_reference(cuda_device, ref_ssm):
    """s5b200_scan_i32 against the reference's recurrent_loop (fxpmodel.py:147-208) under jax.vmap."""
    import torch
    from sparsernns_b200 import _abi

    lib = _abi.lib()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    Bre, Bim, Are, Aim = (dev(ref_ssm[f"rec/{k}"]) for k in ("Bre", "Bim", "Are", "Aim"))
    eBu_re, eBu_im, eA_re, eA_im, ex_re, ex_im = (int(v) for v in ref_ssm["rec/exps"])
    B, L, P = Bre.shape
    xr, xi = torch.empty_like(Bre), torch.empty_like(Bim)
    _abi.check(lib.s5b200_scan_i32(Bre.data_ptr(), Bim.data_ptr(), Are.data_ptr(), Aim.data_ptr(), xr.data_ptr(), xi.data_ptr(),
                                   B, L, P, eA_re, eA_im, eBu_re, eBu_im, ex_re, ex_im, 0,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s5b200_scan_i32")
    assert np.array_equal(xr.cpu().numpy(), ref_ssm["rec/xs"][..., 0]) and np.array_equal(xi.cpu().numpy(), ref_ssm["rec/xs"][..., 1])


def _bound_from_fixture(static, relufication):
    from sparsernns_b200.checkpoint import load_npz
    from sparsernns_b200.model import BoundModel, S5Config

    tree = load_npz(os.path.join(GOLD, "ndns_ds0.5_sparse_w8a16.npz"))
    cfg = S5Config(n_layers=3, d_model=96, ssm_size=64, relufication=relufication, quantization="w8a16" if static else "none")
    return BoundModel(cfg, {"params": tree["params"], "batch_stats": tree["batch_stats"]}), tree


def test_float_ssm_vs_reference(cuda_device, ref_ssm):
    """apply_ssm of block 1 (B projection, scan_c64, ReLU, C projection, + D u) against build_apply_ssm._apply_ssm
    (ssm.py:84-185) evaluated by the reference itself, plus its D * u term (ssm.py:618-622)."""
    import torch

    for relu in (False, True):
        m, tree = _bound_from_fixture(False, relu)
        D = tree["params"]["encoder"]["layers_1"]["mixer"]["D"]
        u = ref_ssm["u"]
        y, xs = m.apply_ssm(1, torch.from_numpy(u[None]).to(cuda_device))
        for scan in ("seq", "assoc"):
            tag = f"float_{scan}_{'relu' if relu else 'lin'}"
            assert _rel(xs.cpu().numpy()[0], ref_ssm[f"{tag}/xs"]) <= TOL_F32, tag
            assert _rel(y.cpu().numpy()[0], ref_ssm[f"{tag}/ys"] + D * u) <= TOL_F32, tag


def test_scan_c64_vs_reference_states(cuda_device, ref_ssm):
    """s5b200_scan_c64 on the reference's own Bu = B_bar u against the states its lax.scan / associative_scan returned."""
    import torch
    from sparsernns_b200.model import scan_c64

    Lb, Bb, u = ref_ssm["zoh/Lambda_bar"], ref_ssm["zoh/B_bar"], ref_ssm["u"]
    Bu = (u @ Bb.real.T + 1j * (u @ Bb.imag.T)).astype(np.complex64)
    xs = scan_c64(torch.from_numpy(Bu).to(cuda_device), torch.from_numpy(Lb).to(cuda_device)).cpu().numpy()
    assert _rel(xs, ref_ssm["float_seq_lin/xs"]) <= TOL_F32 and _rel(xs, ref_ssm["float_assoc_lin/xs"]) <= TOL_F32


def test_static_ssm_vs_reference(cuda_device, ref_ssm):
    """Static w8a16 apply_ssm of block 1 against build_apply_ssm_static_quant._apply_ssm (ssm.py:194-343) evaluated by the
    reference with the fixture's frozen scales: states, and y = FQ16(ys + FQ8(D) u_q) built from the reference's ys / u_q
    with the reference's own _quantdequant semantics (ssm.py:594-598)."""
    import torch

    m, tree = _bound_from_fixture(True, True)
    mx = tree["params"]["encoder"]["layers_1"]["mixer"]
    u = ref_ssm["u"]
    y, xs = m.apply_ssm(1, torch.from_numpy(u[None]).to(cuda_device))
    assert _rel(xs.cpu().numpy()[0], ref_ssm["static_relu_nobias/xs"]) <= TOL_W8A16
    fq = lambda v, s, bits: (np.clip(np.round(v / np.float32(s)), -(2 ** (bits - 1)), 2 ** (bits - 1) - 1) * np.float32(s)).astype(np.float32)
    Dq = fq(mx["D"], mx["quant_D"]["scale"], 8)
    y_ref = fq(ref_ssm["static_relu_nobias/ys"] + Dq * ref_ssm["static_relu_nobias/uq"], mx["quant_yt"]["scale"], 16)
    assert _rel(y.cpu().numpy()[0], y_ref) <= TOL_W8A16
