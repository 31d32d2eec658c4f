"""The oracle against vectors produced by the REFERENCE'S OWN SOURCES.

tests/golden/reference_*.npz were written by tests/golden/make_reference_golden.py, which executes the reference's
fxparray.py / fxputils.py / fxpmodel.py and the pure functions of model/ssm.py and utils/quantization.py in place over a
NumPy stand-in for jax.numpy (tests/golden/jaxshim.py; JAX itself is not installable in the build container).  Integer
results must be equal bit for bit, float results within float32 rounding of a different summation order.
Where /root/reference exists (the build container) the fixtures are also regenerated live and compared.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)

import refcases  # noqa: E402
from oracle import fxp as OF  # noqa: E402
from oracle import fxpmodel as OFM  # noqa: E402
from oracle import quant as OQ  # noqa: E402
from oracle import ssm as OS  # noqa: E402
from oracle.fxpconfig import build_fxp_qconfig  # noqa: E402
from sparsernns_b200.checkpoint import load_npz  # noqa: E402


def _load(name):
    with np.load(os.path.join(GOLD, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def ref_ops():
    return _load("reference_ops.npz")


@pytest.fixture(scope="module")
def ref_model():
    return _load("reference_fxp_model.npz")


@pytest.fixture(scope="module")
def ref_ssm():
    return _load("reference_ssm.npz")


# ---------------------------------------------------------------------------------------------------------------------
def test_fxparray_ops_bit_exact(ref_ops):
    cases = refcases.op_cases()
    assert len(cases) == len(ref_ops) // 2 >= 80
    for case in cases:
        data, bits, exp, signed = refcases.run_case(OF, case)
        cfg = ref_ops[f"{case['name']}/cfg"]
        assert (bits, exp, int(signed)) == tuple(int(v) for v in cfg), case["name"]
        assert np.array_equal(data, ref_ops[f"{case['name']}/data"]), case["name"]


def _check_fxp(got, ref, key, what):
    cfg = ref[f"{key}/cfg"]
    assert (got.bits, got.exp, int(got.signed)) == tuple(int(v) for v in cfg), (what, key)
    assert np.array_equal(np.asarray(got.data, np.int32), ref[f"{key}/data"]), (what, key)


def _oracle_model(dim_scale, shared_exp=True):
    tree = load_npz(os.path.join(GOLD, f"ndns_ds{dim_scale}_sparse_w8a16.npz"))
    md, q = build_fxp_qconfig(tree["params"], tree["cal_stats"], "w8a16", shared_exp=shared_exp)
    return OFM.FxpRegressionModel(md, q, 3), q


def _jsonable(t):
    if isinstance(t, dict):
        return {str(k): _jsonable(v) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return [_jsonable(v) for v in t]
    if isinstance(t, (np.ndarray, np.generic)):
        return np.asarray(t).tolist()
    return t


def test_fxp_qconfig_equals_reference(ref_model):
    """fxputils.load_modeldict / create_fxp_qconfig / add_target_bits_exp: every bit width and exponent."""
    for case, ds, shared in (("A", 0.5, True), ("C", 0.5, False), ("E", 1.0, True)):
        _, q = _oracle_model(ds, shared)
        ref_q = json.loads(str(ref_model[f"{case}/qconfig_json"]))
        assert json.loads(json.dumps(_jsonable(q), sort_keys=True)) == ref_q, case


# reference sow() name -> oracle tap
_LAYER_TAPS = {
    "norm/norm_input_minus_mean": ("norm", "norm_input_minus_mean"), "norm/norm_output_raw": ("norm", "norm_output_raw"),
    "norm/norm_output_scaled": ("norm", "norm_output_scaled"), "norm/norm_output": ("norm", "norm_output"),
    "mixer/Bu_elements": ("mixer", "Bu"), "mixer/xs": ("mixer", "xs"), "mixer/xs_relu": ("mixer", "xs_relu"),
    "mixer/2Cxs": ("mixer", "two_Cxs"), "mixer/Du": ("mixer", "Du"), "mixer/ys": ("mixer", "ys"),
    "pre_s5": ("pre_s5",), "pre_C": ("pre_C",), "pre_GLU": ("pre_GLU",), "out2/__call__": ("out2",),
    "out2_sigmoid": ("out2_sigmoid",), "post_GLU": ("post_GLU",), "residadd": ("residadd",), "output": ("output",),
}


def test_fxp_model_every_intermediate_bit_exact(ref_model):
    """FxpRegressionModel.forward (fxpmodel.py:1431-1439) on the golden dim_scale 0.5 model: the output and every
    sow()-named intermediate of every block equal the reference's, payload and (bits, exp, signed)."""
    m, _ = _oracle_model(0.5)
    taps = {}
    y = m(m.quantize_input(ref_model["A/x"]), taps)
    _check_fxp(y, ref_model, "A/y", "output")
    _check_fxp(taps["encoder_output_relu"], ref_model, "A/intermediates/encoder/encoder_output_relu", "encoder")
    n = 0
    for i in range(3):
        for ref_name, path in _LAYER_TAPS.items():
            t = taps[f"layers_{i}"]
            for p in path:
                t = t[p]
            key = f"A/intermediates/encoder/layers_{i}/{ref_name}"
            if isinstance(t, OF.ComplexFxpArray):
                _check_fxp(t.real, ref_model, key + ".re", i)
                _check_fxp(t.imag, ref_model, key + ".im", i)
            else:
                _check_fxp(t, ref_model, key, i)
            n += 1
    assert n == 3 * len(_LAYER_TAPS)


@pytest.mark.parametrize("case,dim_scale,shared", [("B", 0.5, True), ("C", 0.5, False), ("E", 1.0, True)],
                         ids=["saturating_and_silent_input", "per_layer_exponents", "dim_scale_1.0"])
def test_fxp_model_outputs_bit_exact(ref_model, case, dim_scale, shared):
    m, _ = _oracle_model(dim_scale, shared)
    x = ref_model[f"{case}/x"] if f"{case}/x" in ref_model else ref_model["A/x"]
    taps = {}
    y = m(m.quantize_input(x), taps)
    _check_fxp(y, ref_model, f"{case}/y", case)
    for i in range(3):
        _check_fxp(taps[f"layers_{i}"]["output"], ref_model, f"{case}/layer_{i}_output", (case, i))


def test_fxp_model_unbatched_clip(ref_model):
    """(L, 257) input: the non-vmapped recurrence of fxpmodel.py:683-687."""
    m, _ = _oracle_model(0.5)
    y = m(m.quantize_input(ref_model["D/x"][None]))
    assert np.array_equal(y.data[0], ref_model["D/y/data"]) and y.exp == int(ref_model["D/y/cfg"][1])


# ---------------------------------------------------------------------------------------------------------------------
def _mixer(layer=1):
    tree = load_npz(os.path.join(GOLD, "ndns_ds0.5_sparse_w8a16.npz"))
    return tree["params"]["encoder"][f"layers_{layer}"]["mixer"]


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_discretisation_matches_reference(ref_ssm):
    Lb, Bb, Ct, D = OS.ssm_discrete_params(_mixer(), clip_eigs=True, dtype=np.float32)
    assert _rel(Lb, ref_ssm["zoh/Lambda_bar"]) <= 2e-7 and _rel(Bb, ref_ssm["zoh/B_bar"]) <= 1e-6
    Lb2, Bb2, _, _ = OS.ssm_discrete_params(_mixer(), clip_eigs=True, discretization="bilinear", dtype=np.float32)
    assert _rel(Lb2, ref_ssm["bilinear/Lambda_bar"]) <= 1e-6 and _rel(Bb2, ref_ssm["bilinear/B_bar"]) <= 1e-6


def test_float_ssm_matches_reference(ref_ssm):
    """build_apply_ssm._apply_ssm (ssm.py:84-185), sequential and associative scans, with and without ReLU."""
    Lb, Bb = ref_ssm["zoh/Lambda_bar"], ref_ssm["zoh/B_bar"]
    mx = _mixer()
    Ct = (mx["C"][..., 0] + 1j * mx["C"][..., 1]).astype(np.complex64)
    u = ref_ssm["u"]
    for assoc in (False, True):
        for relu in (False, True):
            ys, xs = OS.apply_ssm_float(Lb, Bb, Ct, u, relufication=relu, associative_scan=assoc)
            tag = f"float_{'assoc' if assoc else 'seq'}_{'relu' if relu else 'lin'}"
            assert _rel(xs, ref_ssm[f"{tag}/xs"]) <= 1e-5, tag
            assert _rel(ys, ref_ssm[f"{tag}/ys"]) <= 1e-5, tag


def test_static_ssm_matches_reference(ref_ssm):
    """build_apply_ssm_static_quant._apply_ssm (ssm.py:194-343) with the fixture's frozen scales: the quantised input is
    identical, states and outputs agree to float32 rounding (the reference multiplies complex numbers, the oracle writes
    the product out in real arithmetic)."""
    mx = _mixer()
    g = lambda *k: np.float32(OS_get(mx, k)["scale"])
    scales = dict(ut=g("quant_ut"), A_re=g("quant_A", "quant_real"), A_im=g("quant_A", "quant_imag"),
                  B_re=g("quant_B", "quant_real"), B_im=g("quant_B", "quant_imag"),
                  C_re=g("quant_C", "quant_real"), C_im=g("quant_C", "quant_imag"),
                  But_re=g("quant_But", "quant_real"), But_im=g("quant_But", "quant_imag"))
    bits = dict(a=16, b=8, c=8, ssm_act=16)
    Ct = (mx["C"][..., 0] + 1j * mx["C"][..., 1]).astype(np.complex64)
    for relu in (False, True):
        for bias in (None, ref_ssm["static/B_bias"]):
            ys, xs, uq = OS.apply_ssm_static(ref_ssm["zoh/Lambda_bar"], ref_ssm["zoh/B_bar"], Ct, ref_ssm["u"], scales, bits,
                                             relufication=relu, B_bias=bias)
            tag = f"static_{'relu' if relu else 'lin'}_{'bias' if bias is not None else 'nobias'}"
            assert np.array_equal(uq, ref_ssm[f"{tag}/uq"]), tag
            assert _rel(xs, ref_ssm[f"{tag}/xs"]) <= 1e-5, tag
            assert _rel(ys, ref_ssm[f"{tag}/ys"]) <= 1e-5, tag


def OS_get(d, keys):
    for k in keys:
        d = d[k]
    return d


def test_quantdequant_and_qparams_bit_exact(ref_ssm):
    """_quantdequant (quantization.py:386-400) incl. ties and outliers clipped by far more than a factor of two, where the
    straight-through form x + (xdq - x) is NOT exactly on the grid; _calculate_qparams power-of-two scales (:352-370)."""
    x = ref_ssm["qdq/x"]
    for bits in (8, 16):
        got = OQ.quantdequant(x, np.float32(0.125), bits)
        assert np.array_equal(got, ref_ssm[f"qdq/y{bits}"]), bits
    for bits in (8, 16):
        got = np.array([OQ.calculate_qparams(a, b, bits) for a, b in zip(ref_ssm["qparams/min"], ref_ssm["qparams/max"])],
                       np.float32)
        assert np.array_equal(got, ref_ssm[f"qparams/scale{bits}"]), bits


def test_fxp_nonlinearities_and_recurrence_bit_exact(ref_ssm):
    """FxpSigmoid (fxpmodel.py:70-144), complex fxp_relu with its float32 round trip (:27-45), recurrent_loop (:147-208)."""
    for x_exp, y_exp in ((6, 14), (4, 14), (6, 6)):
        k = f"sigmoid_{x_exp}_{y_exp}"
        sg = OFM.FxpSigmoid(x_exp=x_exp, y_exp=y_exp)
        r = sg.apply(OF.FxpArray(ref_ssm[f"{k}/x"], 16, int(ref_ssm[f"{k}/x_exp"]), True))
        assert np.array_equal(r.data, ref_ssm[f"{k}/y/data"]), k
        assert (r.bits, r.exp) == tuple(int(v) for v in ref_ssm[f"{k}/y/cfg"][:2]), k
    z = OF.ComplexFxpArray(OF.FxpArray(ref_ssm["crelu/re"], 32, 10, True), OF.FxpArray(ref_ssm["crelu/im"], 32, 10, True))
    zr = OF.fxp_relu(z)
    assert np.array_equal(zr.real.data, ref_ssm["crelu/out_re"]) and np.array_equal(zr.imag.data, ref_ssm["crelu/out_im"])
    assert (ref_ssm["crelu/out_re"] != np.where(ref_ssm["crelu/out_re"] != 0, ref_ssm["crelu/re"], 0)).any(), \
        "the fixture must contain kept values that change in the float32 round trip"
    e = [int(v) for v in ref_ssm["rec/exps"]]
    xr, xi = OFM.recurrent_loop(ref_ssm["rec/Bre"], ref_ssm["rec/Bim"], ref_ssm["rec/Are"], ref_ssm["rec/Aim"], *e)
    assert np.array_equal(xr, ref_ssm["rec/xs"][..., 0]) and np.array_equal(xi, ref_ssm["rec/xs"][..., 1])


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.isdir("/root/reference/sparseRNNs"), reason="the reference sources are not on this box")
def test_fixtures_regenerate_from_the_reference_sources(tmp_path):
    """Live: run the generator against /root/reference now and compare with the committed fixtures array by array."""
    env = dict(os.environ, S5B200_REFGOLDEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(GOLD, "make_reference_golden.py")], env=env, capture_output=True,
                       text=True, cwd=os.path.dirname(HERE), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("reference_ops.npz", "reference_fxp_model.npz", "reference_ssm.npz"):
        new, old = _load(os.path.join(str(tmp_path), name)), _load(name)
        assert set(new) == set(old), name
        for k in old:
            assert np.array_equal(new[k], old[k]), (name, k)
