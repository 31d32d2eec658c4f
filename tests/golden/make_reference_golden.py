"""Golden vectors produced by the REFERENCE'S OWN SOURCES (run from the repo root in the build container:
``python tests/golden/make_reference_golden.py``).

/root/reference is pure Python over JAX, which is not installable here.  ``jaxshim.install()`` registers a NumPy-backed
stand-in for the small part of jax the hot path touches (see its docstring for what that does and does not pin), and
this script then imports and EXECUTES the reference's fxparray.py, fxputils.py, fxpmodel.py and the pure functions of
model/ssm.py and utils/quantization.py, in place, on the shipped model fixtures.  Nothing of the reference is copied:
only inputs and the values it returned are stored.

  reference_ops.npz        every case of refcases.op_cases() through fxparray.py: payload + (bits, exp, signed)
  reference_fxp_model.npz  fxputils.load_modeldict / create_fxp_qconfig / add_target_bits_exp and
                           FxpRegressionModel.forward with store_intermediates (fxprun.py:302-460) on the golden
                           dim_scale 0.5 and 1.0 models: qconfig, output, every sow()-named intermediate
  reference_ssm.npz        discretize_zoh, _calculate_qparams, _quantdequant, build_apply_ssm (both scans),
                           build_apply_ssm_static_quant (frozen scales), FxpSigmoid, fxp_relu, recurrent_loop

The fixtures travel with the repo; tests/test_oracle_vs_reference.py (CPU) checks the oracle against them and, where
/root/reference exists, re-runs the reference live; tests/test_gpu_reference.py checks the CUDA path against them.
"""

import json
import os
import pickle
import sys
import tempfile

import numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("S5B200_REFGOLDEN_OUT", HERE)     # the live regeneration test writes to a scratch directory
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import jaxshim  # noqa: E402

jaxshim.install()

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402

import refcases  # noqa: E402
from sparsernns_b200 import synth  # noqa: E402
from sparsernns_b200.checkpoint import load_npz  # noqa: E402

import logging  # noqa: E402

from sparseRNNs import fxparray as RA  # noqa: E402
from sparseRNNs import fxpmodel as RM  # noqa: E402
from sparseRNNs import fxputils as RU  # noqa: E402
from sparseRNNs.model import ssm as RS  # noqa: E402
from sparseRNNs.utils import quantization as RQ  # noqa: E402
from sparseRNNs.utils.logging import logger as ref_logger  # noqa: E402

ref_logger.setLevel(logging.ERROR)
for h in ref_logger.handlers:
    h.setLevel(logging.ERROR)

RA._asarray = jnp.asarray      # refcases.mod_array: operands become shim arrays for the reference
RA.print_call_stack = lambda: None   # the reference prints its call stack on every saturating clip


def _jtree(t):
    return jax.tree_util.tree_map(lambda v: jnp.asarray(v), t)


def _np(x):
    return onp.asarray(x)


# ---------------------------------------------------------------------------------------------------------------------
def make_ops():
    out = {}
    for case in refcases.op_cases():
        data, bits, exp, signed = refcases.run_case(RA, case)
        out[f"{case['name']}/data"] = data
        out[f"{case['name']}/cfg"] = onp.array([bits, exp, int(signed)], onp.int32)
    onp.savez_compressed(os.path.join(OUT, "reference_ops.npz"), **out)
    print(f"reference_ops.npz: {len(out) // 2} cases")


# ---------------------------------------------------------------------------------------------------------------------
def reference_fxp_model(params, stats, shared_exp=True, store=True):
    """fxprun.py:302-460 with the reference's own functions; returns (model, modeldict, qconfig)."""
    d = tempfile.mkdtemp()
    pf, sf = os.path.join(d, "sc_calibrated_params.pkl"), os.path.join(d, "sc_cal_stats.pkl")
    with open(pf, "wb") as f:
        pickle.dump(_jtree(params), f)
    with open(sf, "wb") as f:
        pickle.dump(_jtree(stats), f)
    md = RU.load_modeldict(pf, sf)
    if shared_exp:
        _, q = RU.create_fxp_qconfig(md, agg="max", store_json=True, json_store_folder=d)
    else:
        q = RU.create_fxp_qconfig(md, agg=None, store_json=True, json_store_folder=d)
    prec = dict(non_ssm_w=8, non_ssm_b=16, non_ssm_act=16, ssm_w=8, ssm_act=16)     # fxprun.py:302-308, w8a16
    q = RU.add_target_bits_exp(md, q, prec, shared_exp=shared_exp)
    n_layers = max(int(e.split("_")[1]) for e in md["encoder"] if e.startswith("layers_")) + 1
    d_model = md["encoder"]["encoder"]["kernel"].shape[-1]
    d_ssm = md["encoder"]["layers_0"]["mixer"]["Lambda_re"].shape[0]
    none = RQ.QuantizationConfig.none()
    mixer_cls = RM.FxpSSM.init_fn(H=d_model, P=d_ssm, discretization="zoh", conj_sym=True, q_config=none,
                                  bidirectional=False, relufication=True, associative_scan=False)
    model = RM.FxpRegressionModel(modeldict=md, fxp_qconfig=q, scope="model", mixer_cls=mixer_cls, n_layers=n_layers,
                                  d_model=d_model, batchnorm=True, prenorm=True, bn_momentum=0.95, glu_variant="half1",
                                  step_rescale=1.0, relufication=True, fuse_batchnorm_linear=False, q_config=none,
                                  dropout=0.0, training=False, d_output=md["decoder"]["kernel"].shape[-1], padded=False,
                                  store_intermediates=store)
    return model, md, q


def _flat_fxp(prefix, node, out):
    """Flatten the export() tree: FxpArray -> data + cfg, ComplexFxpArray -> re / im."""
    if isinstance(node, dict):
        for k, v in node.items():
            _flat_fxp(f"{prefix}/{k}", v, out)
    elif isinstance(node, RA.ComplexFxpArray):
        _flat_fxp(f"{prefix}.re", node.real, out)
        _flat_fxp(f"{prefix}.im", node.imag, out)
    elif isinstance(node, RA.FxpArray):
        out[f"{prefix}/data"] = _np(node.data).astype(onp.int32)
        out[f"{prefix}/cfg"] = onp.array([node.bits, node.exp, int(node.signed)], onp.int32)


def _jsonable(t):
    if isinstance(t, dict):
        return {str(k): _jsonable(v) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return [_jsonable(v) for v in t]
    if isinstance(t, (onp.ndarray, onp.generic)):
        return _np(t).tolist()
    return t


def run_input(model, q, x):
    fx = RA.fxp_from_fp(jnp.asarray(x), bits=q["encoder"]["inp_bits"], exp=q["encoder"]["inp_exp"], signed=True,
                        round_mode=RA.RoundingMode.FLOOR)                               # fxprun.py:69-75
    return model(fx)


def make_fxp_model():
    out = {}
    tree = load_npz(os.path.join(HERE, "ndns_ds0.5_sparse_w8a16.npz"))
    params, stats = tree["params"], tree["cal_stats"]
    # A: shared exponents, every intermediate
    model, md, q = reference_fxp_model(params, stats, shared_exp=True)
    x = synth.synthetic_stft_batch(2, 96, seed=123)
    y = run_input(model, q, x)
    out["A/x"] = x
    out["A/qconfig_json"] = onp.array(json.dumps(_jsonable(q), sort_keys=True))
    _flat_fxp("A/y", y, out)
    _flat_fxp("A/intermediates", model.export()["intermediates"], out)
    # B: input far outside the calibrated range (saturation everywhere) and a silent clip
    xb = onp.concatenate([x[:1] * 60.0, x[1:] * 0.0 - 0.0007], 0).astype(onp.float32)
    model, _, q = reference_fxp_model(params, stats, shared_exp=True)
    yb = run_input(model, q, xb)
    out["B/x"] = xb
    _flat_fxp("B/y", yb, out)
    inter = model.export()["intermediates"]
    for i in range(3):
        _flat_fxp(f"B/layer_{i}_output", inter["encoder"][f"layer_{i}_output"], out)
    # C: per-layer exponents (--separate_exponents)
    model, _, qc = reference_fxp_model(params, stats, shared_exp=False)
    yc = run_input(model, qc, x)
    out["C/qconfig_json"] = onp.array(json.dumps(_jsonable(qc), sort_keys=True))
    _flat_fxp("C/y", yc, out)
    inter = model.export()["intermediates"]
    for i in range(3):
        _flat_fxp(f"C/layer_{i}_output", inter["encoder"][f"layer_{i}_output"], out)
    # D: a single un-batched clip (L, 257): the non-vmapped recurrence (fxpmodel.py:683-687)
    model, _, q = reference_fxp_model(params, stats, shared_exp=True, store=False)
    yd = run_input(model, q, x[1])
    out["D/x"] = x[1]
    _flat_fxp("D/y", yd, out)
    # E: dim_scale 1.0 model (H 192, P 128)
    tree = load_npz(os.path.join(HERE, "ndns_ds1.0_sparse_w8a16.npz"))
    model, _, qe = reference_fxp_model(tree["params"], tree["cal_stats"], shared_exp=True)
    xe = synth.synthetic_stft_batch(1, 80, seed=7)
    ye = run_input(model, qe, xe)
    out["E/x"] = xe
    out["E/qconfig_json"] = onp.array(json.dumps(_jsonable(qe), sort_keys=True))
    _flat_fxp("E/y", ye, out)
    inter = model.export()["intermediates"]
    for i in range(3):
        _flat_fxp(f"E/layer_{i}_output", inter["encoder"][f"layer_{i}_output"], out)
    onp.savez_compressed(os.path.join(OUT, "reference_fxp_model.npz"), **out)
    print(f"reference_fxp_model.npz: {len(out)} arrays")


# ---------------------------------------------------------------------------------------------------------------------
class _Frozen:
    """A frozen FakeQuant: the reference's own _quantdequant with a fixed scale (quantization.py:386-400, :428-446)."""

    def __init__(self, scale, bits):
        self.scale, self.bits = jnp.asarray(onp.float32(scale)), bits

    def __call__(self, x):
        return RQ._quantdequant(x, scale=self.scale, zero_point=0.0, precision=self.bits)

    def has_variable(self, *a):      # no observer in batch_stats: frozen inference (ssm.py:254, :283-284)
        return False


class _FrozenComplex:
    """FakeQuantComplex (quantization.py:451-474): real and imaginary parts quantised separately."""

    def __init__(self, s_re, s_im, bits):
        self.quant_real, self.quant_imag = _Frozen(s_re, bits), _Frozen(s_im, bits)

    def __call__(self, x):
        return self.quant_real(x.real) + 1j * self.quant_imag(x.imag)


def make_ssm():
    out = {}
    rng = onp.random.default_rng(77)
    tree = load_npz(os.path.join(HERE, "ndns_ds0.5_sparse_w8a16.npz"))
    mx = tree["params"]["encoder"]["layers_1"]["mixer"]
    H, P = 96, 64
    # qS5SSM.setup, ssm.py:436-523, on the stored parameters (clip_eigs=True as main.py's NDNS default)
    lre = jnp.clip(jnp.asarray(mx["Lambda_re"]), None, -1e-4)
    Lambda = lre + 1j * jnp.asarray(mx["Lambda_im"])
    B_tilde = jnp.asarray(mx["B"])[..., 0] + 1j * jnp.asarray(mx["B"])[..., 1]
    C_tilde = jnp.asarray(mx["C"])[..., 0] + 1j * jnp.asarray(mx["C"])[..., 1]
    step = 1.0 * jnp.exp(jnp.asarray(mx["log_step"])[:, 0])
    Lb, Bb = RS.discretize_zoh(Lambda, B_tilde, step)
    out["zoh/Lambda_bar"], out["zoh/B_bar"] = _np(Lb), _np(Bb)
    Lb2, Bb2 = RS.discretize_bilinear(Lambda, B_tilde, step)
    out["bilinear/Lambda_bar"], out["bilinear/B_bar"] = _np(Lb2), _np(Bb2)
    u = rng.standard_normal((150, H)).astype(onp.float32)
    out["u"] = u
    # float path, both scans (ssm.py:84-185)
    q_none = RQ.QuantizedOperations(RQ.QuantizationConfig.none())
    for assoc in (False, True):
        for relu in (False, True):
            ys, xs = RS.build_apply_ssm(q_none, associative_scan=assoc)(Lb, Bb, C_tilde, jnp.asarray(u), True, False,
                                                                        relufication=relu)
            tag = f"float_{'assoc' if assoc else 'seq'}_{'relu' if relu else 'lin'}"
            out[f"{tag}/ys"], out[f"{tag}/xs"] = _np(ys), _np(xs)
    # static w8a16 path with the frozen scales of the fixture (ssm.py:194-343)
    g = lambda *k: float(onp.float32(_get(mx, k)["scale"]))
    qf = (_FrozenComplex(g("quant_A", "quant_real"), g("quant_A", "quant_imag"), 16),
          _FrozenComplex(g("quant_B", "quant_real"), g("quant_B", "quant_imag"), 8),
          _FrozenComplex(g("quant_C", "quant_real"), g("quant_C", "quant_imag"), 8),
          _Frozen(g("quant_D"), 8),
          _FrozenComplex(1.0, 1.0, 16),     # quant_xt: its result is discarded (ssm.py:314)
          _Frozen(g("quant_ut"), 16), _Frozen(g("quant_yt"), 16),
          _FrozenComplex(g("quant_But", "quant_real"), g("quant_But", "quant_imag"), 16))
    q_static = RQ.QuantizedOperations(RQ.quantization_recipe_map["w8a16"](static_quant=True))   # np.dot / np.multiply
    B_bias = (rng.standard_normal(P) * 0.01 + 1j * rng.standard_normal(P) * 0.01).astype(onp.complex64)
    out["static/B_bias"] = B_bias
    for relu in (False, True):
        for bias in (None, B_bias):
            ys, xs, uq = RS.build_apply_ssm_static_quant(q_static, associative_scan=False)(
                Lb, Bb, C_tilde, jnp.asarray(u), True, False, relufication=relu,
                B_bias=None if bias is None else jnp.asarray(bias), quant_fns=qf)
            tag = f"static_{'relu' if relu else 'lin'}_{'bias' if bias is not None else 'nobias'}"
            out[f"{tag}/ys"], out[f"{tag}/xs"], out[f"{tag}/uq"] = _np(ys), _np(xs), _np(uq)
    # quantiser primitives (quantization.py:352-400)
    xs_q = onp.concatenate([rng.standard_normal(300).astype(onp.float32) * 2,
                            onp.array([0.5, 1.5, 2.5, -0.5, -1.5, 1e4, -1e4, 3.3, 127.49, 128.5, 2.4e7, -4.1e7, 8e9, -3.3e8],
                                      onp.float32) * 0.125])   # the last four: ulp(x) > scale, x + (xdq - x) is off the grid
    out["qdq/x"] = xs_q
    for bits in (8, 16):
        out[f"qdq/y{bits}"] = _np(RQ._quantdequant(jnp.asarray(xs_q), jnp.asarray(onp.float32(0.125)), 0.0, bits))
    mins = -onp.abs(rng.standard_normal(60)).astype(onp.float32) * onp.exp(rng.uniform(-8, 8, 60)).astype(onp.float32)
    maxs = onp.abs(rng.standard_normal(60)).astype(onp.float32) * onp.exp(rng.uniform(-8, 8, 60)).astype(onp.float32)
    out["qparams/min"], out["qparams/max"] = mins, maxs
    for bits in (8, 16):
        out[f"qparams/scale{bits}"] = onp.array([
            _np(RQ._calculate_qparams(jnp.asarray(a), jnp.asarray(b), bits, RQ.QuantScheme.per_tensor_symmetric,
                                      pow2scale=True)[0]) for a, b in zip(mins, maxs)], onp.float32)
    # fixed-point non-linearities and the integer recurrence (fxpmodel.py:27-208)
    for x_exp, y_exp in ((6, 14), (4, 14), (6, 6)):
        sg = RM.FxpSigmoid(x_exp=x_exp, y_exp=y_exp)
        xin = RA.FxpArray(data=jnp.asarray(rng.integers(-40000, 40000, 500).astype(onp.int32)), bits=16, exp=x_exp + 3,
                          signed=True)
        r = sg.apply(xin) if hasattr(sg, "apply") else sg(xin)
        out[f"sigmoid_{x_exp}_{y_exp}/x"] = _np(xin.data)
        out[f"sigmoid_{x_exp}_{y_exp}/x_exp"] = onp.int32(xin.exp)
        _flat_fxp(f"sigmoid_{x_exp}_{y_exp}/y", r, out)
    re = rng.integers(-2 ** 27, 2 ** 27, 400).astype(onp.int32)
    im = rng.integers(-2 ** 27, 2 ** 27, 400).astype(onp.int32)
    re[:40] = 0
    re[40:60] = rng.integers(2 ** 24, 2 ** 30, 20)            # kept values above 2**24 change in the float32 round trip
    im[:20] = rng.integers(2 ** 24, 2 ** 30, 20)
    z = RA.ComplexFxpArray(real=RA.FxpArray(data=jnp.asarray(re), bits=32, exp=10, signed=True),
                           imag=RA.FxpArray(data=jnp.asarray(im), bits=32, exp=10, signed=True))
    zr = RM.fxp_relu(z)
    out["crelu/re"], out["crelu/im"] = re, im
    out["crelu/out_re"], out["crelu/out_im"] = _np(zr.real.data), _np(zr.imag.data)
    L, Pn = 70, 8
    Bre = rng.integers(-30000, 30000, (2, L, Pn)).astype(onp.int32)
    Bim = rng.integers(-30000, 30000, (2, L, Pn)).astype(onp.int32)
    Are = rng.integers(-32768, 32767, Pn).astype(onp.int32)
    Aim = rng.integers(-32768, 32767, Pn).astype(onp.int32)
    out["rec/Bre"], out["rec/Bim"], out["rec/Are"], out["rec/Aim"] = Bre, Bim, Are, Aim
    exps = dict(eBu_re=15, eBu_im=14, eA_re=15, eA_im=15, ex_re=13, ex_im=16)
    out["rec/exps"] = onp.array([exps[k] for k in ("eBu_re", "eBu_im", "eA_re", "eA_im", "ex_re", "ex_im")], onp.int32)
    loop = jax.vmap(RM.recurrent_loop, in_axes=(0, 0, None, None, 0, 0, *(None,) * 6))
    xs_raw = loop(jnp.asarray(Bre), jnp.asarray(Bim), jnp.asarray(Are), jnp.asarray(Aim),
                  jnp.zeros((2, Pn), dtype=jnp.int32), jnp.zeros((2, Pn), dtype=jnp.int32),
                  exps["eBu_re"], exps["eBu_im"], exps["eA_re"], exps["eA_im"], exps["ex_re"], exps["ex_im"])
    out["rec/xs"] = _np(xs_raw)
    onp.savez_compressed(os.path.join(OUT, "reference_ssm.npz"), **out)
    print(f"reference_ssm.npz: {len(out)} arrays")


def _get(d, keys):
    for k in keys:
        d = d[k]
    return d


if __name__ == "__main__":
    cwd = os.getcwd()
    make_ops()
    make_fxp_model()
    os.chdir(cwd)
    make_ssm()
    for f in sorted(os.listdir(OUT)):
        if f.startswith("reference_"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
