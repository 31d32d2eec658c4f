"""CPU-only tests: the C ABI library loads and exports every declared symbol, host-side logic, oracle vs golden."""
import os
import re

import numpy as np
import pytest

import modelzoo as Z
from oracle import fxpconfig as OC
from oracle import fxpmodel as OFM
from oracle import model as OM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_header_symbol():
    from sparsernns_b200 import _abi
    from sparsernns_b200.build import build_library

    build_library()
    header = open(os.path.join(ROOT, "include", "s5b200.h")).read()
    declared = set(re.findall(r"\b(s5b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_abi.PROTOTYPES), (declared ^ set(_abi.PROTOTYPES))
    lib = _abi.lib()           # raises if a symbol is missing
    assert lib.s5b200_version() == 100
    assert lib.s5b200_scan_c64_workspace_bytes(2, 3751, 64) > 0
    assert lib.s5b200_scan_c64_workspace_bytes(2, 3751, 1000) == 0


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "sparsernns_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_product_fails_loudly_without_cuda():
    import torch
    from sparsernns_b200 import _abi
    from sparsernns_b200.model import BoundModel, S5Config

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_abi.S5Error):
        BoundModel(S5Config(), {"params": {}, "batch_stats": {}})


def test_dims_and_erk():
    from sparsernns_b200 import synth

    d = synth.ndns_dims(0.5)
    assert (d.d_model, d.ssm_size, d.blocks, d.n_layers) == (96, 64, 8, 3)
    d = synth.ndns_dims(1.0)
    assert (d.d_model, d.ssm_size) == (192, 128)
    shapes = {"enc": (257, 96), "B": (64, 96, 2), "C": (96, 64, 2), "out2": (96, 96), "dec": (96, 257), "log_step": (64, 1)}
    sp = synth.erk_sparsities(shapes, 0.9)
    assert sp["log_step"] == 0.0                                   # ERK keeps the (P, 1) tensor dense
    tot = sum(np.prod(s) for s in shapes.values())
    nz = sum(np.prod(s) * (1 - sp[k]) for k, s in shapes.items())
    assert abs(nz / tot - 0.1) < 1e-9                              # overall sparsity is exactly the target
    assert all(0.8 < v < 0.95 for k, v in sp.items() if k != "log_step")


def test_pruned_model_sparsity_and_b_union():
    zoo = Z.build_sparse_calibrated(0.5, n_clips=1, n_frames=64)
    p = zoo["variables"]["params"]["encoder"]["layers_0"]
    assert 0.85 < (p["mixer"]["B"] == 0).mean() < 0.95
    assert 0.80 < (p["out2"]["kernel"] == 0).mean() < 0.95
    from oracle import ssm as OS
    Lb, Bb, _, _ = OS.ssm_discrete_params(p["mixer"], clip_eigs=True)
    # B_bar = s_p * (B_re + i B_im) mixes the parts: zero only where BOTH stored parts are zero
    both = (p["mixer"]["B"][..., 0] == 0) & (p["mixer"]["B"][..., 1] == 0)
    assert np.array_equal(Bb.real == 0, both)


def test_product_fxputils_equals_oracle():
    from sparsernns_b200 import fxputils as PU

    for bn in (True, False):
        for shared in (True, False):
            zoo = Z.build_sparse_calibrated(0.5, n_clips=1, n_frames=96, bn_affine=bn)
            _, qo = OC.build_fxp_qconfig(zoo["cal_params"], zoo["cal_stats"], "w8a16", shared_exp=shared)
            qp = PU.create_fxp_qconfig(PU.load_modeldict(zoo["cal_params"], zoo["cal_stats"]), "w8a16", shared_exp=shared)

            def walk(a, b, path=""):
                if isinstance(a, dict):
                    for k, v in a.items():
                        if k.endswith("signbits") or k.endswith("signbit"):
                            continue
                        assert k in b, path + "/" + k
                        walk(v, b[k], path + "/" + k)
                else:
                    assert a == b, (path, a, b)

            walk(qo, qp)


def test_modeldict_log2_quirk_and_nan_replacement():
    """fxputils.py:128-131 takes log2 of EVERY '*scale*' leaf, BatchNorm gamma included; NaNs -> 1.0 (:711-718)."""
    from sparsernns_b200 import fxputils as PU

    zoo = Z.build_sparse_calibrated(0.5, n_clips=1, n_frames=64)
    zoo["cal_params"]["encoder"]["layers_1"]["norm"]["scale"][3] = -0.5
    md = PU.load_modeldict(zoo["cal_params"], zoo["cal_stats"])
    g = zoo["cal_params"]["encoder"]["layers_0"]["norm"]["scale"]
    assert np.allclose(md["encoder"]["layers_0"]["norm"]["scale"], np.log2(g))
    assert np.isnan(md["encoder"]["layers_1"]["norm"]["scale"][3])
    PU.create_fxp_qconfig(md)
    assert md["encoder"]["layers_1"]["norm"]["scale"][3] == 1.0


def test_checkpoint_roundtrip(tmp_path):
    from sparsernns_b200.checkpoint import load_npz, save_npz

    tree = {"a": {"b": np.arange(3, dtype=np.float32), "c": {"d": np.float32(2.0)}}, "e": np.ones((2, 2), np.int32)}
    save_npz(str(tmp_path / "t.npz"), tree)
    back = load_npz(str(tmp_path / "t.npz"))
    assert np.array_equal(back["a"]["b"], tree["a"]["b"]) and back["a"]["c"]["d"] == 2.0 and back["e"].dtype == np.int32


def test_shard_range():
    from sparsernns_b200.dist import shard_range

    for n, w in [(32, 8), (33, 8), (5, 8), (0, 4), (2048, 3)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def _golden():
    from sparsernns_b200.checkpoint import load_npz

    tree = load_npz(os.path.join(GOLD, "ndns_ds0.5_sparse_w8a16.npz"))
    io = np.load(os.path.join(GOLD, "golden_io_ds0.5.npz"))
    return tree, io


def test_oracle_reproduces_golden_vectors():
    from sparsernns_b200 import synth

    tree, io = _golden()
    dims = synth.ndns_dims(0.5)
    var = {"params": tree["params"], "batch_stats": tree["batch_stats"]}
    x = io["x"]
    y = OM.model_forward(x, var, Z.oracle_cfg(dims, relufication=False, static_quant=False, associative_scan=True))
    assert np.abs(y - io["y_float_gelu"]).max() <= 1e-5 * np.abs(io["y_float_gelu"]).max()
    y = OM.model_forward(x, var, Z.oracle_cfg(dims, relufication=True, static_quant=True))
    assert np.abs(y - io["y_static_w8a16"]).max() <= 1e-3 * np.abs(io["y_static_w8a16"]).max()
    md, q = OC.build_fxp_qconfig(tree["params"], tree["cal_stats"], "w8a16")
    m = OFM.FxpRegressionModel(md, q, dims.n_layers)
    taps = {}
    yq = m(m.quantize_input(x), taps)
    assert yq.exp == int(io["y_fxp_exp"]) and np.array_equal(yq.data, io["y_fxp_int"])   # integer path: bit-exact
    for i in range(dims.n_layers):
        assert np.array_equal(taps[f"layers_{i}"]["output"].data, io[f"fxp_layer{i}_output"])
        assert taps[f"layers_{i}"]["output"].exp == int(io["fxp_layer_out_exp"][i])


def test_batch_dependence_of_compute_best():
    """fxp results depend on the batch through the global compute_best exponents."""
    zoo = Z.build_sparse_calibrated(0.5, n_clips=3, n_frames=64)
    md, q = Z.fxp_setup(zoo)
    m = OFM.FxpRegressionModel(md, q, zoo["dims"].n_layers)
    x = zoo["x"].copy()
    x[1] *= 40.0   # a loud batch mate changes the global maxima
    a = m(m.quantize_input(x[:1])).data
    b = m(m.quantize_input(x)).data[:1]
    assert a.shape == b.shape  # both are valid "shard-local" results; they need not be equal


def test_calibration_scale_formula_matches_oracle():
    """sparsernns_b200.calibrate.calculate_qparams (quantization.py:352-370) against the oracle's restatement."""
    from oracle.quant import calculate_qparams as ref
    from sparsernns_b200.calibrate import calculate_qparams

    rng = np.random.default_rng(0)
    for bits in (8, 16):
        for _ in range(500):
            lo, hi = np.sort(rng.standard_normal(2) * 10.0 ** rng.uniform(-8, 4))
            assert calculate_qparams(lo, hi, bits) == ref(lo, hi, bits)
        for v in (0.0, 1e-12, 127.0, 128.0, 32767.0, 2.0 ** -20):        # eps floor and exact powers of two
            assert calculate_qparams(-v, v, bits) == ref(-v, v, bits)


def test_ndns_and_calibrate_fail_loudly_without_cuda():
    import torch
    from sparsernns_b200 import ndns

    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    with pytest.raises(ValueError):
        ndns.stft_splitter(torch.zeros(2, 1024))        # CPU tensors are rejected: there is no CPU path
    with pytest.raises(ValueError):
        ndns.si_snr(torch.zeros(2, 1024), torch.zeros(2, 1024))


def test_product_fxputils_equals_the_reference_qconfig():
    """sparsernns_b200.fxputils against the qconfig the reference's own fxputils.py produced on the golden models
    (tests/golden/reference_fxp_model.npz, see tests/golden/make_reference_golden.py): shared and per-layer exponents."""
    import json
    from sparsernns_b200 import fxputils as PU
    from sparsernns_b200.checkpoint import load_npz

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with np.load(os.path.join(gold, "reference_fxp_model.npz")) as z:
        ref = {k: json.loads(str(z[f"{k}/qconfig_json"])) for k in ("A", "C", "E")}

    def plain(t):      # the informational sign-bit entries (never read by the model) are not carried by the product
        if isinstance(t, dict):
            return {str(k): plain(v) for k, v in t.items() if not (str(k).endswith("signbit") or str(k).endswith("signbits"))}
        if isinstance(t, (np.ndarray, np.generic)):
            return np.asarray(t).tolist()
        return t

    for case, ds, shared in (("A", 0.5, True), ("C", 0.5, False), ("E", 1.0, True)):
        tree = load_npz(os.path.join(gold, f"ndns_ds{ds}_sparse_w8a16.npz"))
        q = PU.create_fxp_qconfig(PU.load_modeldict(tree["params"], tree["cal_stats"]), "w8a16", shared_exp=shared)
        assert plain(q) == plain(ref[case]), case
