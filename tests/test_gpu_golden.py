"""GPU parity against the committed golden vectors (tests/golden/, produced by the oracle)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gold():
    from sparsernns_b200.checkpoint import load_npz

    return load_npz(os.path.join(GOLD, "ndns_ds0.5_sparse_w8a16.npz")), np.load(os.path.join(GOLD, "golden_io_ds0.5.npz"))


def test_float_and_static_vs_golden(cuda_device, gold):
    import torch
    from sparsernns_b200 import synth
    from sparsernns_b200.model import BoundModel, S5Config

    tree, io = gold
    d = synth.ndns_dims(0.5)
    var = {"params": tree["params"], "batch_stats": tree["batch_stats"]}
    x = torch.from_numpy(io["x"]).to(cuda_device)
    for relu, key, quant, tol in [(False, "y_float_gelu", "none", 1e-5), (True, "y_float_relu", "none", 1e-5),
                                  (True, "y_static_w8a16", "w8a16", 1e-2)]:
        m = BoundModel(S5Config(n_layers=d.n_layers, d_model=d.d_model, ssm_size=d.ssm_size, relufication=relu,
                                quantization=quant), var)
        y = m(x).cpu().numpy()
        err = np.abs(y - io[key]).max() / np.abs(io[key]).max()
        assert err <= tol, (key, err)


def test_fxp_vs_golden_bit_exact(cuda_device, gold, fxp_path):
    import torch
    from sparsernns_b200 import fxpmodel as PF
    from sparsernns_b200 import fxputils, synth

    tree, io = gold
    d = synth.ndns_dims(0.5)
    md = fxputils.load_modeldict(tree["params"], tree["cal_stats"])
    q = fxputils.create_fxp_qconfig(md, "w8a16")
    m = PF.FxpRegressionModel(md, q, d.n_layers)
    xq = m.quantize_input(torch.from_numpy(io["x"]).to(cuda_device))
    y = m(xq)
    assert y.exp == int(io["y_fxp_exp"]) and np.array_equal(y.data.cpu().numpy(), io["y_fxp_int"])
    for i in range(d.n_layers):
        t = m.tap(xq, i, PF.TAP_OUTPUT)
        assert t.exp == int(io["fxp_layer_out_exp"][i]) and np.array_equal(t.data.cpu().numpy(), io[f"fxp_layer{i}_output"])


def test_full_size_properties_dim_scale_1(cuda_device):
    """BASELINE config 4 shapes (H 192, P 128) at full sequence length: size-independent properties."""
    import torch
    from sparsernns_b200 import fxpmodel as PF
    from sparsernns_b200 import fxputils, synth
    from sparsernns_b200.checkpoint import load_npz
    from sparsernns_b200.model import BoundModel, S5Config

    tree = load_npz(os.path.join(GOLD, "ndns_ds1.0_sparse_w8a16.npz"))
    d = synth.ndns_dims(1.0)
    var = {"params": tree["params"], "batch_stats": tree["batch_stats"]}
    m = BoundModel(S5Config(n_layers=d.n_layers, d_model=d.d_model, ssm_size=d.ssm_size, relufication=True,
                            quantization="w8a16"), var)
    x = torch.from_numpy(synth.synthetic_stft_batch(6, 3751, seed=5)).to(cuda_device)
    y = m(x)
    assert torch.isfinite(y).all()
    # clips are independent: any sub-batch gives the same rows (float path has no cross-clip coupling)
    assert torch.equal(m(x[2:5]), y[2:5])
    # causality: changing the tail of a clip leaves the head untouched
    x2 = x.clone()
    x2[:, 3000:] += 0.01
    assert torch.equal(m(x2)[:, :3000], y[:, :3000])
    # deterministic
    assert torch.equal(m(x), y)
    # fixed point: deterministic, saturated to the decoder width, shard-local
    md = fxputils.load_modeldict(tree["params"], tree["cal_stats"])
    q = fxputils.create_fxp_qconfig(md, "w8a16")
    fm = PF.FxpRegressionModel(md, q, d.n_layers)
    yq, mask = fm.forward_float(x)
    assert torch.equal(fm.forward_float(x)[0].data, yq.data)
    assert int(yq.data.abs().max()) <= 32768
    assert torch.equal(mask, yq.data.float() / float(1 << yq.exp))
