"""GPU parity: fixed-point ops and the integer model through the C ABI vs the oracle - BIT-EXACT."""
import ctypes as C

import numpy as np
import pytest

import modelzoo as Z
from oracle import fxp as OF
from oracle import fxpmodel as OFM
from oracle.fxp import FxpArray, RoundingMode

pytestmark = pytest.mark.gpu

RM = {RoundingMode.FLOOR: 0, RoundingMode.CEIL: 1, RoundingMode.ROUND: 2}


def _lib():
    from sparsernns_b200 import _abi
    return _abi, _abi.lib()


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_from_fp_and_change_cfg(cuda_device):
    import torch
    abi, lib = _lib()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(5000) * 3, [0.5, 1.5, 2.5, -0.5, -1.5, 1e9, -1e9, 0.0]]).astype(np.float32)
    for bits, exp in [(16, 12), (8, 0), (16, 0), (8, 5)]:
        for rm in (RoundingMode.FLOOR, RoundingMode.ROUND, RoundingMode.CEIL):
            ref = OF.fxp_from_fp(x, bits=bits, exp=exp, round_mode=rm).data
            y = torch.empty(len(x), dtype=torch.int32, device=cuda_device)
            abi.check(lib.s5b200_fxp_from_fp(_dev(x, cuda_device).data_ptr(), y.data_ptr(), len(x), bits, exp, RM[rm], _stream()))
            assert np.array_equal(y.cpu().numpy(), ref), (bits, exp, rm)
    v = rng.integers(-40000, 40000, 6000).astype(np.int32)
    for (b, e, nb, ne) in [(16, 8, 32, 10), (16, 8, 8, 6), (16, 11, 16, 11), (16, 15, 16, 11), (16, 3, 24, 9), (24, 9, 16, 9)]:
        for rm in (RoundingMode.FLOOR, RoundingMode.ROUND, RoundingMode.CEIL):
            if rm == RoundingMode.ROUND and e - ne < 1 and e > ne:
                continue
            ref = OF.fxp_change_cfg(FxpArray(v, b, e), nb, ne, True, rm).data
            y = torch.empty(len(v), dtype=torch.int32, device=cuda_device)
            abi.check(lib.s5b200_fxp_change_cfg(_dev(v, cuda_device).data_ptr(), y.data_ptr(), len(v), b, e, nb, ne, RM[rm], _stream()))
            assert np.array_equal(y.cpu().numpy(), ref), (b, e, nb, ne, rm)


def test_add_mul_explicit_and_compute_best(cuda_device):
    import torch
    abi, lib = _lib()
    rng = np.random.default_rng(1)
    H = 96
    a = rng.integers(-32768, 32767, (7, 50, H)).astype(np.int32)
    b = rng.integers(-32768, 32767, H).astype(np.int32)
    da, db = _dev(a, cuda_device), _dev(b, cuda_device)
    y = torch.empty_like(da)
    for (ae, be, rb, re_) in [(12, 15, 16, 11), (15, 9, 16, 15), (8, 8, 16, 8), (4, 13, 12, 9)]:
        ref = OF.fxp_add(FxpArray(a, 16, ae), FxpArray(b, 16, be), result_bits=rb, result_exp=re_).data
        abi.check(lib.s5b200_fxp_add(da.data_ptr(), db.data_ptr(), y.data_ptr(), a.size, H, 16, ae, 16, be, rb, re_, 0, _stream()))
        assert np.array_equal(y.cpu().numpy(), ref)
    for (ae, be, rb, re_) in [(12, 15, 16, 11), (15, 9, 16, 15), (8, 8, 16, 16), (4, 6, 12, 0)]:
        for rm in (RoundingMode.FLOOR, RoundingMode.ROUND):
            if rm == RoundingMode.ROUND and ae + be - re_ < 1:
                continue
            ref = OF.fxp_mul(FxpArray(a, 16, ae), FxpArray(b, 16, be), result_exp=re_, result_bits=rb, round_mode=rm).data
            abi.check(lib.s5b200_fxp_mul(da.data_ptr(), db.data_ptr(), y.data_ptr(), a.size, H, ae, be, rb, re_, RM[rm], _stream()))
            assert np.array_equal(y.cpu().numpy(), ref)
    e = C.c_int(0)
    for scale_a, (ae, be) in [(1, (12, 15)), (8, (15, 6)), (64, (14, 14)), (1, (3, 15)), (4096, (15, 15))]:
        a2 = (a // scale_a).astype(np.int32)
        da2 = _dev(a2, cuda_device)
        ref = OF.fxp_add(FxpArray(a2, 16, ae), FxpArray(b, 16, be), result_exp="compute_best")
        abi.check(lib.s5b200_fxp_add_compute_best(da2.data_ptr(), db.data_ptr(), y.data_ptr(), a2.size, H, 16, ae, 16, be, 16, C.byref(e), _stream()))
        assert e.value == ref.exp and np.array_equal(y.cpu().numpy(), ref.data), (scale_a, ae, be)
        ref = OF.fxp_mul(FxpArray(a2, 16, ae), FxpArray(b, 16, be), result_exp="compute_best")
        abi.check(lib.s5b200_fxp_mul_compute_best(da2.data_ptr(), db.data_ptr(), y.data_ptr(), a2.size, H, ae, be, 16, C.byref(e), _stream()))
        assert e.value == ref.exp and np.array_equal(y.cpu().numpy(), ref.data), (scale_a, ae, be)
    # the reference raises ValueError when the product would need a left shift (fxparray.py:619-621)
    with pytest.raises(abi.S5Error):
        abi.check(lib.s5b200_fxp_mul(da.data_ptr(), db.data_ptr(), y.data_ptr(), a.size, H, 3, 3, 16, 10, 0, _stream()))


def test_relu_sigmoid_matmul(cuda_device):
    import torch
    abi, lib = _lib()
    rng = np.random.default_rng(2)
    v = np.concatenate([rng.integers(-32768, 32767, 9000), np.arange(-600, 600)]).astype(np.int32)
    dv = _dev(v, cuda_device)
    y = torch.empty_like(dv)
    abi.check(lib.s5b200_fxp_relu(dv.data_ptr(), y.data_ptr(), len(v), _stream()))
    assert np.array_equal(y.cpu().numpy(), np.maximum(v, 0))
    for (e, out_bits) in [(13, 16), (6, 16), (4, 16), (9, 8)]:
        sx, sy = min(e, 6), out_bits - 2
        s = OFM.FxpSigmoid(x_exp=sx, y_exp=sy)
        ref = s.apply(FxpArray(v, out_bits, e)).data
        lut = np.ascontiguousarray(s.lut, dtype=np.int32)
        abi.check(lib.s5b200_fxp_sigmoid(dv.data_ptr(), y.data_ptr(), len(v), out_bits, e, sx, sy, lut.ctypes.data, _stream()))
        assert np.array_equal(y.cpu().numpy(), ref), (e, out_bits)
    # fxp_matmul incl. int32 wrap, the three rounding modes and saturation; ragged row counts
    for (rows, K, N, dens, wmax, xmax, rs, bits) in [(1, 5, 3, 1.0, 127, 32767, 7, 16), (70, 96, 128, 0.15, 127, 32767, 11, 16),
                                                     (257, 257, 96, 0.1, 127, 32767, 0, 16), (40, 64, 64, 0.5, 2 ** 20, 2 ** 20, 3, 32),
                                                     (129, 128, 96, 0.0, 127, 1000, 2, 8)]:
        W = (rng.integers(-wmax, wmax, (K, N)) * (rng.uniform(size=(K, N)) < dens)).astype(np.int32)
        X = rng.integers(-xmax, xmax, (rows, K)).astype(np.int32)
        h = C.c_void_p()
        abi.check(lib.s5b200_sparse_create(W.ctypes.data, K, N, abi.I32, 0, C.byref(h)))
        dX = _dev(X, cuda_device)
        dY = torch.empty(rows, N, dtype=torch.int32, device=cuda_device)
        for rm in (RoundingMode.FLOOR, RoundingMode.ROUND, RoundingMode.CEIL):
            if rm == RoundingMode.ROUND and rs < 1:
                continue
            ref = OF.fxp_matmul(FxpArray(X, 16, rs), FxpArray(W, 8, 0), result_bits=bits, result_exp=0, round_mode=rm).data
            abi.check(lib.s5b200_spmm_i32(h, dX.data_ptr(), dY.data_ptr(), rows, rs, RM[rm], bits, 0, _stream()))
            assert np.array_equal(dY.cpu().numpy(), ref), (rows, K, N, rm)
        lib.s5b200_sparse_destroy(h)


@pytest.fixture(scope="module", params=[True, False], ids=["bn_affine", "simple_bn"])
def fxp_case(request):
    zoo = Z.build_sparse_calibrated(0.5, n_clips=3, n_frames=200, bn_affine=request.param)
    md, q = Z.fxp_setup(zoo)
    return zoo, md, q


def test_fxp_model_bit_exact(cuda_device, fxp_case, fxp_path):
    import torch
    from sparsernns_b200 import fxpmodel as PF

    zoo, md, q = fxp_case
    n_layers = zoo["dims"].n_layers
    om = OFM.FxpRegressionModel(md, q, n_layers)
    pm = PF.FxpRegressionModel(md, q, n_layers)
    x = zoo["x"]
    xq_ref = om.quantize_input(x)
    xq = pm.quantize_input(torch.from_numpy(x).to(cuda_device))
    assert np.array_equal(xq.data.cpu().numpy(), xq_ref.data)
    taps = {}
    y_ref = om(xq_ref, taps)
    # every sown intermediate that crosses a kernel boundary, block by block
    for i in range(n_layers):
        T = taps[f"layers_{i}"]
        t_in = pm.tap(xq, i, PF.TAP_LAYER_INPUT)
        ref_in = taps["encoder_output_relu"] if i == 0 else taps[f"layers_{i-1}"]["output"]
        assert t_in.exp == ref_in.exp and np.array_equal(t_in.data.cpu().numpy(), ref_in.data), f"layer {i} input"
        t_u = pm.tap(xq, i, PF.TAP_PRE_S5)
        assert np.array_equal(t_u.data.cpu().numpy(), T["mixer"]["u"].data), f"layer {i} pre_s5"
        t_xs = pm.tap(xq, i, PF.TAP_XS).data.cpu().numpy()
        assert np.array_equal(t_xs[..., 0], T["pre_C"].real.data) and np.array_equal(t_xs[..., 1], T["pre_C"].imag.data), f"layer {i} xs"
        t_g = pm.tap(xq, i, PF.TAP_POST_GLU)
        assert np.array_equal(t_g.data.cpu().numpy(), T["post_GLU"].data), f"layer {i} post_GLU"
        t_o = pm.tap(xq, i, PF.TAP_OUTPUT)
        assert t_o.exp == T["output"].exp and np.array_equal(t_o.data.cpu().numpy(), T["output"].data), f"layer {i} output"
    y = pm(xq)
    assert y.exp == y_ref.exp and y.bits == y_ref.bits
    assert np.array_equal(y.data.cpu().numpy(), y_ref.data)
    # float entry point: quantise + forward + to_float in one call
    y2, mask = pm.forward_float(torch.from_numpy(x).to(cuda_device))
    assert np.array_equal(y2.data.cpu().numpy(), y_ref.data)
    assert np.array_equal(mask.cpu().numpy(), y_ref.to_float())


def test_fxp_model_batch_of_one_and_dense_tiles(cuda_device, fxp_case, fxp_path):
    """bsz = 1 (the reference's verification mode, fxprun.py:460-463) and the dense-tile weight format."""
    import torch
    from sparsernns_b200 import fxpmodel as PF

    zoo, md, q = fxp_case
    n_layers = zoo["dims"].n_layers
    om = OFM.FxpRegressionModel(md, q, n_layers)
    x1 = zoo["x"][:1, :77]
    ref = om(om.quantize_input(x1))
    for use_sparse in (True, False):
        pm = PF.FxpRegressionModel(md, q, n_layers, use_sparse=use_sparse)
        y = pm(pm.quantize_input(torch.from_numpy(x1).to(cuda_device)))
        assert np.array_equal(y.data.cpu().numpy(), ref.data)


def test_fxp_states_beyond_16_bits(cuda_device, fxp_case, fxp_path):
    """The integer recurrence does not saturate (fxpmodel.py:147-208): a loud input drives the states past 16 bits.  The
    dense-tile GLU kernel splits states into two byte planes only, so it must hand such tiles to the CUDA-core kernel
    (bad-tile list, fxp_tc.cu); the result stays bit-exact, tile by tile, including clips that stay small."""
    import torch
    from sparsernns_b200 import fxpmodel as PF

    import copy

    zoo, md, q = fxp_case
    n_layers = zoo["dims"].n_layers
    q = copy.deepcopy(q)
    for name in ("x_re", "x_im"):          # three more fractional bits for the states: 8 x larger payloads
        q["blocks"]["ssm"]["activations"][name]["exp"] += 3
    om = OFM.FxpRegressionModel(md, q, n_layers)
    pm = PF.FxpRegressionModel(md, q, n_layers)
    x = zoo["x"].copy()
    x[0] *= 400.0          # clip 0 loud, the others as they were: mixed tiles
    x[1, 100:] *= 400.0    # clip 1 turns loud half way
    xq_ref = om.quantize_input(x)
    taps = {}
    y_ref = om(xq_ref, taps)
    per_clip = [[int(np.abs(taps[f"layers_{i}"]["pre_C"].real.data[c]).max()) for c in range(x.shape[0])] for i in range(n_layers)]
    assert max(map(max, per_clip)) >= 1 << 15, f"test input does not drive the states past 16 bits ({per_clip})"
    assert min(map(min, per_clip)) < 1 << 15, f"every clip is loud in every block: no dense-tile coverage ({per_clip})"
    xq = pm.quantize_input(torch.from_numpy(x).to(cuda_device))
    for i in range(n_layers):
        t_xs = pm.tap(xq, i, PF.TAP_XS).data.cpu().numpy()
        assert np.array_equal(t_xs[..., 0], taps[f"layers_{i}"]["pre_C"].real.data), f"layer {i} xs"
        t_g = pm.tap(xq, i, PF.TAP_POST_GLU)
        assert np.array_equal(t_g.data.cpu().numpy(), taps[f"layers_{i}"]["post_GLU"].data), f"layer {i} post_GLU"
    y = pm(xq)
    assert y.exp == y_ref.exp and np.array_equal(y.data.cpu().numpy(), y_ref.data)


def test_fxp_export_matches_oracle_and_writes_the_reference_files(cuda_device, fxp_case, tmp_path):
    """FxpRegressionModel.export / export_files (fxpmodel.py:1441-1458, fxprun.py:476-495): the reference's nesting, integer
    payloads equal to the oracle model's parameters, intermediates equal to the oracle's taps, five pickles."""
    import pickle
    import torch
    from sparsernns_b200 import fxpmodel as PF

    zoo, md, q = fxp_case
    n_layers = zoo["dims"].n_layers
    om = OFM.FxpRegressionModel(md, q, n_layers)
    pm = PF.FxpRegressionModel(md, q, n_layers)
    xq = pm.quantize_input(torch.from_numpy(zoo["x"]).to(cuda_device))
    data = pm.export(xq)
    assert set(data) == {"params", "qconfig", "intermediates"}
    assert set(data["params"]) == {"encoder", "decoder"} and set(data["params"]["encoder"]) == {"encoder"} | {f"layers_{i}" for i in range(n_layers)}
    assert np.array_equal(data["params"]["decoder"]["weight"], om.decoder.weight.data)
    assert np.array_equal(data["params"]["encoder"]["encoder"]["bias"], om.encoder.bias.data)
    taps = {}
    y_ref = om(om.quantize_input(zoo["x"]), taps)
    for i in range(n_layers):
        lp, lq = data["params"]["encoder"][f"layers_{i}"], data["qconfig"]["encoder"][f"layers_{i}"]
        ol = om.layers[i]
        assert np.array_equal(lp["mixer"]["A_real"], ol.mixer.A.real.data) and np.array_equal(lp["mixer"]["C_imag"], ol.mixer.C.imag.data)
        assert np.array_equal(lp["mixer"]["D"], ol.mixer.D.data) and lq["mixer"]["A_real_exp"] == ol.mixer.A.real.exp
        assert np.array_equal(lp["norm"]["mean"], -ol.norm.minus_mean.data)
        assert np.array_equal(lp["out2"]["weight"], ol.out2.weight.data)
        assert set(lq) == {"mixer", "out2", "norm", "multgate", "sigmoid"}
        it = data["intermediates"]["encoder"][f"layers_{i}"]
        assert np.array_equal(it["output"]["data"], taps[f"layers_{i}"]["output"].data) and it["output"]["exp"] == taps[f"layers_{i}"]["output"].exp
        assert np.array_equal(it["post_GLU"]["data"], taps[f"layers_{i}"]["post_GLU"].data)
    assert np.array_equal(data["intermediates"]["decoder"]["__call__"]["data"], y_ref.data)
    pm.export_files(str(tmp_path), xq)
    names = {"fxpmodel_modeldict.pkl", "fxpmodel_fxp_qconfig.pkl", "fxpmodel_io.pkl", "fxpmodel.pkl", "fxpmodel_activations.pkl"}
    assert names == {p.name for p in tmp_path.iterdir()}
    io = pickle.load(open(tmp_path / "fxpmodel_io.pkl", "rb"))
    assert np.array_equal(io["output"]["data"], y_ref.data) and np.array_equal(io["input"]["data"], xq.data.cpu().numpy())
    assert set(pickle.load(open(tmp_path / "fxpmodel.pkl", "rb"))) == {"params", "qconfig"}


def _global_batch_worker(rank, world, port):
    """One rank of test_fxp_global_batch_parity_two_gpus (module level: torch.multiprocessing.spawn pickles it)."""
    import os
    import torch
    import torch.distributed as dist
    from sparsernns_b200 import fxpmodel as PF
    from sparsernns_b200.dist import shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        zoo = Z.build_sparse_calibrated(0.5, n_clips=4, n_frames=160)
        md, q = Z.fxp_setup(zoo)
        n_layers = zoo["dims"].n_layers
        # clips of very different loudness: the shards' own maxima differ, so shard-local exponents differ from the union's
        xn = zoo["x"].copy()
        xn[2:] *= 0.05
        x = torch.from_numpy(xn).cuda()
        # both code paths under sharding: the dense-tile kernels (column extremes travel in the all-reduce) and the CUDA-core ones
        for min_rows in ("1", "1000000000"):
            os.environ["S5B200_FXP_TC_MIN_ROWS"] = min_rows
            one = PF.FxpRegressionModel(md, q, n_layers)                         # ONE process on the union batch
            y_full = one(one.quantize_input(x))
            lo, hi = shard_range(4, rank, world)
            glob = PF.FxpRegressionModel(md, q, n_layers, global_batch=True)     # sharded, maxima all-reduced over NCCL
            y_g = glob(glob.quantize_input(x[lo:hi]))
            assert y_g.exp == y_full.exp and torch.equal(y_g.data, y_full.data[lo:hi]), f"rank {rank}: global-batch parity"
            for i in range(n_layers):
                t_g, t_f = glob.tap(x_q := glob.quantize_input(x[lo:hi]), i, PF.TAP_OUTPUT), one.tap(one.quantize_input(x), i, PF.TAP_OUTPUT)
                assert t_g.exp == t_f.exp and torch.equal(t_g.data, t_f.data[lo:hi]), f"rank {rank}: block {i} output"
            local = PF.FxpRegressionModel(md, q, n_layers)                       # shard-local parity (the default)
            y_l = local(local.quantize_input(x[lo:hi]))
            differs = torch.tensor([int(not torch.equal(y_l.data, y_full.data[lo:hi]))], device="cuda")
            dist.all_reduce(differs)
            assert int(differs.item()) >= 1, "the test batch should make shard-local and global-batch results differ somewhere"
    finally:
        dist.destroy_process_group()


def test_fxp_global_batch_parity_two_gpus(cuda_device):
    """SURVEY 8e option B: clips sharded over 2 GPUs with the maxima all-reduced (NCCL) give, bit for bit, what one process gives
    on the union batch - output and every block output; the default shard-local mode does not (and is not meant to)."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    mp.spawn(_global_batch_worker, args=(2, 29641), nprocs=2, join=True)


@pytest.mark.gpu
def test_fxp_bn_maxima_paths_agree(cuda_device, fxp_case, monkeypatch):
    """BatchNorm `compute_best` maxima three ways: derived from column extremes (fxp_bn_derive_kernel, default with the
    dense-tile kernels), the derive kernel's whole-tensor fallback, and the three reduce passes - same integers and exponents."""
    import torch
    from sparsernns_b200 import fxpmodel as PF

    zoo, md, q = fxp_case
    n_layers = zoo["dims"].n_layers
    om = OFM.FxpRegressionModel(md, q, n_layers)
    taps = {}
    y_ref = om(om.quantize_input(zoo["x"]), taps)
    monkeypatch.setenv("S5B200_FXP_TC_MIN_ROWS", "1")
    for env in ({}, {"S5B200_FXP_BN_FORCE_SLOW": "1"}, {"S5B200_FXP_BN_DERIVE": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        pm = PF.FxpRegressionModel(md, q, n_layers)
        xq = pm.quantize_input(torch.from_numpy(zoo["x"]).to(cuda_device))
        for i in range(n_layers):
            t_u = pm.tap(xq, i, PF.TAP_PRE_S5)
            assert np.array_equal(t_u.data.cpu().numpy(), taps[f"layers_{i}"]["mixer"]["u"].data), (env, i)
            t_o = pm.tap(xq, i, PF.TAP_OUTPUT)
            assert t_o.exp == taps[f"layers_{i}"]["output"].exp and np.array_equal(t_o.data.cpu().numpy(), taps[f"layers_{i}"]["output"].data), (env, i)
        y = pm(xq)
        assert y.exp == y_ref.exp and np.array_equal(y.data.cpu().numpy(), y_ref.data), env
        for k in env:
            monkeypatch.delenv(k)


@pytest.mark.gpu
@pytest.mark.parametrize("B,L", [(1, 333), (5, 523), (2, 1030), (7, 129)])
def test_fxp_dense_tile_path_equals_cuda_core_path_on_odd_shapes(cuda_device, fxp_case, monkeypatch, B, L):
    """Row counts that are no multiple of 4, 32, 64 or 128 and do not split evenly over the CTAs: the dense-tile kernels
    (balanced row ranges, 32-row TMA slots, half-tile bulk stores, tensor-map boxes that are clipped at the end of a clip)
    give, bit for bit, what the CUDA-core kernels give - payloads, float image, and every block output."""
    import torch
    from sparsernns_b200 import fxpmodel as PF

    zoo, md, q = fxp_case
    n_layers = zoo["dims"].n_layers
    rng = np.random.default_rng(B * 1000 + L)
    x = np.abs(rng.standard_normal((B, L, zoo["x"].shape[-1]))).astype(np.float32) * np.float32(zoo["x"].std() * 3)
    xt = torch.from_numpy(x).to(cuda_device)
    outs = {}
    for name, env in (("cuda_cores", {"S5B200_FXP_TC": "0"}), ("dense_tiles", {"S5B200_FXP_TC": "1", "S5B200_FXP_TC_MIN_ROWS": "1"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        pm = PF.FxpRegressionModel(md, q, n_layers)
        y, mask = pm.forward_float(xt)
        xq = pm.quantize_input(xt)
        outs[name] = (y.data.clone(), y.exp, mask.clone(), [pm.tap(xq, i, PF.TAP_OUTPUT) for i in range(n_layers)])
    a, b = outs["cuda_cores"], outs["dense_tiles"]
    assert a[1] == b[1] and torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    for ta, tb in zip(a[3], b[3]):
        assert ta.exp == tb.exp and torch.equal(ta.data, tb.data)
