"""GPU parity of the bare tcgen05 dense-tile primitives through the C ABI: integer codes in, exact integer sums out.

These pin the UMMA shared-memory layout, the matrix / instruction descriptors and the TMEM read-back that the model
kernels (tc_model.cu) reuse, independently of any model arithmetic.  Bit-exact against NumPy int64 matmul.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(128, 32, 32), (300, 96, 128), (1000, 256, 96), (129, 128, 96), (500, 96, 256), (64, 96, 96), (257, 288, 16)]


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("rows,K,N", SHAPES)
def test_tc_gemm_f16_exact_integer_sums(cuda_device, rows, K, N):
    import torch
    from sparsernns_b200 import _abi

    lib = _abi.lib()
    rng = np.random.default_rng(rows + K + N)
    A = rng.integers(-128, 256, (rows, K))       # hi bytes are in [-128, 127], lo bytes in [0, 255]
    W = rng.integers(-128, 128, (N, K))
    tA = torch.from_numpy(A.astype(np.float16)).to(cuda_device)
    tW = torch.from_numpy(W.astype(np.float16)).to(cuda_device)
    Y = torch.full((rows, N), float("nan"), device=cuda_device)
    _abi.check(lib.s5b200_tc_gemm_f16(tA.data_ptr(), tW.data_ptr(), Y.data_ptr(), rows, K, N, _stream()), "tc_gemm_f16")
    torch.cuda.synchronize()
    ref = A.astype(np.int64) @ W.astype(np.int64).T
    assert np.abs(ref).max() < 2 ** 24
    assert np.array_equal(Y.cpu().numpy().astype(np.int64), ref)


@pytest.mark.parametrize("a_signed", [1, 0])
@pytest.mark.parametrize("rows,K,N", SHAPES)
def test_tc_gemm_i8_exact(cuda_device, rows, K, N, a_signed):
    import torch
    from sparsernns_b200 import _abi

    lib = _abi.lib()
    rng = np.random.default_rng(rows + K + N + a_signed)
    A = rng.integers(-128, 128, (rows, K)) if a_signed else rng.integers(0, 256, (rows, K))
    W = rng.integers(-128, 128, (N, K))
    tA = torch.from_numpy(A.astype(np.int8 if a_signed else np.uint8)).to(cuda_device)
    tW = torch.from_numpy(W.astype(np.int8)).to(cuda_device)
    Y = torch.full((rows, N), -(2 ** 31), dtype=torch.int32, device=cuda_device)
    _abi.check(lib.s5b200_tc_gemm_i8(tA.data_ptr(), tW.data_ptr(), Y.data_ptr(), rows, K, N, a_signed, _stream()), "tc_gemm_i8")
    torch.cuda.synchronize()
    ref = A.astype(np.int64) @ W.astype(np.int64).T
    assert np.array_equal(Y.cpu().numpy().astype(np.int64), ref)


def test_tc_gemm_rejects_bad_shapes(cuda_device):
    import torch
    from sparsernns_b200 import _abi

    lib = _abi.lib()
    t = torch.zeros(128 * 48, dtype=torch.int32, device=cuda_device)
    assert lib.s5b200_tc_gemm_i8(t.data_ptr(), t.data_ptr(), t.data_ptr(), 128, 48, 32, 1, _stream()) != 0   # K % 32
    assert lib.s5b200_tc_gemm_f16(t.data_ptr(), t.data_ptr(), t.data_ptr(), 128, 32, 24, _stream()) != 0     # N % 16
