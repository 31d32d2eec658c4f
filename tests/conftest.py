import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(params=["cuda_cores", "dense_tiles"])
def fxp_path(request, monkeypatch):
    """Both code paths of the integer model: the CUDA-core kernels (fxp_model.cu) and the tcgen05 kind::i8 dense-tile
    kernels (fxp_tc.cu).  The switches are read when a model handle is created."""
    if request.param == "cuda_cores":
        monkeypatch.setenv("S5B200_FXP_TC", "0")
    else:
        monkeypatch.setenv("S5B200_FXP_TC", "1")
        monkeypatch.setenv("S5B200_FXP_TC_MIN_ROWS", "1")
    return request.param
