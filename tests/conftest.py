import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _bind_emu():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    import dsgcn_b200
    dsgcn_b200._lib._testing_use_library(build_emu.build())


def _bind_cuda():
    import dsgcn_b200
    dsgcn_b200._lib._testing_use_library(None)   # back to the product library (lazy load of libdsgcn_b200.so)
    assert dsgcn_b200._lib.is_device_build()


@pytest.fixture(params=["sim", pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    """'sim': kernel sources on the host-side SIMT simulator (CPU tensors) — index/algebra checks
    in the GPU-less container.  'cuda': the real sm_100a library on the B200 (the parity tests proper)."""
    if request.param == "sim":
        _bind_emu()
        return torch.device("cpu")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _bind_cuda()
    return torch.device("cuda:0")
