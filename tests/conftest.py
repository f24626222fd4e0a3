import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle("f64")


@pytest.fixture(scope="session")
def oracle_ld():
    from oracle.oracle_py import Oracle
    return Oracle("ld")


@pytest.fixture(scope="session")
def ms():
    """The product package with its C-ABI library built (nvcc cross-compiles without a GPU)."""
    from mav_trajectory_generation_cmake_b200 import build
    build.build()
    import mav_trajectory_generation_cmake_b200 as pkg
    pkg.load()
    return pkg


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a gpu-marked test ran without a CUDA device")
    torch.cuda.set_device(0)
    return torch
