import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)  # golden_checks.py, emu/


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def orc():
    """CPU oracle binding (test infrastructure)."""
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def emu():
    """CPU warp emulator of the device headers (test infrastructure)."""
    from tests.emu import emu_py
    emu_py.build()
    return emu_py


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library. Built in-tree by __graft_entry__.build(); tests never fall back to anything else."""
    import __graft_entry__ as g
    g.build_cuda()
    from mpc_local_planner_b200 import capi
    return capi.load_library()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
