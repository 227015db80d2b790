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
    """The CPU oracle (test infrastructure): numpy front end of oracle/pg_oracle.c."""
    from oracle import pg_oracle
    pg_oracle.build()
    return pg_oracle


@pytest.fixture(scope="session")
def ops():
    """The product API on the GPU; fails loudly when the CUDA library is missing."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from d3net_b200 import _native, pointgroup_ops
    _native.lib()          # raises if libpg_b200.so was not built
    return pointgroup_ops
