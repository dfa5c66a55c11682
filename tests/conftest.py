import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def syn():
    import fs2_b200
    return fs2_b200.synthetic


@pytest.fixture(scope="session")
def sd32(syn):
    """The seed-0 synthetic state dict every golden fixture was generated with."""
    return syn.synthetic_state_dict(seed=0)


@pytest.fixture(scope="session")
def sd64(sd32):
    from oracle import fs2_oracle
    return fs2_oracle.cast_state_dict(sd32, __import__("torch").float64)
