import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): compile the CUDA library for sm_100a (nvcc
    # cross-compiles without a GPU), the oracle and the host build of the device rules once
    from gym_chess_b200 import _lib

    if not os.path.exists(_lib.SO_PATH):
        _lib.build()


@pytest.fixture(scope="session")
def golden():
    from tests import parity_helpers as ph

    return ph.load_golden()
