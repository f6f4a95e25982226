import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture()
def rng():
    import numpy as np

    return np.random.default_rng(0x5EED)


# the streaming decode kernel serves one sequence by default (faster than the per-op path only there); the GPU tests
# exercise its multi-sequence code as well
import os

os.environ.setdefault("MC_STREAM_MAX_ROWS", "8")
