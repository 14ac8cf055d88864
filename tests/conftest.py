import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_frontend():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "frontend.npz"))


def load_golden_head(model_type):
    import numpy as np
    return np.load(os.path.join(GOLDEN, f"head_{model_type}.npz"))
