import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def digests():
    with open(os.path.join(ROOT, "tests", "golden", "digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixtures():
    return np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import mx_oracle
    mx_oracle.lib()
    return mx_oracle


@pytest.fixture(autouse=True)
def _reset_quantization_env():
    """tests flip env.MX_EXACT_QUANTIZATION the way the reference's do (tests/conftest.py:66-69)."""
    from torchmx_b200 import env_variables as env
    old = env.MX_EXACT_QUANTIZATION
    yield
    env.MX_EXACT_QUANTIZATION = old
