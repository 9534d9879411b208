import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "speigen_cases.npz"))


def colwise_inner(A, B):
    """|<a_c, b_c>| per column (sign/phase invariant)."""
    return np.abs(np.sum(np.conj(A) * B, axis=0))
