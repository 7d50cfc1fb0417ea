import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


def load_golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


CONTRACT_TOL = 1e-5    # the north star's bar: max abs error of full scale (1.0), fp32
REGRESSION_TOL = 1e-6  # measured worst case over the reference goldens is 7.5e-8: a 10x regression fails here


def check_regression_bar(worst, peak, what=""):
    """Second, tighter bar on the worst error of a golden replay (the 1e-5 contract is asserted per
    block by the tests themselves); prints the error relative to the signal peak."""
    rel = worst / peak if peak else 0.0
    print(f"{what}: max abs err {worst:.3g}, peak {peak:.3g}, relative {rel:.3g}")
    assert worst <= REGRESSION_TOL, (what, worst, "regression bar 1e-6 (contract: 1e-5)")

