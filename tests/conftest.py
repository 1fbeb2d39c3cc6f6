import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "cube_N*.npz")))


def golden_cases():
    out = []
    for p in GOLDEN:
        base = os.path.basename(p)[:-4]          # cube_N10_neumann
        _, n, bc = base.split("_")
        out.append(pytest.param(int(n[1:]), bc, p, id=base))
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle.pyoracle import RefLib, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libcsref.so not built (needs /root/reference)")
    return RefLib(1)


@pytest.fixture(scope="session")
def csb():
    """The product library on cuda:0 (gpu tests only)."""
    from code_saturne_b200 import api
    api.init(0)
    return api


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
