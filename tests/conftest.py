import glob
import os
import sys

import numpy as np
import pytest

# The fixtures and the oracle restate the reference run with ONE OpenMP thread (its greedy
# aggregation is confined to one block of rows per thread, so the maps depend on the count);
# libcsb200 follows the process's OpenMP thread count like the reference does.  Tests that
# exercise several blocks ask for them explicitly (set_aggregation_blocks).
os.environ.setdefault("OMP_NUM_THREADS", "1")

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


def equiv_iters_close(n_iter, n_ref, n_iter_per_cycle):
    """North-star bar "iteration counts within +-2".  cs_multigrid reports *equivalent
    fine-grid iterations* (cs_multigrid.cpp:3341): one V-cycle is worth n_iter_per_cycle of
    them, so the count may differ by 2, or by what one cycle more or less adds (+-2)."""
    d = abs(int(n_iter) - int(n_ref))
    return d <= 2 or d <= int(np.ceil(n_iter_per_cycle)) + 2


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
