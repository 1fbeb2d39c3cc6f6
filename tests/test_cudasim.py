"""Desk-check of the kernel-per-phase CUDA path in a container without a GPU: the unmodified
csrc/kernels.cu + csrc/sles.cu are compiled against tests/cudasim (host threads play CUDA
threads) and compared with the reference-built golden fixtures and the oracle.  This is
test infrastructure only -- the product library has no CPU path, and the `-m gpu` tests
remain the parity tests proper."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("check", ["spmv_pcg", "smoothers", "cycle_pc", "fcg_mgpc", "general_pc",
                                   "kernel_per_phase", "block3_jacobi", "galerkin", "pairwise", "compute_coeffs", "k_cycle", "native", "assembler", "rhs"])
def test_cudasim(check):
    env = dict(os.environ)
    env.pop("CSB_LIB", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cudasim", "run_checks.py"),
                        check], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"cudasim {check}: ok" in r.stdout
