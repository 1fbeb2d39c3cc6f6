"""Golden fixtures for the 3x3 block Jacobi solver (_block_3_jacobi, cs_sles_it.cpp:2284-2440)
on the CS_MATRIX_BLOCK_D convection-diffusion systems of block3_N*.npz: outputs of the
reference's own compiled CPU code (oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_block3_jacobi.py   -> tests/golden/block3jac_N{N}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402


def main():
    ref = RefLib(1)
    for N in (6, 10):
        ri, ci, d, xv, rhs = cube_block3_convection_diffusion(N)
        n = len(ri) - 1
        h = ref.system_block(n, 3, ri, ci, d.ravel(), xv, False)
        r_norm = float(np.linalg.norm(rhs))
        out = dict(N=N, r_norm=r_norm)
        r = ref.sles_it(h, ref.JACOBI, rhs, 1000, 1e-8, r_norm)
        for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
            out[k] = r[k]
        for nit in (1, 2, 7):
            f = ref.sles_it(h, ref.JACOBI, rhs, nit, 1e-12, r_norm)
            out[f"sweeps{nit}_vx"] = f["vx"]
            out[f"sweeps{nit}_residual"] = f["residual"]
        x0 = np.cos(0.3 * np.arange(3 * n))
        w = ref.sles_it(h, ref.JACOBI, rhs, 5, 1e-12, r_norm, vx0=x0)
        out["vx0"] = x0
        out["warm_vx"] = w["vx"]
        out["warm_residual"] = w["residual"]
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"block3jac_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "n_iter", r["n_iter"])


if __name__ == "__main__":
    main()
