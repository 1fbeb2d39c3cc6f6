"""Golden fixtures for the flexible CG (cs_sles_it.cpp:832-985) and inexact-preconditioning
CG (:1016-1203) paths, outputs of the reference's own compiled CPU code
(oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_fcg.py   -> tests/golden/{fcg,ipcg}_N{N}_{bc}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

CASES = [(8, "neumann"), (12, "dirichlet"), (16, "neumann")]


def main():
    ref = RefLib(1)
    for name, solver in (("fcg", ref.FCG), ("ipcg", ref.IPCG)):
        for N, bc in CASES:
            s = cube_poisson(N, bc)
            h = ref.system(s)
            r_norm = float(np.sqrt(ref.dot_xx(s.rhs)))
            out = {"N": N, "r_norm": r_norm}
            r = ref.sles_it(h, solver, s.rhs, 1000, 1e-9, r_norm)
            for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
                out[k] = r[k]
            out["history"] = np.array([ref.sles_it(h, solver, s.rhs, k, 1e-9, r_norm)["residual"]
                                       for k in range(1, r["n_iter"] + 1)])
            w = ref.sles_it(h, solver, s.rhs, 5, 1e-12, r_norm, vx0=0.5 * r["vx"])
            for k in ("cvg", "n_iter", "residual", "vx"):
                out["warm_" + k] = w[k]
            ref.destroy(h)
            path = os.path.join(ROOT, "tests", "golden", f"{name}_N{N}_{bc}.npz")
            np.savez_compressed(path, **out)
            print("wrote", path, "n_iter", r["n_iter"])


if __name__ == "__main__":
    main()
