"""Golden fixtures for the Galerkin coarse matrices (p0p1_relax = 0: cs_grid_coarsen ->
_compute_coarse_quantities_msr, cs_grid.cpp:5969-6090, no mesh quantities) and the V-cycle
solve over that hierarchy: outputs of the reference's own compiled CPU code
(oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_galerkin.py   -> tests/golden/galerkin_N{N}_{bc}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402


def main():
    ref = RefLib(1)
    for N, bc in ((8, "neumann"), (12, "dirichlet")):
        s = cube_poisson(N, bc)
        h = ref.system(s)
        r_norm = float(np.sqrt(ref.dot_xx(s.rhs)))
        out = {"N": N, "r_norm": r_norm}
        lv = ref.hierarchy(h, p0p1_relax=0.)
        out["n_levels"] = len(lv)
        for l in range(1, len(lv)):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                out[f"l{l}_{k}"] = lv[l][k]
        r = ref.multigrid(h, s.rhs, 1e-8, r_norm, p0p1_relax=0.)
        for k in ("cvg", "n_iter", "residual", "vx"):
            out[f"mg_{k}"] = r[k]
        r = ref.fcg_mgpc(h, s.rhs, 200, 1e-8, r_norm, p0p1_relax=0.)
        for k in ("cvg", "n_iter", "residual", "vx"):
            out[f"fcg_{k}"] = r[k]
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"galerkin_N{N}_{bc}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "levels", len(lv), "mg n_iter", out["mg_n_iter"])


if __name__ == "__main__":
    main()
