"""Golden fixtures for the pairwise aggregation hierarchy (CS_GRID_COARSENING_SPD_PW,
_pairwise_msr cs_grid.cpp:1744-2020, aggregation limits 2 / 4 / 8 with their recursion
:8305-8346, Galerkin coarse matrices): outputs of the reference's own compiled CPU code
(oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_pairwise.py   -> tests/golden/pairwise_N{N}_{bc}.npz
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
        for lim in (2, 4, 8):
            kw = dict(aggregation_limit=lim, coarsening_type=3, p0p1_relax=0., n_levels_max=10)
            lv = ref.hierarchy(h, **kw)
            out[f"lim{lim}_n_levels"] = len(lv)
            for l in range(1, len(lv)):
                for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                    out[f"lim{lim}_l{l}_{k}"] = lv[l][k]
            r = ref.multigrid(h, s.rhs, 1e-8, r_norm, **kw)
            for k in ("cvg", "n_iter", "residual"):
                out[f"lim{lim}_mg_{k}"] = r[k]
            print(N, bc, lim, [x["n_rows"] for x in lv], "n_iter", r["n_iter"])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"pairwise_N{N}_{bc}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
