"""Golden fixtures for the K-cycle (_multigrid_k_cycle, cs_multigrid.cpp:4051-4480; pairwise
aggregation limit 4, Galerkin coarse matrices, 2 Jacobi sweeps, threshold 0.25 -- what
cs_multigrid_pc_create(CS_MULTIGRID_K_CYCLE) gives with the device smoother substitution) and
FCG preconditioned by it: outputs of the reference's own compiled CPU code
(oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_kcycle.py   -> tests/golden/kcycle_N{N}_{bc}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

KW = dict(aggregation_limit=4, coarsening_type=3, n_levels_max=10, n_g_rows_min=30,
          p0p1_relax=0., mg_type=1)
SM = dict(types=(3, 3, 0), precision_mult=(-1., -1., 10.))


def main():
    ref = RefLib(1)
    for N, bc in ((8, "dirichlet"), (12, "dirichlet")):
        s = cube_poisson(N, bc)
        h = ref.system(s)
        r_norm = float(np.sqrt(ref.dot_xx(s.rhs)))
        out = {"N": N, "r_norm": r_norm}
        r = ref.multigrid(h, s.rhs, 1e-8, r_norm, n_max_cycles=1, n_max_iter=(2, 2, 500), **SM, **KW)
        out["cycle_vx"], out["cycle_n_iter"] = r["vx"], r["n_iter"]
        r = ref.it_mgpc(h, 1, s.rhs, 200, 1e-8, r_norm, mg_n_max_iter=(2, 2, 500), **SM, **KW)
        for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
            out[f"fcg_{k}"] = r[k]
        print(N, bc, "fcg + k-cycle n_iter", r["n_iter"], "cvg", r["cvg"])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"kcycle_N{N}_{bc}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
