"""Golden fixtures for the multigrid smoothers (cs_multigrid_smoother.cpp:143-373, :804-905),
the single-cycle (preconditioner) V-cycle (_multigrid_v_cycle_pc, cs_multigrid.cpp:3569-4000)
and FCG preconditioned by it (cs_multigrid_pc_create + cs_sles_it FCG): outputs of the
reference's own compiled CPU code (oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_mgpc.py   -> tests/golden/mgpc_N{N}_{bc}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

CASES = [(8, "neumann"), (12, "dirichlet")]
SMOOTHERS = {"jacobi": 3, "l1_jacobi": 13, "r_jacobi": 14, "rj2": 15, "rj3": 16}
# (descent, ascent, coarse) types and sweep counts of the V-cycle variants
CYCLES = {"rj3": ((16, 16, 0), (3, 3, 500)), "rj2": ((15, 15, 0), (2, 4, 500)),
          "r_jacobi": ((14, 14, 0), (2, 2, 500)), "l1_jacobi": ((13, 13, 0), (2, 2, 500)),
          "jacobi": ((3, 3, 0), (2, 3, 500))}


def main():
    ref = RefLib(1)
    for N, bc in CASES:
        s = cube_poisson(N, bc)
        h = ref.system(s)
        r_norm = float(np.sqrt(ref.dot_xx(s.rhs)))
        vx0 = np.cos(0.37 * np.arange(s.n_rows))
        out = {"N": N, "r_norm": r_norm, "vx0": vx0}
        for name, st in SMOOTHERS.items():
            for nit in (1, 3):
                out[f"sm_{name}_{nit}_zero"] = ref.smoother(h, st, s.rhs, nit)["vx"]
                w = ref.smoother(h, st, s.rhs, nit, vx0)
                out[f"sm_{name}_{nit}_warm"] = w["vx"]
                out[f"sm_{name}_{nit}_n_iter"] = w["n_iter"]
        for name, (types, nmi) in CYCLES.items():
            r = ref.multigrid(h, s.rhs, 1e-8, r_norm, n_max_cycles=1, n_max_iter=nmi,
                              types=types, precision_mult=(-1., -1., 1.))
            out[f"cyc_{name}_vx"] = r["vx"]
            out[f"cyc_{name}_n_iter"] = r["n_iter"]
            # the same smoothers in the V-cycle's solver mode (cycles to convergence)
            r = ref.multigrid(h, s.rhs, 1e-8, r_norm, n_max_iter=nmi, types=types,
                              precision_mult=(-1., -1., 1.))
            for k in ("cvg", "n_iter", "residual", "vx"):
                out[f"vsm_{name}_{k}"] = r[k]
            r = ref.fcg_mgpc(h, s.rhs, 200, 1e-8, r_norm, types, nmi)
            for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
                out[f"fcg_{name}_{k}"] = r[k]
            out[f"fcg_{name}_history"] = np.array(
                [ref.fcg_mgpc(h, s.rhs, k, 1e-8, r_norm, types, nmi)["residual"]
                 for k in range(1, r["n_iter"] + 1)])
            print(N, bc, name, "fcg+mg n_iter", r["n_iter"])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"mgpc_N{N}_{bc}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
