"""Golden fixtures for the process-local Gauss-Seidel family: the smoothers
(_p_gauss_seidel_msr, _p_sym_gauss_seidel_msr, _ts_f_gauss_seidel_msr, _ts_b_gauss_seidel_msr:
cs_multigrid_smoother.cpp:1053, :1179, :1389, :1519), the cs_sles_it solvers with convergence
test (cs_sles_it.cpp:3940, :4124), and FCG preconditioned by the reference's CPU-default
cs_multigrid_pc (one V-cycle, symmetric Gauss-Seidel x 1 down and up, cs_multigrid.cpp:1622-1636):
outputs of the reference's own compiled CPU code (oracle/_ref/libcsref.so, OMP_NUM_THREADS=1 --
its OpenMP sweeps update vx in place, so only the one-thread run is defined):

    python tests/golden/make_golden_gs.py   -> tests/golden/gs_N{N}_{bc}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

CASES = [(8, "neumann"), (12, "dirichlet")]
SMOOTHERS = {"p_gs": 8, "p_sym_gs": 9, "ts_f_gs": 17, "ts_b_gs": 18}


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
        for name, st in (("p_gs", 8), ("p_sym_gs", 9)):
            for tag, v0 in (("zero", None), ("warm", vx0)):
                r = ref.sles_it(h, st, s.rhs, 400, 1e-6, r_norm, v0)
                for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
                    out[f"it_{name}_{tag}_{k}"] = r[k]
                print(N, bc, name, tag, "n_iter", r["n_iter"], "cvg", r["cvg"])
        types, nmi = (9, 9, 0), (1, 1, 500)
        r = ref.multigrid(h, s.rhs, 1e-8, r_norm, n_max_cycles=1, n_max_iter=nmi, types=types,
                          precision_mult=(-1., -1., 1.))
        out["cyc_sgs_vx"] = r["vx"]
        out["cyc_sgs_n_iter"] = r["n_iter"]
        r = ref.fcg_mgpc(h, s.rhs, 200, 1e-8, r_norm, types, nmi)
        for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
            out[f"fcg_sgs_{k}"] = r[k]
        print(N, bc, "fcg + V-cycle(sym GS) n_iter", r["n_iter"])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"gs_N{N}_{bc}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
