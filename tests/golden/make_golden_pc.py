"""Golden fixtures for cs_sles_it with general preconditioners: single-reduction PCG
(_conjugate_gradient_sr, cs_sles_it.cpp:1229-1415), FCG (:832-985) and IPCG (:1016-1203) with the polynomial
preconditioner of degree 1 and 2 (cs_sles_pc.cpp:549-610), FCG without preconditioner, and
PCG preconditioned by one V-cycle: outputs of the reference's own compiled CPU code
(oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_pc.py   -> tests/golden/pc_N{N}_{bc}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

CASES = [(8, "neumann"), (12, "dirichlet")]


def main():
    ref = RefLib(1)
    for N, bc in CASES:
        s = cube_poisson(N, bc)
        h = ref.system(s)
        r_norm = float(np.sqrt(ref.dot_xx(s.rhs)))
        out = {"N": N, "r_norm": r_norm}

        def record(tag, solve):
            r = solve(500, 1e-9, None)
            for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
                out[f"{tag}_{k}"] = r[k]
            out[f"{tag}_history"] = np.array([solve(k, 1e-9, None)["residual"]
                                              for k in range(1, min(r["n_iter"], 6) + 1)])
            w = solve(4, 1e-12, 0.3 * r["vx"])
            for k in ("cvg", "n_iter", "residual", "vx"):
                out[f"{tag}_warm_{k}"] = w[k]
            print(N, bc, tag, "n_iter", r["n_iter"])

        for sname, solver in (("pcg", ref.PCG), ("fcg", ref.FCG), ("ipcg", ref.IPCG)):
            for deg in (1, 2) + ((-1,) if sname == "fcg" else ()):
                record(f"{sname}_poly{deg}",
                       lambda nmax, eps, v0, solver=solver, deg=deg:
                       ref.sles_it(h, solver, s.rhs, nmax, eps, r_norm, vx0=v0, poly_degree=deg))
        record("pcg_mgpc", lambda nmax, eps, v0: ref.it_mgpc(h, ref.PCG, s.rhs, nmax, eps, r_norm,
                                                              vx0=v0))
        record("ipcg_mgpc", lambda nmax, eps, v0: ref.it_mgpc(h, ref.IPCG, s.rhs, nmax, eps, r_norm,
                                                               vx0=v0))
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"pc_N{N}_{bc}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
