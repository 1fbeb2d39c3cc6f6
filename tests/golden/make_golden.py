"""Generate the golden fixtures from the reference's OWN compiled CPU code.

Run in the authoring container, where /root/reference is mounted and
oracle/_ref/libcsref.so has been built (make -f oracle/Makefile.ref -j8):

    python tests/golden/make_golden.py

Writes tests/golden/cube_N{N}_{bc}.npz.  Everything stored here is an output
of the reference (cs_matrix_vector_multiply, cs_dot*, cs_sles_it_solve,
cs_grid_coarsen, cs_multigrid_solve) with OMP_NUM_THREADS=1 on inputs from
code_saturne_b200.mesh_cartesian; the GPU box never needs /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson, matrix_check_coefficients  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

CASES = [(6, "neumann"), (10, "neumann"), (10, "dirichlet"), (16, "neumann")]


def main():
    ref = RefLib(1)
    for N, bc in CASES:
        s = cube_poisson(N, bc)
        out = {"N": N}
        # SpMV on _matrix_check-style coefficients
        d, xv, x = matrix_check_coefficients(s)
        h = ref.system(s, d, xv, with_mesh=False)
        out["spmv_y"] = ref.spmv(h, x, s.n_rows)
        out["spmv_y_exdiag"] = ref.spmv(h, x, s.n_rows, True)
        out["dot_xx"] = ref.dot_xx(x)
        out["dot_xy"] = ref.dot(x, out["spmv_y"])
        out["dot_xx_xy_yz"] = np.array(ref.dot_xx_xy_yz(x, out["spmv_y"], out["spmv_y_exdiag"]))
        ref.destroy(h)

        h = ref.system(s)
        r_norm = float(np.sqrt(ref.dot_xx(s.rhs)))
        out["r_norm"] = r_norm
        for name, solver, nmax in (("pcg", ref.PCG, 1000), ("jacobi", ref.JACOBI, 40),
                                   ("bicgstab", ref.BICGSTAB, 1000)):
            r = ref.sles_it(h, solver, s.rhs, nmax, 1e-8, r_norm)
            out[f"{name}_cvg"] = r["cvg"]
            out[f"{name}_n_iter"] = r["n_iter"]
            out[f"{name}_residual"] = r["residual"]
            out[f"{name}_initial_residual"] = r["initial_residual"]
            out[f"{name}_vx"] = r["vx"]
            # residual history: rerun with n_max_iter = k (O(K^2), tiny cases only)
            hist = []
            for k in range(1, r["n_iter"] + 1):
                rk = ref.sles_it(h, solver, s.rhs, k, 1e-8, r_norm)
                hist.append(rk["residual"])
            out[f"{name}_history"] = np.array(hist)
        # warm start
        r = ref.sles_it(h, ref.PCG, s.rhs, 1000, 1e-10, r_norm, vx0=out["jacobi_vx"])
        out["pcg_warm_n_iter"] = r["n_iter"]
        out["pcg_warm_residual"] = r["residual"]
        out["pcg_warm_vx"] = r["vx"]

        # hierarchy
        lv = ref.hierarchy(h)
        out["mg_n_levels"] = len(lv)
        for l, g in enumerate(lv):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                out[f"mg_l{l}_{k}"] = g[k]
        # solve (solver-mode defaults of cs_multigrid_define)
        r = ref.multigrid(h, s.rhs, 1e-8, r_norm)
        out["mg_cvg"] = r["cvg"]
        out["mg_n_iter"] = r["n_iter"]
        out["mg_residual"] = r["residual"]
        out["mg_vx"] = r["vx"]
        # fine residual at the start of cycle k+1, for k = 2..: n_max_cycles = k
        cyc = []
        k = 2
        while True:
            rk = ref.multigrid(h, s.rhs, 1e-8, r_norm, n_max_cycles=k)
            cyc.append(rk["residual"])
            if rk["cvg"] == 1 or k > 60:
                break
            k += 1
        out["mg_cycle_residuals_from3"] = np.array(cyc)
        ref.destroy(h)

        path = os.path.join(ROOT, "tests", "golden", f"cube_N{N}_{bc}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes; pcg", out["pcg_n_iter"], "mg", out["mg_n_iter"],
              "levels", len(lv))


if __name__ == "__main__":
    main()
