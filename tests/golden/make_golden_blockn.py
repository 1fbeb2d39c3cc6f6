"""Golden fixtures for block-diagonal (CS_MATRIX_BLOCK_D) matrices with block sizes other
than 3 (and 3 for the smoothers): the SpMV (_b_mat_vec_p_l_msr_6 / _generic,
cs_matrix_spmv.cpp:2478, :2308), block Jacobi (_block_jacobi, cs_sles_it.cpp:2460, with the
block_nn_inverse of cs_sles_it_priv.cpp:178-222), BiCGStab + point Jacobi, and the block forms
of the multigrid smoothers (_block_3_jacobi, _block_jacobi, _block_relaxed_jacobi,
cs_multigrid_smoother.cpp:390-790): outputs of the reference's own
compiled CPU code (oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_blockn.py   -> tests/golden/blockn_db{db}_N{N}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_block_convection_diffusion  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

CASES = [(6, 6), (2, 7), (4, 5), (9, 4), (3, 6)]       # (db, N)


def main():
    ref = RefLib(1)
    for db, N in CASES:
        ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db)
        n = len(ri) - 1
        h = ref.system_block(n, db, ri, ci, d.ravel(), xv, False)
        r_norm = float(np.linalg.norm(rhs))
        x = np.cos(0.3 * np.arange(db * n))
        out = dict(N=N, db=db, r_norm=r_norm, x=x)
        out["y"] = ref.spmv(h, x, db * n)
        out["y_exdiag"] = ref.spmv(h, x, db * n, exclude_diag=True)
        for nit in (1, 2, 4):
            f = ref.sles_it(h, ref.JACOBI, rhs, nit, 1e-12, r_norm)
            out[f"jac{nit}_vx"] = f["vx"]
            out[f"jac{nit}_residual"] = f["residual"]
        w = ref.sles_it(h, ref.JACOBI, rhs, 3, 1e-12, r_norm, vx0=x)
        out["jac_warm_vx"] = w["vx"]
        out["jac_warm_residual"] = w["residual"]
        r = ref.sles_it(h, ref.JACOBI, rhs, 500, 1e-8, r_norm)
        for k in ("cvg", "n_iter", "residual"):
            out[f"jac_{k}"] = r[k]
        b = ref.sles_it(h, ref.BICGSTAB, rhs, 500, 1e-10, r_norm)
        for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
            out[f"bicg_{k}"] = b[k]
        # cs_multigrid_smoother on the block matrix: block Jacobi (3), block relaxed Jacobi (14-16)
        for st in (3, 14, 15, 16):
            for nit in (1, 3):
                out[f"sm{st}_{nit}_zero"] = ref.smoother(h, st, rhs, nit)["vx"]
                w = ref.smoother(h, st, rhs, nit, x)
                out[f"sm{st}_{nit}_warm"] = w["vx"]
                out[f"sm{st}_{nit}_n_iter"] = w["n_iter"]
        print(db, N, "jacobi", r["cvg"], r["n_iter"], "bicgstab", b["cvg"], b["n_iter"])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"blockn_db{db}_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
