"""Golden fixtures for the 3x3 block-diagonal path (config 4): _b_mat_vec_p_l_msr_3
(cs_matrix_spmv.cpp:2394-2455) and _bi_cgstab on a CS_MATRIX_BLOCK_D matrix
(cs_sles_it.cpp:2693-2880), computed by the reference build (oracle/_ref/libcsref.so).

    python tests/golden/make_golden_block3.py      -> tests/golden/block3_N{N}.npz
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion  # noqa: E402
from oracle.pyoracle import RefLib, _dp  # noqa: E402


def ref_spmv(ref, h, x, n, exclude_diag=False):
    xx = np.ascontiguousarray(x, dtype=np.float64).copy()
    y = np.empty(3 * n)
    ref.lib.csref_spmv(h, C.c_int(int(exclude_diag)), _dp(xx), _dp(y))
    return y


def main():
    ref = RefLib(1)
    for N in (6, 10, 14):
        ri, ci, d, xv, rhs = cube_block3_convection_diffusion(N)
        n = len(ri) - 1
        h = ref.system_block(n, 3, ri, ci, d.ravel(), xv, False)
        x = np.cos(0.7 * np.arange(3 * n) + 0.1)
        r_norm = float(np.linalg.norm(rhs))
        out = dict(N=N, x=x, y=ref_spmv(ref, h, x, n), y_exdiag=ref_spmv(ref, h, x, n, True),
                   r_norm=r_norm)
        r = ref.sles_it(h, ref.BICGSTAB, rhs, 1000, 1e-10, r_norm)
        for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
            out["bicg_" + k] = r[k]
        out["bicg_history"] = np.array([ref.sles_it(h, ref.BICGSTAB, rhs, k, 1e-10, r_norm)["residual"]
                                        for k in range(1, r["n_iter"] + 1)])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"block3_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, "n_iter", r["n_iter"], "cvg", r["cvg"])


if __name__ == "__main__":
    main()
