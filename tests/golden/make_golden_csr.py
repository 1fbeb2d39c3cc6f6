"""Golden fixtures for the CSR format (cs_matrix_struct_csr_t, _mat_vec_p_l_csr,
cs_matrix_spmv.cpp:1890-1955): CSR arrays and products computed by the reference build
(oracle/_ref/libcsref.so) from the edge list of a cs_mesh_cartesian-style cube.

    python tests/golden/make_golden_csr.py        -> tests/golden/csr_N{N}_{sym|nonsym}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402


def coefficients(s, symmetric):
    """_matrix_check-style values (cs_benchmark_matrix.cpp:868-896) on the edge list."""
    n = s.n_rows
    edges = s.face_cells.astype(np.int32)
    g = np.arange(1, n + 1, dtype=np.float64)
    da = 1.0 + np.cos(g * 0.1 + 0.3)
    lo = np.minimum(edges[:, 0], edges[:, 1]) + 1.0
    hi = np.maximum(edges[:, 0], edges[:, 1]) + 1.0
    xa = 0.5 * (0.45 + np.cos(lo * 0.13) + np.cos(hi * 0.07))
    if not symmetric:
        xa = np.stack([xa, -0.7 * xa + 0.1], axis=1).copy()
    x = np.sin(g * 0.37 + 0.2)
    return edges, da, xa, x


def main():
    ref = RefLib(1)
    for N, symmetric in ((6, True), (10, False), (14, True)):
        s = cube_poisson(N, "neumann")
        edges, da, xa, x = coefficients(s, symmetric)
        h, ri, ci, val = ref.system_csr(s.n_rows, edges, da, xa, symmetric)
        out = dict(N=N, row_index=ri, col_id=ci, val=val, x=x,
                   y=ref.spmv(h, x, s.n_rows), y_exdiag=ref.spmv(h, x, s.n_rows, True))
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden",
                            f"csr_N{N}_{'sym' if symmetric else 'nonsym'}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, len(ci), "entries")


if __name__ == "__main__":
    main()
