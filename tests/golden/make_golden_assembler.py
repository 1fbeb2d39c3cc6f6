"""Golden fixtures for the matrix assembler (cs_matrix_assembler_create / _add_g_ids /
_compute, cs_matrix_create_from_assembler, cs_matrix_assembler_values_init / _add_g /
_finalize) on one rank with an MSR matrix: outputs of the reference's own compiled CPU code
(oracle/_ref/libcsref.so, OMP_NUM_THREADS=1):

    python tests/golden/make_golden_assembler.py   -> tests/golden/assembler_N{N}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402

CASES = {6: (0, 1), 10: (1000, 2), 14: (5, 3)}      # N: (first global row id, seed)


def assembler_input(N, l0, seed):
    """(n_rows, rows, cols, vals, batches): the 4 coefficients of every cube face -- (i,i),
    (i,j), (j,i), (j,j), so diagonals collect up to 6 terms -- plus seeded long-range couplings
    that no mesh face carries, some declared twice; `batches` adds the whole sequence in uneven
    pieces and a second, partial pass on top (coefficients that differ in magnitude, so that
    every entry's summation order shows in its last bits)."""
    s = cube_poisson(N, "dirichlet")
    rng = np.random.default_rng(seed)
    f = s.face_cells.astype(np.int64)
    i, j = f[:, 0], f[:, 1]
    rows = np.stack([i, i, j, j], 1).ravel()
    cols = np.stack([i, j, i, j], 1).ravel()
    ne = s.n_rows // 3
    er, ec = rng.integers(0, s.n_rows, ne), rng.integers(0, s.n_rows, ne)
    rows = np.concatenate([rows, er, ec, er]) + l0
    cols = np.concatenate([cols, ec, er, er]) + l0
    v = rng.random(len(f)) + 0.5
    ev = rng.standard_normal(ne) * 0.01
    vals = np.concatenate([np.stack([v, -v, -v, v], 1).ravel(), ev, ev, np.abs(ev) * 3])
    cuts = [0, 7, len(rows) // 3, len(rows) // 3 + 1, len(rows)]
    batches = [(rows[a:b], cols[a:b], vals[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    batches.append((rows[::5], cols[::5], 0.1 * vals[::5]))
    return s.n_rows, rows, cols, vals, batches


def main():
    from oracle.pyoracle import RefLib
    ref = RefLib(1)
    for N, (l0, seed) in CASES.items():
        n, rows, cols, vals, batches = assembler_input(N, l0, seed)
        h, ri, ci, d, x = ref.assembler(n, rows, cols, batches, l0)
        xv = np.sin(np.arange(n) * 0.37)
        y = ref.spmv(h, xv, n)
        ref.destroy(h)
        # the declared sequence added once (what values_add_declared does)
        h, _, _, d1, x1 = ref.assembler(n, rows, cols, [(rows, cols, vals)], l0)
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"assembler_N{N}.npz")
        np.savez_compressed(path, row_index=ri, col_id=ci, d_val=d, x_val=x, y=y,
                            d_val_declared=d1, x_val_declared=x1)
        print("wrote", path, os.path.getsize(path), "nnz", len(ci))


if __name__ == "__main__":
    main()
