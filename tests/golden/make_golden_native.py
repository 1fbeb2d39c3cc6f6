"""Golden fixtures for the hierarchy of a matrix "mapped from native" (face-based) arrays:
cs_matrix_set_coefficients(edges, da, xa) -> cs_grid_create_from_shared (use_faces / use_xa,
cs_grid.cpp:7485-7491) -> _automatic_aggregation_fc (one thread) or _automatic_aggregation_dx_msr
(several), _coarsen_face_cell, _compute_coarse_cell_quantities,
_compute_coarse_quantities_native, _matrix_from_native -- and the V-cycle solve over it.
Outputs of the reference's own compiled CPU code (oracle/_ref/libcsref.so):

    python tests/golden/make_golden_native.py   -> tests/golden/native_{case}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib, native_arrays  # noqa: E402

# name: (N, bc, reference threads, perturbed face values, p0p1_relax)
CASES = {"N8_neumann_t1": (8, "neumann", 1, False, 0.95),
         "N12_dirichlet_t1p": (12, "dirichlet", 1, True, 0.95),
         "N12_neumann_t3p": (12, "neumann", 3, True, 0.95),
         "N12_neumann_t1p_galerkin": (12, "neumann", 1, True, 0.)}


def native_system(N, bc, perturb):
    """(system, da, xa): the cube's face arrays; `perturb` makes the face values differ from
    face to face (seeded), so that every ordered sum of the coarse assembly matters."""
    s = cube_poisson(N, bc)
    da, xa = native_arrays(s)
    if perturb:
        rng = np.random.default_rng(1234)
        xa = xa * (1 + 0.3 * rng.random(len(xa)))
        da = np.zeros(s.n_rows)
        np.add.at(da, s.face_cells[:, 0], -xa)
        np.add.at(da, s.face_cells[:, 1], -xa)
        da += 1e-3 * np.abs(xa).mean()
    return s, da, xa


def main():
    for name, (N, bc, nt, perturb, relax) in CASES.items():
        ref = RefLib(nt)
        s, da, xa = native_system(N, bc, perturb)
        h = ref.system_native(s, da, xa)
        r_norm = float(np.sqrt(ref.dot_xx(s.rhs)))
        out = {"N": N, "r_norm": r_norm, "n_threads": nt, "da": da, "xa": xa}
        lv = ref.hierarchy(h, p0p1_relax=relax)
        out["n_levels"] = len(lv)
        for k in ("row_index", "col_id", "x_val"):
            out[f"l0_{k}"] = lv[0][k]
        for l in range(1, len(lv)):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                out[f"l{l}_{k}"] = lv[l][k]
        r = ref.multigrid(h, s.rhs, 1e-8, r_norm, p0p1_relax=relax)
        for k in ("cvg", "n_iter", "residual", "vx"):
            out[f"mg_{k}"] = r[k]
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"native_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "levels", [l["n_rows"] for l in lv],
              "mg n_iter", out["mg_n_iter"])


if __name__ == "__main__":
    main()
