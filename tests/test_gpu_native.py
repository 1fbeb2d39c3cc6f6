"""Matrix "mapped from native" (face-based) arrays -- SURVEY.md section 8 rows a29 / a35:
csb_matrix_create_native -> csb_multigrid_setup follows the reference's path for such a matrix
(_automatic_aggregation_fc, _coarsen_face_cell, _compute_coarse_cell_quantities,
_compute_coarse_quantities_native, _matrix_from_native; src/alge/cs_grid.cpp:3997, :637,
:4778, :5144, :7105).  Bar: maps, structures and coarse values bit-exact (face-order sums);
cycle counts within +-2 equivalent iterations; solutions to the solver tolerance."""
import os
import sys

import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import cube_poisson
from conftest import equiv_iters_close, rel_l2

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
from make_golden_native import CASES, native_system  # noqa: E402


def _setup(csb, s, da, xa, blocks, relax):
    m = csb.Matrix.from_system_native(s, da, xa)
    mg = csb.Multigrid()
    mg.set_aggregation_blocks(blocks)
    mg.set_coarsening_options(p0p1_relax=relax)
    mg.setup(m)
    return m, mg


@pytest.mark.parametrize("name,case", sorted(CASES.items()))
def test_native_hierarchy_matches_reference_golden(csb, name, case):
    N, bc, nt, perturb, relax = case
    g = np.load(os.path.join(GOLD, f"native_{name}.npz"))
    s, da, xa = native_system(N, bc, perturb)
    m, mg = _setup(csb, s, da, xa, nt, relax)
    # level 0: the MSR matrix cs_matrix_set_coefficients(edges, da, xa) builds
    d0, x0 = m.get_coefficients()
    assert np.array_equal(d0, da) and np.array_equal(x0, g["l0_x_val"])
    lv = mg.levels()
    assert len(lv) == int(g["n_levels"])
    assert np.array_equal(lv[0]["row_index"], g["l0_row_index"])
    assert np.array_equal(lv[0]["col_id"], g["l0_col_id"])
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], g[f"l{l}_{k}"]), (l, k)
    r_norm = float(g["r_norm"])
    r = mg.solve(m, s.rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == int(g["mg_cvg"]) == 1
    assert equiv_iters_close(r["n_iter"], int(g["mg_n_iter"]), r["n_iter"] / max(r["n_cycles"], 1))
    res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
    assert res <= 1.0001e-8 * r_norm
    mg.destroy(); m.destroy()


@pytest.mark.parametrize("N,bc,blocks", [(40, "neumann", 1), (48, "dirichlet", 1), (40, "dirichlet", 4)])
def test_native_hierarchy_vs_oracle_larger(csb, oracle, N, bc, blocks):
    """larger cubes with face values that differ from face to face: every ordered sum of the
    coarse assembly matters; tight solve compared with the oracle's"""
    s, da, xa = native_system(N, bc, True)
    omg = oracle.mg_setup_native(s, n_threads=blocks, da=da, xa=xa)
    olv = oracle.mg_levels(omg)
    m, mg = _setup(csb, s, da, xa, blocks, 0.95)
    lv = mg.levels()
    assert [x["n_rows"] for x in lv] == [x["n_rows"] for x in olv]
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], olv[l][k]), (l, k)
    r_norm = float(np.linalg.norm(s.rhs))
    o = oracle.mg_solve(omg, s.rhs, 1e-12, r_norm)
    r = mg.solve(m, s.rhs, 1e-12, r_norm)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_cycles"] - o["n_cycles"]) <= 2
    assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= 1e-10
    oracle.mg_free(omg); mg.destroy(); m.destroy()


def test_native_fcg_mg_production_layout(csb, oracle):
    """the production solver (FCG + V-cycle preconditioner, RJ3 smoothers) over a native
    hierarchy: same iteration count as the oracle's run over the oracle's native hierarchy"""
    s, da, xa = native_system(32, "neumann", True)
    omg = oracle.mg_setup_native(s, da=da, xa=xa)
    m = csb.Matrix.from_system_native(s, da, xa)
    r_norm = float(np.linalg.norm(s.rhs))
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    pc = csb.Multigrid(pc=True)
    pc.set_aggregation_blocks(1)
    c.transfer_pc(pc)
    r = c.solve(m, s.rhs, 1e-10, r_norm)
    o = oracle.fcg_mgpc(omg, s.rhs, 200, 1e-10, r_norm)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= 1e-8
    oracle.mg_free(omg); c.destroy(); m.destroy()


def test_native_new_values_and_errors(csb):
    """csb_matrix_set_coefficients_native replaces the values on the same edges; duplicate
    edges, non-symmetric input and MSR-style updates are refused loudly"""
    s, da, xa = native_system(10, "dirichlet", False)
    m = csb.Matrix.from_system_native(s, da, xa)
    x = np.sin(np.arange(s.n_rows) * 0.37)
    y0 = m.vector_multiply(x.copy())
    m.set_coefficients_native(2. * da, 2. * xa)
    assert np.array_equal(m.vector_multiply(x.copy()), 2. * y0)
    with pytest.raises(csb.CsbError):
        m.set_coefficients(da, np.zeros(m.nnz))
    m.destroy()
    dup = np.vstack([s.face_cells, s.face_cells[:1]])
    with pytest.raises(csb.CsbError):
        csb.Matrix.from_native(s.n_rows, dup, da, np.append(xa, xa[0]))
    with pytest.raises(csb.CsbError):
        csb.Matrix.from_native(s.n_rows, np.array([[0, s.n_rows]]), da, xa[:1])


def test_native_spmv_matches_msr_entry(csb):
    """the matrix built from the face arrays multiplies like the MSR-built one"""
    s = cube_poisson(12, "neumann")
    a = csb.Matrix.from_system(s)
    b = csb.Matrix.from_system_native(s)
    x = np.cos(np.arange(s.n_rows) * 0.11)
    assert np.array_equal(a.vector_multiply(x.copy()), b.vector_multiply(x.copy()))
    a.destroy(); b.destroy()
