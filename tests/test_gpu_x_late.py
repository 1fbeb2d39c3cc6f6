"""GPU tests of the two changes made after the round's last visit to a B200 (the GPU budget
of the round was spent): the K-cycle with its inner Krylov coefficients kept on the device,
and the gradient limiters.  Both were checked in the GPU-less container against the same
fixtures through tests/cudasim (host threads play CUDA threads; see tests/test_cudasim.py);
they sit in a file that sorts last among the GPU tests so that the files validated on the
device run first.  Bars as everywhere: bit-identical for the per-row arithmetic, iteration
counts +-2, solutions to the solver tolerance."""
import os
import sys

import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import cube_poisson
from conftest import rel_l2

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
from make_golden_rhs import CASES, rc_input  # noqa: E402


def _rc_matrix(csb, s, bfc, b_geo, i_geo):
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
    return m


# ------------------------------------------------------ K-cycle

def _kpc(csb, min_g_rows=30, limit=4):
    pc = csb.Multigrid(pc=True, mg_type=1)           # cs_multigrid_pc_create(CS_MULTIGRID_K_CYCLE)
    pc.set_coarsening_options(aggregation_limit=limit, coarsening_type=3, n_max_levels=10,
                              min_g_rows=min_g_rows, p0p1_relax=0.)
    return pc


@pytest.mark.parametrize("N,bc", [(8, "dirichlet"), (12, "dirichlet")])
def test_k_cycle_matches_reference_golden(csb, N, bc):
    """One K-cycle (pairwise aggregation, Galerkin, 2 Jacobi sweeps, inner 2-step Krylov with
    the 0.25 threshold) and FCG preconditioned by it."""
    g = np.load(os.path.join(GOLD, f"kcycle_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    mg = _kpc(csb)
    mg.setup(m)
    r = mg.solve(m, s.rhs, 1e-8, r_norm)
    assert rel_l2(r["vx"][:s.n_rows], g["cycle_vx"]) <= 1e-6       # coarse PCG stops at 1e-7
    assert abs(r["n_iter"] - int(g["cycle_n_iter"])) <= 1
    mg.destroy()
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(_kpc(csb))
    r = c.solve(m, s.rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == int(g["fcg_cvg"]) == 1
    assert abs(r["n_iter"] - int(g["fcg_n_iter"])) <= 2
    res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
    assert res <= 1.01e-8 * r_norm
    assert rel_l2(r["vx"][:s.n_rows], g["fcg_vx"]) <= 1e-6
    c.destroy(); m.destroy()


def test_k_cycle_deep_hierarchy_vs_oracle(csb, oracle):
    """24^3, pairs only (limit 2): 9 levels, both branches of the inner iteration occur."""
    s = cube_poisson(24, "dirichlet", with_mesh=False)
    r_norm = float(np.linalg.norm(s.rhs))
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    omg = oracle.mg_setup_pw(cube_poisson(24, "dirichlet"), aggregation_limit=2, n_levels_max=10,
                             n_g_rows_min=30)
    mg = _kpc(csb, limit=2)
    mg.setup(m)
    o = oracle.mg_k_cycle(omg, s.rhs, 1e-13, r_norm)
    r = mg.solve(m, s.rhs, 1e-13, r_norm)
    assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= 1e-9
    mg.destroy()
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(_kpc(csb, limit=2))
    r = c.solve(m, s.rhs, 1e-10, r_norm)
    o = oracle.it_kpc(1, omg, s.rhs, 200, 1e-10, r_norm)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    c.destroy(); oracle.mg_free(omg); m.destroy()



# --------------------------------------------------------------- gradient limiters

@pytest.mark.parametrize("N", CASES)
def test_gradient_limiter_matches_reference_golden(csb, N):
    """eqp->imligr 0 / 1 / 2 (cell, face and I'/J' limiters; factors 0.3 and 1.5) on gradient
    types 0 / 1 / 4, with the limiter of the boundary values at I', two cases with the
    hydrostatic treatment: limited gradient and divergence bit-identical to the reference;
    switching the limiter off gives the plain result back"""
    from make_golden_rhs import CLIP_CASES, diipb_input, green_input, hyd_input
    g = np.load(os.path.join(GOLD, f"rhs_clip_N{N}.npz"))
    s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    fx = hyd_input(s.n_rows)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    m.set_face_cog(ifc, b_geo[:, 3:6])
    m.set_boundary_iprime(diipb_input(N, len(bfc)))
    plain = m.gradient(4, f["pvar"], f["coefa"], f["coefb"])
    for imligr, climgr, imrgra, hyd in CLIP_CASES:
        m.set_gradient_type(imrgra)
        m.set_gradient_limiter(imligr, climgr)
        tag = f"l{imligr}_c{climgr}_g{imrgra}_h{int(hyd)}"
        if hyd:
            gr = m.gradient_hyd(imrgra, fx, f["pvar"], f["coefa"], f["coefb"])
            d = m.diffusion_potential_hyd(fx, f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                          f["cofbf"], f["i_visc"], f["b_visc"], f["visel"])
        else:
            gr = m.gradient(imrgra, f["pvar"], f["coefa"], f["coefb"])
            d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                         f["cofbf"], f["i_visc"], f["b_visc"], f["visel"])
        assert np.array_equal(gr, g["grad_" + tag]), tag
        assert np.array_equal(d, g["div_" + tag]), tag
    with pytest.raises(csb.CsbError):
        m.set_gradient_limiter(3, 1.0)
    m.set_gradient_limiter(-1, -1.)
    assert np.array_equal(m.gradient(4, f["pvar"], f["coefa"], f["coefb"]), plain)
    m.destroy()


# ------------------------------------------------------------------- irregular mesh

def test_face_operators_on_irregular_mesh_vs_oracle(csb, oracle):
    """rows of 3 to 7 entries, arbitrary face orientation and order, unordered boundary faces
    (tests/irregular_mesh.py; the oracle is pinned to the reference on this very mesh in
    tests/test_oracle.py): gradients 0 / 1 / 4, hydrostatic treatment, limiters, reconstructed
    divergence, face fluxes and their divergence, all bit-exact"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import check_device_against_oracle
    check_device_against_oracle(csb, oracle)


def test_multigrid_on_irregular_mesh_vs_oracle(csb, oracle):
    """the hot path beyond the hex cube: hierarchy of an irregular system (3000 rows of 4 to 10
    entries) bit-exact against the oracle -- itself pinned to the reference on this system in
    tests/test_oracle.py --, V-cycle solve: cycles +-1, first residuals, solution"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import check_multigrid_against_oracle
    check_multigrid_against_oracle(csb, oracle)


def test_solvers_on_irregular_mesh_vs_oracle(csb, oracle):
    """SpMV and smoothers bit-exact, PCG and the production layout FCG + cs_multigrid_pc
    (k_fcg_mg_persistent) on the irregular system: iteration counts +-2, true residual under
    the threshold, solution to the solver tolerance"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import check_solvers_against_oracle
    check_solvers_against_oracle(csb, oracle)


def test_native_entry_on_irregular_mesh_vs_oracle(csb, oracle):
    """the irregular system entered through its face arrays: level-0 MSR matrix and hierarchy
    (face-list aggregation, face-order coarse values) bit-exact against the oracle"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import check_native_against_oracle
    check_native_against_oracle(csb, oracle)


def test_coefficient_build_and_k_cycle_on_irregular_mesh_vs_oracle(csb, oracle):
    """cs_matrix_compute_coeffs (symmetric and convection forms) bit-exact, pairwise hierarchies
    bit-exact, the K-cycle and FCG + K-cycle on the irregular system against the oracle (pinned
    to the reference on this system in tests/test_oracle.py)"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import check_coeffs_and_k_cycle_against_oracle
    check_coeffs_and_k_cycle_against_oracle(csb, oracle)


def test_block3_on_irregular_mesh_vs_oracle(csb, oracle):
    """3x3-block SpMV bit-exact and BiCGStab + Jacobi on a non-symmetric block system over the
    irregular graph: early residual history to 1e-9, true residual under the threshold, count
    within 15 % (BiCGStab amplifies rounding differences exponentially on this system: see
    tests/irregular_mesh.py)"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import check_block3_against_oracle
    check_block3_against_oracle(csb, oracle)

