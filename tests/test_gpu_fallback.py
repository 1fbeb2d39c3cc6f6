"""GPU parity of GCR and of the solver fallback of cs_sles_it_solve (a22;
_fallback cs_sles_it.cpp:4464-4528, call site :5383-5418, device choice GCR) through the C-ABI,
against fixtures built from the reference's own runs (tests/golden/make_golden_gcr.py).

GCR's coefficients are all global sums (taken in another order here): iteration counts +-2,
solutions to <= 1e-8 relative L2 at precision 1e-10.
"""
import os

import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import (cube_block3_convection_diffusion,
                                              cube_block_convection_diffusion, cube_poisson)
from conftest import rel_l2

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "gcr.npz"))


def test_gcr_scalar_vs_reference_golden(csb):
    s = cube_poisson(12, "dirichlet")
    rn = float(G["s_r_norm"])
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    for tag, pd, v0, nmi, prec in (("jac_zero", 0, None, 200, 1e-10), ("jac_warm", 0, G["s_vx0"], 200, 1e-10),
                                   ("npc_warm", -1, G["s_vx0"], 200, 1e-10),
                                   ("jac_27", 0, G["s_vx0"], 27, 1e-14)):
        c = csb.SlesIt(csb.SlesItType.GCR, nmi, poly_degree=pd)
        r = c.solve(m, s.rhs, prec, rn, vx=None if v0 is None else v0.copy(),
                    use_vx_ini=v0 is not None)
        assert int(r["cvg"]) == int(G[f"s_{tag}_cvg"]), tag
        assert abs(r["n_iter"] - int(G[f"s_{tag}_n_iter"])) <= 2, tag
        assert r["initial_residual"] == pytest.approx(float(G[f"s_{tag}_initial_residual"]), rel=1e-12)
        assert rel_l2(r["vx"][:s.n_rows], G[f"s_{tag}_vx"]) <= 1e-8, tag
        c.destroy()
    m.destroy()


def test_gcr_restart_interval(csb, oracle):
    s = cube_poisson(16, "dirichlet", with_mesh=False)
    rn = float(np.linalg.norm(s.rhs))
    vx0 = np.cos(0.3 * np.arange(s.n_rows))
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    c = csb.SlesIt(csb.SlesItType.GCR, 400)
    c.set_restart_interval(7)
    r = c.solve(m, s.rhs, 1e-9, rn, vx=vx0.copy(), use_vx_ini=True)
    o = oracle.gcr(1, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 400, 1e-9, rn, vx0, restart=7)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= 1e-7
    k = min(len(r["history"]), len(o["history"])) - 2
    assert k > 7 and np.allclose(r["history"][:k], o["history"][:k], rtol=1e-6)
    c.destroy()
    m.destroy()


def test_bicgstab_falls_back_to_gcr_from_its_iterate(csb):
    """Threshold CONVERGED: a BiCGStab run that ends at its iteration limit hands its vx to
    GCR (prev_state >= BREAKDOWN keeps vx, cs_sles_it.cpp:4503); counts are added (:4521)."""
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(8)
    n = len(ri) - 1
    rn = float(G["b3_r_norm"])
    m = csb.Matrix.from_block(n, 3, ri, ci, d.ravel(), xv)
    c = csb.SlesIt(csb.SlesItType.BICGSTAB, 3)
    r = c.solve(m, rhs, 1e-10, rn)
    assert int(r["cvg"]) == -1 and r["n_iter"] == 3 and not c.used_fallback()   # default: no fallback
    c.set_fallback_threshold(csb.ConvergenceState.CONVERGED, 200)
    r = c.solve(m, rhs, 1e-10, rn)
    assert c.used_fallback()
    assert int(r["cvg"]) == int(G["b3_fb200_cvg"]) == 1
    assert abs(r["n_iter"] - (3 + int(G["b3_fb200_n_iter"]))) <= 2
    assert rel_l2(r["vx"], G["b3_fb200_vx"]) <= 1e-8
    c.set_fallback_threshold(csb.ConvergenceState.CONVERGED, 3)                 # fallback stops too
    r = c.solve(m, rhs, 1e-10, rn)
    assert int(r["cvg"]) == int(G["b3_fb_cvg"]) == -1 and r["n_iter"] == 3 + int(G["b3_fb_n_iter"])
    assert rel_l2(r["vx"], G["b3_fb_vx"]) <= 1e-9
    c.destroy()
    m.destroy()


def test_diverged_block_jacobi_falls_back_to_gcr_from_zero(csb):
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(5, 4)
    n = len(ri) - 1
    rn = float(G["b4_r_norm"])
    m = csb.Matrix.from_block(n, 4, ri, ci, d.ravel(), xv)
    c = csb.SlesIt(csb.SlesItType.JACOBI, 500)
    r = c.solve(m, rhs, 1e-8, rn)
    assert int(r["cvg"]) == int(G["b4_fb_first_cvg"]) == -3 and not c.used_fallback()
    n_first = r["n_iter"]
    c.set_fallback_threshold(csb.ConvergenceState.BREAKDOWN, 500)
    r = c.solve(m, rhs, 1e-8, rn)
    assert c.used_fallback() and int(r["cvg"]) == 1
    assert abs(r["n_iter"] - (n_first + int(G["b4_fb_n_iter"]))) <= 2
    assert rel_l2(r["vx"], G["b4_fb_vx"]) <= 1e-6
    c.destroy()
    m.destroy()
