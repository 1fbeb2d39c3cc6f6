"""GPU parity of block-diagonal (CS_MATRIX_BLOCK_D) matrices with block sizes other than 3
(a4 _b_mat_vec_p_l_msr_6 / _generic, a17 _block_jacobi) through the C-ABI,
against the reference-built golden fixtures (tests/golden/blockn_*.npz) and the oracle.

SpMV and fixed numbers of block Jacobi sweeps have no reduction in them -> BIT-EXACT; the
residual norms to reduction-order rounding (1e-13 relative); BiCGStab (its scalars are global
sums): iteration counts +-2, solutions <= 1e-8 relative L2 at precision 1e-10.  Block sizes 4
and 9 keep the reference's behaviour of applying the LU factors of _fact_lu as an inverse
(cs_sles_it_priv.cpp:118-222): block Jacobi diverges there exactly as the reference does.
"""
import os

import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import cube_block_convection_diffusion
from conftest import rel_l2

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = [(6, 6), (2, 7), (4, 5), (9, 4)]       # (db, N)


@pytest.mark.parametrize("db,N", CASES)
def test_blockn_spmv_and_jacobi_bit_exact_vs_reference_golden(csb, db, N):
    g = np.load(os.path.join(GOLD, f"blockn_db{db}_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db)
    n = len(ri) - 1
    r_norm = float(g["r_norm"])
    m = csb.Matrix.from_block(n, db, ri, ci, d.ravel(), xv)
    assert np.array_equal(m.vector_multiply(g["x"]), g["y"])
    assert np.array_equal(m.vector_multiply(g["x"], exclude_diag=True), g["y_exdiag"])
    for nit in (1, 2, 4):
        c = csb.SlesIt(csb.SlesItType.JACOBI, nit)
        r = c.solve(m, rhs, 1e-12, r_norm)
        assert r["n_iter"] == nit
        assert np.array_equal(r["vx"], g[f"jac{nit}_vx"]), nit
        assert r["residual"] == pytest.approx(float(g[f"jac{nit}_residual"]), rel=1e-12)
        c.destroy()
    c = csb.SlesIt(csb.SlesItType.JACOBI, 3)
    w = c.solve(m, rhs, 1e-12, r_norm, vx=g["x"].copy(), use_vx_ini=True)
    assert np.array_equal(w["vx"], g["jac_warm_vx"])
    c.destroy()
    c = csb.SlesIt(csb.SlesItType.JACOBI, 500)
    r = c.solve(m, rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == int(g["jac_cvg"])                   # converged, or DIVERGED (-3)
    assert abs(r["n_iter"] - int(g["jac_n_iter"])) <= 1
    c.destroy()
    m.destroy()


@pytest.mark.parametrize("db,N", CASES)
def test_blockn_bicgstab_vs_reference_golden(csb, db, N):
    g = np.load(os.path.join(GOLD, f"blockn_db{db}_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db)
    n = len(ri) - 1
    m = csb.Matrix.from_block(n, db, ri, ci, d.ravel(), xv)
    c = csb.SlesIt(csb.SlesItType.BICGSTAB, 500)
    r = c.solve(m, rhs, 1e-10, float(g["r_norm"]))
    assert int(r["cvg"]) == 1 == int(g["bicg_cvg"])
    assert abs(r["n_iter"] - int(g["bicg_n_iter"])) <= 2
    assert rel_l2(r["vx"], g["bicg_vx"]) <= 1e-8
    c.destroy()
    m.destroy()


@pytest.mark.parametrize("db,N", [(6, 24), (5, 17)])
def test_blockn_vs_oracle_larger(csb, oracle, db, N):
    """Several thread blocks, ragged last block; set_coefficients refreshes whole blocks."""
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db, seed=5)
    n = len(ri) - 1
    x = np.sin(0.21 * np.arange(db * n))
    m = csb.Matrix.from_block(n, db, ri, ci, np.zeros_like(d).ravel(), np.zeros_like(xv))
    m.set_coefficients(d.ravel(), xv, symmetric=False)
    assert np.array_equal(m.vector_multiply(x), oracle.spmv_bn(db, ri, ci, d.ravel(), xv, x))
    r_norm = float(np.linalg.norm(rhs))
    c = csb.SlesIt(csb.SlesItType.JACOBI, 3)
    r = c.solve(m, rhs, 1e-12, r_norm)
    o = oracle.block_jacobi(db, ri, ci, d.ravel(), xv, rhs, 3, 1e-12, r_norm)
    assert np.array_equal(r["vx"], o["vx"])
    c.destroy()
    c = csb.SlesIt(csb.SlesItType.BICGSTAB, 300)
    r = c.solve(m, rhs, 1e-10, r_norm)
    o = oracle.bicgstab_bn(db, ri, ci, d.ravel(), xv, rhs, 300, 1e-10, r_norm)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"], o["vx"]) <= 1e-8
    c.destroy()
    m.destroy()


@pytest.mark.parametrize("db,N", CASES + [(3, 6)])
def test_block_smoothers_bit_exact_vs_reference_golden(csb, db, N):
    """cs_multigrid_smoother on block-diagonal matrices (the block variants of a26): block
    Jacobi and block relaxed Jacobi (R_JACOBI, RJ2, RJ3) -- per-row arithmetic, no reduction."""
    g = np.load(os.path.join(GOLD, f"blockn_db{db}_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db)
    n = len(ri) - 1
    m = csb.Matrix.from_block(n, db, ri, ci, d.ravel(), xv)
    for st in (3, 14, 15, 16):
        for nit in (1, 3):
            sm = csb.Smoother(st, nit)
            r = sm.solve(m, rhs)
            assert int(r["cvg"]) == -1
            assert np.array_equal(r["vx"], g[f"sm{st}_{nit}_zero"]), (st, nit)
            r = sm.solve(m, rhs, vx=g["x"].copy(), use_vx_ini=True)
            assert np.array_equal(r["vx"], g[f"sm{st}_{nit}_warm"]), (st, nit)
            assert r["n_iter"] == int(g[f"sm{st}_{nit}_n_iter"])
            sm.destroy()
    with pytest.raises(csb.CsbError):
        csb.Smoother(csb.SlesItType.L1_JACOBI, 2).solve(m, rhs)      # no block form (reference: error)
    m.destroy()
