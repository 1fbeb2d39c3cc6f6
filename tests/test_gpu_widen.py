"""GPU parity of the rows added after the first V-cycle path (a14/a15 FCG with a
general preconditioner, a24 _multigrid_v_cycle_pc, a26 Jacobi-family multigrid smoothers),
through the C-ABI, against the reference-built golden fixtures (tests/golden/mgpc_*.npz) and
the oracle.

Tolerances: the smoothers are per-row arithmetic with no reduction -> bit-exact.  The
single-cycle V-cycle contains a coarsest-level PCG whose dot products are summed in another
order -> agreement to the coarse solver's tolerance; FCG + multigrid: iteration counts +-2,
solutions <= 1e-10 relative L2 when both are solved tightly.
"""
import os

import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import cube_poisson
from conftest import equiv_iters_close, rel_l2

pytestmark = pytest.mark.gpu

SOL_TOL = 1e-10
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MGPC_CASES = [(8, "neumann"), (12, "dirichlet")]
SMOOTHERS = {"jacobi": 3, "l1_jacobi": 13, "r_jacobi": 14, "rj2": 15, "rj3": 16}
CYCLES = {"rj3": ((16, 16, 0), (3, 3, 500)), "rj2": ((15, 15, 0), (2, 4, 500)),
          "r_jacobi": ((14, 14, 0), (2, 2, 500)), "l1_jacobi": ((13, 13, 0), (2, 2, 500)),
          "jacobi": ((3, 3, 0), (2, 3, 500))}


def _mg(csb, types, nmi, precision_mult=(-1., -1., 1.)):
    mg = csb.Multigrid()
    mg.set_solver_options(descent=types[0], ascent=types[1], coarse=types[2], n_max_cycles=1,
                          n_max_iter_descent=nmi[0], n_max_iter_ascent=nmi[1],
                          n_max_iter_coarse=nmi[2], precision_mult=precision_mult)
    return mg


# ------------------------------------------------------------------------------ smoothers

@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_smoothers_bit_exact_vs_reference_golden(csb, N, bc):
    """cs_multigrid_smoother_solve: every Jacobi-family type, from zero and from a given
    iterate; one fused kernel per sweep, same IEEE operations as the reference -> identical."""
    g = np.load(os.path.join(GOLD, f"mgpc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    for name, st in SMOOTHERS.items():
        for nit in (1, 3):
            sm = csb.Smoother(st, nit)
            r = sm.solve(m, s.rhs)
            assert int(r["cvg"]) == -1                       # CS_SLES_MAX_ITERATION
            assert r["residual"] == -1.0
            assert np.array_equal(r["vx"][:s.n_rows], g[f"sm_{name}_{nit}_zero"]), (name, nit)
            r = sm.solve(m, s.rhs, vx=g["vx0"].copy(), use_vx_ini=True)
            assert np.array_equal(r["vx"][:s.n_rows], g[f"sm_{name}_{nit}_warm"]), (name, nit)
            assert r["n_iter"] == int(g[f"sm_{name}_{nit}_n_iter"])
            sm.destroy()
    m.destroy()


@pytest.mark.parametrize("N", [33, 64])
def test_smoothers_bit_exact_vs_oracle(csb, oracle, N):
    """Sizes with several thread blocks and a ragged last block; even and odd sweep counts
    (result in either ping-pong buffer)."""
    s = cube_poisson(N, "neumann", with_mesh=False)
    rng = np.random.default_rng(N)
    vx0 = rng.standard_normal(s.n_rows)
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    for st in SMOOTHERS.values():
        for nit in (2, 5):
            sm = csb.Smoother(st, nit)
            for v0 in (None, vx0):
                o = oracle.smoother(st, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nit, v0)
                r = sm.solve(m, s.rhs, vx=None if v0 is None else v0.copy(),
                             use_vx_ini=v0 is not None)
                assert r["n_iter"] == o["n_iter"]
                assert np.array_equal(r["vx"][:s.n_rows], o["vx"]), (st, nit, v0 is None)
            sm.destroy()
    m.destroy()


def test_smoother_rejects_other_types(csb):
    with pytest.raises(csb.CsbError):
        csb.Smoother(csb.SlesItType.PCG, 2)
    with pytest.raises(csb.CsbError):
        csb.SlesIt(csb.SlesItType.RJ3, 3)


# ------------------------------------------------- single-cycle V-cycle (preconditioner form)

@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_cycle_pc_matches_reference_golden(csb, N, bc):
    """cs_multigrid_solve with n_max_cycles == 1 = _multigrid_v_cycle_pc: no convergence
    test, ascent smoother on level 0 too, returns CS_SLES_ITERATING."""
    g = np.load(os.path.join(GOLD, f"mgpc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    m = csb.Matrix.from_system(s)
    for name, (types, nmi) in CYCLES.items():
        mg = _mg(csb, types, nmi)
        mg.setup(m)
        r = mg.solve(m, s.rhs, 1e-8, r_norm)
        assert int(r["cvg"]) == 0
        assert abs(r["n_iter"] - int(g[f"cyc_{name}_n_iter"])) <= 1   # coarse PCG count +-1
        # the coarsest-level PCG stops at 1e-8 in both
        assert rel_l2(r["vx"][:s.n_rows], g[f"cyc_{name}_vx"]) <= 1e-6, name
        mg.destroy()
    m.destroy()


@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_v_cycle_solver_mode_with_smoothers_matches_reference_golden(csb, N, bc):
    """cs_multigrid V-cycle in solver mode (cycles to convergence, fine-level residual test every
    cycle) with cs_multigrid_smoother smoothers: kernel-per-phase cycle, no reduction except
    the residual norm and the coarsest solve."""
    g = np.load(os.path.join(GOLD, f"mgpc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    m = csb.Matrix.from_system(s)
    for name, (types, nmi) in CYCLES.items():
        mg = csb.Multigrid()
        mg.set_solver_options(descent=types[0], ascent=types[1], coarse=types[2],
                              n_max_cycles=100, n_max_iter_descent=nmi[0],
                              n_max_iter_ascent=nmi[1], n_max_iter_coarse=nmi[2],
                              precision_mult=(-1., -1., 1.))
        mg.setup(m)
        r = mg.solve(m, s.rhs, 1e-8, r_norm)
        n_ref = int(g[f"vsm_{name}_n_iter"])
        assert int(r["cvg"]) == int(g[f"vsm_{name}_cvg"]) == 1, name
        assert equiv_iters_close(r["n_iter"], n_ref, r["n_iter"] / max(r["n_cycles"], 1)), name
        assert abs(r["residual"] - float(g[f"vsm_{name}_residual"])) <= 1e-6 * r["residual"]
        assert rel_l2(r["vx"][:s.n_rows], g[f"vsm_{name}_vx"]) <= (1e-7 if bc == "dirichlet" else 1e-3)
        mg.destroy()
    m.destroy()


def test_cycle_pc_tight_vs_oracle(csb, oracle):
    """With the coarsest level solved to 1e-14 the whole cycle agrees to rounding; also from a
    nonzero initial vx (level-0 descent smoother starts from it)."""
    s = cube_poisson(24, "neumann")
    r_norm = float(np.linalg.norm(s.rhs))
    vx0 = np.sin(0.11 * np.arange(s.n_rows))
    m = csb.Matrix.from_system(s)
    omg = oracle.mg_setup(s)
    for name, (types, nmi) in CYCLES.items():
        mg = _mg(csb, types, nmi)
        mg.setup(m)
        for v0 in (None, vx0):
            o = oracle.mg_cycle_pc(omg, s.rhs, 1e-14, r_norm, types, nmi, vx0=v0)
            r = mg.solve(m, s.rhs, 1e-14, r_norm, vx=None if v0 is None else
                         np.concatenate([v0, np.zeros(m.n_cols_ext - s.n_rows)]),
                         use_vx_ini=v0 is not None)
            assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= SOL_TOL, (name, v0 is None)
        mg.destroy()
    oracle.mg_free(omg)
    m.destroy()


def test_cycle_pc_with_pcg_smoothers_vs_reference(csb, oracle):
    """n_max_cycles == 1 with the cs_multigrid_create default PCG smoothers also takes the
    single-cycle form in the reference (cs_multigrid.cpp:5484)."""
    s = cube_poisson(16, "neumann")
    r_norm = float(np.linalg.norm(s.rhs))
    m = csb.Matrix.from_system(s)
    omg = oracle.mg_setup(s)
    mg = _mg(csb, (0, 0, 0), (2, 10, 500), precision_mult=(1., 1., 1.))
    mg.setup(m)
    o = oracle.mg_cycle_pc(omg, s.rhs, 1e-14, r_norm, (0, 0, 0), (2, 10, 500), (1., 1., 1.))
    r = mg.solve(m, s.rhs, 1e-14, r_norm)
    assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= 1e-9
    mg.destroy(); oracle.mg_free(omg); m.destroy()


# ----------------------------------------------------------- FCG + multigrid preconditioner

@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_fcg_mgpc_matches_reference_golden(csb, N, bc):
    """The reference's production layout: cs_sles_it FCG + cs_multigrid_pc (one V-cycle per
    iteration), for each smoother family."""
    g = np.load(os.path.join(GOLD, f"mgpc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    m = csb.Matrix.from_system(s)
    for name, (types, nmi) in CYCLES.items():
        c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
        c.transfer_pc(_mg(csb, types, nmi))
        r = c.solve(m, s.rhs, 1e-8, r_norm)
        assert int(r["cvg"]) == int(g[f"fcg_{name}_cvg"]) == 1
        assert abs(r["n_iter"] - int(g[f"fcg_{name}_n_iter"])) <= 2, name
        ir = float(g[f"fcg_{name}_initial_residual"])
        assert abs(r["initial_residual"] - ir) <= 1e-12 * ir
        k = min(len(r["history"]) - 1, len(g[f"fcg_{name}_history"]), 4)
        assert np.allclose(r["history"][1:1 + k], g[f"fcg_{name}_history"][:k], rtol=1e-5, atol=0)
        res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
        assert res <= 1.01e-8 * r_norm
        assert rel_l2(r["vx"][:s.n_rows], g[f"fcg_{name}_vx"]) <= (1e-6 if bc == "dirichlet" else 1e-3)
        c.destroy()
    m.destroy()


def test_fcg_mgpc_default_pc_tight_tolerance_solution_parity(csb, oracle):
    """csb_multigrid_pc_create defaults (RJ3, 3 sweeps, PCG coarse): both solved to 1e-13 ->
    solutions within the north-star tolerance; warm start with a max-iteration exit."""
    s = cube_poisson(32, "dirichlet")
    r_norm = float(np.sqrt(oracle.dot_xx(s.rhs)))
    omg = oracle.mg_setup(s)
    o = oracle.fcg_mgpc(omg, s.rhs, 200, 1e-13, r_norm)
    m = csb.Matrix.from_system(s)
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(csb.Multigrid(pc=True))
    r = c.solve(m, s.rhs, 1e-13, r_norm)
    assert int(r["cvg"]) == 1 and o["cvg"] == 1
    assert abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= SOL_TOL
    # second solve with the same object (hierarchy kept), zero right-hand side
    z = c.solve(m, np.zeros(s.n_rows), 1e-8, 1.0)
    assert not z["vx"].any()
    # warm start, 3 iterations allowed: same iterate as the oracle's
    c3 = csb.SlesIt(csb.SlesItType.FCG, 3, poly_degree=-1)
    c3.transfer_pc(csb.Multigrid(pc=True))
    ow = oracle.fcg_mgpc(omg, s.rhs, 3, 1e-14, r_norm, vx0=0.5 * o["vx"])
    w = c3.solve(m, s.rhs, 1e-14, r_norm,
                 vx=np.concatenate([0.5 * o["vx"], np.zeros(m.n_cols_ext - s.n_rows)]),
                 use_vx_ini=True)
    assert (int(w["cvg"]), w["n_iter"]) == (ow["cvg"], ow["n_iter"]) == (-1, 3)
    assert rel_l2(w["vx"][:s.n_rows], ow["vx"]) <= 1e-9
    c.destroy(); c3.destroy(); oracle.mg_free(omg); m.destroy()


def test_fcg_without_preconditioner_converges(csb):
    """poly_degree < 0 and nothing transferred: identity preconditioner."""
    s = cube_poisson(20, "neumann", with_mesh=False)
    r_norm = float(np.linalg.norm(s.rhs))
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    c = csb.SlesIt(csb.SlesItType.FCG, 2000, poly_degree=-1)
    r = c.solve(m, s.rhs, 1e-10, r_norm)
    assert int(r["cvg"]) == 1 and r["n_iter"] > 5
    res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
    assert res <= 1.01e-10 * r_norm and abs(res - r["residual"]) <= 1e-2 * res + 1e-14 * r_norm
    c.destroy(); m.destroy()


def test_fcg_mgpc_128_true_residual(csb):
    """Config-2 size (128^3): FCG + V-cycle(RJ3) converges in a few tens of iterations and the
    reported residual is the true one."""
    s = cube_poisson(128, "neumann")
    r_norm = float(np.linalg.norm(s.rhs))
    m = csb.Matrix.from_system(s)
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(csb.Multigrid(pc=True))
    r = c.solve(m, s.rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == 1 and r["n_iter"] <= 60
    res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
    assert res <= 1.01e-8 * r_norm and abs(res - r["residual"]) <= 1e-2 * res + 1e-14 * r_norm
    c.destroy(); m.destroy()


# -------------------------------- cs_sles_it with general preconditioners (a14, a21)

PC_TAGS = [("pcg_poly1", 0, 1), ("pcg_poly2", 0, 2), ("fcg_poly1", 1, 1), ("fcg_poly2", 1, 2),
           ("fcg_poly-1", 1, -1), ("ipcg_poly1", 2, 1), ("ipcg_poly2", 2, 2)]


@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_general_pc_matches_reference_golden(csb, N, bc):
    """Single-reduction PCG (_conjugate_gradient_sr), FCG and IPCG with the polynomial
    preconditioner (degree+1 fused Jacobi sweeps from zero), FCG without preconditioner, PCG
    preconditioned by one V-cycle: iteration counts +-2, first residuals to reduction-order
    rounding, warm start with a max-iteration exit reproduces the same iterate."""
    g = np.load(os.path.join(GOLD, f"pc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    m = csb.Matrix.from_system(s)

    def make(tag, so, deg, nmax):
        c = csb.SlesIt((csb.SlesItType.PCG, csb.SlesItType.FCG, csb.SlesItType.IPCG)[so], nmax,
                       poly_degree=deg)
        if tag.endswith("_mgpc"):
            c.transfer_pc(csb.Multigrid(pc=True))
        return c

    for tag, so, deg in PC_TAGS + [("pcg_mgpc", 0, -1), ("ipcg_mgpc", 2, -1)]:
        c = make(tag, so, deg, 500)
        r = c.solve(m, s.rhs, 1e-9, r_norm)
        assert int(r["cvg"]) == int(g[f"{tag}_cvg"]) == 1, tag
        assert abs(r["n_iter"] - int(g[f"{tag}_n_iter"])) <= 2, tag
        ir = float(g[f"{tag}_initial_residual"])
        assert abs(r["initial_residual"] - ir) <= 1e-12 * ir
        res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
        assert res <= 1.01e-9 * r_norm, tag
        w = make(tag, so, deg, 4)
        rw = w.solve(m, s.rhs, 1e-12, r_norm,
                     vx=np.concatenate([0.3 * g[f"{tag}_vx"], np.zeros(m.n_cols_ext - s.n_rows)]),
                     use_vx_ini=True)
        assert (int(rw["cvg"]), rw["n_iter"]) == (int(g[f"{tag}_warm_cvg"]),
                                                  int(g[f"{tag}_warm_n_iter"])), tag
        assert rel_l2(rw["vx"][:s.n_rows], g[f"{tag}_warm_vx"]) <= 1e-9, tag
        c.destroy(); w.destroy()
    m.destroy()


def test_general_pc_tight_tolerance_solution_parity(csb, oracle):
    s = cube_poisson(32, "dirichlet")
    r_norm = float(np.sqrt(oracle.dot_xx(s.rhs)))
    A = (s.row_index, s.col_id, s.d_val, s.x_val)
    m = csb.Matrix.from_system(s)
    for so, T in ((0, csb.SlesItType.PCG), (1, csb.SlesItType.FCG), (2, csb.SlesItType.IPCG)):
        for deg in (1, 2):
            o = oracle.it_poly(so, *A, s.rhs, deg, 2000, 1e-13, r_norm)
            c = csb.SlesIt(T, 2000, poly_degree=deg)
            r = c.solve(m, s.rhs, 1e-13, r_norm)
            assert int(r["cvg"]) == 1 and o["cvg"] == 1
            assert abs(r["n_iter"] - o["n_iter"]) <= 2
            assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= SOL_TOL, (so, deg)
            c.destroy()
    omg = oracle.mg_setup(s)
    o = oracle.it_mgpc(0, omg, s.rhs, 200, 1e-13, r_norm)
    c = csb.SlesIt(csb.SlesItType.PCG, 200, poly_degree=-1)
    c.transfer_pc(csb.Multigrid(pc=True))
    r = c.solve(m, s.rhs, 1e-13, r_norm)
    assert int(r["cvg"]) == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"][:s.n_rows], o["vx"]) <= SOL_TOL
    c.destroy(); oracle.mg_free(omg); m.destroy()


# ------------------------------------------------------------ 3x3 block Jacobi (a17)

@pytest.mark.parametrize("N", [6, 10])
def test_block3_jacobi_matches_reference_golden(csb, N):
    """_block_3_jacobi on a CS_MATRIX_BLOCK_D matrix: the iterates have no reduction in them
    (Gauss-Jordan block inverses, block SpMV, D^-1.(rhs - E.x)) -> bit-exact for a fixed
    number of sweeps; the residual norm to reduction-order rounding."""
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion
    g = np.load(os.path.join(GOLD, f"block3jac_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(N)
    n = len(ri) - 1
    r_norm = float(g["r_norm"])
    m = csb.Matrix.from_block(n, 3, ri, ci, d.ravel(), xv)
    for nit in (1, 2, 7):
        c = csb.SlesIt(csb.SlesItType.JACOBI, nit)
        r = c.solve(m, rhs, 1e-12, r_norm)
        assert (int(r["cvg"]), r["n_iter"]) == (-1, nit)
        assert np.array_equal(r["vx"], g[f"sweeps{nit}_vx"]), nit
        assert abs(r["residual"] - float(g[f"sweeps{nit}_residual"])) <= 1e-13 * r["residual"]
        c.destroy()
    c = csb.SlesIt(csb.SlesItType.JACOBI, 1000)
    r = c.solve(m, rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == int(g["cvg"]) == 1
    assert abs(r["n_iter"] - int(g["n_iter"])) <= 1
    assert rel_l2(r["vx"], g["vx"]) <= 1e-7
    c.destroy()
    c = csb.SlesIt(csb.SlesItType.JACOBI, 5)
    w = c.solve(m, rhs, 1e-12, r_norm, vx=g["vx0"].copy(), use_vx_ini=True)
    assert np.array_equal(w["vx"], g["warm_vx"])
    c.destroy(); m.destroy()


# ------------------------------------------- Galerkin coarse matrices (a36, p0p1_relax = 0)

@pytest.mark.parametrize("N,bc", [(8, "neumann"), (12, "dirichlet")])
def test_galerkin_hierarchy_matches_reference_golden(csb, N, bc):
    """p0p1_relax <= 0 selects _compute_coarse_quantities_msr: no mesh association needed;
    coarse matrices bit-exact (thread per coarse row, the reference's summation order),
    V-cycle and FCG + V-cycle over that hierarchy within the usual tolerances."""
    g = np.load(os.path.join(GOLD, f"galerkin_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)     # no mesh
    mg = csb.Multigrid()
    mg.set_coarsening_options(p0p1_relax=0.)
    mg.setup(m)
    lv = mg.levels()
    assert len(lv) == int(g["n_levels"])
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], g[f"l{l}_{k}"]), (l, k)
    r = mg.solve(m, s.rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == int(g["mg_cvg"]) == 1
    assert equiv_iters_close(r["n_iter"], int(g["mg_n_iter"]), r["n_iter"] / max(r["n_cycles"], 1))
    res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
    assert res <= 1.0001e-8 * r_norm
    mg.destroy()
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    pc = csb.Multigrid(pc=True)
    pc.set_coarsening_options(p0p1_relax=0.)
    c.transfer_pc(pc)
    r = c.solve(m, s.rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == 1 and abs(r["n_iter"] - int(g["fcg_n_iter"])) <= 2
    c.destroy(); m.destroy()


def test_galerkin_vs_oracle_larger(csb, oracle):
    s = cube_poisson(40, "neumann", with_mesh=False)
    r_norm = float(np.linalg.norm(s.rhs))
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    omg = oracle.mg_setup(cube_poisson(40, "neumann"), p0p1_relax=0.)
    olv = oracle.mg_levels(omg)
    mg = csb.Multigrid()
    mg.set_coarsening_options(p0p1_relax=0.)
    mg.setup(m)
    lv = mg.levels()
    assert [x["n_rows"] for x in lv] == [x["n_rows"] for x in olv]
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], olv[l][k]), (l, k)
    o = oracle.mg_solve(omg, s.rhs, 1e-8, r_norm)
    r = mg.solve(m, s.rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_cycles"] - o["n_cycles"]) <= 1
    oracle.mg_free(omg); mg.destroy(); m.destroy()


# ------------------------------------------ pairwise aggregation (a30, CS_GRID_COARSENING_SPD_PW)

@pytest.mark.parametrize("N,bc", [(8, "neumann"), (12, "dirichlet")])
def test_pairwise_hierarchy_matches_reference_golden(csb, N, bc):
    """coarsening_type 3 with Galerkin coarse matrices, aggregation limits 2 / 4 / 8: maps,
    structures and values bit-exact against the reference; the V-cycle over each hierarchy
    converges with the reference's iteration count (+- the usual slack)."""
    g = np.load(os.path.join(GOLD, f"pairwise_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    for lim in (2, 4, 8):
        mg = csb.Multigrid()
        mg.set_coarsening_options(aggregation_limit=lim, coarsening_type=3, n_max_levels=10,
                                  p0p1_relax=0.)
        mg.setup(m)
        lv = mg.levels()
        assert len(lv) == int(g[f"lim{lim}_n_levels"])
        for l in range(1, len(lv)):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                assert np.array_equal(lv[l][k], g[f"lim{lim}_l{l}_{k}"]), (lim, l, k)
        r = mg.solve(m, s.rhs, 1e-8, r_norm)
        n_ref = int(g[f"lim{lim}_mg_n_iter"])
        assert int(r["cvg"]) == int(g[f"lim{lim}_mg_cvg"]) == 1
        assert equiv_iters_close(r["n_iter"], n_ref, r["n_iter"] / max(r["n_cycles"], 1))
        res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
        assert res <= 1.0001e-8 * r_norm
        mg.destroy()
    # the with-faces coarse values are not available with pairwise aggregation
    mg = csb.Multigrid()
    mg.set_coarsening_options(aggregation_limit=4, coarsening_type=3, p0p1_relax=0.95)
    with pytest.raises(csb.CsbError):
        mg.setup(m)
    mg.destroy(); m.destroy()


# --------------------------------- matrix coefficient build on the device

COEFF_OPTIONS = {"default": dict(), "no_dirichlet": dict(ndircp=0),
                 "theta_relax": dict(thetap=0.5, relaxp=0.7), "no_diffusion": dict(idiffp=0)}


@pytest.mark.parametrize("N", [5, 7])
def test_matrix_compute_coeffs_bit_exact_vs_reference_golden(csb, oracle, N):
    """cs_matrix_compute_coeffs (iconvp = 0) in HBM: thread per cell, the reference's per-cell
    summation order (interior faces in MSR order, boundary faces in face order) -> the MSR
    coefficients are bit-identical; the matrix is then used without any upload."""
    g = np.load(os.path.join(GOLD, f"coeffs_N{N}.npz"))
    s = cube_poisson(N, "neumann")
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(g["b_face_cells"])
    for name, kw in COEFF_OPTIONS.items():
        m.compute_coeffs(g["rovsdt"], g["i_visc"], g["b_visc"], g["cofbfp"], **kw)
        d, x = m.get_coefficients()
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name
    m.compute_coeffs(g["rovsdt"], g["i_visc"], g["b_visc"], g["cofbfp"])
    xx = np.cos(0.3 * np.arange(s.n_rows))
    assert np.array_equal(m.vector_multiply(xx),
                          oracle.spmv(s.row_index, s.col_id, g["default_d"], g["default_x"], xx))
    # solve with the rebuilt coefficients: hierarchy from the device-resident values
    rhs = s.rhs
    r_norm = float(np.linalg.norm(rhs))
    mg = csb.Multigrid()
    mg.setup(m)
    r = mg.solve(m, rhs, 1e-8, r_norm)
    assert int(r["cvg"]) == 1
    res = np.linalg.norm(rhs - oracle.spmv(s.row_index, s.col_id, g["default_d"], g["default_x"],
                                           r["vx"][:s.n_rows]))
    assert res <= 1.0001e-8 * r_norm
    mg.destroy(); m.destroy()


def test_matrix_compute_coeffs_large_device_inputs(csb, oracle):
    """128^3 with device-resident inputs (torch tensors): no host array of the matrix size is
    touched; checked against the oracle."""
    import torch
    from code_saturne_b200.mesh_cartesian import cube_boundary_face_cells
    N = 128
    s = cube_poisson(N, "neumann")
    bfc = cube_boundary_face_cells(N)
    rng = np.random.default_rng(5)
    iv, bv = 0.5 + rng.random(s.n_faces), 0.5 + rng.random(len(bfc))
    cb, ro = rng.random(len(bfc)), 1e-3 * rng.random(s.n_rows)
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    t = [torch.from_numpy(a).cuda() for a in (ro, iv, bv, cb)]
    m.compute_coeffs(t[0], t[1], t[2], t[3], ndircp=0)
    d, x = m.get_coefficients()
    d0, x0 = oracle.matrix_compute_coeffs_sym(s.row_index, s.c2f, bfc, cb, ro, iv, bv, ndircp=0)
    assert np.array_equal(d, d0) and np.array_equal(x, x0)
    m.destroy()


CONV_OPTIONS = {"conv": dict(), "conv_theta_relax": dict(ndircp=0, thetap=0.5, relaxp=0.7),
                "conv_only": dict(idiffp=0), "conv_thermal": dict(thermal=True, thetap=0.5)}


@pytest.mark.parametrize("N", [5, 7])
def test_matrix_compute_coeffs_convection_bit_exact_vs_reference_golden(csb, oracle, N):
    """iconvp = 1 (_coeffs_msr): upwind convection + diffusion, optional Cp weighting; the
    non-symmetric matrix is then solved with BiCGStab on the device-resident coefficients."""
    g = np.load(os.path.join(GOLD, f"coeffs_N{N}.npz"))
    s = cube_poisson(N, "neumann")
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(g["b_face_cells"])
    for name, kw in CONV_OPTIONS.items():
        kw = dict(kw)
        xcpp = g["xcpp"] if kw.pop("thermal", False) else None
        m.compute_coeffs(g["rovsdt"], g["i_visc"], g["b_visc"], g["cofbfp"], iconvp=1,
                         coefbp=g["coefbp"], i_massflux=g["i_massflux"],
                         b_massflux=g["b_massflux"], xcpp=xcpp, **kw)
        d, x = m.get_coefficients()
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name
    # a diagonally dominant convection-diffusion system (small fluxes, unsteady term)
    m.compute_coeffs(1.0 + g["rovsdt"], g["i_visc"], g["b_visc"], g["cofbfp"], iconvp=1,
                     coefbp=g["coefbp"], i_massflux=0.1 * g["i_massflux"],
                     b_massflux=0.0 * g["b_massflux"])
    d, x = m.get_coefficients()
    rhs = np.sin(0.2 * np.arange(s.n_rows))
    r_norm = float(np.linalg.norm(rhs))
    c = csb.SlesIt(csb.SlesItType.BICGSTAB, 500)
    r = c.solve(m, rhs, 1e-10, r_norm)
    assert int(r["cvg"]) == 1
    res = np.linalg.norm(rhs - oracle.spmv(s.row_index, s.col_id, d, x, r["vx"][:s.n_rows]))
    assert res <= 1.01e-10 * r_norm
    c.destroy(); m.destroy()


V3_OPTIONS = {"v3_sym": dict(iconvp=0), "v3_sym_nofimp": dict(iconvp=0, fimp=False, ndircp=0, relaxp=0.5),
              "v3_conv": dict(iconvp=1), "v3_conv_theta": dict(iconvp=1, ndircp=0, thetap=0.5, relaxp=0.7),
              "v3_conv_only": dict(iconvp=1, idiffp=0, fimp=False)}


@pytest.mark.parametrize("N", [5, 7])
def test_matrix_compute_coeffs_vector_bit_exact_vs_reference_golden(csb, N):
    """cs_matrix_compute_coeffs<3> on a CS_MATRIX_BLOCK_D matrix (BASELINE config 4's format):
    3x3 diagonal blocks and scalar extra-diagonal terms filled in HBM, bit-identical; the
    convection-diffusion system is then solved with BiCGStab."""
    g = np.load(os.path.join(GOLD, f"coeffs_N{N}.npz"))
    s = cube_poisson(N, "neumann")
    m = csb.Matrix.from_block(s.n_rows, 3, s.row_index, s.col_id,
                              np.tile(np.eye(3).ravel(), s.n_rows), np.zeros(s.nnz))
    m.set_mesh_association(s.n_faces, s.c2f, s.c2f_sgn, s.cell_cen, s.cell_vol, s.face_u_normal,
                           s.face_surf)
    m.set_boundary_faces(g["b_face_cells"])
    for name, kw in V3_OPTIONS.items():
        kw = dict(kw)
        fimp = g["v3_fimp"] if kw.pop("fimp", True) else None
        m.compute_coeffs_vector(g["i_visc"], g["b_visc"], g["v3_cofbfp"], g["v3_coefbp"], fimp,
                                g["i_massflux"], g["b_massflux"], **kw)
        d, x = m.get_coefficients()
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name
    # diagonally dominant velocity-like system: unsteady term on the block diagonals
    fimp = np.tile((2.0 * np.eye(3)).ravel(), s.n_rows)
    m.compute_coeffs_vector(g["i_visc"], g["b_visc"], 0.0 * g["v3_cofbfp"], 0.0 * g["v3_coefbp"],
                            fimp, 0.1 * g["i_massflux"], 0.0 * g["b_massflux"], iconvp=1)
    rhs = np.sin(0.2 * np.arange(3 * s.n_rows))
    r_norm = float(np.linalg.norm(rhs))
    c = csb.SlesIt(csb.SlesItType.BICGSTAB, 500)
    r = c.solve(m, rhs, 1e-10, r_norm)
    assert int(r["cvg"]) == 1
    res = np.linalg.norm(rhs - m.vector_multiply(r["vx"][:3 * s.n_rows].copy()))
    assert res <= 1.01e-10 * r_norm
    c.destroy(); m.destroy()
