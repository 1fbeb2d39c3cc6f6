"""CPU tests of the checker itself: the C restatement (oracle/cs_oracle.c) against
(1) the committed golden fixtures generated from the reference's own compiled
code and (2) that code directly when oracle/_ref/libcsref.so is present.
Integer/index data must match bit-exactly; floating-point data too, since the
restatement keeps the reference's evaluation order."""
import os
import sys

import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import cube_poisson, matrix_check_coefficients
from conftest import golden_cases


@pytest.mark.parametrize("N,bc,path", golden_cases())
def test_oracle_kernels_match_golden(oracle, N, bc, path):
    g = np.load(path)
    s = cube_poisson(N, bc)
    d, xv, x = matrix_check_coefficients(s)
    y = oracle.spmv(s.row_index, s.col_id, d, xv, x)
    ye = oracle.spmv(s.row_index, s.col_id, d, xv, x, True)
    assert np.array_equal(y, g["spmv_y"])
    assert np.array_equal(ye, g["spmv_y_exdiag"])
    assert oracle.dot_xx(x) == float(g["dot_xx"])
    assert oracle.dot(x, y) == float(g["dot_xy"])
    assert np.array_equal(np.array(oracle.dot_xx_xy_yz(x, y, ye)), g["dot_xx_xy_yz"])


@pytest.mark.parametrize("N,bc,path", golden_cases())
@pytest.mark.parametrize("solver,nmax", [("pcg", 1000), ("jacobi", 40), ("bicgstab", 1000)])
def test_oracle_solvers_match_golden(oracle, N, bc, path, solver, nmax):
    g = np.load(path)
    s = cube_poisson(N, bc)
    r = getattr(oracle, solver)(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nmax, 1e-8,
                                float(g["r_norm"]))
    assert r["cvg"] == int(g[f"{solver}_cvg"])
    assert r["n_iter"] == int(g[f"{solver}_n_iter"])
    assert r["residual"] == float(g[f"{solver}_residual"])
    assert r["initial_residual"] == float(g[f"{solver}_initial_residual"])
    assert np.array_equal(r["vx"], g[f"{solver}_vx"])
    # history: golden[k-1] = residual reported with n_max_iter = k
    hist = g[f"{solver}_history"]
    for k in (1, 2, len(hist) // 2, len(hist)):
        if k < 1:
            continue
        rk = getattr(oracle, solver)(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, k, 1e-8,
                                     float(g["r_norm"]))
        assert rk["residual"] == hist[k - 1]


@pytest.mark.parametrize("N,bc,path", golden_cases())
def test_oracle_pcg_warm_start_matches_golden(oracle, N, bc, path):
    g = np.load(path)
    s = cube_poisson(N, bc)
    r = oracle.pcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 1000, 1e-10,
                   float(g["r_norm"]), vx0=g["jacobi_vx"])
    assert r["n_iter"] == int(g["pcg_warm_n_iter"])
    assert np.array_equal(r["vx"], g["pcg_warm_vx"])


@pytest.mark.parametrize("N,bc,path", golden_cases())
def test_oracle_hierarchy_matches_golden(oracle, N, bc, path):
    g = np.load(path)
    s = cube_poisson(N, bc)
    mg = oracle.mg_setup(s)
    lv = oracle.mg_levels(mg)
    assert len(lv) == int(g["mg_n_levels"])
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], g[f"mg_l{l}_{k}"]), (l, k)
    r = oracle.mg_solve(mg, s.rhs, 1e-8, float(g["r_norm"]))
    assert r["cvg"] == int(g["mg_cvg"])
    assert r["n_iter"] == int(g["mg_n_iter"])
    assert r["residual"] == float(g["mg_residual"])
    assert np.array_equal(r["vx"], g["mg_vx"])
    cyc = g["mg_cycle_residuals_from3"]
    assert np.array_equal(r["cycle_residuals"][2:2 + len(cyc)], cyc[:len(r["cycle_residuals"]) - 2])
    oracle.mg_free(mg)


@pytest.mark.parametrize("N,bc,n_threads", [(12, "neumann", 1), (20, "dirichlet", 1),
                                            (24, "neumann", 4), (30, "dirichlet", 3)])
def test_oracle_vs_reference_direct(oracle, reflib, N, bc, n_threads):
    """Same comparison against the live reference build, incl. the per-thread
    aggregation blocks (cs_grid.cpp:4374-4393)."""
    reflib.set_threads(n_threads)
    try:
        s = cube_poisson(N, bc)
        h = reflib.system(s)
        lv_ref = reflib.hierarchy(h)
        mg = oracle.mg_setup(s, n_threads=n_threads)
        lv = oracle.mg_levels(mg)
        assert [g["n_rows"] for g in lv] == [g["n_rows"] for g in lv_ref]
        for l in range(1, len(lv)):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                assert np.array_equal(lv[l][k], lv_ref[l][k]), (l, k)
        if n_threads == 1:
            r_norm = float(np.sqrt(oracle.dot_xx(s.rhs)))
            a = reflib.multigrid(h, s.rhs, 1e-8, r_norm)
            b = oracle.mg_solve(mg, s.rhs, 1e-8, r_norm)
            assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
            assert np.array_equal(a["vx"], b["vx"])
            a = reflib.sles_it(h, reflib.PCG, s.rhs, 2000, 1e-9, r_norm)
            b = oracle.pcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 2000, 1e-9, r_norm)
            assert (a["n_iter"], a["residual"]) == (b["n_iter"], b["residual"])
            assert np.array_equal(a["vx"], b["vx"])
            # flexible CG (cs_sles_it.cpp:832-985), cold and warm start
            a = reflib.sles_it(h, reflib.FCG, s.rhs, 2000, 1e-9, r_norm)
            b = oracle.fcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 2000, 1e-9, r_norm)
            assert (a["cvg"], a["n_iter"], a["residual"], a["initial_residual"]) \
                == (b["cvg"], b["n_iter"], b["residual"], b["initial_residual"])
            assert np.array_equal(a["vx"], b["vx"])
            vx0 = 0.5 * a["vx"]
            a = reflib.sles_it(h, reflib.FCG, s.rhs, 7, 1e-12, r_norm, vx0=vx0)
            b = oracle.fcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 7, 1e-12, r_norm, vx0=vx0)
            assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
            assert np.array_equal(a["vx"], b["vx"])
        oracle.mg_free(mg)
        reflib.destroy(h)
    finally:
        reflib.set_threads(1)


@pytest.mark.parametrize("N,bc", [(8, "neumann"), (12, "dirichlet"), (16, "neumann")])
def test_oracle_fcg_matches_golden(oracle, N, bc):
    """Flexible CG restatement against fixtures produced by the reference build
    (tests/golden/make_golden_fcg.py): bit-exact solution, counts and residuals."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"fcg_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    o = oracle.fcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 1000, 1e-9, r_norm)
    assert (o["cvg"], o["n_iter"]) == (int(g["cvg"]), int(g["n_iter"]))
    assert o["residual"] == float(g["residual"])
    assert o["initial_residual"] == float(g["initial_residual"])
    assert np.array_equal(o["vx"], g["vx"])
    # history[0] is the initial residual, then one entry per convergence test
    assert np.array_equal(o["history"][1:1 + len(g["history"])], g["history"])
    w = oracle.fcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 5, 1e-12, r_norm,
                   vx0=0.5 * g["vx"])
    assert (w["cvg"], w["n_iter"], w["residual"]) \
        == (int(g["warm_cvg"]), int(g["warm_n_iter"]), float(g["warm_residual"]))
    assert np.array_equal(w["vx"], g["warm_vx"])


@pytest.mark.parametrize("name", ["csr_N6_sym", "csr_N10_nonsym", "csr_N14_sym"])
def test_oracle_csr_spmv_matches_golden(oracle, name):
    """_mat_vec_p_l_csr restatement against arrays and products of the reference build
    (tests/golden/make_golden_csr.py): bit-exact, diagonal included and excluded."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    y = oracle.spmv_csr(g["row_index"], g["col_id"], g["val"], g["x"])
    assert np.array_equal(y, g["y"])
    y = oracle.spmv_csr(g["row_index"], g["col_id"], g["val"], g["x"], exclude_diag=True)
    assert np.array_equal(y, g["y_exdiag"])


def test_oracle_csr_vs_reference_direct(oracle, reflib):
    sys_path = os.path.join(os.path.dirname(__file__), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(sys_path, "make_golden_csr.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    for N, symmetric in ((9, True), (12, False)):
        s = cube_poisson(N, "dirichlet")
        edges, da, xa, x = mk.coefficients(s, symmetric)
        h, ri, ci, val = reflib.system_csr(s.n_rows, edges, da, xa, symmetric)
        assert np.array_equal(reflib.spmv(h, x, s.n_rows), oracle.spmv_csr(ri, ci, val, x))
        assert np.array_equal(reflib.spmv(h, x, s.n_rows, True),
                              oracle.spmv_csr(ri, ci, val, x, exclude_diag=True))
        reflib.destroy(h)


@pytest.mark.parametrize("N", [6, 10, 14])
def test_oracle_block3_matches_golden(oracle, N):
    """3x3 block-diagonal SpMV and BiCGStab restatements against the reference build's
    outputs (tests/golden/make_golden_block3.py): bit-exact."""
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"block3_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(N)
    assert np.array_equal(oracle.spmv_b3(ri, ci, d.ravel(), xv, g["x"]), g["y"])
    assert np.array_equal(oracle.spmv_b3(ri, ci, d.ravel(), xv, g["x"], True), g["y_exdiag"])
    o = oracle.bicgstab_b3(ri, ci, d.ravel(), xv, rhs, 1000, 1e-10, float(g["r_norm"]))
    assert (o["cvg"], o["n_iter"], o["residual"]) \
        == (int(g["bicg_cvg"]), int(g["bicg_n_iter"]), float(g["bicg_residual"]))
    assert np.array_equal(o["vx"], g["bicg_vx"])


def test_oracle_block3_vs_reference_direct(oracle, reflib):
    import ctypes as C
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion
    from oracle.pyoracle import _dp
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(9, seed=7)
    n = len(ri) - 1
    h = reflib.system_block(n, 3, ri, ci, d.ravel(), xv, False)
    x = np.sin(0.3 * np.arange(3 * n))
    y = np.empty(3 * n)
    reflib.lib.csref_spmv(h, C.c_int(0), _dp(x.copy()), _dp(y))
    assert np.array_equal(y, oracle.spmv_b3(ri, ci, d.ravel(), xv, x))
    r_norm = float(np.linalg.norm(rhs))
    a = reflib.sles_it(h, reflib.BICGSTAB, rhs, 7, 1e-12, r_norm, vx0=0.1 * x)
    b = oracle.bicgstab_b3(ri, ci, d.ravel(), xv, rhs, 7, 1e-12, r_norm, vx0=0.1 * x)
    assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
    assert np.array_equal(a["vx"], b["vx"])
    reflib.destroy(h)


def test_oracle_block3_on_irregular_mesh_matches_reference(oracle, reflib):
    """3x3-block SpMV and BiCGStab (BASELINE config 4's solver) on a non-symmetric block
    system over the irregular graph against the reference run live: 72 iterations, bit for bit"""
    import ctypes as C
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import irregular_block3_system
    from oracle.pyoracle import _dp
    ri, ci, d, xv, rhs = irregular_block3_system()
    n = len(ri) - 1
    h = reflib.system_block(n, 3, ri, ci, d.ravel(), xv, False)
    x = np.sin(0.3 * np.arange(3 * n))
    y = np.empty(3 * n)
    reflib.lib.csref_spmv(h, C.c_int(0), _dp(x.copy()), _dp(y))
    assert np.array_equal(y, oracle.spmv_b3(ri, ci, d.ravel(), xv, x))
    rn = float(np.linalg.norm(rhs))
    a = reflib.sles_it(h, reflib.BICGSTAB, rhs, 200, 1e-8, rn)
    b = oracle.bicgstab_b3(ri, ci, d.ravel(), xv, rhs, 200, 1e-8, rn)
    assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
    assert a["n_iter"] > 20 and np.array_equal(a["vx"], b["vx"])
    reflib.destroy(h)


@pytest.mark.parametrize("N,bc", [(8, "neumann"), (12, "dirichlet"), (16, "neumann")])
def test_oracle_ipcg_matches_golden(oracle, N, bc):
    """_conjugate_gradient_ip restatement against fixtures produced by the reference build."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"ipcg_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    o = oracle.ipcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 1000, 1e-9, r_norm)
    assert (o["cvg"], o["n_iter"], o["residual"], o["initial_residual"]) \
        == (int(g["cvg"]), int(g["n_iter"]), float(g["residual"]), float(g["initial_residual"]))
    assert np.array_equal(o["vx"], g["vx"])
    w = oracle.ipcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 5, 1e-12, r_norm,
                    vx0=0.5 * g["vx"])
    assert (w["cvg"], w["n_iter"], w["residual"]) \
        == (int(g["warm_cvg"]), int(g["warm_n_iter"]), float(g["warm_residual"]))
    assert np.array_equal(w["vx"], g["warm_vx"])


def test_oracle_pcg_npc_vs_reference_direct(oracle, reflib):
    """_conjugate_gradient_sr_npc (poly_degree < 0) restatement against the live reference."""
    for N, bc in ((10, "neumann"), (13, "dirichlet")):
        s = cube_poisson(N, bc)
        d = s.d_val * (1.0 + 0.3 * np.cos(0.2 * np.arange(s.n_rows)))
        h = reflib.system(s, d_val=d, with_mesh=False)
        r_norm = float(np.sqrt(oracle.dot_xx(s.rhs)))
        a = reflib.sles_it(h, reflib.PCG, s.rhs, 2000, 1e-9, r_norm, poly_degree=-1)
        b = oracle.pcg_npc(s.row_index, s.col_id, d, s.x_val, s.rhs, 2000, 1e-9, r_norm)
        assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
        assert np.array_equal(a["vx"], b["vx"])
        a = reflib.sles_it(h, reflib.PCG, s.rhs, 4, 1e-12, r_norm, vx0=0.2 * a["vx"], poly_degree=-1)
        b = oracle.pcg_npc(s.row_index, s.col_id, d, s.x_val, s.rhs, 4, 1e-12, r_norm, vx0=0.2 * b["vx"])
        assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
        assert np.array_equal(a["vx"], b["vx"])
        reflib.destroy(h)


# ---- multigrid smoothers, single-cycle V-cycle (preconditioner form), FCG + multigrid pc ----

MGPC_CASES = [(8, "neumann"), (12, "dirichlet")]
SMOOTHERS = {"jacobi": 3, "l1_jacobi": 13, "r_jacobi": 14, "rj2": 15, "rj3": 16}
MGPC_CYCLES = {"rj3": ((16, 16, 0), (3, 3, 500)), "rj2": ((15, 15, 0), (2, 4, 500)),
               "r_jacobi": ((14, 14, 0), (2, 2, 500)), "l1_jacobi": ((13, 13, 0), (2, 2, 500)),
               "jacobi": ((3, 3, 0), (2, 3, 500))}


@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_oracle_smoothers_match_golden(oracle, N, bc):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"mgpc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    for name, st in SMOOTHERS.items():
        for nit in (1, 3):
            r = oracle.smoother(st, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nit)
            assert np.array_equal(r["vx"], g[f"sm_{name}_{nit}_zero"]), (name, nit)
            r = oracle.smoother(st, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nit, g["vx0"])
            assert np.array_equal(r["vx"], g[f"sm_{name}_{nit}_warm"]), (name, nit)
            assert r["n_iter"] == int(g[f"sm_{name}_{nit}_n_iter"])


@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_oracle_cycle_pc_and_fcg_mgpc_match_golden(oracle, N, bc):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"mgpc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    mg = oracle.mg_setup(s)
    for name, (types, nmi) in MGPC_CYCLES.items():
        r = oracle.mg_cycle_pc(mg, s.rhs, 1e-8, r_norm, types, nmi)
        assert np.array_equal(r["vx"], g[f"cyc_{name}_vx"]), name
        assert r["n_iter"] == int(g[f"cyc_{name}_n_iter"])
        r = oracle.fcg_mgpc(mg, s.rhs, 200, 1e-8, r_norm, types, nmi)
        assert r["cvg"] == int(g[f"fcg_{name}_cvg"]) == 1
        assert r["n_iter"] == int(g[f"fcg_{name}_n_iter"])
        assert r["residual"] == float(g[f"fcg_{name}_residual"])
        assert r["initial_residual"] == float(g[f"fcg_{name}_initial_residual"])
        assert np.array_equal(r["vx"], g[f"fcg_{name}_vx"])
        hist = g[f"fcg_{name}_history"]
        assert np.array_equal(r["history"][1:], hist)     # history[0] = initial residual
        r = oracle.mg_solve_sm(mg, s.rhs, 1e-8, r_norm, types, nmi)      # V-cycle solver mode
        assert (r["cvg"], r["n_iter"], r["residual"]) == (
            int(g[f"vsm_{name}_cvg"]), int(g[f"vsm_{name}_n_iter"]), float(g[f"vsm_{name}_residual"]))
        assert np.array_equal(r["vx"], g[f"vsm_{name}_vx"])
    oracle.mg_free(mg)


def test_oracle_mgpc_vs_reference_direct(oracle, reflib):
    """Sizes beyond the fixtures, warm starts and every sweep count, against the reference's
    own code when it is built here."""
    rng = np.random.default_rng(11)
    for N, bc in ((10, "dirichlet"), (21, "neumann")):
        s = cube_poisson(N, bc)
        h = reflib.system(s)
        r_norm = float(np.linalg.norm(s.rhs))
        vx0 = rng.standard_normal(s.n_rows)
        for st in SMOOTHERS.values():
            for nit in (1, 2, 4, 7):
                for v0 in (None, vx0):
                    a = oracle.smoother(st, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nit, v0)
                    b = reflib.smoother(h, st, s.rhs, nit, v0)
                    assert a["n_iter"] == b["n_iter"] and a["cvg"] == b["cvg"]
                    assert np.array_equal(a["vx"], b["vx"]), (N, st, nit)
        mg = oracle.mg_setup(s)
        for types, nmi in MGPC_CYCLES.values():
            for v0 in (None, vx0):
                a = oracle.mg_cycle_pc(mg, s.rhs, 1e-8, r_norm, types, nmi, vx0=v0)
                b = reflib.multigrid(h, s.rhs, 1e-8, r_norm, n_max_cycles=1, n_max_iter=nmi,
                                     types=types, precision_mult=(-1., -1., 1.), vx0=v0)
                assert np.array_equal(a["vx"], b["vx"]) and a["n_iter"] == b["n_iter"]
            a = oracle.fcg_mgpc(mg, s.rhs, 200, 1e-8, r_norm, types, nmi)
            b = reflib.fcg_mgpc(h, s.rhs, 200, 1e-8, r_norm, types, nmi)
            assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
            assert np.array_equal(a["vx"], b["vx"])
            a = oracle.fcg_mgpc(mg, s.rhs, 3, 1e-12, r_norm, types, nmi, vx0=0.5 * a["vx"])
            b = reflib.fcg_mgpc(h, s.rhs, 3, 1e-12, r_norm, types, nmi, vx0=0.5 * b["vx"])
            assert a["cvg"] == b["cvg"] == -1 and np.array_equal(a["vx"], b["vx"])
        oracle.mg_free(mg)
        reflib.destroy(h)


# ---- cs_sles_it with general preconditioners (polynomial, none, multigrid) ----

PC_TAGS = [("pcg_poly1", 0, 1), ("pcg_poly2", 0, 2), ("fcg_poly1", 1, 1), ("fcg_poly2", 1, 2),
           ("fcg_poly-1", 1, -1), ("ipcg_poly1", 2, 1), ("ipcg_poly2", 2, 2)]


@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_oracle_general_pc_matches_golden(oracle, N, bc):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"pc_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    A = (s.row_index, s.col_id, s.d_val, s.x_val)
    mg = oracle.mg_setup(s)
    runs = [(tag, lambda nmax, eps, v0, so=so, deg=deg:
             oracle.it_poly(so, *A, s.rhs, deg, nmax, eps, r_norm, vx0=v0))
            for tag, so, deg in PC_TAGS]
    runs.append(("pcg_mgpc", lambda nmax, eps, v0: oracle.it_mgpc(0, mg, s.rhs, nmax, eps, r_norm,
                                                                    vx0=v0)))
    runs.append(("ipcg_mgpc", lambda nmax, eps, v0: oracle.it_mgpc(2, mg, s.rhs, nmax, eps, r_norm,
                                                                     vx0=v0)))
    for tag, solve in runs:
        r = solve(500, 1e-9, None)
        assert r["cvg"] == int(g[f"{tag}_cvg"]) and r["n_iter"] == int(g[f"{tag}_n_iter"]), tag
        assert r["residual"] == float(g[f"{tag}_residual"])
        assert r["initial_residual"] == float(g[f"{tag}_initial_residual"])
        assert np.array_equal(r["vx"], g[f"{tag}_vx"]), tag
        w = solve(4, 1e-12, 0.3 * r["vx"])
        assert (w["cvg"], w["n_iter"]) == (int(g[f"{tag}_warm_cvg"]), int(g[f"{tag}_warm_n_iter"]))
        assert np.array_equal(w["vx"], g[f"{tag}_warm_vx"]), tag
    oracle.mg_free(mg)


def test_oracle_general_pc_vs_reference_direct(oracle, reflib):
    s = cube_poisson(17, "neumann")
    h = reflib.system(s)
    r_norm = float(np.linalg.norm(s.rhs))
    A = (s.row_index, s.col_id, s.d_val, s.x_val)
    for tag, so, deg in PC_TAGS + [("pcg_poly3", 0, 3)]:
        a = oracle.it_poly(so, *A, s.rhs, deg, 500, 1e-10, r_norm)
        b = reflib.sles_it(h, so, s.rhs, 500, 1e-10, r_norm, poly_degree=deg)
        assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"]), tag
        assert np.array_equal(a["vx"], b["vx"]), tag
    mg = oracle.mg_setup(s)
    for types, nmi in ((16, 16, 0), (3, 3, 500)), ((13, 13, 0), (2, 2, 500)):
        a = oracle.it_mgpc(0, mg, s.rhs, 200, 1e-10, r_norm, types, nmi)
        b = reflib.it_mgpc(h, 0, s.rhs, 200, 1e-10, r_norm, types, nmi)
        assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
        assert np.array_equal(a["vx"], b["vx"])
    oracle.mg_free(mg)
    reflib.destroy(h)


@pytest.mark.parametrize("N", [6, 10])
def test_oracle_block3_jacobi_matches_golden(oracle, N):
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"block3jac_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(N)
    r_norm = float(g["r_norm"])
    r = oracle.block3_jacobi(ri, ci, d.ravel(), xv, rhs, 1000, 1e-8, r_norm)
    assert (r["cvg"], r["n_iter"]) == (int(g["cvg"]), int(g["n_iter"]))
    assert r["residual"] == float(g["residual"])
    assert r["initial_residual"] == float(g["initial_residual"])
    assert np.array_equal(r["vx"], g["vx"])
    for nit in (1, 2, 7):
        f = oracle.block3_jacobi(ri, ci, d.ravel(), xv, rhs, nit, 1e-12, r_norm)
        assert np.array_equal(f["vx"], g[f"sweeps{nit}_vx"])
        assert f["residual"] == float(g[f"sweeps{nit}_residual"])
    w = oracle.block3_jacobi(ri, ci, d.ravel(), xv, rhs, 5, 1e-12, r_norm, vx0=g["vx0"])
    assert np.array_equal(w["vx"], g["warm_vx"]) and w["residual"] == float(g["warm_residual"])


@pytest.mark.parametrize("N,bc", [(8, "neumann"), (12, "dirichlet")])
def test_oracle_galerkin_hierarchy_matches_golden(oracle, N, bc):
    """p0p1_relax = 0: Galerkin coarse matrices (_compute_coarse_quantities_msr)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"galerkin_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    mg = oracle.mg_setup(s, p0p1_relax=0.)
    lv = oracle.mg_levels(mg)
    assert len(lv) == int(g["n_levels"])
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], g[f"l{l}_{k}"]), (l, k)
    r = oracle.mg_solve(mg, s.rhs, 1e-8, r_norm)
    assert (r["cvg"], r["n_iter"], r["residual"]) == (int(g["mg_cvg"]), int(g["mg_n_iter"]),
                                                      float(g["mg_residual"]))
    assert np.array_equal(r["vx"], g["mg_vx"])
    r = oracle.fcg_mgpc(mg, s.rhs, 200, 1e-8, r_norm)
    assert (r["cvg"], r["n_iter"], r["residual"]) == (int(g["fcg_cvg"]), int(g["fcg_n_iter"]),
                                                      float(g["fcg_residual"]))
    assert np.array_equal(r["vx"], g["fcg_vx"])
    oracle.mg_free(mg)


PW_CASES = [(8, "neumann"), (12, "dirichlet")]


@pytest.mark.parametrize("N,bc", PW_CASES)
def test_oracle_pairwise_hierarchy_matches_golden(oracle, N, bc):
    """CS_GRID_COARSENING_SPD_PW: pairs, pairs of pairs (limit 4), and of those (limit 8)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"pairwise_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    for lim in (2, 4, 8):
        mg = oracle.mg_setup_pw(s, aggregation_limit=lim, n_levels_max=10)
        lv = oracle.mg_levels(mg)
        assert len(lv) == int(g[f"lim{lim}_n_levels"])
        for l in range(1, len(lv)):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                assert np.array_equal(lv[l][k], g[f"lim{lim}_l{l}_{k}"]), (lim, l, k)
        r = oracle.mg_solve(mg, s.rhs, 1e-8, float(g["r_norm"]))
        assert (r["cvg"], r["n_iter"], r["residual"]) == (
            int(g[f"lim{lim}_mg_cvg"]), int(g[f"lim{lim}_mg_n_iter"]), float(g[f"lim{lim}_mg_residual"]))
        oracle.mg_free(mg)


COEFF_OPTIONS = {"default": dict(), "no_dirichlet": dict(ndircp=0),
                 "theta_relax": dict(thetap=0.5, relaxp=0.7), "no_diffusion": dict(idiffp=0)}


@pytest.mark.parametrize("N", [5, 7])
def test_oracle_matrix_compute_coeffs_matches_golden(oracle, N):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"coeffs_N{N}.npz"))
    s = cube_poisson(N, "neumann")
    for name, kw in COEFF_OPTIONS.items():
        d, x = oracle.matrix_compute_coeffs_sym(s.row_index, s.c2f, g["b_face_cells"], g["cofbfp"],
                                                g["rovsdt"], g["i_visc"], g["b_visc"], **kw)
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name


def test_oracle_matrix_compute_coeffs_vs_reference_direct(oracle, reflib):
    from code_saturne_b200.mesh_cartesian import cube_boundary_face_cells
    rng = np.random.default_rng(3)
    s = cube_poisson(13, "neumann")
    h = reflib.system(s)
    bfc = cube_boundary_face_cells(13)
    iv, bv = 0.1 + rng.random(s.n_faces), 0.1 + rng.random(len(bfc))
    cb, ro = rng.random(len(bfc)), rng.random(s.n_rows)
    for kw in COEFF_OPTIONS.values():
        a = oracle.matrix_compute_coeffs_sym(s.row_index, s.c2f, bfc, cb, ro, iv, bv, **kw)
        b = reflib.matrix_compute_coeffs_sym(h, s, bfc, cb, ro, iv, bv, **kw)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    reflib.destroy(h)


CONV_OPTIONS = {"conv": dict(), "conv_theta_relax": dict(ndircp=0, thetap=0.5, relaxp=0.7),
                "conv_only": dict(idiffp=0), "conv_thermal": dict(thermal=True, thetap=0.5)}


@pytest.mark.parametrize("N", [5, 7])
def test_oracle_matrix_compute_coeffs_convection_matches_golden(oracle, N):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"coeffs_N{N}.npz"))
    s = cube_poisson(N, "neumann")
    for name, kw in CONV_OPTIONS.items():
        kw = dict(kw)
        xcpp = g["xcpp"] if kw.pop("thermal", False) else None
        d, x = oracle.matrix_compute_coeffs(s, g["b_face_cells"], g["coefbp"], g["cofbfp"],
                                            g["rovsdt"], g["i_massflux"], g["b_massflux"],
                                            g["i_visc"], g["b_visc"], xcpp=xcpp, **kw)
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name


@pytest.mark.parametrize("N,bc", [(8, "dirichlet"), (12, "dirichlet")])
def test_oracle_k_cycle_matches_golden(oracle, N, bc):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"kcycle_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    mg = oracle.mg_setup_pw(s, aggregation_limit=4, n_levels_max=10, n_g_rows_min=30)
    r = oracle.mg_k_cycle(mg, s.rhs, 1e-8, r_norm)
    assert np.array_equal(r["vx"], g["cycle_vx"]) and r["n_iter"] == int(g["cycle_n_iter"])
    r = oracle.it_kpc(1, mg, s.rhs, 200, 1e-8, r_norm)
    assert (r["cvg"], r["n_iter"], r["residual"]) == (int(g["fcg_cvg"]), int(g["fcg_n_iter"]),
                                                      float(g["fcg_residual"]))
    assert np.array_equal(r["vx"], g["fcg_vx"])
    oracle.mg_free(mg)


def test_oracle_k_cycle_vs_reference_direct(oracle, reflib):
    """deeper hierarchies (both inner Krylov branches taken), pairs only, Neumann problem"""
    for N, bc, lim in ((16, "neumann", 4), (14, "dirichlet", 2)):
        s = cube_poisson(N, bc)
        h = reflib.system(s)
        r_norm = float(np.linalg.norm(s.rhs))
        kw = dict(aggregation_limit=lim, coarsening_type=3, n_levels_max=10, n_g_rows_min=30,
                  p0p1_relax=0., mg_type=1)
        mg = oracle.mg_setup_pw(s, aggregation_limit=lim, n_levels_max=10, n_g_rows_min=30)
        a = oracle.mg_k_cycle(mg, s.rhs, 1e-8, r_norm)
        b = reflib.multigrid(h, s.rhs, 1e-8, r_norm, n_max_cycles=1, n_max_iter=(2, 2, 500),
                             types=(3, 3, 0), precision_mult=(-1., -1., 10.), **kw)
        assert np.array_equal(a["vx"], b["vx"]) and a["n_iter"] == b["n_iter"]
        a = oracle.it_kpc(1, mg, s.rhs, 60, 1e-8, r_norm)
        b = reflib.it_mgpc(h, 1, s.rhs, 60, 1e-8, r_norm, types=(3, 3, 0),
                           mg_n_max_iter=(2, 2, 500), precision_mult=(-1., -1., 10.), **kw)
        assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
        assert np.array_equal(a["vx"], b["vx"])
        oracle.mg_free(mg)
        reflib.destroy(h)


V3_OPTIONS = {"v3_sym": dict(iconvp=0), "v3_sym_nofimp": dict(iconvp=0, fimp=False, ndircp=0, relaxp=0.5),
              "v3_conv": dict(iconvp=1), "v3_conv_theta": dict(iconvp=1, ndircp=0, thetap=0.5, relaxp=0.7),
              "v3_conv_only": dict(iconvp=1, idiffp=0, fimp=False)}


@pytest.mark.parametrize("N", [5, 7])
def test_oracle_matrix_compute_coeffs_vector_matches_golden(oracle, N):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"coeffs_N{N}.npz"))
    s = cube_poisson(N, "neumann")
    for name, kw in V3_OPTIONS.items():
        kw = dict(kw)
        fimp = g["v3_fimp"] if kw.pop("fimp", True) else None
        d, x = oracle.matrix_compute_coeffs_v3(s, g["b_face_cells"], g["v3_coefbp"], g["v3_cofbfp"],
                                               fimp, g["i_massflux"], g["b_massflux"], g["i_visc"],
                                               g["b_visc"], **kw)
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name


# ---- process-local Gauss-Seidel family (a18, GS part of a26) ----

GS_SMOOTHERS = {"p_gs": 8, "p_sym_gs": 9, "ts_f_gs": 17, "ts_b_gs": 18}


@pytest.mark.parametrize("N,bc", MGPC_CASES)
def test_oracle_gauss_seidel_matches_golden(oracle, N, bc):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"gs_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    r_norm = float(g["r_norm"])
    for name, st in GS_SMOOTHERS.items():
        for nit in (1, 3):
            r = oracle.smoother(st, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nit)
            assert np.array_equal(r["vx"], g[f"sm_{name}_{nit}_zero"]), (name, nit)
            r = oracle.smoother(st, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nit, g["vx0"])
            assert np.array_equal(r["vx"], g[f"sm_{name}_{nit}_warm"]), (name, nit)
            assert r["n_iter"] == int(g[f"sm_{name}_{nit}_n_iter"])
    for name, sym in (("p_gs", False), ("p_sym_gs", True)):
        for tag, v0 in (("zero", None), ("warm", g["vx0"])):
            r = oracle.gauss_seidel(sym, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 400,
                                    1e-6, r_norm, v0)
            for k in ("cvg", "n_iter", "residual", "initial_residual"):
                assert r[k] == g[f"it_{name}_{tag}_{k}"], (name, tag, k)
            assert np.array_equal(r["vx"], g[f"it_{name}_{tag}_vx"])
    mg = oracle.mg_setup(s)
    types, nmi = (9, 9, 0), (1, 1, 500)
    r = oracle.mg_cycle_pc(mg, s.rhs, 1e-8, r_norm, types, nmi)
    assert np.array_equal(r["vx"], g["cyc_sgs_vx"]) and r["n_iter"] == int(g["cyc_sgs_n_iter"])
    r = oracle.fcg_mgpc(mg, s.rhs, 200, 1e-8, r_norm, types, nmi)
    for k in ("cvg", "n_iter", "residual", "initial_residual"):
        assert r[k] == g[f"fcg_sgs_{k}"], k
    assert np.array_equal(r["vx"], g["fcg_sgs_vx"])
    oracle.mg_free(mg)


def test_oracle_gauss_seidel_vs_reference_direct(oracle, reflib):
    rng = np.random.default_rng(5)
    for N, bc in ((10, "dirichlet"), (21, "neumann")):
        s = cube_poisson(N, bc)
        h = reflib.system(s)
        r_norm = float(np.linalg.norm(s.rhs))
        vx0 = rng.standard_normal(s.n_rows)
        for st in GS_SMOOTHERS.values():
            for nit in (1, 2, 4):
                for v0 in (None, vx0):
                    a = oracle.smoother(st, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, nit, v0)
                    b = reflib.smoother(h, st, s.rhs, nit, v0)
                    assert a["n_iter"] == b["n_iter"] and a["cvg"] == b["cvg"]
                    assert np.array_equal(a["vx"], b["vx"]), (N, st, nit)
        for sym, st in ((False, 8), (True, 9)):
            for v0 in (None, vx0):
                a = oracle.gauss_seidel(sym, s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 150,
                                        1e-6, r_norm, v0)
                b = reflib.sles_it(h, st, s.rhs, 150, 1e-6, r_norm, v0)
                assert (a["cvg"], a["n_iter"], a["residual"], a["initial_residual"]) == (
                    b["cvg"], b["n_iter"], b["residual"], b["initial_residual"])
                assert np.array_equal(a["vx"], b["vx"])
        reflib.destroy(h)


# ---- block-diagonal matrices of any block size (a4 _6 / _generic, a17 _block_jacobi) ----

BLOCKN_CASES = [(6, 6), (2, 7), (4, 5), (9, 4)]       # (db, N)


@pytest.mark.parametrize("db,N", BLOCKN_CASES)
def test_oracle_blockn_matches_golden(oracle, db, N):
    from code_saturne_b200.mesh_cartesian import cube_block_convection_diffusion
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"blockn_db{db}_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db)
    r_norm = float(g["r_norm"])
    assert np.array_equal(oracle.spmv_bn(db, ri, ci, d.ravel(), xv, g["x"]), g["y"])
    assert np.array_equal(oracle.spmv_bn(db, ri, ci, d.ravel(), xv, g["x"], True), g["y_exdiag"])
    for nit in (1, 2, 4):
        f = oracle.block_jacobi(db, ri, ci, d.ravel(), xv, rhs, nit, 1e-12, r_norm)
        assert np.array_equal(f["vx"], g[f"jac{nit}_vx"]), nit
        assert f["residual"] == float(g[f"jac{nit}_residual"])
    w = oracle.block_jacobi(db, ri, ci, d.ravel(), xv, rhs, 3, 1e-12, r_norm, vx0=g["x"])
    assert np.array_equal(w["vx"], g["jac_warm_vx"]) and w["residual"] == float(g["jac_warm_residual"])
    r = oracle.block_jacobi(db, ri, ci, d.ravel(), xv, rhs, 500, 1e-8, r_norm)
    assert (r["cvg"], r["n_iter"], r["residual"]) == (int(g["jac_cvg"]), int(g["jac_n_iter"]),
                                                      float(g["jac_residual"]))
    b = oracle.bicgstab_bn(db, ri, ci, d.ravel(), xv, rhs, 500, 1e-10, r_norm)
    assert (b["cvg"], b["n_iter"], b["residual"], b["initial_residual"]) == (
        int(g["bicg_cvg"]), int(g["bicg_n_iter"]), float(g["bicg_residual"]),
        float(g["bicg_initial_residual"]))
    assert np.array_equal(b["vx"], g["bicg_vx"])


def test_oracle_blockn_vs_reference_direct(oracle, reflib):
    from code_saturne_b200.mesh_cartesian import cube_block_convection_diffusion
    for db, N in ((5, 6), (6, 8), (8, 5)):
        ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db, seed=77)
        n = len(ri) - 1
        h = reflib.system_block(n, db, ri, ci, d.ravel(), xv, False)
        x = np.sin(0.21 * np.arange(db * n))
        assert np.array_equal(oracle.spmv_bn(db, ri, ci, d.ravel(), xv, x), reflib.spmv(h, x, db * n))
        r_norm = float(np.linalg.norm(rhs))
        for nit in (1, 3):
            a = oracle.block_jacobi(db, ri, ci, d.ravel(), xv, rhs, nit, 1e-12, r_norm)
            b = reflib.sles_it(h, reflib.JACOBI, rhs, nit, 1e-12, r_norm)
            assert np.array_equal(a["vx"], b["vx"]) and a["residual"] == b["residual"]
        a = oracle.bicgstab_bn(db, ri, ci, d.ravel(), xv, rhs, 300, 1e-9, r_norm)
        b = reflib.sles_it(h, reflib.BICGSTAB, rhs, 300, 1e-9, r_norm)
        assert (a["cvg"], a["n_iter"], a["residual"]) == (b["cvg"], b["n_iter"], b["residual"])
        assert np.array_equal(a["vx"], b["vx"])
        reflib.destroy(h)


# ---- GCR (the fallback solver on a device), a22 ----

def test_oracle_gcr_matches_golden(oracle):
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "gcr.npz"))
    s = cube_poisson(12, "dirichlet")
    A = (1, s.row_index, s.col_id, s.d_val, s.x_val)
    rn = float(g["s_r_norm"])
    for tag, jac, v0, nmi, prec in (("jac_zero", True, None, 200, 1e-10),
                                    ("jac_warm", True, g["s_vx0"], 200, 1e-10),
                                    ("npc_warm", False, g["s_vx0"], 200, 1e-10),
                                    ("jac_27", True, g["s_vx0"], 27, 1e-14)):
        r = oracle.gcr(*A, s.rhs, nmi, prec, rn, v0, jacobi=jac)
        for k in ("cvg", "n_iter", "residual", "initial_residual"):
            assert r[k] == g[f"s_{tag}_{k}"], (tag, k)
        assert np.array_equal(r["vx"], g[f"s_{tag}_vx"]), tag
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(8)
    r = oracle.gcr(3, ri, ci, d.ravel(), xv, rhs, 200, 1e-10, float(g["b3_r_norm"]))
    assert (r["cvg"], r["n_iter"], r["residual"]) == (int(g["b3_gcr_cvg"]), int(g["b3_gcr_n_iter"]),
                                                      float(g["b3_gcr_residual"]))
    assert np.array_equal(r["vx"], g["b3_gcr_vx"])
    # the fallback sequence: BiCGStab stopped after 3 iterations, GCR from its iterate
    b = oracle.bicgstab_bn(3, ri, ci, d.ravel(), xv, rhs, 3, 1e-10, float(g["b3_r_norm"]))
    assert (b["cvg"], b["n_iter"]) == (int(g["b3_fb_first_cvg"]), int(g["b3_fb_first_n_iter"]))
    f = oracle.gcr(3, ri, ci, d.ravel(), xv, rhs, 200, 1e-10, float(g["b3_r_norm"]), vx0=b["vx"])
    assert (f["cvg"], f["n_iter"]) == (int(g["b3_fb200_cvg"]), int(g["b3_fb200_n_iter"]))
    assert np.array_equal(f["vx"], g["b3_fb200_vx"])


@pytest.mark.parametrize("db,N", BLOCKN_CASES + [(3, 6)])
def test_oracle_block_smoothers_match_golden(oracle, db, N):
    """Block forms of the multigrid smoothers (a26): block Jacobi, block relaxed Jacobi."""
    from code_saturne_b200.mesh_cartesian import cube_block_convection_diffusion
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"blockn_db{db}_N{N}.npz"))
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(N, db)
    for st in (3, 14, 15, 16):
        for nit in (1, 3):
            r = oracle.smoother_block(st, db, ri, ci, d.ravel(), xv, rhs, nit)
            assert np.array_equal(r["vx"], g[f"sm{st}_{nit}_zero"]), (st, nit)
            r = oracle.smoother_block(st, db, ri, ci, d.ravel(), xv, rhs, nit, g["x"])
            assert np.array_equal(r["vx"], g[f"sm{st}_{nit}_warm"]), (st, nit)
            assert r["n_iter"] == int(g[f"sm{st}_{nit}_n_iter"])


# ----------------- matrix mapped from native (face-based) arrays: a29, a35 (cs_oracle_native.c)

def _native_cases():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_native import CASES
    return sorted(CASES.items())


@pytest.mark.parametrize("name,case", _native_cases())
def test_oracle_native_hierarchy_matches_golden(oracle, name, case):
    """_automatic_aggregation_fc / _coarsen_face_cell / _compute_coarse_quantities_native /
    _matrix_from_native restated: every level bit-exact against the reference's own run, and
    the V-cycle over that hierarchy reproduces its iterates."""
    from make_golden_native import native_system
    N, bc, nt, perturb, relax = case
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"native_{name}.npz"))
    s, da, xa = native_system(N, bc, perturb)
    assert np.array_equal(da, g["da"]) and np.array_equal(xa, g["xa"])
    mg = oracle.mg_setup_native(s, n_threads=nt, da=da, xa=xa, p0p1_relax=relax)
    lv = oracle.mg_levels(mg)
    assert len(lv) == int(g["n_levels"])
    for k in ("row_index", "col_id", "x_val"):
        assert np.array_equal(lv[0][k], g[f"l0_{k}"]), k
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], g[f"l{l}_{k}"]), (l, k)
    r = oracle.mg_solve(mg, s.rhs, 1e-8, float(g["r_norm"]))
    assert (r["cvg"], r["n_iter"]) == (int(g["mg_cvg"]), int(g["mg_n_iter"]))
    if nt == 1:       # the dot products of a several-thread reference run sum in another order
        assert r["residual"] == float(g["mg_residual"])
        assert np.array_equal(r["vx"], g["mg_vx"])
    else:
        assert np.allclose(r["vx"], g["mg_vx"], rtol=0, atol=1e-12 * np.abs(g["mg_vx"]).max())
    oracle.mg_free(mg)


@pytest.mark.parametrize("N,bc,nt", [(20, "neumann", 1), (24, "dirichlet", 1), (20, "dirichlet", 4)])
def test_oracle_native_hierarchy_matches_reference(oracle, N, bc, nt):
    """the same against the reference library run here (larger cubes, perturbed faces)"""
    from oracle.pyoracle import RefLib, ref_available
    from make_golden_native import native_system
    if not ref_available():
        pytest.skip("oracle/_ref/libcsref.so not built (needs /root/reference)")
    s, da, xa = native_system(N, bc, True)
    ref = RefLib(nt)
    h = ref.system_native(s, da, xa)
    lr = ref.hierarchy(h)
    ref.destroy(h)
    ref.set_threads(1)
    mg = oracle.mg_setup_native(s, n_threads=nt, da=da, xa=xa)
    lo = oracle.mg_levels(mg)
    assert [x["n_rows"] for x in lo] == [x["n_rows"] for x in lr]
    for a, b in zip(lo, lr):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val", "cell_cen", "cell_vol"):
            assert np.array_equal(a[k], b[k]), (a["n_rows"], k)
    oracle.mg_free(mg)


def test_oracle_native_hierarchy_on_irregular_mesh_matches_reference(oracle):
    """the face-array entry on the irregular system (faces in arbitrary orientation and order):
    face-list aggregation, first-seen coarse faces, face-order coarse values -- the oracle's
    hierarchy against the reference run live, bit for bit"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import irregular_system
    from oracle.pyoracle import RefLib, native_arrays, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    s, _, _, _ = irregular_system(n=3000, k=4)
    da, xa = native_arrays(s)
    ref = RefLib(1)
    h = ref.system_native(s, da, xa)
    lr = ref.hierarchy(h)
    ref.destroy(h)
    mg = oracle.mg_setup_native(s, n_threads=1, da=da, xa=xa)
    lo = oracle.mg_levels(mg)
    assert [x["n_rows"] for x in lo] == [x["n_rows"] for x in lr] and len(lo) >= 4
    for a, b in zip(lo, lr):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val", "cell_cen", "cell_vol"):
            assert np.array_equal(a[k], b[k]), (a["n_rows"], k)
    oracle.mg_free(mg)


def test_oracle_k_cycle_and_coefficient_build_on_irregular_mesh_match_reference(oracle):
    """pairwise aggregation (limits 2 and 4: 8 and 4 levels), the K-cycle and FCG + K-cycle,
    and cs_matrix_compute_coeffs (symmetric and convection forms) of the oracle on the
    irregular system against the reference run live, bit for bit"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import irregular_system
    from oracle.pyoracle import RefLib, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ref = RefLib(1)
    s, bfc, _, _ = irregular_system(n=3000, k=4)
    rn = float(np.linalg.norm(s.rhs))
    h = ref.system(s)
    for lim in (2, 4):
        kw = dict(aggregation_limit=lim, coarsening_type=3, n_levels_max=10, n_g_rows_min=30,
                  p0p1_relax=0., mg_type=1)
        mg = oracle.mg_setup_pw(s, aggregation_limit=lim, n_levels_max=10, n_g_rows_min=30)
        a = oracle.mg_k_cycle(mg, s.rhs, 1e-8, rn)
        b = ref.multigrid(h, s.rhs, 1e-8, rn, n_max_cycles=1, n_max_iter=(2, 2, 500),
                          types=(3, 3, 0), precision_mult=(-1., -1., 10.), **kw)
        assert np.array_equal(a["vx"], b["vx"]) and a["n_iter"] == b["n_iter"], lim
        a = oracle.it_kpc(1, mg, s.rhs, 60, 1e-8, rn)
        b = ref.it_mgpc(h, 1, s.rhs, 60, 1e-8, rn, types=(3, 3, 0), mg_n_max_iter=(2, 2, 500),
                        precision_mult=(-1., -1., 10.), **kw)
        assert (a["cvg"], a["n_iter"]) == (b["cvg"], b["n_iter"]) == (1, a["n_iter"]), lim
        assert np.array_equal(a["vx"], b["vx"]), lim
        oracle.mg_free(mg)
    rng = np.random.default_rng(3)
    iv, bv = 0.1 + rng.random(s.n_faces), 0.1 + rng.random(len(bfc))
    cb, cf, ro = rng.random(len(bfc)), rng.random(len(bfc)), rng.random(s.n_rows)
    for kw in COEFF_OPTIONS.values():
        a = oracle.matrix_compute_coeffs_sym(s.row_index, s.c2f, bfc, cf, ro, iv, bv, **kw)
        b = ref.matrix_compute_coeffs_sym(h, s, bfc, cf, ro, iv, bv, **kw)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), kw
    ref.destroy(h)
    h = ref.system(s, symmetric=False)
    imf, bmf = rng.standard_normal(s.n_faces), rng.standard_normal(len(bfc))
    for kw in (dict(), dict(ndircp=0, thetap=0.5, relaxp=0.7), dict(idiffp=0)):
        a = oracle.matrix_compute_coeffs(s, bfc, cb, cf, ro, imf, bmf, iv, bv, **kw)
        b = ref.matrix_compute_coeffs(h, s, bfc, cb, cf, ro, imf, bmf, iv, bv, **kw)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), kw
    ref.destroy(h)


def test_oracle_aggregation_fc_properties(oracle):
    """face-list aggregation: every row mapped, aggregates within the limit, penalised rows
    (diagonal above 1e4 x the extra-diagonal mass) left out at level 0"""
    from oracle.pyoracle import native_arrays
    s = cube_poisson(9, "neumann")
    da, xa = native_arrays(s)
    da = da.copy()
    da[[5, 77]] = 1e6
    nc, f_c = oracle.aggregation_fc(0, s.n_rows, s.face_cells, da, xa, 3)
    assert f_c[5] == -1 and f_c[77] == -1
    kept = f_c[f_c >= 0]
    assert kept.size == s.n_rows - 2 and kept.max() == nc - 1
    assert np.bincount(kept).max() <= 4 and np.bincount(kept).min() >= 1
    nc1, f_c1 = oracle.aggregation_fc(1, s.n_rows, s.face_cells, da, xa, 3)
    assert (f_c1 >= 0).all()          # no penalisation below level 0
    # no faces at all: at level 0 every row is penalised (no extra-diagonal mass), below
    # level 0 every row becomes a coarse row of its own
    nc0, f_c0 = oracle.aggregation_fc(0, 4, np.zeros((0, 2), dtype=np.int32), np.ones(4),
                                      np.zeros(0), 3)
    assert nc0 == 0 and (f_c0 == -1).all()
    nc0, f_c0 = oracle.aggregation_fc(1, 4, np.zeros((0, 2), dtype=np.int32), np.ones(4),
                                      np.zeros(0), 3)
    assert nc0 == 4 and np.array_equal(f_c0, np.arange(4))


# ------------------------------ matrix assembler, one rank, MSR (f4; cs_oracle_assembler.c)

def _assembler_cases():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_assembler import CASES
    return sorted(CASES.items())


@pytest.mark.parametrize("N,case", _assembler_cases())
def test_oracle_assembler_matches_golden(oracle, N, case):
    """structure (sorted unique columns, diagonal apart) and values (every entry the sum of its
    coefficients in arrival order) bit-exact against the reference's assembler"""
    from make_golden_assembler import assembler_input
    l0, seed = case
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"assembler_N{N}.npz"))
    n, rows, cols, vals, batches = assembler_input(N, l0, seed)
    ri, ci, d, x = oracle.assembler(n, rows, cols, batches, l0)
    for a, k in ((ri, "row_index"), (ci, "col_id"), (d, "d_val"), (x, "x_val")):
        assert np.array_equal(a, g[k]), k
    assert np.array_equal(oracle.spmv(ri, ci, d, x, np.sin(np.arange(n) * 0.37)), g["y"])
    ri, ci, d, x = oracle.assembler(n, rows, cols, [(rows, cols, vals)], l0)
    assert np.array_equal(d, g["d_val_declared"]) and np.array_equal(x, g["x_val_declared"])


def test_oracle_assembler_matches_reference_live(oracle):
    from oracle.pyoracle import RefLib, ref_available
    from make_golden_assembler import assembler_input
    if not ref_available():
        pytest.skip("oracle/_ref/libcsref.so not built (needs /root/reference)")
    ref = RefLib(1)
    n, rows, cols, vals, batches = assembler_input(18, 77, 9)
    h, ri, ci, d, x = ref.assembler(n, rows, cols, batches, 77)
    ref.destroy(h)
    ri2, ci2, d2, x2 = oracle.assembler(n, rows, cols, batches, 77)
    assert all(np.array_equal(a, b) for a, b in ((ri, ri2), (ci, ci2), (d, d2), (x, x2)))


# ------------------- cs_diffusion_potential without reconstruction (f2; cs_oracle_rhs.c)

@pytest.mark.parametrize("N", [5, 9])
def test_oracle_diffusion_potential_matches_golden(oracle, N):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import rhs_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_N{N}.npz"))
    s, bfc, _, _, f = rhs_input(N)
    for init, inc in ((1, 1), (0, 1), (1, 0)):
        d = oracle.diffusion_potential(s, bfc, f["pvar"], f["cofaf"], f["cofbf"], f["i_visc"],
                                       f["b_visc"], init=init, inc=inc, diverg=f["div0"])
        assert np.array_equal(d, g[f"div_init{init}_inc{inc}"]), (init, inc)


def test_oracle_diffusion_potential_matches_reference_live(oracle):
    from oracle.pyoracle import RefLib, ref_available
    from make_golden_rhs import rhs_input
    if not ref_available():
        pytest.skip("oracle/_ref/libcsref.so not built (needs /root/reference)")
    ref = RefLib(4)          # the face loops follow the mesh numbering, not the thread count
    s, bfc, b_geo, ifc, f = rhs_input(14, seed=11)
    h = ref.system(s)
    r = ref.diffusion_potential(h, s, bfc, b_geo, ifc, f["pvar"], f["cofaf"], f["cofbf"],
                                f["i_visc"], f["b_visc"])
    ref.destroy(h)
    ref.set_threads(1)
    o = oracle.diffusion_potential(s, bfc, f["pvar"], f["cofaf"], f["cofbf"], f["i_visc"],
                                   f["b_visc"])
    assert np.array_equal(r, o)
    # a Poisson matrix applied to pvar is this operator with the matrix's own face values
    xa = np.zeros(s.n_faces)
    up = np.asarray(s.c2f_sgn) > 0
    xa[np.asarray(s.c2f)[up]] = np.asarray(s.x_val)[up]
    zb = np.zeros(len(bfc))
    d = oracle.diffusion_potential(s, bfc, f["pvar"], zb, zb, -xa, zb)
    y = oracle.spmv(s.row_index, s.col_id, s.d_val * 0., s.x_val, f["pvar"])
    rowsum = np.add.reduceat(np.append(s.x_val, 0.), s.row_index[:-1])
    assert np.allclose(d, y - rowsum * f["pvar"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("N", [5, 9])
def test_oracle_gradient_lsq_and_reconstructed_flux_match_golden(oracle, N):
    """least-squares gradient (float cocg, float Gauss-Jordan inverse) and the reconstructed
    branch of cs_diffusion_potential, both formulas: bit-exact against the reference"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import rc_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_rc_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    grad = oracle.gradient_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], f["pvar"], f["coefa"], f["coefb"])
    assert np.array_equal(grad, g["grad"])
    for irecmf in (0, 1):
        d = oracle.diffusion_potential_rc(s, bfc, i_geo, grad, f["pvar"], f["cofaf"], f["cofbf"],
                                          f["i_visc"], f["b_visc"], f["visel"], irecmf=irecmf)
        assert np.array_equal(d, g[f"div_irecmf{irecmf}"]), irecmf
    grad0 = oracle.gradient_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], f["pvar"], f["coefa"],
                                f["coefb"], inc=0)
    d = oracle.diffusion_potential_rc(s, bfc, i_geo, grad0, f["pvar"], f["cofaf"], f["cofbf"],
                                      f["i_visc"], f["b_visc"], f["visel"], irecmf=0, init=0, inc=0,
                                      diverg=f["div0"])
    assert np.array_equal(d, g["div_init0_inc0"])


@pytest.mark.parametrize("N", [5, 9])
def test_oracle_boundary_iprime_reconstruction_matches_golden(oracle, N):
    """II' of the boundary faces not zero: cs_gradient_boundary_iprime_lsq_s values at I' feed
    the face values of the gradient and the boundary fluxes (both irecmf formulas, inc 0 / 1):
    bit-exact against the reference"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import diipb_input, rc_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_rcb_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    diipb = diipb_input(N, len(bfc))
    vip = oracle.b_iprime_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], diipb, f["pvar"], f["coefa"],
                              f["coefb"])
    # the two short cuts of the reference: I' = I gives the cell value
    assert np.array_equal(vip[::7], f["pvar"][bfc][::7])
    assert np.array_equal(vip[3::11], f["pvar"][bfc][3::11])
    assert not np.array_equal(vip[1::7], f["pvar"][bfc][1::7])
    for inc in (1, 0):
        grad = oracle.gradient_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], f["pvar"], f["coefa"],
                                   f["coefb"], inc=inc, diipb=diipb)
        assert np.array_equal(grad, g[f"grad_inc{inc}"].reshape(-1, 3)), inc
        for irecmf in (0, 1):
            d = oracle.diffusion_potential_rc(s, bfc, i_geo, grad, f["pvar"], f["cofaf"],
                                              f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                              irecmf=irecmf, inc=inc, val_ip=vip)
            assert np.array_equal(d, g[f"div_irecmf{irecmf}_inc{inc}"]), (irecmf, inc)


@pytest.mark.parametrize("N", [5, 9])
def test_oracle_green_lsq_gradient_matches_golden(oracle, N):
    """imrgra = 4 (the reference's default gradient): least-squares gradient, then the
    Green-Gauss reconstruction pass; plain, and with boundary II' + warped-face correction:
    gradient and reconstructed divergence bit-exact against the reference"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import diipb_input, green_input, rc_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_green_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    for tag, diipb, corr in (("plain", None, None), ("full", diipb_input(N, len(bfc)), cgl)):
        vip = None
        if diipb is not None:
            vip = oracle.b_iprime_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], diipb, f["pvar"],
                                      f["coefa"], f["coefb"])
        for inc in (1, 0):
            grad = oracle.gradient_green_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], g_geo, b_surf,
                                             f["pvar"], f["coefa"], f["coefb"], inc=inc,
                                             diipb=diipb, corr_grad_lin=corr)
            assert np.array_equal(grad, g[f"{tag}_grad_inc{inc}"].reshape(-1, 3)), (tag, inc)
            for irecmf in (0, 1):
                d = oracle.diffusion_potential_rc(s, bfc, i_geo, grad, f["pvar"], f["cofaf"],
                                                  f["cofbf"], f["i_visc"], f["b_visc"],
                                                  f["visel"], irecmf=irecmf, inc=inc, val_ip=vip)
                assert np.array_equal(d, g[f"{tag}_div_irecmf{irecmf}_inc{inc}"]), (tag, irecmf, inc)


def test_oracle_green_lsq_gradient_is_exact_for_linear_fields(oracle):
    """property: on the orthogonal cube (OF = 0, weight 1/2) the Green-Gauss gradient of a
    linear field is its gradient when the boundary values are those of the field"""
    from make_golden_rhs import rhs_input
    N = 8
    s, bfc, b_geo, _, _ = rhs_input(N)
    h = 1.0 / N
    cen = s.cell_cen.reshape(-1, 3)
    a = np.array([0.7, -1.3, 2.1])
    p = cen @ a + 0.4
    vf = p[bfc] + b_geo[:, 6] * (b_geo[:, :3] @ a)
    g_geo = np.hstack([np.full((s.n_faces, 1), 0.5), np.zeros((s.n_faces, 3))])
    grad = oracle.gradient_green_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], g_geo,
                                     np.full(len(bfc), h * h), p, vf, np.zeros(len(bfc)))
    assert np.abs(grad - a).max() < 1e-12 * np.abs(a).max() * N


@pytest.mark.parametrize("N", [5, 9])
def test_oracle_face_diffusion_potential_and_divergence_match_golden(oracle, N):
    """cs_face_diffusion_potential (per-face fluxes, ircflu 0 / 1, imrgra 1 / 4, boundary II')
    and cs_divergence of them: bit-exact against the reference; the divergence of the fluxes is
    cs_diffusion_potential's result (irecmf 0)"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import (FACE_CASES, FACE_INIT_INC, diipb_input, face_input, green_input,
                                 rc_input)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_face_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    diipb = diipb_input(N, len(bfc))
    mf0 = face_input(s, len(bfc))
    bun, bd = b_geo[:, :3], b_geo[:, 6]
    vip = oracle.b_iprime_lsq(s, bfc, bun, bd, diipb, f["pvar"], f["coefa"], f["coefb"])
    for ircflu, imrgra in FACE_CASES:
        for init, inc in FACE_INIT_INC:
            kw = {}
            if ircflu:
                if imrgra == 4:
                    grad = oracle.gradient_green_lsq(s, bfc, bun, bd, g_geo, b_surf, f["pvar"],
                                                     f["coefa"], f["coefb"], inc=inc, diipb=diipb,
                                                     corr_grad_lin=cgl)
                else:
                    grad = oracle.gradient_lsq(s, bfc, bun, bd, f["pvar"], f["coefa"], f["coefb"],
                                               inc=inc, diipb=diipb)
                kw = dict(i_geo=i_geo, grad=grad, visel=f["visel"], val_ip=vip)
            im, bm = oracle.face_diffusion_potential(s, bfc, f["pvar"], f["cofaf"], f["cofbf"],
                                                     f["i_visc"], f["b_visc"], init=init, inc=inc,
                                                     massflux=mf0, **kw)
            tag = f"rc{ircflu}_g{imrgra}_init{init}_inc{inc}"
            assert np.array_equal(im, g["i_" + tag]) and np.array_equal(bm, g["b_" + tag]), tag
            d = oracle.divergence(s, bfc, im, bm, init=init, diverg=f["div0"])
            assert np.array_equal(d, g["div_" + tag]), tag
            if init == 1 and ircflu:
                d2 = oracle.diffusion_potential_rc(s, bfc, i_geo, grad, f["pvar"], f["cofaf"],
                                                   f["cofbf"], f["i_visc"], f["b_visc"],
                                                   f["visel"], irecmf=0, inc=inc, val_ip=vip)
                assert np.array_equal(d, d2), tag


@pytest.mark.parametrize("N", [5, 9])
def test_oracle_hydrostatic_treatment_matches_golden(oracle, N):
    """iphydp = 1 with an exterior force: values at I' / face values, least-squares gradient,
    Green-Gauss pass, reconstructed divergence: bit-exact against the reference"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import HYD_CASES, diipb_input, green_input, hyd_input, rc_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_hyd_N{N}.npz"))
    s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    fx = hyd_input(s.n_rows)
    for bt, imrgra in HYD_CASES:
        diipb = diipb_input(N, len(bfc)) if bt == "b" else None
        for inc in (1, 0):
            grad, vip = oracle.gradient_hyd(imrgra, s, bfc, b_geo[:, :3], b_geo[:, 6], ifc,
                                            b_geo[:, 3:6], fx, f["pvar"], f["coefa"], f["coefb"],
                                            inc=inc, diipb=diipb, g_geo=g_geo, b_face_surf=b_surf,
                                            corr_grad_lin=cgl)
            assert np.array_equal(grad, g[f"{bt}{imrgra}_grad_inc{inc}"]), (bt, imrgra, inc)
            for irecmf in (0, 1):
                d = oracle.diffusion_potential_rc(s, bfc, i_geo, grad, f["pvar"], f["cofaf"],
                                                  f["cofbf"], f["i_visc"], f["b_visc"],
                                                  f["visel"], irecmf=irecmf, inc=inc, val_ip=vip)
                assert np.array_equal(d, g[f"{bt}{imrgra}_div_irecmf{irecmf}_inc{inc}"])
    # a zero force takes the other code path to the same mathematical gradient
    z, _ = oracle.gradient_hyd(1, s, bfc, b_geo[:, :3], b_geo[:, 6], ifc, b_geo[:, 3:6],
                               np.zeros((s.n_rows, 3)), f["pvar"], f["coefa"], f["coefb"])
    p = oracle.gradient_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], f["pvar"], f["coefa"], f["coefb"])
    assert np.abs(z - p).max() <= 1e-12 * np.abs(p).max()


@pytest.mark.parametrize("N", [5, 9])
def test_oracle_iterative_green_gauss_matches_golden(oracle, N):
    """imrgra = 0: initial Green-Gauss gradient + reconstruction sweeps to the reference's
    convergence test (5 to 13 sweeps here): gradient and reconstructed divergence bit-exact"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import diipb_input, green_input, rc_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_iter_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    g_geo, b_surf, _ = green_input(N, s.n_faces, len(bfc), s.n_rows)
    sweeps = set()
    for bt in ("p", "b"):
        diipb = diipb_input(N, len(bfc)) if bt == "b" else None
        vip = None
        if diipb is not None:
            vip = oracle.b_iprime_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], diipb, f["pvar"],
                                      f["coefa"], f["coefb"])
        for inc in (1, 0):
            grad, ns = oracle.gradient_green_iter(s, bfc, b_geo[:, :3], g_geo, b_surf, f["pvar"],
                                                  f["coefa"], f["coefb"], inc=inc, diipb=diipb)
            sweeps.add(ns)
            assert np.array_equal(grad, g[f"{bt}_grad_inc{inc}"]), (bt, inc)
            for irecmf in (0, 1):
                d = oracle.diffusion_potential_rc(s, bfc, i_geo, grad, f["pvar"], f["cofaf"],
                                                  f["cofbf"], f["i_visc"], f["b_visc"],
                                                  f["visel"], irecmf=irecmf, inc=inc, val_ip=vip)
                assert np.array_equal(d, g[f"{bt}_div_irecmf{irecmf}_inc{inc}"]), (bt, irecmf, inc)
    assert min(sweeps) >= 2 and max(sweeps) < 99            # really iterated, really converged


@pytest.mark.parametrize("N", [5, 9])
def test_oracle_gradient_limiter_matches_golden(oracle, N):
    """eqp->imligr 0 / 1 / 2 (cell, face, I'/J' limiters) on every gradient type, with the
    limiter of the boundary values at I': limited gradient and divergence bit-exact"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_rhs import (CLIP_CASES, diipb_input, green_input, hyd_input, rc_input)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rhs_clip_N{N}.npz"))
    s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    diipb = diipb_input(N, len(bfc))
    fx = hyd_input(s.n_rows)
    bun, bd = b_geo[:, :3], b_geo[:, 6]
    n_clipped = 0
    for imligr, climgr, imrgra, hyd in CLIP_CASES:
        if hyd:
            go, vip, vf = oracle.gradient_hyd(imrgra, s, bfc, bun, bd, ifc, b_geo[:, 3:6], fx,
                                              f["pvar"], f["coefa"], f["coefb"], diipb=diipb,
                                              g_geo=g_geo, b_face_surf=b_surf, corr_grad_lin=cgl,
                                              clip_coeff=1., with_val_f=True)
        else:
            vip = oracle.b_iprime_lsq(s, bfc, bun, bd, diipb, f["pvar"], f["coefa"], f["coefb"],
                                      clip_coeff=1.)
            vf = f["coefa"] + f["coefb"] * vip
            if imrgra == 4:
                go = oracle.gradient_green_lsq(s, bfc, bun, bd, g_geo, b_surf, f["pvar"],
                                               f["coefa"], f["coefb"], diipb=diipb,
                                               corr_grad_lin=cgl, clip_coeff=1.)
            elif imrgra == 1:
                go = oracle.gradient_lsq(s, bfc, bun, bd, f["pvar"], f["coefa"], f["coefb"],
                                         diipb=diipb, clip_coeff=1.)
            else:
                go, _ = oracle.gradient_green_iter(s, bfc, bun, g_geo, b_surf, f["pvar"],
                                                   f["coefa"], f["coefb"], diipb=diipb)
        gc = oracle.gradient_clip(s, bfc, i_geo, imligr, climgr, f["pvar"], vf, go)
        n_clipped += int((gc != go).any(axis=1).sum())
        tag = f"l{imligr}_c{climgr}_g{imrgra}_h{int(hyd)}"
        assert np.array_equal(gc, g["grad_" + tag]), tag
        d = oracle.diffusion_potential_rc(s, bfc, i_geo, gc, f["pvar"], f["cofaf"], f["cofbf"],
                                          f["i_visc"], f["b_visc"], f["visel"], irecmf=0,
                                          val_ip=vip)
        assert np.array_equal(d, g["div_" + tag]), tag
    assert n_clipped > 100                              # the limiters really act here


def test_oracle_face_operators_on_irregular_mesh_match_reference(oracle):
    """an irregular mesh (tests/irregular_mesh.py: rows of 3 to 7 entries, arbitrary face
    orientation and order, unordered boundary faces, several per cell): gradient types
    0 / 1 / 4, hydrostatic treatment, limiters, reconstructed divergence and face fluxes of the
    oracle against the reference run live (oracle/_ref), bit for bit"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import irregular_cases, irregular_system, oracle_operators
    from oracle.pyoracle import RefLib, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ref = RefLib(1)
    s, bfc, G, f = irregular_system()
    assert np.diff(s.row_index).min() < 6 < np.diff(s.row_index).max()
    h = ref.system(s)
    for imrgra, hyd, imligr, climgr in irregular_cases():
        d, g = ref.diffusion_potential_rc(
            h, s, bfc, G["b_geo"], G["i_face_cog"], G["i_geo"], f["pvar"], f["coefa"], f["coefb"],
            f["cofaf"], f["cofbf"], f["i_visc"], f["b_visc"], f["visel"], irecmf=0, inc=1,
            diipb=G["diipb"], imrgra=imrgra, g_geo=G["g_geo"], b_face_surf=G["b_surf"],
            corr_grad_lin=G["cgl"], frcxt=(f["frcxt"] if hyd else None), imligr=imligr,
            climgr=climgr)
        go, do, _ = oracle_operators(oracle, s, bfc, G, f, imrgra, hyd, imligr, climgr)
        assert np.array_equal(g, go), (imrgra, hyd, imligr)
        assert np.array_equal(d, do), (imrgra, hyd, imligr)
    ref.destroy(h)


@pytest.mark.parametrize("n,k", [(300, 3), (3000, 4)])
def test_oracle_multigrid_on_irregular_mesh_matches_reference(oracle, n, k):
    """the hot path beyond the hex cube: hierarchy (aggregation maps, coarse structures and
    values) and the V-cycle solve of the oracle on the irregular system against the reference
    run live (oracle/_ref), bit for bit"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import irregular_system
    from oracle.pyoracle import RefLib, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ref = RefLib(1)
    s, _, _, _ = irregular_system(n=n, k=k)
    h = ref.system(s)
    lv_ref = ref.hierarchy(h)
    mg = oracle.mg_setup(s)
    lv = oracle.mg_levels(mg)
    assert [g["n_rows"] for g in lv] == [g["n_rows"] for g in lv_ref] and len(lv) >= 3
    for l in range(1, len(lv)):
        for key in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][key], lv_ref[l][key]), (l, key)
    r_norm = float(np.linalg.norm(s.rhs))
    a = ref.multigrid(h, s.rhs, 1e-8, r_norm)
    r = oracle.mg_solve(mg, s.rhs, 1e-8, r_norm)
    assert a["cvg"] == r["cvg"] == 1 and a["n_iter"] == r["n_iter"]
    assert np.array_equal(a["vx"], r["vx"])
    oracle.mg_free(mg)
    ref.destroy(h)


def test_oracle_solvers_on_irregular_mesh_match_reference(oracle):
    """SpMV, PCG, FCG, BiCGStab, Gauss-Seidel / RJ3 smoothers and FCG + cs_multigrid_pc of the
    oracle on the irregular system (3000 rows of 4 to 10 entries) against the reference run
    live: iterates and iteration counts bit for bit"""
    sys.path.insert(0, os.path.dirname(__file__))
    from irregular_mesh import irregular_system
    from oracle.pyoracle import RefLib, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    ref = RefLib(1)
    s, _, _, _ = irregular_system(n=3000, k=4)
    A = (s.row_index, s.col_id, s.d_val, s.x_val)
    rn = float(np.linalg.norm(s.rhs))
    h = ref.system(s)
    x = np.sin(np.arange(s.n_rows) * 0.37)
    assert np.array_equal(ref.spmv(h, x, s.n_rows), oracle.spmv(*A, x))
    for code, fn in ((ref.PCG, oracle.pcg), (ref.FCG, oracle.fcg), (ref.BICGSTAB, oracle.bicgstab)):
        a = ref.sles_it(h, code, s.rhs, 500, 1e-10, rn)
        r = fn(*A, s.rhs, 500, 1e-10, rn)
        assert a["cvg"] == r["cvg"] == 1 and a["n_iter"] == r["n_iter"], code
        assert np.array_equal(a["vx"], r["vx"]), code
    for st in (8, 9, 16):
        assert np.array_equal(ref.smoother(h, st, s.rhs, 3)["vx"],
                              oracle.smoother(st, *A, s.rhs, 3)["vx"]), st
    mg = oracle.mg_setup(s)
    a = ref.fcg_mgpc(h, s.rhs, 200, 1e-8, rn)
    r = oracle.fcg_mgpc(mg, s.rhs, 200, 1e-8, rn)
    assert a["cvg"] == r["cvg"] == 1 and a["n_iter"] == r["n_iter"]
    assert np.array_equal(a["vx"], r["vx"])
    oracle.mg_free(mg)
    ref.destroy(h)


def test_oracle_gradient_lsq_is_exact_for_linear_fields(oracle):
    """property: a least-squares gradient reproduces the gradient of a linear field (to the
    single precision of its cocg matrices) when the boundary values are those of the field"""
    from make_golden_rhs import rc_input
    s, bfc, b_geo, _, _, _ = rc_input(8)
    cen = s.cell_cen.reshape(-1, 3)
    a = np.array([0.7, -1.3, 2.1])
    p = cen @ a + 0.4
    # boundary value of the field at distance b_dist along the normal from the cell centre
    vf = p[bfc] + b_geo[:, 6] * (b_geo[:, :3] @ a)
    grad = oracle.gradient_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], p, vf, np.zeros(len(bfc)))
    assert np.abs(grad - a).max() < 5e-6 * np.abs(a).max()
