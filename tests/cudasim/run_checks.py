"""tests/cudasim/run_checks.py -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).

Runs in a process of its own (tests/test_cudasim.py) with the simulated library bound in
place of libcsb200.so: the UNMODIFIED kernels of csrc/kernels.cu and the launch sequences of
csrc/sles.cu execute on host threads and are compared with the reference-built golden
fixtures and the oracle.  Sizes are small: everything runs on one host thread.
Usage: python tests/cudasim/run_checks.py [check ...]   (default: all)
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import build as simbuild  # noqa: E402

os.environ["CSB_LIB"] = simbuild.build()
# the persistent cooperative kernels cannot run here: kernel-per-phase variants
os.environ["CSB_PCG_MODE_MULTI"] = "1"
os.environ["CSB_MG_MODE_MULTI"] = "1"
os.environ["CSB_FCG_MG_MODE_MULTI"] = "1"

import numpy as np  # noqa: E402

from code_saturne_b200 import api as csb  # noqa: E402
from code_saturne_b200.mesh_cartesian import cube_poisson, matrix_check_coefficients  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SMOOTHERS = {"jacobi": 3, "l1_jacobi": 13, "r_jacobi": 14, "rj2": 15, "rj3": 16}
CYCLES = {"rj3": ((16, 16, 0), (3, 3, 500)), "rj2": ((15, 15, 0), (2, 4, 500)),
          "r_jacobi": ((14, 14, 0), (2, 2, 500)), "l1_jacobi": ((13, 13, 0), (2, 2, 500)),
          "jacobi": ((3, 3, 0), (2, 3, 500))}


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _mg(types, nmi, precision_mult=(-1., -1., 1.)):
    mg = csb.Multigrid()
    mg.set_solver_options(descent=types[0], ascent=types[1], coarse=types[2], n_max_cycles=1,
                          n_max_iter_descent=nmi[0], n_max_iter_ascent=nmi[1],
                          n_max_iter_coarse=nmi[2], precision_mult=precision_mult)
    return mg


def check_spmv_pcg(ora):
    """the simulator itself, on paths the B200 runs have already validated"""
    s = cube_poisson(8, "neumann")
    d, xv, x = matrix_check_coefficients(s)
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, d, xv)
    assert np.array_equal(m.vector_multiply(x), ora.spmv(s.row_index, s.col_id, d, xv, x))
    assert np.array_equal(m.vector_multiply(x, exclude_diag=True),
                          ora.spmv(s.row_index, s.col_id, d, xv, x, True))
    m.destroy()
    m = csb.Matrix.from_system(s)
    rn = float(np.linalg.norm(s.rhs))
    c = csb.SlesIt(csb.SlesItType.PCG, 200)
    r = c.solve(m, s.rhs, 1e-8, rn)
    o = ora.pcg(s.row_index, s.col_id, s.d_val, s.x_val, s.rhs, 200, 1e-8, rn)
    assert r["n_iter"] == o["n_iter"] and rel(r["vx"][:s.n_rows], o["vx"]) < 1e-12
    c.destroy(); m.destroy()


def check_kernel_per_phase_paths(ora):
    """the other kernel-per-phase solvers that share k_spmv_stream's epilogues: PCG warm
    start (EPI_PCG_INIT), Jacobi, BiCGStab, the multi-kernel V-cycle (EPI_RESIDUAL_NORM)"""
    s = cube_poisson(8, "neumann")
    rn = float(np.linalg.norm(s.rhs))
    A = (s.row_index, s.col_id, s.d_val, s.x_val)
    m = csb.Matrix.from_system(s)
    o = ora.pcg(*A, s.rhs, 200, 1e-10, rn, vx0=0.5 * s.rhs)
    c = csb.SlesIt(csb.SlesItType.PCG, 200)
    r = c.solve(m, s.rhs, 1e-10, rn, vx=0.5 * s.rhs, use_vx_ini=True)
    assert abs(r["n_iter"] - o["n_iter"]) <= 1 and rel(r["vx"][:s.n_rows], o["vx"]) < 1e-9
    c.destroy()
    o = ora.jacobi(*A, s.rhs, 20, 1e-12, rn)
    c = csb.SlesIt(csb.SlesItType.JACOBI, 20)
    r = c.solve(m, s.rhs, 1e-12, rn)
    assert r["n_iter"] == o["n_iter"] and rel(r["vx"][:s.n_rows], o["vx"]) < 1e-13
    c.destroy()
    o = ora.bicgstab(*A, s.rhs, 200, 1e-10, rn)
    c = csb.SlesIt(csb.SlesItType.BICGSTAB, 200)
    r = c.solve(m, s.rhs, 1e-10, rn)
    assert abs(r["n_iter"] - o["n_iter"]) <= 2 and rel(r["vx"][:s.n_rows], o["vx"]) < 1e-7
    c.destroy()
    omg = ora.mg_setup(s)
    o = ora.mg_solve(omg, s.rhs, 1e-8, rn)
    mg = csb.Multigrid()
    mg.setup(m)
    lv, olv = mg.levels(), ora.mg_levels(omg)
    for a, b in zip(lv[1:], olv[1:]):
        assert np.array_equal(a["coarse_row"], b["coarse_row"])
        assert np.array_equal(a["x_val"], b["x_val"]) and np.array_equal(a["d_val"], b["d_val"])
    r = mg.solve(m, s.rhs, 1e-8, rn)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_cycles"] - o["n_cycles"]) <= 1
    k = min(len(r["cycle_residuals"]), len(o["cycle_residuals"]), 4)
    assert np.allclose(r["cycle_residuals"][:k], o["cycle_residuals"][:k], rtol=1e-7, atol=0)
    mg.destroy(); ora.mg_free(omg); m.destroy()
    # the same on an irregular system (rows of 3 to 9 entries, arbitrary face order)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from irregular_mesh import check_multigrid_against_oracle, check_solvers_against_oracle
    check_multigrid_against_oracle(csb, ora, n=600, k=4)
    check_solvers_against_oracle(csb, ora, n=600, k=4)


def check_smoothers(ora):
    for N, bc in ((8, "neumann"), (12, "dirichlet")):
        g = np.load(os.path.join(GOLD, f"mgpc_N{N}_{bc}.npz"))
        s = cube_poisson(N, bc)
        m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
        for name, st in SMOOTHERS.items():
            for nit in (1, 3):
                sm = csb.Smoother(st, nit)
                r = sm.solve(m, s.rhs)
                assert int(r["cvg"]) == -1 and r["residual"] == -1.0
                assert np.array_equal(r["vx"][:s.n_rows], g[f"sm_{name}_{nit}_zero"]), (name, nit)
                r = sm.solve(m, s.rhs, vx=g["vx0"].copy(), use_vx_ini=True)
                assert np.array_equal(r["vx"][:s.n_rows], g[f"sm_{name}_{nit}_warm"]), (name, nit)
                assert r["n_iter"] == int(g[f"sm_{name}_{nit}_n_iter"])
                sm.destroy()
        m.destroy()


def check_cycle_pc(ora):
    g = np.load(os.path.join(GOLD, "mgpc_N8_neumann.npz"))
    s = cube_poisson(8, "neumann")
    rn = float(g["r_norm"])
    m = csb.Matrix.from_system(s)
    for name in ("rj3", "l1_jacobi", "jacobi"):
        types, nmi = CYCLES[name]
        mg = _mg(types, nmi)
        mg.setup(m)
        r = mg.solve(m, s.rhs, 1e-8, rn)
        assert int(r["cvg"]) == 0
        assert abs(r["n_iter"] - int(g[f"cyc_{name}_n_iter"])) <= 1
        assert rel(r["vx"][:s.n_rows], g[f"cyc_{name}_vx"]) <= 1e-6, name
        mg.destroy()
    # the V-cycle's solver mode with smoothers (cycles to convergence)
    for name in ("rj3", "l1_jacobi"):
        types, nmi = CYCLES[name]
        mg = csb.Multigrid()
        mg.set_solver_options(descent=types[0], ascent=types[1], coarse=types[2],
                              n_max_cycles=100, n_max_iter_descent=nmi[0],
                              n_max_iter_ascent=nmi[1], n_max_iter_coarse=nmi[2],
                              precision_mult=(-1., -1., 1.))
        mg.setup(m)
        r = mg.solve(m, s.rhs, 1e-8, rn)
        n_ref = int(g[f"vsm_{name}_n_iter"])
        assert int(r["cvg"]) == 1 and abs(r["n_iter"] - n_ref) <= max(2, n_ref // 20), name
        assert abs(r["residual"] - float(g[f"vsm_{name}_residual"])) <= 1e-6 * r["residual"]
        mg.destroy()
    # PCG smoothers in the single-cycle form, warm start
    omg = ora.mg_setup(s)
    vx0 = np.sin(0.11 * np.arange(s.n_rows))
    mg = _mg((0, 0, 0), (2, 10, 500), (1., 1., 1.))
    mg.setup(m)
    o = ora.mg_cycle_pc(omg, s.rhs, 1e-14, rn, (0, 0, 0), (2, 10, 500), (1., 1., 1.), vx0=vx0)
    r = mg.solve(m, s.rhs, 1e-14, rn, vx=vx0.copy(), use_vx_ini=True)
    assert rel(r["vx"][:s.n_rows], o["vx"]) <= 1e-9
    mg.destroy(); ora.mg_free(omg); m.destroy()


def check_fcg_mgpc(ora):
    g = np.load(os.path.join(GOLD, "mgpc_N8_neumann.npz"))
    s = cube_poisson(8, "neumann")
    rn = float(g["r_norm"])
    m = csb.Matrix.from_system(s)
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(csb.Multigrid(pc=True))          # defaults = RJ3, 2 -> 3 sweeps
    r = c.solve(m, s.rhs, 1e-8, rn)
    assert int(r["cvg"]) == int(g["fcg_rj3_cvg"]) == 1
    assert abs(r["n_iter"] - int(g["fcg_rj3_n_iter"])) <= 2
    k = min(len(r["history"]) - 1, len(g["fcg_rj3_history"]), 4)
    assert np.allclose(r["history"][1:1 + k], g["fcg_rj3_history"][:k], rtol=1e-5, atol=0)
    assert rel(r["vx"][:s.n_rows], g["fcg_rj3_vx"]) <= 1e-6
    z = c.solve(m, np.zeros(s.n_rows), 1e-8, 1.0)
    assert not z["vx"].any()
    # warm start with a max-iteration exit
    omg = ora.mg_setup(s)
    c3 = csb.SlesIt(csb.SlesItType.FCG, 3, poly_degree=-1)
    c3.transfer_pc(csb.Multigrid(pc=True))
    ow = ora.fcg_mgpc(omg, s.rhs, 3, 1e-14, rn, vx0=0.5 * g["fcg_rj3_vx"])
    w = c3.solve(m, s.rhs, 1e-14, rn, vx=0.5 * g["fcg_rj3_vx"], use_vx_ini=True)
    assert (int(w["cvg"]), w["n_iter"]) == (ow["cvg"], ow["n_iter"]) == (-1, 3)
    assert rel(w["vx"][:s.n_rows], ow["vx"]) <= 1e-9
    # copy keeps the preconditioner settings
    c.destroy(); c3.destroy(); ora.mg_free(omg)
    # identity preconditioner
    c = csb.SlesIt(csb.SlesItType.FCG, 500, poly_degree=-1)
    r = c.solve(m, s.rhs, 1e-10, rn)
    assert int(r["cvg"]) == 1
    res = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
    assert res <= 1.01e-10 * rn
    c.destroy(); m.destroy()


def check_general_pc(ora):
    g = np.load(os.path.join(GOLD, "pc_N8_neumann.npz"))
    s = cube_poisson(8, "neumann")
    rn = float(g["r_norm"])
    m = csb.Matrix.from_system(s)
    tags = [("pcg_poly1", 0, 1), ("pcg_poly2", 0, 2), ("fcg_poly1", 1, 1), ("fcg_poly2", 1, 2),
            ("fcg_poly-1", 1, -1), ("pcg_mgpc", 0, -1), ("ipcg_poly1", 2, 1),
            ("ipcg_poly2", 2, 2), ("ipcg_mgpc", 2, -1)]

    def make(tag, so, deg, nmax):
        c = csb.SlesIt((csb.SlesItType.PCG, csb.SlesItType.FCG, csb.SlesItType.IPCG)[so], nmax,
                       poly_degree=deg)
        if tag.endswith("_mgpc"):
            c.transfer_pc(csb.Multigrid(pc=True))
        return c

    for tag, so, deg in tags:
        c = make(tag, so, deg, 500)
        r = c.solve(m, s.rhs, 1e-9, rn)
        assert int(r["cvg"]) == int(g[f"{tag}_cvg"]) == 1, tag
        assert abs(r["n_iter"] - int(g[f"{tag}_n_iter"])) <= 2, tag
        assert rel(r["vx"][:s.n_rows], g[f"{tag}_vx"]) <= 1e-6, tag
        w = make(tag, so, deg, 4)
        rw = w.solve(m, s.rhs, 1e-12, rn, vx=0.3 * g[f"{tag}_vx"], use_vx_ini=True)
        assert (int(rw["cvg"]), rw["n_iter"]) == (int(g[f"{tag}_warm_cvg"]),
                                                  int(g[f"{tag}_warm_n_iter"])), tag
        assert rel(rw["vx"][:s.n_rows], g[f"{tag}_warm_vx"]) <= 1e-9, tag
        c.destroy(); w.destroy()
    m.destroy()


def check_block3_jacobi(ora):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from irregular_mesh import check_block3_against_oracle
    check_block3_against_oracle(csb, ora)
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion
    g = np.load(os.path.join(GOLD, "block3jac_N6.npz"))
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(6)
    n = len(ri) - 1
    rn = float(g["r_norm"])
    m = csb.Matrix.from_block(n, 3, ri, ci, d.ravel(), xv)
    for nit in (1, 2, 7):
        c = csb.SlesIt(csb.SlesItType.JACOBI, nit)
        r = c.solve(m, rhs, 1e-12, rn)
        assert (int(r["cvg"]), r["n_iter"]) == (-1, nit)
        assert np.array_equal(r["vx"], g[f"sweeps{nit}_vx"]), nit
        c.destroy()
    c = csb.SlesIt(csb.SlesItType.JACOBI, 1000)
    r = c.solve(m, rhs, 1e-8, rn)
    assert int(r["cvg"]) == 1 and abs(r["n_iter"] - int(g["n_iter"])) <= 1
    c.destroy()
    c = csb.SlesIt(csb.SlesItType.JACOBI, 5)
    w = c.solve(m, rhs, 1e-12, rn, vx=g["vx0"].copy(), use_vx_ini=True)
    assert np.array_equal(w["vx"], g["warm_vx"])
    c.destroy(); m.destroy()


def check_galerkin(ora):
    g = np.load(os.path.join(GOLD, "galerkin_N8_neumann.npz"))
    s = cube_poisson(8, "neumann")
    rn = float(g["r_norm"])
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)     # no mesh
    mg = csb.Multigrid()
    mg.set_coarsening_options(p0p1_relax=0.)
    mg.setup(m)
    lv = mg.levels()
    assert len(lv) == int(g["n_levels"])
    for l in range(1, len(lv)):
        for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
            assert np.array_equal(lv[l][k], g[f"l{l}_{k}"]), (l, k)
    r = mg.solve(m, s.rhs, 1e-8, rn)
    assert int(r["cvg"]) == 1 and abs(r["n_iter"] - int(g["mg_n_iter"])) <= 5
    assert rel(r["vx"][:s.n_rows], g["mg_vx"]) <= 1e-6
    mg.destroy(); m.destroy()


def check_native(ora):
    """matrix mapped from native arrays: face-list aggregation, face-order coarse values
    (oracle/cs_oracle_native.c is pinned bit for bit to the reference, tests/test_oracle.py)"""
    from oracle.pyoracle import native_arrays
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from irregular_mesh import check_native_against_oracle
    check_native_against_oracle(csb, ora, n=600, k=4)      # faces in arbitrary orientation / order
    for N, bc, blocks, perturb, relax in ((8, "neumann", 1, False, 0.95),
                                          (12, "dirichlet", 1, True, 0.95),
                                          (12, "neumann", 3, True, 0.95),
                                          (12, "neumann", 1, True, 0.)):
        s = cube_poisson(N, bc)
        da, xa = native_arrays(s)
        if perturb:
            rng = np.random.default_rng(1234)
            xa = xa * (1 + 0.3 * rng.random(len(xa)))
            da = np.zeros(s.n_rows)
            np.add.at(da, s.face_cells[:, 0], -xa)
            np.add.at(da, s.face_cells[:, 1], -xa)
            da += 1e-3 * np.abs(xa).mean()
        omg = ora.mg_setup_native(s, n_threads=blocks, da=da, xa=xa, p0p1_relax=relax)
        lo = ora.mg_levels(omg)
        m = csb.Matrix.from_system_native(s, da, xa)
        assert np.array_equal(m.get_coefficients()[1], lo[0]["x_val"])
        mg = csb.Multigrid()
        mg.set_aggregation_blocks(blocks)
        mg.set_coarsening_options(p0p1_relax=relax)
        mg.setup(m)
        lv = mg.levels()
        assert len(lv) == len(lo), (len(lv), len(lo))
        for l in range(1, len(lv)):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                assert np.array_equal(lv[l][k], lo[l][k]), (N, bc, blocks, relax, l, k)
        rn = float(np.linalg.norm(s.rhs))
        r = mg.solve(m, s.rhs, 1e-8, rn)
        o = ora.mg_solve(omg, s.rhs, 1e-8, rn)
        assert int(r["cvg"]) == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2, (r["n_iter"], o["n_iter"])
        assert rel(r["vx"][:s.n_rows], o["vx"]) <= 1e-6
        ora.mg_free(omg)
        mg.destroy(); m.destroy()


def check_assembler(ora):
    """matrix assembler: structure and values (sums in arrival order) against the reference's
    fixtures; the declared-sequence plan; SpMV on the assembled matrix"""
    sys.path.insert(0, GOLD)
    from make_golden_assembler import CASES, assembler_input
    for N, (l0, seed) in CASES.items():
        g = np.load(os.path.join(GOLD, f"assembler_N{N}.npz"))
        n, rows, cols, vals, batches = assembler_input(N, l0, seed)
        ma = csb.MatrixAssembler((l0, l0 + n))
        ma.add_g_ids(rows[:100], cols[:100]); ma.add_g_ids(rows[100:], cols[100:])
        ma.compute()
        ri, ci = ma.structure()
        assert np.array_equal(ri, g["row_index"]) and np.array_equal(ci, g["col_id"])
        m = ma.create_matrix()
        ma.values_init(m)
        for b in batches:
            ma.values_add_g(*b)
        ma.values_finalize()
        d, x = m.get_coefficients()
        assert np.array_equal(d, g["d_val"]) and np.array_equal(x, g["x_val"])
        xv = np.sin(np.arange(n) * 0.37)
        assert np.array_equal(m.vector_multiply(xv.copy()), g["y"])
        ma.values_init(m)
        ma.values_add_declared(vals)
        ma.values_finalize()
        d, x = m.get_coefficients()
        assert np.array_equal(d, g["d_val_declared"]) and np.array_equal(x, g["x_val_declared"])
        m.destroy(); ma.destroy()


def check_rhs(ora):
    """cs_diffusion_potential without reconstruction: per-cell gathers in face order against
    the reference's fixtures"""
    sys.path.insert(0, GOLD)
    from make_golden_rhs import CASES, rhs_input
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_N{N}.npz"))
        s, bfc, _, _, f = rhs_input(N)
        for native in (False, True):
            m = csb.Matrix.from_system_native(s) if native else csb.Matrix.from_system(s)
            m.set_boundary_faces(bfc)
            for init, inc in ((1, 1), (0, 1), (1, 0)):
                d = m.diffusion_potential(f["pvar"], f["i_visc"], f["b_visc"], f["cofaf"], f["cofbf"],
                                          init=init, inc=inc, diverg=f["div0"])
                assert np.array_equal(d, g[f"div_init{init}_inc{inc}"]), (N, native, init, inc)
            m.destroy()
    # non-orthogonal variant: least-squares gradient and the reconstructed branch
    from make_golden_rhs import rc_input
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_rc_N{N}.npz"))
        s, bfc, b_geo, _, i_geo, f = rc_input(N)
        m = csb.Matrix.from_system(s)
        m.set_boundary_faces(bfc)
        m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
        assert np.array_equal(m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"]), g["grad"])
        for irecmf in (0, 1):
            d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                         f["i_visc"], f["b_visc"], f["visel"], irecmf=irecmf)
            assert np.array_equal(d, g[f"div_irecmf{irecmf}"]), (N, irecmf)
        d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                     f["i_visc"], f["b_visc"], f["visel"], irecmf=0, init=0, inc=0,
                                     diverg=f["div0"])
        assert np.array_equal(d, g["div_init0_inc0"])
        m.destroy()
    # II' of the boundary faces not zero: values at I' (cs_gradient_boundary_iprime_lsq_s)
    from make_golden_rhs import diipb_input
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_rcb_N{N}.npz"))
        s, bfc, b_geo, _, i_geo, f = rc_input(N)
        m = csb.Matrix.from_system(s)
        m.set_boundary_faces(bfc)
        m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
        g0 = m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"])       # cocg cached with II' = 0
        m.set_boundary_iprime(diipb_input(N, len(bfc)))
        for inc in (1, 0):
            gr = m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"], inc=inc)
            assert np.array_equal(gr, g[f"grad_inc{inc}"].reshape(-1, 3)), (N, inc)
            for irecmf in (0, 1):
                d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                             f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                             irecmf=irecmf, inc=inc)
                assert np.array_equal(d, g[f"div_irecmf{irecmf}_inc{inc}"]), (N, irecmf, inc)
        m.set_boundary_iprime(None)
        assert np.array_equal(m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"]), g0)
        m.destroy()
    # cs_face_diffusion_potential + cs_divergence
    from make_golden_rhs import FACE_CASES, FACE_INIT_INC, face_input, green_input
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_face_N{N}.npz"))
        s, bfc, b_geo, _, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        mf0 = face_input(s, len(bfc))
        m = csb.Matrix.from_system(s)
        m.set_boundary_faces(bfc)
        m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
        m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
        m.set_boundary_iprime(diipb_input(N, len(bfc)))
        for ircflu, imrgra in FACE_CASES:
            m.set_gradient_type(imrgra)
            for init, inc in FACE_INIT_INC:
                im, bm = m.face_diffusion_potential(
                    s.n_faces, len(bfc), f["pvar"], f["cofaf"], f["cofbf"], f["i_visc"], f["b_visc"],
                    coefa=f["coefa"], coefb=f["coefb"], visel=f["visel"], ircflu=ircflu, init=init,
                    inc=inc, massflux=mf0)
                tag = f"rc{ircflu}_g{imrgra}_init{init}_inc{inc}"
                assert np.array_equal(im, g["i_" + tag]) and np.array_equal(bm, g["b_" + tag]), (N, tag)
                d = m.divergence(im, bm, init=init, diverg=f["div0"])
                assert np.array_equal(d, g["div_" + tag]), (N, tag)
        m.destroy()
    # hydrostatic pressure treatment (iphydp = 1 with an exterior force)
    from make_golden_rhs import HYD_CASES, hyd_input
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_hyd_N{N}.npz"))
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        fx = hyd_input(s.n_rows)
        mf0 = face_input(s, len(bfc))
        m = csb.Matrix.from_system(s)
        m.set_boundary_faces(bfc)
        m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
        m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
        m.set_face_cog(ifc, b_geo[:, 3:6])
        for bt, imrgra in HYD_CASES:
            m.set_boundary_iprime(diipb_input(N, len(bfc)) if bt == "b" else None)
            m.set_gradient_type(imrgra)
            for inc in (1, 0):
                gr = m.gradient_hyd(imrgra, fx, f["pvar"], f["coefa"], f["coefb"], inc=inc)
                assert np.array_equal(gr, g[f"{bt}{imrgra}_grad_inc{inc}"]), (N, bt, imrgra, inc)
                for irecmf in (0, 1):
                    d = m.diffusion_potential_hyd(fx, f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                                  f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                                  irecmf=irecmf, inc=inc)
                    assert np.array_equal(d, g[f"{bt}{imrgra}_div_irecmf{irecmf}_inc{inc}"]), \
                        (N, bt, imrgra, irecmf, inc)
            im, bm = m.face_diffusion_potential_hyd(
                s.n_faces, len(bfc), fx, f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                f["i_visc"], f["b_visc"], f["visel"], init=0, massflux=mf0)
            assert np.array_equal(im, g[f"{bt}{imrgra}_i_massflux"]), (N, bt, imrgra)
            assert np.array_equal(bm, g[f"{bt}{imrgra}_b_massflux"]), (N, bt, imrgra)
        m.destroy()
    # imrgra = 0: iterative Green-Gauss gradient
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_iter_N{N}.npz"))
        s, bfc, b_geo, _, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        m = csb.Matrix.from_system(s)
        m.set_boundary_faces(bfc)
        m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
        m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
        m.set_gradient_type(0)
        for bt in ("p", "b"):
            diipb = diipb_input(N, len(bfc)) if bt == "b" else None
            m.set_boundary_iprime(diipb)
            for inc in (1, 0):
                gr = m.gradient(0, f["pvar"], f["coefa"], f["coefb"], inc=inc)
                assert np.array_equal(gr, g[f"{bt}_grad_inc{inc}"]), (N, bt, inc)
                _, ns = ora.gradient_green_iter(s, bfc, b_geo[:, :3], g_geo, b_surf, f["pvar"],
                                                f["coefa"], f["coefb"], inc=inc, diipb=diipb)
                assert m.gradient_sweeps() == ns, (N, bt, inc, m.gradient_sweeps(), ns)
                for irecmf in (0, 1):
                    d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                                 f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                                 irecmf=irecmf, inc=inc)
                    assert np.array_equal(d, g[f"{bt}_div_irecmf{irecmf}_inc{inc}"]), \
                        (N, bt, irecmf, inc)
        for nsw in (1, 3):                   # sweep limit: no sweep at all, two sweeps
            m.set_gradient_sweeps(nsw, 1e-4)
            go, ns = ora.gradient_green_iter(s, bfc, b_geo[:, :3], g_geo, b_surf, f["pvar"],
                                             f["coefa"], f["coefb"], diipb=diipb, nswrgr=nsw)
            assert np.array_equal(m.gradient(0, f["pvar"], f["coefa"], f["coefb"]), go), (N, nsw)
            assert m.gradient_sweeps() == ns == nsw
        m.destroy()
    # irregular mesh: arbitrary face orientation / order, unordered boundary faces
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from irregular_mesh import check_device_against_oracle
    check_device_against_oracle(csb, ora)
    # gradient limiter (imligr 0 / 1 / 2) on every gradient type
    from make_golden_rhs import CLIP_CASES
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_clip_N{N}.npz"))
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        fx = hyd_input(s.n_rows)
        m = csb.Matrix.from_system(s)
        m.set_boundary_faces(bfc)
        m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
        m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
        m.set_face_cog(ifc, b_geo[:, 3:6])
        m.set_boundary_iprime(diipb_input(N, len(bfc)))
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
            assert np.array_equal(gr, g["grad_" + tag]), (N, tag)
            assert np.array_equal(d, g["div_" + tag]), (N, tag)
        m.destroy()
    # imrgra = 4: Green-Gauss gradient with least-squares face values
    for N in CASES:
        g = np.load(os.path.join(GOLD, f"rhs_green_N{N}.npz"))
        s, bfc, b_geo, _, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        m = csb.Matrix.from_system(s)
        m.set_boundary_faces(bfc)
        m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
        for tag, diipb, corr in (("plain", None, None), ("full", diipb_input(N, len(bfc)), cgl)):
            m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, corr)
            m.set_boundary_iprime(diipb)
            m.set_gradient_type(4)
            for inc in (1, 0):
                gr = m.gradient(4, f["pvar"], f["coefa"], f["coefb"], inc=inc)
                assert np.array_equal(gr, g[f"{tag}_grad_inc{inc}"].reshape(-1, 3)), (N, tag, inc)
                for irecmf in (0, 1):
                    d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                                 f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                                 irecmf=irecmf, inc=inc)
                    assert np.array_equal(d, g[f"{tag}_div_irecmf{irecmf}_inc{inc}"]), \
                        (N, tag, irecmf, inc)
        m.destroy()


def check_pairwise(ora):
    g = np.load(os.path.join(GOLD, "pairwise_N8_neumann.npz"))
    s = cube_poisson(8, "neumann")
    rn = float(g["r_norm"])
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    for lim in (4,):
        mg = csb.Multigrid()
        mg.set_coarsening_options(aggregation_limit=lim, coarsening_type=3, n_max_levels=10,
                                  p0p1_relax=0.)
        mg.setup(m)
        lv = mg.levels()
        assert len(lv) == int(g[f"lim{lim}_n_levels"])
        for l in range(1, len(lv)):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                assert np.array_equal(lv[l][k], g[f"lim{lim}_l{l}_{k}"]), (lim, l, k)
        r = mg.solve(m, s.rhs, 1e-8, rn)
        n_ref = int(g[f"lim{lim}_mg_n_iter"])
        assert int(r["cvg"]) == 1 and abs(r["n_iter"] - n_ref) <= max(2, n_ref // 20)
        mg.destroy()
    m.destroy()


def check_compute_coeffs(ora):
    g = np.load(os.path.join(GOLD, "coeffs_N5.npz"))
    s = cube_poisson(5, "neumann")
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(g["b_face_cells"])
    for name, kw in (("default", dict()), ("no_dirichlet", dict(ndircp=0)),
                     ("theta_relax", dict(thetap=0.5, relaxp=0.7)),
                     ("no_diffusion", dict(idiffp=0))):
        m.compute_coeffs(g["rovsdt"], g["i_visc"], g["b_visc"], g["cofbfp"], **kw)
        d, x = m.get_coefficients()
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name
    for name, kw in (("conv", dict()), ("conv_theta_relax", dict(ndircp=0, thetap=0.5, relaxp=0.7)),
                     ("conv_only", dict(idiffp=0)), ("conv_thermal", dict(xcpp=g["xcpp"], thetap=0.5))):
        m.compute_coeffs(g["rovsdt"], g["i_visc"], g["b_visc"], g["cofbfp"], iconvp=1,
                         coefbp=g["coefbp"], i_massflux=g["i_massflux"],
                         b_massflux=g["b_massflux"], **kw)
        d, x = m.get_coefficients()
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name
    # the non-symmetric matrix solves with BiCGStab straight from the device coefficients
    m.compute_coeffs(1.0 + g["rovsdt"], g["i_visc"], g["b_visc"], g["cofbfp"], iconvp=1,
                     coefbp=g["coefbp"], i_massflux=0.1 * g["i_massflux"],
                     b_massflux=0.0 * g["b_massflux"])
    d, x = m.get_coefficients()
    rhs = np.sin(0.2 * np.arange(s.n_rows))
    rn = float(np.linalg.norm(rhs))
    c = csb.SlesIt(csb.SlesItType.BICGSTAB, 500)
    r = c.solve(m, rhs, 1e-10, rn)
    assert int(r["cvg"]) == 1
    res = np.linalg.norm(rhs - ora.spmv(s.row_index, s.col_id, d, x, r["vx"][:s.n_rows]))
    assert res <= 1.01e-10 * rn
    c.destroy(); m.destroy()
    # vector variable: 3x3 diagonal blocks
    mb = csb.Matrix.from_block(s.n_rows, 3, s.row_index, s.col_id,
                               np.tile(np.eye(3).ravel(), s.n_rows), np.zeros(s.nnz))
    mb.set_mesh_association(s.n_faces, s.c2f, s.c2f_sgn, s.cell_cen, s.cell_vol,
                            s.face_u_normal, s.face_surf)
    mb.set_boundary_faces(g["b_face_cells"])
    for name, kw in (("v3_sym", dict(iconvp=0)),
                     ("v3_sym_nofimp", dict(iconvp=0, fimp=False, ndircp=0, relaxp=0.5)),
                     ("v3_conv", dict(iconvp=1)),
                     ("v3_conv_theta", dict(iconvp=1, ndircp=0, thetap=0.5, relaxp=0.7)),
                     ("v3_conv_only", dict(iconvp=1, idiffp=0, fimp=False))):
        kw = dict(kw)
        fimp = g["v3_fimp"] if kw.pop("fimp", True) else None
        mb.compute_coeffs_vector(g["i_visc"], g["b_visc"], g["v3_cofbfp"], g["v3_coefbp"], fimp,
                                 g["i_massflux"], g["b_massflux"], **kw)
        d, x = mb.get_coefficients()
        assert np.array_equal(d, g[f"{name}_d"]) and np.array_equal(x, g[f"{name}_x"]), name
    mb.destroy()


def check_k_cycle(ora):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from irregular_mesh import check_coeffs_and_k_cycle_against_oracle
    check_coeffs_and_k_cycle_against_oracle(csb, ora, n=600, k=4)
    g = np.load(os.path.join(GOLD, "kcycle_N8_dirichlet.npz"))
    s = cube_poisson(8, "dirichlet")
    rn = float(g["r_norm"])
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)

    def kpc():
        pc = csb.Multigrid(pc=True, mg_type=1)
        pc.set_coarsening_options(aggregation_limit=4, coarsening_type=3, n_max_levels=10,
                                  min_g_rows=30, p0p1_relax=0.)
        return pc

    mg = kpc()
    mg.setup(m)
    r = mg.solve(m, s.rhs, 1e-8, rn)
    assert rel(r["vx"][:s.n_rows], g["cycle_vx"]) <= 1e-6
    assert abs(r["n_iter"] - int(g["cycle_n_iter"])) <= 1
    mg.destroy()
    c = csb.SlesIt(csb.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(kpc())
    r = c.solve(m, s.rhs, 1e-8, rn)
    assert int(r["cvg"]) == 1 and abs(r["n_iter"] - int(g["fcg_n_iter"])) <= 2
    assert rel(r["vx"][:s.n_rows], g["fcg_vx"]) <= 1e-6
    c.destroy(); m.destroy()


CHECKS = {"spmv_pcg": check_spmv_pcg, "rhs": check_rhs, "native": check_native, "assembler": check_assembler, "k_cycle": check_k_cycle, "compute_coeffs": check_compute_coeffs, "pairwise": check_pairwise, "galerkin": check_galerkin, "block3_jacobi": check_block3_jacobi, "general_pc": check_general_pc,
          "kernel_per_phase": check_kernel_per_phase_paths, "smoothers": check_smoothers,
          "cycle_pc": check_cycle_pc, "fcg_mgpc": check_fcg_mgpc}


def main(argv):
    csb.init(0)
    ora = Oracle()
    for name in (argv or list(CHECKS)):
        t0 = time.time()
        CHECKS[name](ora)
        print(f"cudasim {name}: ok ({time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
