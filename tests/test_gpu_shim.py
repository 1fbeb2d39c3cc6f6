"""Drop-in test: reference -> cs_sles registry -> integration/cs_sles_b200.cpp -> libcsb200.so
-> back.

oracle/_ref/libcsref_b200.so holds the reference's own objects, the reference-side binding
compiled against the reference's headers, and a driver that (1) lets the reference build a
cs_matrix_t (MSR + mesh association), (2) registers a B200 solver with cs_sles_define
(src/alge/cs_sles.cpp:1385-1425) through cs_sles_b200_define, (3) finds the system by name
and calls cs_sles_solve (:1726-1860) on it, as cs_sles_solve_native and the physics modules
do.  The result is compared with the reference's own cs_multigrid / cs_sles_it solve of the
same cs_matrix_t.
"""
import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import cube_poisson
from conftest import equiv_iters_close, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refb():
    from oracle.pyoracle import RefLibB200, ref_b200_available
    if not ref_b200_available():
        pytest.skip("oracle/_ref/libcsref_b200.so not built (needs /root/reference)")
    return RefLibB200(1)


def test_registry_roundtrip_matches_reference_solvers(csb, refb):
    s = cube_poisson(32, "dirichlet")
    r_norm = float(np.linalg.norm(s.rhs))
    h = refb.system(s)
    # cs_multigrid V-cycle in solver mode
    o = refb.multigrid(h, s.rhs, 1e-12, r_norm)
    r = refb.b200_sles_solve(h, refb.MULTIGRID, s.rhs, 100, 1e-12, r_norm)
    assert r["cvg"] == o["cvg"] == 1
    assert equiv_iters_close(r["n_iter"], o["n_iter"], 6)
    assert rel_l2(r["vx"], o["vx"]) <= 1e-10
    assert r["uploads"] == 1 and r["value_updates"] == 0
    # FCG + multigrid preconditioner (the default pressure solver)
    o = refb.fcg_mgpc(h, s.rhs, 200, 1e-13, r_norm)
    r = refb.b200_sles_solve(h, refb.FCG_MG, s.rhs, 200, 1e-13, r_norm)
    assert r["cvg"] == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"], o["vx"]) <= 1e-10
    # PCG / FCG + Jacobi
    for kind, ref_type in ((refb.PCG_JACOBI, refb.PCG), (refb.FCG_JACOBI, refb.FCG)):
        o = refb.sles_it(h, ref_type, s.rhs, 2000, 1e-13, r_norm)
        r = refb.b200_sles_solve(h, kind, s.rhs, 2000, 1e-13, r_norm)
        assert r["cvg"] == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2, kind
        assert rel_l2(r["vx"], o["vx"]) <= 1e-10, kind
        assert abs(r["residual"] - o["residual"]) <= 0.5 * o["residual"] + 1e-16
    refb.destroy(h)


def test_registry_native_matrix_follows_the_reference_face_path(csb, refb):
    """A cs_matrix_t the reference built with cs_matrix_set_coefficients(edges, da, xa)
    ("mapped from native"): the binding creates the device matrix from the faces, and the
    hierarchy is the reference's face-list one (1 thread: _automatic_aggregation_fc +
    _compute_coarse_quantities_native) -- same cycle counts, same solution."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_native import native_system
    s, da, xa = native_system(24, "dirichlet", True)
    r_norm = float(np.linalg.norm(s.rhs))
    h = refb.system_native(s, da, xa)
    o = refb.multigrid(h, s.rhs, 1e-12, r_norm)
    r = refb.b200_sles_solve(h, refb.MULTIGRID, s.rhs, 100, 1e-12, r_norm)
    assert r["cvg"] == o["cvg"] == 1
    assert equiv_iters_close(r["n_iter"], o["n_iter"], 6)
    assert rel_l2(r["vx"], o["vx"]) <= 1e-10
    assert r["uploads"] == 1 and r["value_updates"] == 0
    o = refb.fcg_mgpc(h, s.rhs, 200, 1e-13, r_norm)
    r = refb.b200_sles_solve(h, refb.FCG_MG, s.rhs, 200, 1e-13, r_norm)
    assert r["cvg"] == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"], o["vx"]) <= 1e-10
    # the MSR-built matrix of the same system gives another hierarchy (the MSR aggregation):
    # the two entries are really different paths
    h2 = refb.system(s, d_val=da, x_val=xa[np.asarray(s.c2f)])
    o2 = refb.fcg_mgpc(h2, s.rhs, 200, 1e-13, r_norm)
    assert rel_l2(o2["vx"], o["vx"]) <= 1e-9
    refb.destroy(h2)
    refb.destroy(h)


def test_registry_reuses_device_matrix_and_frees(csb, refb):
    """Second solve of the same cs_matrix_t: no second upload; after cs_sles_free the next
    solve goes through setup (and one upload) again; warm start through vx_ini_0 = false."""
    s = cube_poisson(20, "neumann")
    r_norm = float(np.linalg.norm(s.rhs))
    h = refb.system(s)
    a = refb.b200_sles_solve(h, refb.FCG_MG, s.rhs, 200, 1e-9, r_norm, n_solves=3)
    assert a["cvg"] == 1 and a["uploads"] == 1
    b = refb.b200_sles_solve(h, refb.FCG_MG, s.rhs, 200, 1e-9, r_norm, n_solves=2,
                             free_between=True)
    assert b["cvg"] == 1 and b["uploads"] == 2
    assert np.array_equal(a["vx"], b["vx"])          # deterministic run to run
    w = refb.b200_sles_solve(h, refb.FCG_MG, s.rhs, 200, 1e-9, r_norm, vx0=a["vx"])
    assert w["cvg"] == 1 and w["n_iter"] <= 1
    refb.destroy(h)
    refb.finalize()                                  # destroy callbacks of every system

