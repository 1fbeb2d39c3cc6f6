"""Round-2 GPU tests, through the C-ABI.

* k_fcg_mg_persistent (FCG + one V-cycle per iteration in ONE launch) against the
  kernel-per-phase path it replaces, the oracle and the reference build;
* parity at the benchmarked sizes (BASELINE configs 1-3: 50^3, 128^3, 256^3) against the
  reference's own code (oracle/_ref/libcsref.so): level count, every aggregation map
  bit-exact, cycle / iteration counts, converged solutions <= 1e-10 relative L2 at
  precision 1e-12;
* the fixes of round 1's review: csb_matrix_set_coefficients on block / CSR matrices, the
  one-level hierarchy in the kernel-per-phase cycle loop, a failed setup leaving the object
  clean.

Tolerances: aggregation maps bit-exact; Krylov iteration counts +-2; cs_multigrid cycle
counts +-1 (equivalent iterations: conftest.equiv_iters_close); solutions 1e-10 relative L2.
"""
import os

import numpy as np
import pytest

from code_saturne_b200.mesh_cartesian import cube_poisson
from conftest import equiv_iters_close, rel_l2

pytestmark = pytest.mark.gpu

SOL_TOL = 1e-10
CYCLES = {"rj3": ((16, 16, 0), (3, 3, 500)), "rj2": ((15, 15, 0), (2, 4, 500)),
          "r_jacobi": ((14, 14, 0), (2, 2, 500)), "l1_jacobi": ((13, 13, 0), (2, 2, 500)),
          "jacobi": ((3, 3, 0), (2, 3, 500)), "jacobi_1": ((3, 3, 0), (1, 1, 500)),
          "l1_4": ((13, 13, 0), (4, 4, 500))}


def _pc(csb, types, nmi):
    mg = csb.Multigrid()
    mg.set_solver_options(descent=types[0], ascent=types[1], coarse=types[2], n_max_cycles=1,
                          n_max_iter_descent=nmi[0], n_max_iter_ascent=nmi[1],
                          n_max_iter_coarse=nmi[2], precision_mult=(-1., -1., 1.))
    return mg


def _fcg(csb, m, s, precision, r_norm, pc, multi, n_max_iter=200, **kw):
    """FCG + multigrid preconditioner; multi: kernel per phase instead of one launch."""
    old = os.environ.get("CSB_FCG_MG_MODE_MULTI")
    os.environ["CSB_FCG_MG_MODE_MULTI"] = "1" if multi else "0"
    try:
        c = csb.SlesIt(csb.SlesItType.FCG, n_max_iter, poly_degree=-1)
        c.transfer_pc(pc)
        c.setup(m)                       # hierarchy construction launches are not the solve's
        csb.synchronize()
        csb.launch_count(reset=True)
        r = c.solve(m, s.rhs, precision, r_norm, **kw)
        r["launches"] = csb.launch_count()
        c.destroy()
    finally:
        if old is None:
            os.environ.pop("CSB_FCG_MG_MODE_MULTI", None)
        else:
            os.environ["CSB_FCG_MG_MODE_MULTI"] = old
    return r


@pytest.mark.parametrize("N,bc", [(24, "neumann"), (40, "dirichlet")])
def test_fcg_mg_persistent_matches_kernel_per_phase(csb, N, bc):
    """Same algorithm, same per-row arithmetic; only the order of the 4 global sums differs."""
    s = cube_poisson(N, bc)
    r_norm = float(np.linalg.norm(s.rhs))
    m = csb.Matrix.from_system(s)
    for name, (types, nmi) in CYCLES.items():
        p = _fcg(csb, m, s, 1e-12, r_norm, _pc(csb, types, nmi), multi=False)
        q = _fcg(csb, m, s, 1e-12, r_norm, _pc(csb, types, nmi), multi=True)
        assert int(p["cvg"]) == int(q["cvg"]) == 1, name
        assert p["launches"] <= 3 < q["launches"], (name, p["launches"], q["launches"])
        assert abs(p["n_iter"] - q["n_iter"]) <= 1, (name, p["n_iter"], q["n_iter"])
        k = min(len(p["history"]), len(q["history"]), 6)
        assert np.allclose(p["history"][:k], q["history"][:k], rtol=1e-9, atol=0), name
        assert rel_l2(p["vx"][:s.n_rows], q["vx"][:s.n_rows]) <= SOL_TOL, name
        res = np.linalg.norm(s.rhs - m.vector_multiply(p["vx"][:s.n_rows].copy()))
        assert res <= 1.01e-12 * r_norm, name
    m.destroy()


def test_fcg_mg_persistent_warm_start_and_max_iter(csb, oracle):
    s = cube_poisson(32, "dirichlet")
    r_norm = float(np.sqrt(oracle.dot_xx(s.rhs)))
    omg = oracle.mg_setup(s)
    o = oracle.fcg_mgpc(omg, s.rhs, 200, 1e-13, r_norm)
    ow = oracle.fcg_mgpc(omg, s.rhs, 3, 1e-14, r_norm, vx0=0.5 * o["vx"])
    m = csb.Matrix.from_system(s)
    vx0 = np.concatenate([0.5 * o["vx"], np.zeros(m.n_cols_ext - s.n_rows)])
    w = _fcg(csb, m, s, 1e-14, r_norm, csb.Multigrid(pc=True), multi=False, n_max_iter=3,
             vx=vx0, use_vx_ini=True)
    assert (int(w["cvg"]), w["n_iter"]) == (ow["cvg"], ow["n_iter"]) == (-1, 3)
    assert w["launches"] <= 3
    assert rel_l2(w["vx"][:s.n_rows], ow["vx"]) <= 1e-9
    oracle.mg_free(omg); m.destroy()


def _reference_parity(csb, reflib, N, n_threads, vcycle=True, vcycle_precision=1e-12):
    """Hierarchy, V-cycle solver and FCG + V-cycle(RJ3) of an N^3 cube against the reference's
    own code run with n_threads OpenMP threads (its aggregation never crosses a thread's
    block of rows, cs_grid.cpp:4374-4393: the maps depend on the thread count)."""
    s = cube_poisson(N, "neumann")
    r_norm = float(np.linalg.norm(s.rhs))
    reflib.set_threads(n_threads)
    out = {}
    try:
        h = reflib.system(s)
        rlv = reflib.hierarchy(h)
        m = csb.Matrix.from_system(s)
        mg = csb.Multigrid()
        mg.set_aggregation_blocks(n_threads)
        mg.setup(m)
        assert mg.n_levels() == len(rlv), (mg.n_levels(), len(rlv))
        lv = mg.levels()
        # With more than 4 threads the reference lists the fine rows of a coarse row in the
        # order its threads happen to run (atomic capture, cs_grid.cpp:2930-2958): its own
        # coarse diagonal / cell sums then vary in the last bit from run to run (x_val and
        # the maps were identical in every run seen).  Values are compared bit for bit up
        # to 4 threads, to 1e-14 beyond; the maps bit for bit as long as the values feeding
        # them were.
        exact = (n_threads <= 4)
        bits_equal = True
        for l in range(1, len(lv)):
            assert lv[l]["n_rows"] == rlv[l]["n_rows"], l
            if bits_equal:
                assert np.array_equal(lv[l]["coarse_row"], rlv[l]["coarse_row"]), l
            if lv[l]["n_rows"] < 3000000:
                for k in ("x_val", "d_val"):
                    if exact:
                        assert np.array_equal(lv[l][k], rlv[l][k]), (l, k)
                    else:
                        assert lv[l][k].shape == rlv[l][k].shape, (l, k)
                        assert np.allclose(lv[l][k], rlv[l][k], rtol=1e-14, atol=0), (l, k)
                        bits_equal = bits_equal and np.array_equal(lv[l][k], rlv[l][k])
        out["maps_bit_exact_through_level"] = len(lv) - 1 if bits_equal else "first differing values"
        del lv, rlv
        out["n_levels"] = mg.n_levels()
        if vcycle:
            o = reflib.multigrid(h, s.rhs, vcycle_precision, r_norm)
            r = mg.solve(m, s.rhs, vcycle_precision, r_norm)
            assert int(r["cvg"]) == o["cvg"] == 1, (int(r["cvg"]), o["cvg"], r["n_cycles"])
            assert equiv_iters_close(r["n_iter"], o["n_iter"], r["n_iter"] / r["n_cycles"]), \
                (r["n_iter"], o["n_iter"], r["n_cycles"])
            out["vcycle"] = (r["n_cycles"], r["n_iter"], o["n_iter"],
                             rel_l2(r["vx"][:s.n_rows], o["vx"]))
            assert abs(r["residual"] - o["residual"]) <= 0.05 * o["residual"], out
            # solutions: 1e-10 when both are converged to 1e-12; a solve stopped at 1e-8 only
            # determines the (eps-regularised) constant mode to ~1e-6
            assert out["vcycle"][3] <= (SOL_TOL if vcycle_precision <= 1e-12 else 1e-5), out
        mg.destroy()
        o = reflib.fcg_mgpc(h, s.rhs, 200, 1e-12, r_norm)
        pc = csb.Multigrid(pc=True)
        pc.set_aggregation_blocks(n_threads)
        r = _fcg(csb, m, s, 1e-12, r_norm, pc, multi=False)
        assert int(r["cvg"]) == o["cvg"] == 1
        assert abs(r["n_iter"] - o["n_iter"]) <= 2, (r["n_iter"], o["n_iter"])
        out["fcg"] = (r["n_iter"], o["n_iter"], rel_l2(r["vx"][:s.n_rows], o["vx"]))
        assert out["fcg"][2] <= SOL_TOL, out
        assert r["launches"] <= 3
        m.destroy()
        reflib.destroy(h)
    finally:
        reflib.set_threads(1)
    return out


def test_reference_parity_config1_50_cube_threaded_aggregation(csb, reflib):
    """BASELINE config 1 (50^3) with the reference on 4 threads = 4 aggregation blocks."""
    _reference_parity(csb, reflib, 50, 4)


def test_reference_parity_config1_50_cube_single_thread(csb, reflib):
    _reference_parity(csb, reflib, 50, 1)


def test_reference_parity_config2_128_cube(csb, reflib):
    """BASELINE config 2 size, reference at OMP_NUM_THREADS=1."""
    _reference_parity(csb, reflib, 128, 1)


def test_reference_parity_bench_256_cube(csb, reflib):
    """The benchmarked workload (BASELINE config 3, 256^3): the reference on all host cores,
    this library with as many aggregation blocks."""
    t = max(1, min(os.cpu_count() or 1, 64))
    # the PCG-smoothed V-cycle in solver mode needs ~55 cycles per 1e-8 at this size (more
    # than its 100-cycle budget for 1e-12): it is compared at the benchmark's precision, the
    # 1e-10 solution bar is checked with FCG + V-cycle, solved to 1e-12
    out = _reference_parity(csb, reflib, 256, t, vcycle_precision=1e-8)
    print("256^3 parity:", out)


# ----------------------------------------------------------------- fixes of round 1's review

def test_set_coefficients_block3_updates_whole_blocks(csb):
    s = cube_poisson(8, "dirichlet", with_mesh=False)
    rng = np.random.default_rng(1234)
    n = s.n_rows
    def blocks(seed_shift):
        b = 0.1 * rng.standard_normal((n, 3, 3)) + seed_shift
        return (b + np.eye(3)[None] * (4.0 + seed_shift)).reshape(-1)
    d1, d2 = blocks(0.0), blocks(0.5)
    x2 = s.x_val * 1.25
    x = np.sin(0.3 * np.arange(3 * n))
    m = csb.Matrix.from_block(n, 3, s.row_index, s.col_id, d1, s.x_val)
    y1 = m.vector_multiply(x.copy())
    m.set_coefficients(d2, x2, symmetric=False)
    y2 = m.vector_multiply(x.copy())
    f = csb.Matrix.from_block(n, 3, s.row_index, s.col_id, d2, x2)
    yf = f.vector_multiply(x.copy())
    assert np.array_equal(y2, yf) and not np.array_equal(y1, y2)
    m.destroy(); f.destroy()


def test_set_coefficients_refused_on_csr_matrix(csb):
    s = cube_poisson(6, "dirichlet", with_mesh=False)
    # CSR with the diagonal first in each row
    ri = s.row_index + np.arange(s.n_rows + 1, dtype=np.int32)
    ci = np.empty(ri[-1], dtype=np.int32); v = np.empty(ri[-1])
    for i in range(s.n_rows):
        k0, k1 = s.row_index[i], s.row_index[i+1]
        ci[ri[i]] = i; v[ri[i]] = s.d_val[i]
        ci[ri[i]+1:ri[i+1]] = s.col_id[k0:k1]; v[ri[i]+1:ri[i+1]] = s.x_val[k0:k1]
    m = csb.Matrix.from_csr(s.n_rows, ri, ci, v)
    with pytest.raises(csb.CsbError):
        m.set_coefficients(s.d_val, s.x_val)
    m.destroy()


def test_one_level_hierarchy_kernel_per_phase_terminates(csb, oracle, monkeypatch):
    """n_max_levels = 1: no descent, no cycle test; the coarsest solver's result is the
    cycle's (k_mg_persistent does the same, persistent.cu)."""
    s = cube_poisson(12, "dirichlet")
    r_norm = float(np.linalg.norm(s.rhs))
    m = csb.Matrix.from_system(s)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CSB_MG_MODE_MULTI", mode)
        mg = csb.Multigrid()
        mg.set_coarsening_options(n_max_levels=1)
        mg.setup(m)
        assert mg.n_levels() == 1
        r = mg.solve(m, s.rhs, 1e-10, r_norm)
        assert int(r["cvg"]) == 1, mode
        tr = np.linalg.norm(s.rhs - m.vector_multiply(r["vx"][:s.n_rows].copy()))
        assert tr <= 1.01e-10 * r_norm and abs(tr - r["residual"]) <= 1e-2 * tr + 1e-14 * r_norm
        res[mode] = r
        mg.destroy()
    assert rel_l2(res["0"]["vx"], res["1"]["vx"]) <= 1e-9
    m.destroy()


def test_failed_setup_leaves_multigrid_reusable(csb):
    s = cube_poisson(10, "dirichlet")
    r_norm = float(np.linalg.norm(s.rhs))
    n = s.n_rows
    d3 = np.tile(np.eye(3).reshape(-1) * 6.0, n)
    mb = csb.Matrix.from_block(n, 3, s.row_index, s.col_id, d3, s.x_val, symmetric=True)
    m = csb.Matrix.from_system(s)
    mg = csb.Multigrid()
    with pytest.raises(csb.CsbError):
        mg.setup(mb)                      # block matrices are refused after level 0 is pushed
    mg.setup(m)
    fresh = csb.Multigrid()
    fresh.setup(m)
    assert mg.n_levels() == fresh.n_levels()
    a, b = mg.solve(m, s.rhs, 1e-9, r_norm), fresh.solve(m, s.rhs, 1e-9, r_norm)
    assert int(a["cvg"]) == 1 and a["n_cycles"] == b["n_cycles"]
    assert np.array_equal(a["vx"], b["vx"])
    mg.destroy(); fresh.destroy(); m.destroy(); mb.destroy()


def test_shared_memory_coarse_pcg_is_bit_identical_to_streamed(csb, monkeypatch):
    """pcg_small (coarsest level resident in shared memory) follows the arithmetic and the
    summation order of the streamed single-CTA PCG: same bits, V-cycle solver and FCG."""
    out = {}
    for N, bc in ((20, "neumann"), (32, "dirichlet")):
        s = cube_poisson(N, bc)
        r_norm = float(np.linalg.norm(s.rhs))
        m = csb.Matrix.from_system(s)
        for small in ("0", "448"):
            monkeypatch.setenv("CSB_PCG_SMALL_ROWS", small)
            mg = csb.Multigrid()
            mg.setup(m)
            r = mg.solve(m, s.rhs, 1e-10, r_norm)
            f = _fcg(csb, m, s, 1e-10, r_norm, csb.Multigrid(pc=True), multi=False)
            out[(N, small)] = (r["vx"].copy(), r["n_cycles"], f["vx"].copy(), f["n_iter"])
            mg.destroy()
        a, b = out[(N, "0")], out[(N, "448")]
        assert a[1] == b[1] and a[3] == b[3]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
        m.destroy()


def test_shared_host_arrays_give_the_same_hierarchy(csb):
    """csb_matrix_share_host_arrays: level 0 is aggregated from the caller's arrays instead of
    a device-to-host mirror; the hierarchy and the solve are identical, a value update drops
    the lent values, and setups after the first reuse the host pool's blocks."""
    s = cube_poisson(20, "neumann")
    r_norm = float(np.linalg.norm(s.rhs))
    out = []
    for share in (False, True, True):
        m = csb.Matrix.from_system(s, share_host=share)
        mg = csb.Multigrid()
        mg.set_aggregation_blocks(1)
        mg.setup(m)
        lv = mg.levels()
        r = mg.solve(m, s.rhs, 1e-8, r_norm)
        out.append((lv, r))
        if share:
            m.set_coefficients(s.d_val * 2.0, s.x_val * 2.0)      # lent values dropped
            mg.setup(m)
            r2 = mg.solve(m, 2.0 * s.rhs, 1e-8, 2.0 * r_norm)
            assert r2["n_cycles"] == r["n_cycles"]
            assert np.allclose(r2["vx"][:s.n_rows], r["vx"][:s.n_rows], rtol=1e-12, atol=0)
        mg.destroy()
        m.destroy()
    for lv, r in out[1:]:
        assert len(lv) == len(out[0][0])
        for a, b in zip(lv[1:], out[0][0][1:]):
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(r["vx"], out[0][1]["vx"]) and r["n_cycles"] == out[0][1]["n_cycles"]
    csb.load_library().csb_host_pool_trim()


_WIDE_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
from code_saturne_b200 import api
from code_saturne_b200.mesh_cartesian import cube_poisson
from oracle.pyoracle import Oracle
api.init(0)
ora = Oracle()
s = cube_poisson(20, "neumann")
m = api.Matrix.from_system(s)
mg = api.Multigrid()
mg.set_aggregation_blocks(1)
mg.setup(m)
omg = ora.mg_setup(s)
lv, olv = mg.levels(), ora.mg_levels(omg)
assert len(lv) == len(olv)
for a, b in zip(lv[1:], olv[1:]):
    for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
        assert np.array_equal(a[k], b[k]), k
print("WIDE-OK", len(lv))
"""


@pytest.mark.parametrize("limit", [0, 9])
def test_coarse_assembly_wide_row_kernel_is_bit_exact(limit):
    """Coarse rows with more entries than a warp's lanes own go to k_coarse_sums_wide; the cube
    hierarchies never have such rows, so the limit is lowered (CSB_CA_WIDE_LIMIT, read once per
    process): all rows (0) or the longer ones (9) take that kernel, the matrices stay those of
    the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CSB_CA_WIDE_LIMIT=str(limit))
    p = subprocess.run([sys.executable, "-c", _WIDE_SCRIPT.format(root=root)], env=env,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "WIDE-OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
