"""Multi-rank GPU scenarios, run as a subprocess by tests/test_gpu_dist.py (a rank that
fails leaves its peers waiting, so every scenario runs under the parent's timeout):

    python tests/dist_worker.py comm R
    python tests/dist_worker.py pcg  N R bc
    python tests/dist_worker.py mg   N R bc [precision]

R ranks run as host threads of this process (one library context and stream each) and,
with a single visible GPU, share it: the same kernels, flags and peer stores as with one
process per GPU, minus NVLink.  The checker is the multi-rank oracle
(oracle/dist_oracle.py) on the same partition; solutions are also compared with the
single-rank oracle.  Tolerances as in tests/test_gpu_parity.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from code_saturne_b200 import api  # noqa: E402
from code_saturne_b200.dist import build_rank, gather_global, run_ranks_in_threads  # noqa: E402
from code_saturne_b200.mesh_cartesian import cube_poisson, cube_poisson_local  # noqa: E402


def device_of_rank_fn():
    import torch
    n = torch.cuda.device_count()
    return (lambda r: r % n) if n > 1 else None


def scenario_comm(R):
    import torch
    N = 12

    def fn(rank, pc):
        out = {}
        out["sum"] = pc.allreduce_sum([rank + 1.0, 0.5 * rank, 1.0])
        for _ in range(5):
            pc.barrier()
        s, h, gids = cube_poisson_local(N, R, rank)
        halo = api.Halo(h, rank, pcomm=pc)
        x = np.full(s.meta["n_cols_ext"], -1.0)
        x[:s.n_rows] = np.sin(gids[:s.n_rows] * 0.37)
        x_d = torch.from_numpy(x).cuda(torch.cuda.current_device())
        for _ in range(3):                       # repeated: acknowledgement counters
            halo.sync_var(x_d)
        torch.cuda.synchronize()
        out["x"] = x_d.cpu().numpy()
        out["gids"] = gids
        pc.barrier()
        halo.destroy()
        return out

    res = run_ranks_in_threads(R, fn, device_of_rank_fn())
    for r, o in enumerate(res):
        assert np.array_equal(o["sum"], [R * (R + 1) / 2.0, 0.5 * R * (R - 1) / 2.0, float(R)]), o["sum"]
        assert np.array_equal(o["x"], np.sin(o["gids"] * 0.37)), f"ghosts of rank {r}"
    print(f"OK comm R={R}")


def scenario_pcg(N, R, bc):
    from oracle.dist_oracle import DistOracle
    from oracle.pyoracle import Oracle
    ora = Oracle()
    g = cube_poisson(N, bc)
    r_norm = float(np.sqrt(ora.dot_xx(g.rhs)))
    parts = [cube_poisson_local(N, R, r, bc) for r in range(R)]
    D = DistOracle([p[0] for p in parts], [p[1] for p in parts], ora)
    o = D.pcg_solve([p[0].rhs for p in parts], 2000, 1e-10, r_norm)
    o1 = ora.pcg(g.row_index, g.col_id, g.d_val, g.x_val, g.rhs, 2000, 1e-10, r_norm)

    def fn(rank, pc):
        ro = build_rank(N, R, rank, pc, bc)
        c = ro.pcg(2000)
        r = c.solve(ro.matrix, ro.system.rhs, 1e-10, r_norm)
        r2 = c.solve(ro.matrix, ro.system.rhs, 1e-10, r_norm)       # counters persist
        assert np.array_equal(r["vx"][:ro.system.n_rows], r2["vx"][:ro.system.n_rows])
        # warm start from the solution: converged at once
        r3 = c.solve(ro.matrix, ro.system.rhs, 1e-8, r_norm, vx=r["vx"].copy(), use_vx_ini=True)
        assert int(r3["cvg"]) == 1 and r3["n_iter"] <= 1, r3["n_iter"]
        out = dict(cvg=int(r["cvg"]), n_iter=r["n_iter"], residual=r["residual"],
                   vx=r["vx"][:ro.system.n_rows].copy(), hist=r["history"])
        pc.barrier()
        c.destroy(); ro.destroy()
        return out

    res = run_ranks_in_threads(R, fn, device_of_rank_fn())
    for r, x in enumerate(res):
        assert x["cvg"] == o["cvg"] == 1
        assert x["n_iter"] == res[0]["n_iter"] and x["residual"] == res[0]["residual"]
        assert abs(x["n_iter"] - o["n_iter"]) <= 2, (x["n_iter"], o["n_iter"])
        k = min(len(x["hist"]), len(o["history"]), 6)
        assert np.allclose(x["hist"][:k], o["history"][:k], rtol=1e-10, atol=0)
    vx = gather_global([x["vx"] for x in res], [p[2] for p in parts], g.n_rows)
    err = np.linalg.norm(vx - o1["vx"]) / np.linalg.norm(o1["vx"])
    assert err <= 1e-9, err                    # both converged to 1e-10 only
    res_true = np.linalg.norm(ora.spmv(g.row_index, g.col_id, g.d_val, g.x_val, vx) - g.rhs)
    assert res_true <= 1.5e-10 * r_norm
    print(f"OK pcg N={N} R={R} {bc}: n_iter {res[0]['n_iter']} (oracle {o['n_iter']}, "
          f"single-rank {o1['n_iter']}), rel diff {err:.2e}")
    D.free()


def scenario_mg(N, R, bc, precision):
    from oracle.dist_oracle import DistOracle
    from oracle.pyoracle import Oracle
    ora = Oracle()
    g = cube_poisson(N, bc)
    r_norm = float(np.sqrt(ora.dot_xx(g.rhs)))
    parts = [cube_poisson_local(N, R, r, bc) for r in range(R)]
    D = DistOracle([p[0] for p in parts], [p[1] for p in parts], ora)
    nl = D.mg_setup()
    o = D.mg_solve([p[0].rhs for p in parts], precision, r_norm)
    omg = ora.mg_setup(g)
    o1 = ora.mg_solve(omg, g.rhs, precision, r_norm)

    def fn(rank, pc):
        ro = build_rank(N, R, rank, pc, bc)
        mg = ro.multigrid()
        mg.setup(ro.matrix)
        lv = mg.levels()
        halos = [mg.level_halo(l) for l in range(len(lv))]
        r = mg.solve(ro.matrix, ro.system.rhs, precision, r_norm)
        out = dict(levels=lv, halos=halos, cvg=int(r["cvg"]), n_iter=r["n_iter"],
                   n_cycles=r["n_cycles"], residual=r["residual"],
                   cyc=r["cycle_residuals"], vx=r["vx"][:ro.system.n_rows].copy())
        pc.barrier()
        mg.destroy(); ro.destroy()
        return out

    res = run_ranks_in_threads(R, fn, device_of_rank_fn())
    for r, x in enumerate(res):
        lv = x["levels"]
        assert len(lv) == nl, (len(lv), nl)
        for l in range(1, nl):
            a = D.level_arrays(l, r)
            assert lv[l]["n_rows"] == a["n_rows"] and lv[l]["n_cols_ext"] == a["n_cols_ext"]
            for k in ("coarse_row", "row_index", "col_id"):             # index data: bit-exact
                assert np.array_equal(lv[l][k], a[k]), (r, l, k)
            for k in ("d_val", "x_val"):                                # per-row arithmetic: bit-exact
                assert np.array_equal(lv[l][k], a[k]), (r, l, k)
            h, ho = x["halos"][l], a["halo"]
            if R > 1:
                assert np.array_equal(h["ranks"], ho.ranks) and np.array_equal(h["index"], ho.index)
                assert np.array_equal(h["send_index"], ho.send_index)
                assert np.array_equal(h["send_list"], ho.send_list)
        assert x["cvg"] == o["cvg"] == 1
        assert (x["n_iter"], x["n_cycles"], x["residual"]) \
            == (res[0]["n_iter"], res[0]["n_cycles"], res[0]["residual"])   # ranks agree exactly
        assert abs(x["n_cycles"] - o["n_cycles"]) <= 1, (x["n_cycles"], o["n_cycles"])
        d_it = abs(x["n_iter"] - o["n_iter"])
        assert d_it <= 2 or d_it <= int(np.ceil(o["n_iter"] / o["n_cycles"])) + 2
        k = min(len(x["cyc"]), len(o["cycle_residuals"]), 5)
        assert np.allclose(x["cyc"][:k], o["cycle_residuals"][:k], rtol=1e-7, atol=0)
    vx = gather_global([x["vx"] for x in res], [p[2] for p in parts], g.n_rows)
    res_true = np.linalg.norm(ora.spmv(g.row_index, g.col_id, g.d_val, g.x_val, vx) - g.rhs)
    assert res_true <= 1.0001 * precision * r_norm, res_true / r_norm
    err1 = np.linalg.norm(vx - o1["vx"]) / np.linalg.norm(o1["vx"])
    vxo = gather_global(o["vx"], [p[2] for p in parts], g.n_rows)
    erro = np.linalg.norm(vx - vxo) / np.linalg.norm(vxo)
    if bc == "dirichlet":
        # SPD, well conditioned: vs the single-rank reference solution
        assert err1 <= (1e-10 if precision <= 1e-12 else 1e-7), err1
        assert erro <= (1e-10 if precision <= 1e-12 else 1e-7), erro
    else:
        # all-Neumann + eps: the constant mode is determined to precision/eps only; the
        # same algorithm on the same partition must still agree closely
        assert erro <= 1e-5, erro
    print(f"OK mg N={N} R={R} {bc}: {nl} levels, cycles {res[0]['n_cycles']} "
          f"(oracle {o['n_cycles']}, single-rank {o1['n_cycles']}), "
          f"rel diff vs single-rank {err1:.2e}")
    ora.mg_free(omg)
    D.free()


def scenario_fcgmg(N, R, bc, precision, merge_mean):
    """FCG + V-cycle(RJ3) preconditioner in one persistent launch per rank, optionally with
    the coarse levels merged onto rank 0 (merge_mean > 0: mean rows per rank under which a
    level is merged).  The merged levels keep the partitioned aggregates and the order of
    every row's entries, so the merged run must reproduce the partitioned algorithm."""
    from oracle.dist_oracle import DistOracle
    from oracle.pyoracle import Oracle
    ora = Oracle()
    g = cube_poisson(N, bc)
    r_norm = float(np.sqrt(ora.dot_xx(g.rhs)))
    parts = [cube_poisson_local(N, R, r, bc) for r in range(R)]
    D = DistOracle([p[0] for p in parts], [p[1] for p in parts], ora)
    nl = D.mg_setup()
    o = D.fcg_mgpc([p[0].rhs for p in parts], 200, precision, r_norm)
    omg = ora.mg_setup(g)
    o1 = ora.fcg_mgpc(omg, g.rhs, 200, precision, r_norm)

    def fn(rank, pc):
        ro = build_rank(N, R, rank, pc, bc)
        c = api.SlesIt(api.SlesItType.FCG, 200, poly_degree=-1)
        mgp = api.Multigrid(pc=True)
        mgp.set_aggregation_blocks(1)
        if R > 1:
            c.set_pcomm(pc)
            mgp.set_pcomm(pc)
            if merge_mean > 0:
                mgp.set_merge_options(R, merge_mean, 0)
        c.transfer_pc(mgp)
        c.setup(ro.matrix)
        api.synchronize()
        api.launch_count(reset=True)
        r = c.solve(ro.matrix, ro.system.rhs, precision, r_norm)
        launches = api.launch_count()
        r2 = c.solve(ro.matrix, ro.system.rhs, precision, r_norm)     # counters persist
        assert np.array_equal(r["vx"][:ro.system.n_rows], r2["vx"][:ro.system.n_rows])
        mg = c.pc()
        out = dict(cvg=int(r["cvg"]), n_iter=r["n_iter"], residual=r["residual"],
                   hist=r["history"], vx=r["vx"][:ro.system.n_rows].copy(),
                   n_levels=mg.n_levels(), merge_level=mg.merge_level(), launches=launches)
        pc.barrier()
        c.destroy(); ro.destroy()
        return out

    res = run_ranks_in_threads(R, fn, device_of_rank_fn())
    for r, x in enumerate(res):
        assert x["cvg"] == o["cvg"] == 1, (x["cvg"], o["cvg"])
        assert x["launches"] <= 3, x["launches"]
        assert (x["n_iter"], x["residual"]) == (res[0]["n_iter"], res[0]["residual"])
        assert abs(x["n_iter"] - o["n_iter"]) <= 1, (x["n_iter"], o["n_iter"])
        k = min(len(x["hist"]), len(o["history"]), 6)
        assert np.allclose(x["hist"][:k], o["history"][:k], rtol=1e-9, atol=0)
        if merge_mean > 0 and R > 1:
            assert x["merge_level"] >= 1, x["merge_level"]
            assert x["n_levels"] == (nl if r == 0 else x["merge_level"] + 1), (r, x["n_levels"], nl)
        else:
            assert x["n_levels"] == nl and x["merge_level"] == -1
    vx = gather_global([x["vx"] for x in res], [p[2] for p in parts], g.n_rows)
    res_true = np.linalg.norm(ora.spmv(g.row_index, g.col_id, g.d_val, g.x_val, vx) - g.rhs)
    assert res_true <= 1.01 * precision * r_norm, res_true / r_norm
    vxo = gather_global(o["vx"], [p[2] for p in parts], g.n_rows)
    erro = np.linalg.norm(vx - vxo) / np.linalg.norm(vxo)
    err1 = np.linalg.norm(vx - o1["vx"]) / np.linalg.norm(o1["vx"])
    tol = 1e-10 if precision <= 1e-12 else 1e-6
    assert erro <= tol, erro
    if bc == "dirichlet":
        assert err1 <= tol, err1
    print(f"OK fcgmg N={N} R={R} {bc} merge_mean={merge_mean}: levels {res[0]['n_levels']}, "
          f"merge level {res[0]['merge_level']}, n_iter {res[0]['n_iter']} (oracle {o['n_iter']}, "
          f"single-rank {o1['n_iter']}), rel diff vs partitioned oracle {erro:.2e}, "
          f"vs single-rank {err1:.2e}")
    ora.mg_free(omg)
    D.free()


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "comm":
        scenario_comm(int(sys.argv[2]))
    elif what == "pcg":
        scenario_pcg(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
    elif what == "fcgmg":
        scenario_fcgmg(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], float(sys.argv[5]),
                       int(sys.argv[6]) if len(sys.argv) > 6 else 0)
    elif what == "mg":
        scenario_mg(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4],
                    float(sys.argv[5]) if len(sys.argv) > 5 else 1e-8)
    else:
        raise SystemExit(f"unknown scenario {what}")
