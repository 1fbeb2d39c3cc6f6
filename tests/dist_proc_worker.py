"""One process per GPU (torchrun): the distributed multigrid solve over cudaIpc peer memory,
checked on rank 0 against the multi-rank oracle on the same partition.

    python -m torch.distributed.run --nproc-per-node R --master-addr 127.0.0.1 \
        tests/dist_proc_worker.py N bc [precision]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fcg_mg(N, R, rank, bc, precision, merge_glob, r_norm, s, gids, pc, ro):
    """FCG + V-cycle(RJ3) preconditioner, coarse levels under merge_glob global rows merged
    onto rank 0, against the partitioned oracle (the merged run is the same algorithm)."""
    import torch.distributed as dist
    from code_saturne_b200 import api
    from code_saturne_b200.dist import gather_global
    from code_saturne_b200.mesh_cartesian import cube_poisson, cube_poisson_local
    c = api.SlesIt(api.SlesItType.FCG, 200, poly_degree=-1)
    mgp = api.Multigrid(pc=True)
    mgp.set_aggregation_blocks(1)
    c.set_pcomm(pc)
    mgp.set_pcomm(pc)
    if merge_glob > 0:
        mgp.set_merge_options(R, 0, merge_glob)
    c.transfer_pc(mgp)
    c.setup(ro.matrix)
    r = c.solve(ro.matrix, s.rhs, precision, r_norm)
    r2 = c.solve(ro.matrix, s.rhs, precision, r_norm)
    assert np.array_equal(r["vx"][:s.n_rows], r2["vx"][:s.n_rows])
    mine = dict(cvg=int(r["cvg"]), n_iter=r["n_iter"], residual=r["residual"],
                merge_level=c.pc().merge_level(), n_levels=c.pc().n_levels(),
                vx=r["vx"][:s.n_rows].copy(), gids=gids[:s.n_rows].copy())
    out = [None] * R
    dist.all_gather_object(out, mine)
    dist.barrier()
    c.destroy(); ro.destroy(); pc.destroy()
    if rank == 0:
        from oracle.dist_oracle import DistOracle
        from oracle.pyoracle import Oracle
        ora = Oracle()
        g = cube_poisson(N, bc)
        parts = [cube_poisson_local(N, R, q, bc) for q in range(R)]
        D = DistOracle([p[0] for p in parts], [p[1] for p in parts], ora)
        D.mg_setup()
        o = D.fcg_mgpc([p[0].rhs for p in parts], 200, precision, r_norm)
        for x in out:
            assert x["cvg"] == o["cvg"] == 1
            assert (x["n_iter"], x["residual"]) == (out[0]["n_iter"], out[0]["residual"])
            assert abs(x["n_iter"] - o["n_iter"]) <= 1, (x["n_iter"], o["n_iter"])
        if merge_glob > 0:
            assert out[0]["merge_level"] >= 1 and out[1]["n_levels"] == out[1]["merge_level"] + 1
        vx = gather_global([x["vx"] for x in out], [x["gids"] for x in out], g.n_rows)
        res_true = np.linalg.norm(ora.spmv(g.row_index, g.col_id, g.d_val, g.x_val, vx) - g.rhs)
        assert res_true <= 1.01 * precision * r_norm, res_true / r_norm
        vxo = gather_global(o["vx"], [p[2] for p in parts], g.n_rows)
        err = np.linalg.norm(vx - vxo) / np.linalg.norm(vxo)
        assert err <= 1e-6, err
        print(f"OK proc-fcgmg N={N} R={R} {bc}: levels {out[0]['n_levels']}, merge level "
              f"{out[0]['merge_level']}, n_iter {out[0]['n_iter']} (oracle {o['n_iter']}), "
              f"rel diff {err:.2e}")
        D.free()


def nccl_halo(N, R, rank, bc, s, h, gids):
    """cs_halo_sync_var over the reference-shaped NCCL path (csb_comm_init_rank +
    csb_halo_create(nccl_comm): pack kernel, grouped ncclSend / ncclRecv; cs_halo.cpp:518-587)
    and the distributed SpMV that runs on it, against the global product."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson
    from oracle.pyoracle import Oracle
    L = api.load_library()
    uid = (C.c_char * 128)()
    if rank == 0:
        assert L.csb_comm_get_unique_id(C.cast(uid, C.c_void_p)) == 0, L.csb_last_error()
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)
    uid = (C.c_char * 128).from_buffer_copy(box[0])
    comm = L.csb_comm_init_rank(C.cast(uid, C.c_void_p), R, rank)
    assert comm, L.csb_last_error()
    halo = api.Halo(h, rank, nccl_comm=comm)
    m = api.Matrix.from_system(s)
    m.set_halo(halo)
    n_ext = s.meta.get("n_cols_ext", s.n_rows)
    g = cube_poisson(N, bc)
    xg = np.sin(0.37 * np.arange(g.n_rows) + 0.2)
    x = torch.zeros(n_ext, dtype=torch.float64, device="cuda")
    x[:s.n_rows] = torch.from_numpy(xg[gids[:s.n_rows]]).cuda()
    halo.sync_var(x)
    api.synchronize()
    assert np.array_equal(x.cpu().numpy(), xg[gids[:n_ext]]), "ghost values after the NCCL exchange"
    x[s.n_rows:] = 0.
    y = torch.empty(s.n_rows, dtype=torch.float64, device="cuda")
    m.vector_multiply(x, y)                      # syncs the halo first, like the reference
    api.synchronize()
    # the local rows sum their entries in local column order (ghosts last): the oracle on the
    # local matrix with the exchanged ghost values is the bit-exact reference; the global
    # product agrees to rounding
    yl = Oracle().spmv(s.row_index, s.col_id, s.d_val, s.x_val, xg[gids[:n_ext]])
    assert np.array_equal(y.cpu().numpy(), yl), "distributed SpMV over NCCL"
    yg = Oracle().spmv(g.row_index, g.col_id, g.d_val, g.x_val, xg)
    assert np.allclose(yl, yg[gids[:s.n_rows]], rtol=1e-13, atol=1e-15)
    ok = [None] * R
    dist.all_gather_object(ok, True)
    m.destroy(); halo.destroy()
    L.csb_comm_destroy(C.c_void_p(comm))
    if rank == 0:
        print(f"OK proc-nccl N={N} R={R} {bc}: halo exchange and SpMV over ncclSend/ncclRecv bit-exact")


def main():
    import torch
    import torch.distributed as dist
    from code_saturne_b200 import api
    from code_saturne_b200.dist import RankObjects, gather_global, pcomm_from_torch_dist
    from code_saturne_b200.mesh_cartesian import cube_poisson, cube_poisson_local

    N, bc = int(sys.argv[1]), sys.argv[2]
    precision = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8
    mode = sys.argv[4] if len(sys.argv) > 4 else "mg"
    merge_glob = int(sys.argv[5]) if len(sys.argv) > 5 else 3000
    rank, R = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.init(local)

    s, h, gids = cube_poisson_local(N, R, rank, bc)
    rr = torch.tensor([float(np.dot(s.rhs, s.rhs))], dtype=torch.float64, device="cuda")
    dist.all_reduce(rr)
    r_norm = float(torch.sqrt(rr[0]).item())
    if mode == "nccl":
        nccl_halo(N, R, rank, bc, s, h, gids)
        dist.destroy_process_group()
        return
    pc = pcomm_from_torch_dist(h.n_ghost)
    assert np.array_equal(pc.allreduce_sum([rank + 1.0, 1.0]), [R * (R + 1) / 2.0, float(R)])
    ro = RankObjects(s, h, gids, pc, rank)
    if mode == "fcgmg":
        fcg_mg(N, R, rank, bc, precision, merge_glob, r_norm, s, gids, pc, ro)
        dist.destroy_process_group()
        return
    mg = ro.multigrid()
    mg.setup(ro.matrix)
    r = mg.solve(ro.matrix, s.rhs, precision, r_norm)
    r2 = mg.solve(ro.matrix, s.rhs, precision, r_norm)          # flags and epochs persist
    assert np.array_equal(r["vx"][:s.n_rows], r2["vx"][:s.n_rows])
    mine = dict(cvg=int(r["cvg"]), n_cycles=r["n_cycles"], n_iter=r["n_iter"],
                residual=r["residual"], vx=r["vx"][:s.n_rows].copy(), gids=gids[:s.n_rows].copy())
    out = [None] * R
    dist.all_gather_object(out, mine)
    dist.barrier()
    mg.destroy(); ro.destroy(); pc.destroy()
    if rank == 0:
        from oracle.dist_oracle import DistOracle
        from oracle.pyoracle import Oracle
        ora = Oracle()
        g = cube_poisson(N, bc)
        parts = [cube_poisson_local(N, R, q, bc) for q in range(R)]
        D = DistOracle([p[0] for p in parts], [p[1] for p in parts], ora)
        D.mg_setup()
        o = D.mg_solve([p[0].rhs for p in parts], precision, r_norm)
        for x in out:
            assert x["cvg"] == o["cvg"] == 1
            assert (x["n_cycles"], x["n_iter"], x["residual"]) == \
                (out[0]["n_cycles"], out[0]["n_iter"], out[0]["residual"])
            assert abs(x["n_cycles"] - o["n_cycles"]) <= 1, (x["n_cycles"], o["n_cycles"])
        vx = gather_global([x["vx"] for x in out], [x["gids"] for x in out], g.n_rows)
        res_true = np.linalg.norm(ora.spmv(g.row_index, g.col_id, g.d_val, g.x_val, vx) - g.rhs)
        assert res_true <= 1.0001 * precision * r_norm, res_true / r_norm
        vxo = gather_global(o["vx"], [p[2] for p in parts], g.n_rows)
        err = np.linalg.norm(vx - vxo) / np.linalg.norm(vxo)
        assert err <= (1e-7 if bc == "dirichlet" else 1e-5), err
        print(f"OK proc-mg N={N} R={R} {bc}: cycles {out[0]['n_cycles']} (oracle {o['n_cycles']}), "
              f"rel diff {err:.2e}")
        D.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
