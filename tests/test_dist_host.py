"""N > 1 host logic on CPU: two gloo processes each build their own part of the cube
(the layout an MPI run of the reference holds: local rows, ghost sections per neighbour,
send lists), exchange what a halo exchange would exchange with torch.distributed, and
check it against the global system; the host-side aggregation of the product on a local
matrix with ghost columns must equal the pinned oracle's."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from code_saturne_b200 import api
        from code_saturne_b200.mesh_cartesian import cube_poisson, cube_poisson_local
        from oracle.pyoracle import Oracle

        s, h, gids = cube_poisson_local(N, world, rank, "neumann")
        g = cube_poisson(N, "neumann")
        # halo exchange of x through gloo: send my send lists, receive my ghost sections
        x = np.full(s.meta["n_cols_ext"], np.nan)
        x[:s.n_rows] = np.sin(gids[:s.n_rows] * 0.37)
        reqs, bufs = [], []
        for i, d in enumerate(h.c_domain_rank):
            snd = torch.from_numpy(x[h.send_list[h.send_index[i]:h.send_index[i + 1]]].copy())
            rcv = torch.empty(int(h.index[i + 1] - h.index[i]), dtype=torch.float64)
            reqs.append(dist.isend(snd, int(d)))
            reqs.append(dist.irecv(rcv, int(d)))
            bufs.append((i, rcv))
        for r in reqs:
            r.wait()
        for i, rcv in bufs:
            x[s.n_rows + h.index[i]:s.n_rows + h.index[i + 1]] = rcv.numpy()
        assert np.array_equal(x, np.sin(gids * 0.37))
        # local SpMV with ghosts == rows of the global SpMV (oracle) up to the order of
        # the in-row sum: ghost columns sort last locally, as in an MPI run of the reference
        ora = Oracle()
        y = ora.spmv(s.row_index, s.col_id, s.d_val, s.x_val, x)
        yg = ora.spmv(g.row_index, g.col_id, g.d_val, g.x_val, np.sin(np.arange(g.n_rows) * 0.37))
        assert np.allclose(y, yg[gids[:s.n_rows]], rtol=0, atol=4e-16 * np.abs(yg).max() * 8)
        # global sum in rank order == what every rank computes
        part = torch.tensor([ora.dot_xx(x[:s.n_rows])], dtype=torch.float64)
        allp = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allp, part)
        tot = 0.0
        for p in allp:
            tot += float(p)
        assert abs(tot - ora.dot_xx(np.sin(np.arange(g.n_rows) * 0.37))) <= 1e-13 * tot
        # product host aggregation on the local matrix (ghost columns never aggregate)
        nc, f_c_row = api.host_aggregate(0, s.row_index, s.col_id, s.d_val, s.x_val, 3, 1)
        nco, ref = ora.aggregation(0, s.row_index, s.col_id, s.d_val, s.x_val, 3, 1)
        assert nc == nco and np.array_equal(f_c_row, ref)
        counts = [None] * world
        dist.all_gather_object(counts, nc)
        assert sum(counts) < 0.5 * g.n_rows
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("N", [8, 12])
def test_two_rank_partition_and_halo_with_gloo(N):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + N) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in out:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_partition_single_rank_equals_global():
    sys.path.insert(0, ROOT)
    from code_saturne_b200.mesh_cartesian import cube_poisson, cube_poisson_local, partition_dims
    g = cube_poisson(7, "dirichlet")
    s, h, gids = cube_poisson_local(7, 1, 0, "dirichlet")
    for k in ("row_index", "col_id", "d_val", "x_val", "c2f", "c2f_sgn", "cell_cen", "cell_vol",
              "face_u_normal", "face_surf", "rhs", "face_cells"):
        assert np.array_equal(getattr(s, k), getattr(g, k)), k
    assert h.n_c_domains == 0 and h.n_ghost == 0
    assert partition_dims(8) == (2, 2, 2) and partition_dims(2) == (1, 1, 2)


def test_multirank_oracle_reduces_to_single_rank_oracle():
    """R = 1 of the multi-rank emulation is bit-identical to the pinned oracle; R = 4
    converges to the same solution."""
    sys.path.insert(0, ROOT)
    from code_saturne_b200.mesh_cartesian import cube_poisson, cube_poisson_local
    from oracle.dist_oracle import DistOracle
    from oracle.pyoracle import Oracle
    ora = Oracle()
    N, bc = 12, "dirichlet"
    g = cube_poisson(N, bc)
    r_norm = float(np.sqrt(ora.dot_xx(g.rhs)))
    omg = ora.mg_setup(g)
    o = ora.mg_solve(omg, g.rhs, 1e-12, r_norm)
    for R in (1, 4):
        parts = [cube_poisson_local(N, R, r, bc) for r in range(R)]
        D = DistOracle([p[0] for p in parts], [p[1] for p in parts], ora)
        D.mg_setup()
        res = D.mg_solve([p[0].rhs for p in parts], 1e-12, r_norm)
        vx = np.zeros(g.n_rows)
        for r, p in enumerate(parts):
            vx[p[2][:p[0].n_rows]] = res["vx"][r]
        if R == 1:
            assert np.array_equal(vx, o["vx"]) and res["n_iter"] == o["n_iter"]
        else:
            assert res["cvg"] == 1
            assert np.linalg.norm(vx - o["vx"]) <= 1e-10 * np.linalg.norm(o["vx"])
        D.free()
    ora.mg_free(omg)
