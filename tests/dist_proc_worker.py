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


def main():
    import torch
    import torch.distributed as dist
    from code_saturne_b200 import api
    from code_saturne_b200.dist import RankObjects, gather_global, pcomm_from_torch_dist
    from code_saturne_b200.mesh_cartesian import cube_poisson, cube_poisson_local

    N, bc = int(sys.argv[1]), sys.argv[2]
    precision = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8
    rank, R = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.init(local)

    s, h, gids = cube_poisson_local(N, R, rank, bc)
    rr = torch.tensor([float(np.dot(s.rhs, s.rhs))], dtype=torch.float64, device="cuda")
    dist.all_reduce(rr)
    r_norm = float(torch.sqrt(rr[0]).item())
    pc = pcomm_from_torch_dist(h.n_ghost)
    assert np.array_equal(pc.allreduce_sum([rank + 1.0, 1.0]), [R * (R + 1) / 2.0, float(R)])
    ro = RankObjects(s, h, gids, pc, rank)
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
