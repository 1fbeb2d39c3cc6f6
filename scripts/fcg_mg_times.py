#!/usr/bin/env python
"""Time the reference's production pressure-solver layout on one B200: cs_sles_it FCG
preconditioned by one cs_multigrid V-cycle with the device smoother settings (RJ3, 3 sweeps;
csb_multigrid_pc_create), kernel-per-phase path.

    python scripts/fcg_mg_times.py [--cube 256] [--steps 3] [--warmup 1] [--precision 1e-8]

Prints one JSON object.  bench.py runs it in a process of its own (so that nothing here can
disturb the headline measurement) and attaches the result as `production_solver`.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--precision", type=float, default=1e-8)
    args = ap.parse_args()

    import torch
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson

    torch.cuda.set_device(0)
    api.init(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    api.set_stream(stream.cuda_stream)

    s = cube_poisson(args.cube)
    r_norm = float(np.linalg.norm(s.rhs))
    m = api.Matrix.from_system(s)
    c = api.SlesIt(api.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(api.Multigrid(pc=True))
    t0 = time.perf_counter()
    c.setup(m)
    api.synchronize()
    t_setup = time.perf_counter() - t0

    rhs_d = torch.from_numpy(s.rhs).cuda()
    vx_d = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
    r = None
    for _ in range(args.warmup):
        r = c.solve(m, rhs_d, args.precision, r_norm, vx=vx_d)
    torch.cuda.synchronize()
    api.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        r = c.solve(m, rhs_d, args.precision, r_norm, vx=vx_d)
    e1.record(stream)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / args.steps
    launches = api.launch_count() // max(args.steps, 1)
    y = torch.empty(s.n_rows, dtype=torch.float64, device="cuda")
    m.vector_multiply(vx_d, y)
    torch.cuda.synchronize()
    true_res = float(torch.linalg.norm(rhs_d - y).item())
    print(json.dumps({
        "solver": "cs_sles_it FCG + cs_multigrid_pc (V-cycle, RJ3 x3 sweeps, PCG coarse), "
                  "kernel per phase",
        "cube": args.cube, "precision": args.precision, "s_per_solve": t,
        "n_iter": int(r["n_iter"]), "cvg": int(r["cvg"]),
        "residual_over_rnorm": float(r["residual"] / r_norm),
        "true_residual_over_rnorm": true_res / r_norm,
        "launches_per_solve": int(launches), "setup_s": t_setup,
        "steps": args.steps, "warmup": args.warmup}))
    c.destroy()
    m.destroy()


if __name__ == "__main__":
    main()
