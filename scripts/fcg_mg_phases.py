#!/usr/bin/env python
"""Per-level / per-phase breakdown of one FCG + V-cycle(RJ3) solve (device-side timers of
k_fcg_mg_persistent) against each phase's algorithmic HBM bytes (code_saturne_b200/traffic.py).

    python scripts/fcg_mg_phases.py [--cube 256] [--out file.md] [--solves 2] [--precision 1e-8]

With --solves N the script is also what an `ncu -k regex:k_fcg_mg_persistent` capture runs.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from code_saturne_b200.traffic import fcg_mg_iter_bytes, fcg_mg_phase_bytes, level_sizes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", type=int, default=256)
    ap.add_argument("--out", default=None)
    ap.add_argument("--solves", type=int, default=2)
    ap.add_argument("--precision", type=float, default=1e-8)
    ap.add_argument("--max-iter", type=int, default=200)
    args = ap.parse_args()

    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson
    import torch

    api.init(0)
    s = cube_poisson(args.cube)
    r_norm = float(np.linalg.norm(s.rhs))
    m = api.Matrix.from_system(s)
    c = api.SlesIt(api.SlesItType.FCG, args.max_iter, poly_degree=-1)
    c.transfer_pc(api.Multigrid(pc=True))
    c.setup(m)
    mg = c.pc()
    rhs_d = torch.from_numpy(s.rhs).cuda()
    vx_d = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.solves):
        torch.cuda.synchronize()
        e0.record()
        r = c.solve(m, rhs_d, args.precision, r_norm, vx=vx_d)
        e1.record()
        torch.cuda.synchronize()
    t_solve = e0.elapsed_time(e1) * 1e-3
    pt = mg.phase_times()
    it = mg.smoother_times()        # [0][0] init, [0][1] A.vk + 4 dots, [0][2] update
    n_it = max(int(r["n_iter"]), 1)
    n_cyc = n_it + 1                # one more preconditioner application than iterations
    info = level_sizes(mg)
    peak = 6544.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    names = ["descent", "residual", "restrict", "coarsest", "prolong", "ascent"]
    lines = [f"# FCG + V-cycle(RJ3 x 3) phase breakdown, {args.cube}^3, {n_it} iterations "
             f"({n_cyc} V-cycles), solve {t_solve*1e3:.1f} ms (events), in-kernel timers "
             f"{(pt.sum() + it.sum())*1e3:.1f} ms; peak {peak:.0f} GB/s (measured copy)", "",
             "| level | rows | nnz | " + " | ".join(f"{n} us/cyc (GB/s, frac)" for n in names)
             + " | level total us/cyc |", "|---|---|---|" + "---|" * 7]
    coarsest = len(info) - 1
    for l, (n, nnz) in enumerate(info):
        nc = info[l + 1][0] if l < coarsest else 0
        b = fcg_mg_phase_bytes(n, nnz, nc, 3, 3, l, coarsest)
        cells = []
        for k in range(6):
            us = pt[l, k] / n_cyc * 1e6
            if us <= 0:
                cells.append("-")
            elif b[k] > 0:
                gbs = b[k] / (pt[l, k] / n_cyc) / 1e9
                cells.append(f"{us:.1f} ({gbs:.0f}, {gbs/peak:.2f})")
            else:
                cells.append(f"{us:.1f}")
        lines.append(f"| {l} | {n} | {nnz} | " + " | ".join(cells)
                     + f" | {pt[l].sum()/n_cyc*1e6:.1f} |")
    n0, nnz0 = info[0]
    b_spmv = 12.0 * nnz0 + 4.0 * (n0 + 1) + 40.0 * n0
    b_upd = 96.0 * n0
    t_spmv, t_upd = it[0, 1] / n_cyc, it[0, 2] / n_it
    lines += ["", f"FCG A.vk + 4 dots: {t_spmv*1e6:.1f} us each "
              f"({b_spmv/t_spmv/1e9:.0f} GB/s, {b_spmv/t_spmv/1e9/peak:.2f}); fused update + first "
              f"sweep: {t_upd*1e6:.1f} us each ({b_upd/t_upd/1e9:.0f} GB/s, "
              f"{b_upd/t_upd/1e9/peak:.2f}); init {it[0, 0]*1e6:.1f} us"]
    per_it = fcg_mg_iter_bytes(info)
    ach = per_it * n_it / t_solve / 1e9
    lines += ["", f"algorithmic bytes per iteration {per_it/1e9:.2f} GB; whole solve achieved "
              f"{ach:.0f} GB/s = {ach/peak:.3f} of measured peak; {t_solve/n_it*1e3:.3f} ms per "
              "iteration"]
    txt = "\n".join(lines)
    print(txt)
    print("iterations", r["n_iter"], "cvg", int(r["cvg"]), "residual", r["residual"] / r_norm)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
