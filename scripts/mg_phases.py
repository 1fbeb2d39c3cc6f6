#!/usr/bin/env python
"""Per-level / per-phase breakdown of one multigrid solve (device-side timers of the
persistent V-cycle kernel) against each phase's algorithmic HBM bytes (DESIGN.md,
"algorithmic bytes").  Usage: python scripts/mg_phases.py [--cube 256] [--out file.md]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


from code_saturne_b200.traffic import level_sizes, phase_bytes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", type=int, default=256)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson
    import torch

    api.init(0)
    s = cube_poisson(args.cube)
    r_norm = float(np.linalg.norm(s.rhs))
    m = api.Matrix.from_system(s)
    mg = api.Multigrid()
    mg.setup(m)
    rhs_d = torch.from_numpy(s.rhs).cuda()
    vx_d = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
    for _ in range(2):
        r = mg.solve(m, rhs_d, 1e-8, r_norm, vx=vx_d)
    pt = mg.phase_times()
    n_cycles = r["n_cycles"]
    info = level_sizes(mg)
    peak = 6543.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    names = ["descent", "residual", "restrict", "coarsest", "prolong", "ascent"]
    lines = [f"# multigrid phase breakdown, {args.cube}^3, {n_cycles} cycles, "
             f"kernel time {pt.sum()*1e3:.1f} ms ({pt.sum()/n_cycles*1e3:.3f} ms/cycle); "
             f"peak {peak:.0f} GB/s (measured copy)", "",
             "| level | rows | nnz | " + " | ".join(f"{n} us/cyc (GB/s, frac)" for n in names)
             + " | level total us/cyc |", "|---|---|---|" + "---|" * 7]
    coarsest = len(info) - 1
    tot_bytes = 0.0
    for l, (n, nnz) in enumerate(info):
        nc = info[l + 1][0] if l < coarsest else 0
        b = phase_bytes(n, nnz, nc, 2, 10, l, coarsest)
        tot_bytes += b.sum()
        cells = []
        for k in range(6):
            us = pt[l, k] / n_cycles * 1e6
            if us <= 0:
                cells.append("-")
            elif b[k] > 0:
                gbs = b[k] / (pt[l, k] / n_cycles) / 1e9
                cells.append(f"{us:.1f} ({gbs:.0f}, {gbs/peak:.2f})")
            else:
                cells.append(f"{us:.1f}")
        lines.append(f"| {l} | {n} | {nnz} | " + " | ".join(cells)
                     + f" | {pt[l].sum()/n_cycles*1e6:.1f} |")
    ach = tot_bytes / (pt.sum() / n_cycles) / 1e9
    lines += ["", f"algorithmic bytes per cycle (levels above the coarsest): {tot_bytes/1e9:.2f} GB; "
              f"whole-cycle achieved {ach:.0f} GB/s = {ach/peak:.3f} of measured peak"]
    # inside the PCG smoothers: residual SpMV / SpMV + 3 dots / vector updates
    stt = mg.smoother_times()
    lines += ["", "| level | r = A.vx - rhs: us/cyc (us each, GB/s, frac) | A.p + 3 dots: us/cyc (us each, "
              "GB/s, frac) | vector updates us/cyc (GB/s, frac) |", "|---|---|---|---|"]
    for l, (n, nnz) in enumerate(info):
        if l == coarsest:
            break
        M = 12.0 * nnz + 12.0 * n
        n0, n1 = 1, (2 if l == 0 else 12)
        b0, b1 = M + 32 * n, M + 24 * n
        upd = (64 + 32) * n if l == 0 else (40 + 64 + 32 + 64 + 8 * 96 + 32) * n
        t0, t1, t2 = (stt[l, k] / n_cycles for k in range(3))
        def cell(t, cnt, b):
            if t <= 0:
                return "-"
            g = cnt * b / t / 1e9
            return f"{t*1e6:.1f} ({t/cnt*1e6:.1f}, {g:.0f}, {g/peak:.2f})"
        g2 = upd / t2 / 1e9 if t2 > 0 else 0.0
        lines.append(f"| {l} | {cell(t0, n0, b0)} | {cell(t1, n1, b1)} | {t2*1e6:.1f} ({g2:.0f}, {g2/peak:.2f}) |")
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
