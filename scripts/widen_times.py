#!/usr/bin/env python
"""Device times of the rows added after the first V-cycle path, against their
algorithmic bytes and the measured HBM peak: fused smoother sweeps, the single-cycle V-cycle,
FCG / PCG with the multigrid and polynomial preconditioners, Galerkin / pairwise setup, the
matrix coefficient build.  Prints a markdown table (profiles/rNN/widen_times_*.md).

    python scripts/widen_times.py [--cube 256] [--reps 5]

Times are host clocks around csb_stream_synchronize (every measured region is >= 1 ms of
device work or is repeated --reps times inside the region); no torch needed, so the same
script runs under tests/cudasim at a tiny --cube as a smoke check.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from code_saturne_b200 import api  # noqa: E402
from code_saturne_b200.mesh_cartesian import cube_boundary_face_cells, cube_poisson  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, reps):
    fn()
    api.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    api.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--precision", type=float, default=1e-8)
    args = ap.parse_args()
    api.init(0)
    peak = peak_gbs()
    N = args.cube
    s = cube_poisson(N)
    n, nnz = s.n_rows, s.nnz
    r_norm = float(np.linalg.norm(s.rhs))
    m = api.Matrix.from_system(s)
    rows = []

    def row(name, t, nbytes=None, note=""):
        gbs = (nbytes / t / 1e9) if nbytes else None
        rows.append((name, t * 1e3, gbs, (gbs / peak) if gbs else None, note))

    # ---- smoother sweeps: n_sweeps fused passes (+ the elementwise first sweep from zero)
    sweep_bytes = 12.0 * nnz + 4.0 * (n + 1) + 8.0 * n * 5     # x gather, b, 1/d, x_old, write
    for name, st, nit in (("Jacobi x3", api.SlesItType.JACOBI, 3), ("RJ3 x3", api.SlesItType.RJ3, 3),
                          ("L1 Jacobi x3", api.SlesItType.L1_JACOBI, 3)):
        sm = api.Smoother(st, nit)
        sm.setup(m)
        vx = np.zeros(n)
        t = timed(lambda: sm.solve(m, s.rhs, vx=vx, use_vx_ini=True), args.reps)
        row(f"smoother {name} (host rhs/vx staged: e2e)", t, nit * sweep_bytes,
            "includes H2D/D2H of rhs, vx")
        sm.destroy()

    # ---- single-cycle V-cycle and the solvers built on it
    mg = api.Multigrid(pc=True)
    t0 = time.perf_counter(); mg.setup(m); api.synchronize(); t_setup = time.perf_counter() - t0
    row("V-cycle hierarchy setup (DX aggregation, with-faces values)", t_setup)
    t = timed(lambda: mg.solve(m, s.rhs, args.precision, r_norm), args.reps)
    row("one V-cycle, RJ3 3+3 sweeps (csb_multigrid_solve, n_max_cycles 1; e2e)", t,
        None, f"{mg.n_levels()} levels")
    mg.destroy()

    # V-cycle in solver mode with RJ3 smoothers (cycles to convergence, kernel per phase)
    g = api.Multigrid()
    g.set_solver_options(descent=api.SlesItType.RJ3, ascent=api.SlesItType.RJ3, n_max_cycles=100,
                         n_max_iter_descent=3, n_max_iter_ascent=3, precision_mult=(-1., -1., 1.))
    g.setup(m)
    out = {}
    t = timed(lambda: out.update(g.solve(m, s.rhs, args.precision, r_norm)), max(1, args.reps // 2))
    row(f"V-cycle solver mode, RJ3 3+3 sweeps, solve to {args.precision:g} (e2e)", t, None,
        f"{out['n_cycles']} cycles, cvg {int(out['cvg'])}")
    g.destroy()

    # FCG + K-cycle (pairwise aggregation, Galerkin, 2 Jacobi sweeps)
    c = api.SlesIt(api.SlesItType.FCG, 500, poly_degree=-1)
    c.transfer_pc(api.Multigrid(pc=True, mg_type=1))
    t0 = time.perf_counter(); c.setup(m); api.synchronize(); t_s = time.perf_counter() - t0
    out = {}
    t = timed(lambda: out.update(c.solve(m, s.rhs, args.precision, r_norm)), 1)
    row("K-cycle hierarchy setup (pairwise, limit 4, Galerkin; host)", t_s)
    row(f"FCG + K-cycle solve to {args.precision:g} (e2e)", t, None,
        f"{out['n_iter']} iterations, cvg {int(out['cvg'])}")
    c.destroy()

    for label, T, kw in (("FCG + V-cycle(RJ3)", api.SlesItType.FCG, dict(poly_degree=-1)),
                         ("PCG + V-cycle(RJ3)", api.SlesItType.PCG, dict(poly_degree=-1)),
                         ("FCG + poly(1)", api.SlesItType.FCG, dict(poly_degree=1)),
                         ("PCG + poly(2)", api.SlesItType.PCG, dict(poly_degree=2))):
        if "poly(" in label and N > 128:
            continue          # thousands of iterations at this size: not a production choice
        c = api.SlesIt(T, 10000, **kw)
        if "V-cycle" in label:
            c.transfer_pc(api.Multigrid(pc=True))
        c.setup(m)
        out = {}
        t = timed(lambda: out.update(c.solve(m, s.rhs, args.precision, r_norm)), max(1, args.reps // 2))
        row(f"{label} solve to {args.precision:g} (e2e)", t, None,
            f"{out['n_iter']} iterations, cvg {int(out['cvg'])}, launches {api.launch_count(True)}")
        c.destroy()

    # ---- Galerkin / pairwise hierarchies
    for label, opts in (("Galerkin (DX aggregation)", dict(p0p1_relax=0.)),
                        ("pairwise, limit 4, Galerkin", dict(aggregation_limit=4, coarsening_type=3,
                                                             n_max_levels=10, p0p1_relax=0.))):
        g = api.Multigrid()
        g.set_coarsening_options(**opts)
        t0 = time.perf_counter(); g.setup(m); api.synchronize(); t_s = time.perf_counter() - t0
        out = {}
        t = timed(lambda: out.update(g.solve(m, s.rhs, args.precision, r_norm)), 1)
        row(f"{label}: setup", t_s, None, f"{g.n_levels()} levels")
        row(f"{label}: V-cycle solve to {args.precision:g} (e2e)", t, None,
            f"{out['n_cycles']} cycles, cvg {int(out['cvg'])}")
        g.destroy()

    # ---- coefficient build
    bfc = cube_boundary_face_cells(N)
    rng = np.random.default_rng(5)
    iv, bv = 0.5 + rng.random(s.n_faces), 0.5 + rng.random(len(bfc))
    cb, ro = rng.random(len(bfc)), 1e-3 * rng.random(n)
    m.set_boundary_faces(bfc)
    t = timed(lambda: m.compute_coeffs(ro, iv, bv, cb), args.reps)
    row("cs_matrix_compute_coeffs on device (host face arrays: e2e)", t,
        4.0 * (n + 1) + 12.0 * nnz + 8.0 * n + 8.0 * nnz + 8.0 * n, "includes H2D of 8 B/face")
    t = timed(lambda: m.set_coefficients(s.d_val, s.x_val), args.reps)
    row("csb_matrix_set_coefficients (upload of d_val, x_val) for comparison", t)

    print(f"# widen_times: {N}^3 cube ({n} rows), peak {peak:.0f} GB/s\n")
    print("| measurement | ms | GB/s (algorithmic) | of peak | note |")
    print("|---|---:|---:|---:|---|")
    for name, ms, gbs, frac, note in rows:
        print(f"| {name} | {ms:.3f} | {'' if gbs is None else f'{gbs:.0f}'} | "
              f"{'' if frac is None else f'{frac:.2f}'} | {note} |")
    m.destroy()


if __name__ == "__main__":
    main()
