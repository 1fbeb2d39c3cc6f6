#!/usr/bin/env python
"""Device times of the BASELINE.json configurations other than the headline one, with the
reference's own CPU code (oracle/_ref/libcsref.so, all host cores) timed beside them on the
same systems.  CUDA events around the C-ABI solve call with rhs / vx resident in HBM.

    python scripts/config_times.py [--out profiles/r01/config_times.md]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    import torch
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion, cube_poisson
    from oracle.pyoracle import RefLib, ref_available

    api.init(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    api.set_stream(stream.cuda_stream)
    ref = RefLib(os.cpu_count() or 1) if (ref_available() and not args.no_ref) else None
    rows = []

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / n, r

    def krylov(label, N, bc, kind, ref_kind, nmax, prec):
        s = cube_poisson(N, bc, with_mesh=False)
        r_norm = float(np.linalg.norm(s.rhs))
        m = api.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
        c = api.SlesIt(kind, nmax)
        rhs_d = torch.from_numpy(s.rhs).cuda()
        vx_d = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
        t, r = timed(lambda: c.solve(m, rhs_d, prec, r_norm, vx=vx_d))
        t_ref, it_ref = None, None
        if ref is not None:
            h = ref.system(s, with_mesh=False)
            a = ref.sles_it(h, ref_kind, s.rhs, nmax, prec, r_norm)
            t_ref, it_ref = a["t_solve"], a["n_iter"]
            ref.destroy(h)
        rows.append((label, s.n_rows, r["n_iter"], t, it_ref, t_ref))
        c.destroy(); m.destroy()

    # config 1: 50^3 cavity-size pressure Poisson, multigrid
    def multigrid(label, N):
        s = cube_poisson(N, "neumann")
        r_norm = float(np.linalg.norm(s.rhs))
        m = api.Matrix.from_system(s)
        mg = api.Multigrid()
        mg.setup(m)
        rhs_d = torch.from_numpy(s.rhs).cuda()
        vx_d = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
        t, r = timed(lambda: mg.solve(m, rhs_d, 1e-8, r_norm, vx=vx_d))
        t_ref, it_ref = None, None
        if ref is not None:
            h = ref.system(s)
            a = ref.multigrid(h, s.rhs, 1e-8, r_norm)
            t_ref, it_ref = a["t_solve"], a["n_iter"]
            ref.destroy(h)
        rows.append((label, s.n_rows, r["n_iter"], t, it_ref, t_ref))
        mg.destroy(); m.destroy()

    multigrid("config 1: 50^3 Poisson, V-cycle multigrid (n_equiv_iter)", 50)
    krylov("config 2: 128^3 Poisson, PCG + Jacobi", 128, "dirichlet", api.SlesItType.PCG, 0, 5000, 1e-8)
    krylov("128^3 Poisson, FCG + Jacobi", 128, "dirichlet", api.SlesItType.FCG, 1, 5000, 1e-8)
    krylov("128^3 Poisson, IPCG + Jacobi", 128, "dirichlet", api.SlesItType.IPCG, 2, 5000, 1e-8)

    # config 4 in the small: 3x3 block-diagonal non-symmetric system, BiCGStab
    N = 96
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(N)
    n = len(ri) - 1
    r_norm = float(np.linalg.norm(rhs))
    m = api.Matrix.from_block(n, 3, ri, ci, d.ravel(), xv)
    c = api.SlesIt(api.SlesItType.BICGSTAB, 2000)
    rhs_d = torch.from_numpy(rhs).cuda()
    vx_d = torch.zeros(3 * n, dtype=torch.float64, device="cuda")
    t, r = timed(lambda: c.solve(m, rhs_d, 1e-8, r_norm, vx=vx_d))
    x_d = torch.from_numpy(np.cos(0.1 * np.arange(3 * n))).cuda()
    y_d = torch.empty_like(x_d)
    ts, _ = timed(lambda: m.vector_multiply(x_d, y_d), n=20)
    b3_bytes = 12.0 * len(ci) + 4.0 * n + 72.0 * n + 24.0 * n + 24.0 * n
    t_ref, it_ref = None, None
    if ref is not None:
        h = ref.system_block(n, 3, ri, ci, d.ravel(), xv, False)
        a = ref.sles_it(h, ref.BICGSTAB, rhs, 2000, 1e-8, r_norm)
        t_ref, it_ref = a["t_solve"], a["n_iter"]
        ref.destroy(h)
    rows.append((f"config 4 (small): {N}^3 x 3 block system, BiCGStab + Jacobi", n, r["n_iter"], t,
                 it_ref, t_ref))
    c.destroy(); m.destroy()

    cores = os.cpu_count()
    lines = ["# Other BASELINE configurations: device solve time vs the reference's CPU code",
             "",
             f"rhs / vx resident in HBM, CUDA events, mean of 3 solves after 1 warm-up; reference = "
             f"oracle/_ref/libcsref.so with {cores} OpenMP threads on the same box.", "",
             "| case | rows | iterations (B200) | s/solve (B200) | iterations (reference) | "
             "s/solve (reference CPU) | ratio |", "|---|---|---|---|---|---|---|"]
    for label, nr, it, t, itr, tr in rows:
        ratio = f"{tr/t:.1f}x" if tr else "-"
        lines.append(f"| {label} | {nr} | {it} | {t:.5f} | {itr if itr is not None else '-'} | "
                     f"{tr if tr is not None else float('nan'):.4f} | {ratio} |")
    lines += ["", f"3x3 block SpMV (k_spmv_b3, {N}^3 rows): {ts*1e3:.3f} ms per launch = "
              f"{b3_bytes/ts/1e9:.0f} GB/s algorithmic ({b3_bytes/1e6:.0f} MB)"]
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
