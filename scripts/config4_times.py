#!/usr/bin/env python
"""BASELINE config 4 at full size: a ~50 M-cell 3x3 block-diagonal (velocity convection-diffusion)
system solved by BiCGStab + Jacobi on one B200, the reference's own CPU code beside it on the
same system (oracle/_ref/libcsref.so, all host cores).  The mesh is the hex cube of
mesh_cartesian (368^3 = 49.8 M cells; no tetrahedral mesh generator here: 6 neighbours per cell
instead of 4, i.e. MORE matrix entries per cell than BASELINE's tets).

    python scripts/config4_times.py [--cube 368] [--out profiles/r02/config4_times.md]
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
    ap.add_argument("--cube", type=int, default=368)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--peak", type=float, default=6544.3)
    args = ap.parse_args()
    import torch
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_block3_convection_diffusion
    from oracle.pyoracle import RefLib, ref_available

    api.init(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    api.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    N = args.cube
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(N)
    n = len(ri) - 1
    t_gen = time.perf_counter() - t0
    r_norm = float(np.linalg.norm(rhs))
    m = api.Matrix.from_block(n, 3, ri, ci, d.ravel(), xv)
    c = api.SlesIt(api.SlesItType.BICGSTAB, 2000)
    rhs_d = torch.from_numpy(rhs).cuda()
    vx_d = torch.zeros(3 * n, dtype=torch.float64, device="cuda")

    def timed(fn, k):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / k, r

    t, r = timed(lambda: c.solve(m, rhs_d, 1e-8, r_norm, vx=vx_d), 3)
    x_d = torch.from_numpy(np.cos(0.1 * np.arange(3 * n))).cuda()
    y_d = torch.empty_like(x_d)
    ts, _ = timed(lambda: m.vector_multiply(x_d, y_d), 20)
    spmv_bytes = 12.0 * len(ci) + 4.0 * n + 72.0 * n + 24.0 * n + 24.0 * n
    # BiCGStab iteration: 2 block SpMV + the vector passes of the kernel-per-phase path
    # (pk/zk 5 vectors, upd1 7, upd2 7: 19 x 24 B/cell) -- the 172 B/cell of SURVEY 8d is the SpMV
    it_bytes = 2.0 * spmv_bytes + 19.0 * 24.0 * n
    lines = [f"# BASELINE config 4 at full size: {N}^3 = {n} cells, 3x3 block-diagonal system "
             f"({len(ci)} extra-diagonal entries), BiCGStab + Jacobi to 1e-8", "",
             f"kernel for the 3x3 block SpMV: {'k_spmv_bn (thread per component)' if os.environ.get('CSB_SPMV_B3_BN') == '1' else 'k_spmv_b3 (thread per block row)'}; "
             f"peak = {args.peak} GB/s (measured copy)", "",
             f"* B200: {t:.4f} s per solve, {r['n_iter']} iterations ({t / max(r['n_iter'], 1) * 1e3:.2f} ms each, "
             f"{it_bytes * r['n_iter'] / t / 1e9:.0f} GB/s algorithmic = "
             f"{it_bytes * r['n_iter'] / t / 1e9 / args.peak:.2f} of peak), cvg {int(r['cvg'])}, "
             f"residual/r_norm {r['residual'] / r_norm:.2e}",
             f"* block SpMV alone: {ts * 1e3:.3f} ms per launch = {spmv_bytes / ts / 1e9:.0f} GB/s "
             f"({spmv_bytes / ts / 1e9 / args.peak:.2f} of peak; {spmv_bytes / n:.0f} B/cell)"]
    if ref_available() and not args.no_ref:
        ref = RefLib(os.cpu_count() or 1)
        h = ref.system_block(n, 3, ri, ci, d.ravel(), xv, False)
        a = ref.sles_it(h, ref.BICGSTAB, rhs, 2000, 1e-8, r_norm)
        ref.destroy(h)
        err = float(np.linalg.norm(vx_d.cpu().numpy() - a["vx"]) / np.linalg.norm(a["vx"]))
        lines.append(f"* reference CPU ({os.cpu_count()} OpenMP threads): {a['t_solve']:.2f} s per solve, "
                     f"{a['n_iter']} iterations, cvg {a['cvg']}; ratio {a['t_solve'] / t:.1f}x; "
                     f"solutions differ by {err:.2e} (relative L2)")
    lines.append(f"* input generation {t_gen:.0f} s (numpy, host)")
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
