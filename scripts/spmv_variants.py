"""Stand-alone SpMV kernels timed side by side on the same cube (one process per variant: the
choice is read once per process): k_spmv_rows (batches of 8 / of 4, plain / streaming matrix
loads), the bulk-copy pipeline, the shared-memory streaming kernel -- and the reference's own
`_mat_vect_p_l_msr` compiled for sm_100a (oracle/_ref/libref_spmv_cuda.so) when present."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [("k_spmv_rows b8", {}), ("k_spmv_rows b4 (32 regs)", {"CSB_SPMV_ROWS_B4": "1"}),
            ("k_spmv_rows b8 .cs", {"CSB_SPMV_ROWS_CS": "1"}),
            ("k_spmv_rows b4 .cs", {"CSB_SPMV_ROWS_B4": "1", "CSB_SPMV_ROWS_CS": "1"}),
            ("k_spmv_pipeline", {"CSB_SPMV_KERNEL": "pipeline"}),
            ("k_spmv_stream", {"CSB_SPMV_KERNEL": "stream"})]


def child(cube, launches):
    import numpy as np
    import torch
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson
    api.init(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    api.set_stream(stream.cuda_stream)        # the library launches on this stream
    s = cube_poisson(cube, "neumann", with_mesh=False)
    m = api.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)
    x = torch.from_numpy(np.sin(np.arange(s.n_rows) * 0.37)).cuda()
    y = torch.empty(s.n_rows, dtype=torch.float64, device="cuda")

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(launches):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / launches

    out = {"kernel": api.load_library().csb_spmv_kernel_name().decode(),
           "ms": timed(lambda: m.vector_multiply(x, y)), "bytes": m.spmv_bytes()}
    if os.environ.get("SPMV_REF") == "1":
        p = os.path.join(ROOT, "oracle", "_ref", "libref_spmv_cuda.so")
        if os.path.exists(p):
            RL = C.CDLL(p, mode=os.RTLD_LAZY)
            t = lambda a: C.c_void_p(a.data_ptr())
            ri = torch.from_numpy(s.row_index).cuda(); ci = torch.from_numpy(s.col_id).cuda()
            dv = torch.from_numpy(s.d_val).cuda(); xv = torch.from_numpy(s.x_val).cuda()
            yr = torch.empty_like(y)
            sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            out["ref_ms"] = timed(lambda: RL.csref_cuda_spmv_msr(s.n_rows, t(ri), t(ci), t(dv), t(xv),
                                                                 t(x), t(yr), sp))
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", type=int, default=256)
    ap.add_argument("--launches", type=int, default=100)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--peak", type=float, default=6544.3)
    a = ap.parse_args()
    if a.child:
        child(a.cube, a.launches)
        sys.exit(0)
    print(f"# stand-alone y = A.x, {a.cube}^3 cube, {a.launches} launches each; peak {a.peak} GB/s (measured copy)\n")
    print("| kernel | ms per launch | GB/s | fraction of peak |\n|---|---|---|---|")
    first = True
    for name, env in VARIANTS:
        e = dict(os.environ); e.update(env)
        if first:
            e["SPMV_REF"] = "1"
        r = subprocess.run([sys.executable, __file__, "--child", "--cube", str(a.cube), "--launches",
                            str(a.launches)], env=e, capture_output=True, text=True)
        try:
            o = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(f"| {name} | failed: {r.stderr[-200:]} | | |")
            first = False
            continue
        gbs = o["bytes"] / o["ms"] / 1e6
        print(f"| {name} | {o['ms']:.4f} | {gbs:.0f} | {gbs / a.peak:.3f} |")
        if first and "ref_ms" in o:
            g = o["bytes"] / o["ref_ms"] / 1e6
            print(f"| reference `_mat_vect_p_l_msr` (cs_matrix_spmv_cuda.cu:415-439, FMA) | {o['ref_ms']:.4f} | {g:.0f} | {g / a.peak:.3f} |")
        first = False
