#!/usr/bin/env python
"""bench.py -- multigrid pressure-Poisson solve on a cs_mesh_cartesian-style hex cube.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--cube 256]

One "step" = one cs_multigrid_solve (V-cycle, solver-mode defaults of
cs_multigrid_define: PCG+Jacobi smoothers 2/10, PCG coarse solve) of the
synthetic pressure system to precision 1e-8, hierarchy already set up (the
reference times setup and solve separately too: cs_multigrid.cpp:5324-5332,
:5583-5586; setup seconds are reported beside the headline).

Prints ONE JSON line (rank 0).  `value` = seconds per solve with rhs/vx resident
in HBM; `e2e` = the same solve through the C-ABI with HOST rhs/vx buffers
(H2D + D2H inside the timed region); `roofline` = level-0 SpMV kernel, achieved
algorithmic bytes / CUDA-event time against the measured HBM copy bandwidth;
`cpu_baseline` = the reference's own CPU code (oracle/_ref/libcsref.so) on the
box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "multigrid Poisson solve, seconds per solve"
UNIT = "s/solve"
PRECISION = 1e-8


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except Exception:
                continue
            for k, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(k)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm

def run_reference(args, sample_cube=None, quiet=False):
    """Time the reference's own CPU implementation (cs_multigrid_setup/_solve from
    oracle/_ref/libcsref.so) on the host cores, on a bounded sample of the workload."""
    from code_saturne_b200.mesh_cartesian import cube_poisson
    from oracle.pyoracle import RefLib, ref_available
    cores = os.cpu_count() or 1
    n_cube = sample_cube or args.ref_cube
    if not ref_available():
        # the oracle port always exists: single-threaded C restatement
        from oracle.pyoracle import Oracle
        ora = Oracle()
        s = cube_poisson(n_cube)
        r_norm = float(np.linalg.norm(s.rhs))
        t0 = time.perf_counter(); mg = ora.mg_setup(s); t_setup = time.perf_counter() - t0
        times = []
        for i in range(max(1, min(args.steps, 2))):
            t0 = time.perf_counter(); r = ora.mg_solve(mg, s.rhs, PRECISION, r_norm)
            times.append(time.perf_counter() - t0)
        ora.mg_free(mg)
        return dict(kind="port", cores=1, cube=n_cube, n_rows=s.n_rows, t_solve=min(times),
                    t_setup=t_setup, n_cycles=r["n_cycles"], n_iter=r["n_iter"])
    ref = RefLib(cores)
    s = cube_poisson(n_cube)
    h = ref.system(s)
    r_norm = float(np.linalg.norm(s.rhs))
    times, setups = [], []
    n_runs = args.warmup_ref + max(1, args.steps_ref)
    r = None
    for i in range(n_runs):
        r = ref.multigrid(h, s.rhs, PRECISION, r_norm)
        if i >= args.warmup_ref:
            times.append(r["t_solve"]); setups.append(r["t_setup"])
    ref.destroy(h)
    return dict(kind="reference", cores=ref.n_threads, cube=n_cube, n_rows=s.n_rows,
                t_solve=float(np.mean(times)), t_setup=float(np.mean(setups)),
                n_iter=r["n_iter"], cvg=r["cvg"], residual=r["residual"])


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = run_reference(args)
    scale = (args.cube ** 3) / float(b["n_rows"])
    value = b["t_solve"] * scale
    sample = (f"{b['cube']}^3 cube ({b['n_rows']} rows) full solve to {PRECISION:g}, "
              f"{b['cores']} OpenMP threads; seconds scaled x{scale:g} (row count) to the "
              f"{args.cube}^3 workload")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": args.warmup_ref,
        "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.cube}^3 hex cube pressure Poisson, cs_multigrid V-cycle "
                               "(PCG+Jacobi smoothers), precision 1e-8",
                   "sample_cube": b["cube"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": b["cores"], "kind": b["kind"],
                         "sample": sample, "sample_s_per_solve": b["t_solve"],
                         "sample_setup_s": b["t_setup"], "sample_n_equiv_iter": b["n_iter"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_production_solver(args):
    """Side measurement, never part of `value`: the reference's production layout (FCG
    preconditioned by one V-cycle, scripts/fcg_mg_times.py) on the same cube, in a process of
    its own after the headline numbers are taken; a failure there only shows up here."""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fcg_mg_times.py"),
                              "--cube", str(args.cube)], capture_output=True, text=True,
                             timeout=300, cwd=ROOT)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": (out.stderr.strip().splitlines() or ["no output"])[-1][:300]}
    except Exception as e:
        return {"unavailable": str(e)[:300]}


# --------------------------------------------------------------------------- our arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cube", type=int, default=256, help="cells per edge of the hex cube")
    ap.add_argument("--ref-cube", type=int, default=128, help="CPU sample size (bounded)")
    ap.add_argument("--steps-ref", type=int, default=2)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spmv-launches", type=int, default=50)
    ap.add_argument("--no-production-solver", action="store_true",
                    help="skip the FCG + multigrid-preconditioner side measurement")
    args = ap.parse_args()

    if args.impl == "reference":
        reference_main(args)
        return

    import torch
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    api.init(local_rank)
    # a dedicated (non-default) stream shared by torch events and the library
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    api.set_stream(stream.cuda_stream)

    # ---- synthetic input (host), matrix + hierarchy setup ------------------
    # N > 1: the SAME cube is partitioned into N bricks (what the reference's Morton-box
    # partitioning yields, cs_partition.cpp:2947), one brick per GPU: strong scaling.
    # Ghost exchange and dot products run inside the persistent solver kernel over
    # NVLink peer memory; torch.distributed only bootstraps.
    t0 = time.perf_counter()
    ro = None
    if world > 1:
        from code_saturne_b200.dist import RankObjects, pcomm_from_torch_dist
        from code_saturne_b200.mesh_cartesian import cube_poisson_local
        s, halo_desc, gids = cube_poisson_local(args.cube, world, rank)
        n_cols_ext = s.meta["n_cols_ext"]
    else:
        s = cube_poisson(args.cube)
        n_cols_ext = s.n_rows
    t_gen = time.perf_counter() - t0
    rr = torch.tensor([float(np.dot(s.rhs, s.rhs)), float(s.n_rows)], dtype=torch.float64,
                      device="cuda")
    if world > 1:
        dist.all_reduce(rr)
    r_norm = float(torch.sqrt(rr[0]).item())
    n_rows_global = int(rr[1].item())
    t0 = time.perf_counter()
    if world > 1:
        pc = pcomm_from_torch_dist(halo_desc.n_ghost)
        ro = RankObjects(s, halo_desc, gids, pc, rank)
        m = ro.matrix
        mg = ro.multigrid()
    else:
        m = api.Matrix.from_system(s)
        mg = api.Multigrid()
    t_upload = time.perf_counter() - t0
    t0 = time.perf_counter()
    mg.setup(m)
    t_setup = time.perf_counter() - t0
    st = mg.setup_times()
    n_levels = mg.n_levels()

    rhs_d = torch.from_numpy(s.rhs).cuda()
    vx_d = torch.zeros(n_cols_ext, dtype=torch.float64, device="cuda")
    rhs_h = torch.from_numpy(s.rhs).pin_memory()
    vx_h = torch.zeros(n_cols_ext, dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve_device():
        return mg.solve(m, rhs_d, PRECISION, r_norm, vx=vx_d)

    def solve_host():
        return mg.solve(m, rhs_h.numpy(), PRECISION, r_norm, vx=vx_h.numpy())

    for _ in range(args.warmup):
        r = solve_device()
    assert int(r["cvg"]) == 1, f"solve did not converge: {r['cvg']}"

    sampler = ClockSampler(local_rank)
    with sampler:
        # ---- value: inputs resident in HBM --------------------------------
        barrier()
        api.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            r = solve_device()
        e1.record(stream)
        barrier()
        launches = api.launch_count()
        t_dev = e0.elapsed_time(e1) * 1e-3 / args.steps

        # ---- e2e: host buffers through the C-ABI --------------------------
        for _ in range(min(args.warmup, 2)):
            solve_host()
        barrier()
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            rh = solve_host()
        e1.record(stream)
        barrier()
        t_e2e = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - w0) / args.steps

        # ---- roofline: level-0 SpMV kernel, inputs (1.7 GB) >> L2 ----------
        x_d = torch.from_numpy(np.sin(np.arange(n_cols_ext) * 0.37)).cuda()
        y_d = torch.empty(s.n_rows, dtype=torch.float64, device="cuda")
        for _ in range(5):
            m.vector_multiply(x_d, y_d)
        barrier()
        e0.record(stream)
        for _ in range(args.spmv_launches):
            m.vector_multiply(x_d, y_d)
        e1.record(stream)
        barrier()
        t_spmv = e0.elapsed_time(e1) * 1e-3 / args.spmv_launches
    clocks = sampler.summary()

    if world > 1:
        t = torch.tensor([t_dev, t_e2e, t_spmv], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_spmv = (float(v) for v in t.cpu())

    spmv_bytes = m.spmv_bytes()          # this rank's brick; all bricks are equal
    peak, peak_src = measured_peak_gbs()
    spmv_gbs = spmv_bytes / t_spmv / 1e9
    # dominant kernel = k_mg_persistent, ONE launch per solve: algorithmic bytes of the
    # cycles it ran (code_saturne_b200/traffic.py) over its CUDA-event time
    from code_saturne_b200.traffic import cycle_bytes, level_sizes
    info = level_sizes(mg)
    solve_bytes = cycle_bytes(info) * int(r["n_cycles"])       # per rank
    achieved = solve_bytes / t_dev / 1e9
    # in-solve SpMV + 3 dots phase of the finest level, from the device-side timers
    stt = mg.smoother_times()
    n0, nnz0 = info[0]
    t_sp = stt[0, 1] / max(int(r["n_cycles"]), 1) / 2.0
    in_solve_spmv = ((12.0 * nnz0 + 36.0 * n0) / t_sp / 1e9) if t_sp > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01", "ncu_mg_persistent_traffic.json")
    if world == 1 and args.cube == 256 and os.path.exists(tp):
        try:      # dram__bytes_read+write of one ncu --set full capture, per cycle
            traffic = float(json.load(open(tp))["dram_bytes_per_cycle"]) * int(r["n_cycles"])
        except Exception:
            traffic = None

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                b = run_reference(args)
                scale = s.n_rows / float(b["n_rows"])
                cpu = {"value": b["t_solve"] * scale, "unit": UNIT, "cores": b["cores"],
                       "kind": b["kind"],
                       "sample": (f"{b['cube']}^3 cube ({b['n_rows']} rows) full solve to "
                                  f"{PRECISION:g} with {b['cores']} OpenMP threads; seconds "
                                  f"scaled x{scale:g} (row count) to the {args.cube}^3 workload"),
                       "sample_s_per_solve": b["t_solve"], "sample_setup_s": b["t_setup"],
                       "sample_n_equiv_iter": b["n_iter"]}
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                       "sample": f"unavailable: {e}"}
        line = {
            "metric": METRIC, "value": t_dev, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.cube}^3 hex cube ({n_rows_global} cells) pressure Poisson, "
                                   "cs_multigrid V-cycle (PCG+Jacobi smoothers 2/10, coarse PCG), "
                                   f"precision {PRECISION:g}, all-Neumann + eps",
                       "rows_per_gpu": n_rows_global // world, "n_levels": n_levels,
                       "n_cycles": int(r["n_cycles"]), "n_equiv_iter": int(r["n_iter"]),
                       "final_residual_over_rnorm": float(r["residual"] / r_norm),
                       "l2_policy": "inputs (1.7 GB matrix) larger than L2; no flush needed",
                       "setup_s": t_setup, "setup_aggregation_host_s": st["aggregation"],
                       "setup_assembly_device_s": st["assembly"],
                       "setup_transfers_s": st["transfers"],
                       "matrix_upload_s": t_upload, "input_generation_s": t_gen,
                       "parallelism": "1 GPU" if world == 1 else
                       f"{world} bricks (Morton-box partition), one per GPU; ghost exchange + "
                       "all-reduce fused in the persistent kernel over NVLink peer memory"},
            "e2e": {"value": t_e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * n_rows_global,
                    "d2h_bytes_per_step": 8 * n_rows_global + 4096 * world},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm",
                         "kernel": "k_mg_persistent (whole V-cycle solve: all cycles and levels "
                                   "in one launch" + (", per GPU)" if world > 1 else ")"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": solve_bytes,
                         "ms_per_launch": t_dev * 1e3, "peak_source": peak_src,
                         "launches_timed": args.steps},
            "spmv": {"kernel": "k_spmv_stream (level-0 MSR SpMV, stand-alone"
                               + (", per GPU, incl. ghost exchange)" if world > 1 else ")"),
                     "achieved": spmv_gbs, "frac": spmv_gbs / peak, "unit": "GB/s",
                     "algorithmic_bytes_per_launch": spmv_bytes, "ms_per_launch": t_spmv * 1e3,
                     "in_solve_spmv_3dots_gbs": in_solve_spmv,
                     "in_solve_spmv_3dots_frac": (in_solve_spmv / peak) if in_solve_spmv else None},
            "spmv_hbm_gbs": spmv_gbs,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_production_solver:
            line["production_solver"] = run_production_solver(args)
        print(json.dumps(line))

    if world > 1:
        dist.barrier()
    mg.destroy()
    if ro is not None:
        ro.destroy()
        pc.destroy()
    else:
        m.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
