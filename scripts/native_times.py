"""Device / host times of the rows added late in round 2, beside the MSR entry on the same cube:
* hierarchy setup + FCG + V-cycle(RJ3) solve of a matrix entered through its face arrays
  (csb_matrix_create_native: face-list aggregation, face-order coarse values);
* matrix assembler: structure, first assembly (values_add_g, host planning) and re-assembly of
  the declared sequence from values resident in HBM (one fold kernel), with its algorithmic
  bandwidth (8 B value + 4 B plan entry per coefficient, 4 B index + 16 B read-modify-write per
  entry)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from code_saturne_b200 import api
from code_saturne_b200.mesh_cartesian import cube_poisson

ap = argparse.ArgumentParser()
ap.add_argument("--cubes", type=int, nargs="+", default=[128, 256])
ap.add_argument("--asm-cube", type=int, default=128)
a = ap.parse_args()
api.init(0)


def fcg_mg(m, s, blocks):
    r_norm = float(np.linalg.norm(s.rhs))
    c = api.SlesIt(api.SlesItType.FCG, 200, poly_degree=-1)
    pc = api.Multigrid(pc=True)
    pc.set_aggregation_blocks(blocks)
    c.transfer_pc(pc)
    t0 = time.perf_counter(); c.setup(m); api.synchronize(); ts = time.perf_counter() - t0
    c.free()
    t0 = time.perf_counter(); c.setup(m); api.synchronize(); ts2 = time.perf_counter() - t0
    r = c.solve(m, s.rhs, 1e-8, r_norm)
    t0 = time.perf_counter()
    r = c.solve(m, s.rhs, 1e-8, r_norm)
    t1 = time.perf_counter() - t0
    n_lev = -1
    c.destroy()
    return ts, ts2, t1, r, n_lev


print("| cube | entry | aggregation | setup s (first / repeated) | solve ms (e2e) | iterations | cvg |")
print("|---|---|---|---|---|---|---|")
for N in a.cubes:
    s = cube_poisson(N, "neumann")
    for entry, blocks in (("MSR", 1), ("native", 1), ("native", 16)):
        t0 = time.perf_counter()
        m = api.Matrix.from_system(s) if entry == "MSR" else api.Matrix.from_system_native(s)
        api.synchronize(); tc = time.perf_counter() - t0
        ts, ts2, t1, r, n_lev = fcg_mg(m, s, blocks)
        agg = ("face list (1 block)" if blocks == 1 else f"MSR, {blocks} blocks") if entry == "native" \
            else f"MSR, {blocks} block"
        print(f"| {N}^3 | {entry} (create {tc:.2f} s) | {agg} | {ts:.3f} / {ts2:.3f} | {t1 * 1e3:.1f} | "
              f"{r['n_iter']} | {int(r['cvg'])} |", flush=True)
        m.destroy()

# ---- assembler
N = a.asm_cube
s = cube_poisson(N, "dirichlet")
f = s.face_cells.astype(np.int64)
i, j = f[:, 0], f[:, 1]
rows = np.stack([i, i, j, j], 1).ravel()
cols = np.stack([i, j, i, j], 1).ravel()
rng = np.random.default_rng(7)
v = rng.random(len(f)) + 0.5
vals = np.stack([v, -v, -v, v], 1).ravel()
n = s.n_rows
t0 = time.perf_counter()
ma = api.MatrixAssembler((0, n)); ma.add_g_ids(rows, cols); ma.compute()
t_struct = time.perf_counter() - t0
m = ma.create_matrix()
t0 = time.perf_counter()
ma.values_init(m); ma.values_add_g(rows, cols, vals); ma.values_done(); api.synchronize()
t_first = time.perf_counter() - t0
vals_d = torch.from_numpy(vals).cuda()
ma.values_init(m); ma.values_add_declared(vals_d); api.synchronize()      # builds the plan
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    ma.values_init(m); ma.values_add_declared(vals_d)
api.synchronize()
t_re = (time.perf_counter() - t0) / reps
ri, ci = ma.structure()
nnz = len(ci)
byts = 12. * len(vals) + 20. * (n + nnz) + 8. * (n + nnz)      # + the zero fill of values_init
d, x = m.get_coefficients()
xv = np.sin(np.arange(n) * 0.37)
y = m.vector_multiply(xv.copy())
# same matrix through the MSR entry for a consistency check of the sums (face order = CSR order?)
print()
print(f"assembler, {N}^3 cube, {len(vals)} coefficients -> {n} rows, {nnz} extra-diagonal entries:")
print(f"* structure (host: declare + sort/unique): {t_struct:.3f} s")
print(f"* first assembly (values_add_g: host slots + bucket pass, upload, fold kernel): {t_first:.3f} s")
print(f"* re-assembly of the declared sequence, values in HBM (zero fill + fold kernel): {t_re * 1e3:.3f} ms "
      f"= {byts / t_re / 1e9:.0f} GB/s algorithmic")
ma.values_finalize(); m.destroy(); ma.destroy()
