"""Device times of the level-scheduled Gauss-Seidel sweeps and of FCG + V-cycle(sym GS) -- the
reference's CPU-default cs_multigrid_pc -- beside FCG + V-cycle(RJ3) on the same cube."""
import argparse
import time

import numpy as np
import torch

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from code_saturne_b200 import api
from code_saturne_b200.mesh_cartesian import cube_poisson

ap = argparse.ArgumentParser()
ap.add_argument("--cube", type=int, default=128)
a = ap.parse_args()
api.init(0)
s = cube_poisson(a.cube, "neumann")
m = api.Matrix.from_system(s)
r_norm = float(np.linalg.norm(s.rhs))
rhs_d = torch.from_numpy(s.rhs).cuda()
print(f"# {a.cube}^3, {s.n_rows} rows")
for name, st, nit in (("sym GS x1", 9, 1), ("GS x1", 8, 1), ("RJ3 x3", 16, 3)):
    sm = api.Smoother(st, nit)
    sm.setup(m)
    vx = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
    sm.solve(m, rhs_d, vx=vx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        sm.solve(m, rhs_d, vx=vx, use_vx_ini=True)
    torch.cuda.synchronize()
    print(f"smoother {name}: {(time.perf_counter() - t0) / 3 * 1e3:.3f} ms per call")
    sm.destroy()
for name, types, nmi in (("sym GS 1/1", (9, 9, 0), (1, 1, 500)), ("RJ3 3/3", (16, 16, 0), (3, 3, 500))):
    mg = api.Multigrid(pc=True)
    mg.set_solver_options(descent=types[0], ascent=types[1], coarse=types[2], n_max_cycles=1,
                          n_max_iter_descent=nmi[0], n_max_iter_ascent=nmi[1],
                          n_max_iter_coarse=nmi[2], precision_mult=(-1., -1., 1.))
    c = api.SlesIt(api.SlesItType.FCG, 200, poly_degree=-1)
    c.transfer_pc(mg)
    t0 = time.perf_counter(); c.setup(m); ts = time.perf_counter() - t0
    r = c.solve(m, s.rhs, 1e-8, r_norm)
    t0 = time.perf_counter()
    r = c.solve(m, s.rhs, 1e-8, r_norm)
    t1 = time.perf_counter() - t0
    print(f"FCG + V-cycle({name}): setup {ts:.3f} s, solve {t1 * 1e3:.1f} ms, {r['n_iter']} iterations, "
          f"cvg {int(r['cvg'])}, residual {r['residual']:.3e}")
    c.destroy()
