"""Device time of csb_diffusion_potential (no reconstruction) with every array resident in HBM."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from code_saturne_b200 import api
from code_saturne_b200.mesh_cartesian import cube_boundary_face_cells, cube_poisson

api.init(0)
for N in [int(a) for a in sys.argv[1:]] or [128, 256]:
    s = cube_poisson(N, "neumann")
    bfc = cube_boundary_face_cells(N)
    m = api.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    rng = np.random.default_rng(1)
    t = {k: torch.from_numpy(v).cuda() for k, v in dict(
        pvar=rng.standard_normal(s.n_rows), cofaf=rng.standard_normal(len(bfc)),
        cofbf=rng.random(len(bfc)), i_visc=rng.random(s.n_faces) + .5,
        b_visc=rng.random(len(bfc)) + .5, div=np.zeros(s.n_rows)).items()}
    L = api.load_library()
    L.csb_diffusion_potential.argtypes = [api.C.c_void_p, api.C.c_int, api.C.c_int] + [api.C.c_void_p] * 6
    args = [api._ptr(t[k]) for k in ("pvar", "cofaf", "cofbf", "i_visc", "b_visc", "div")]
    for _ in range(3):
        L.csb_diffusion_potential(m.handle, 1, 1, *args)
    api.synchronize()
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        L.csb_diffusion_potential(m.handle, 1, 1, *args)
    api.synchronize()
    dt = (time.perf_counter() - t0) / reps
    nnz = s.nnz
    byts = 4. * (s.n_rows + 1) + nnz * (4 + 4 + 2) + 8. * s.n_faces + 8. * 2 * s.n_rows
    print(f"csb_diffusion_potential {N}^3 ({s.n_rows} cells, {s.n_faces} interior faces): {dt * 1e3:.3f} ms per call "
          f"(includes one stream synchronize); {byts / dt / 1e9:.0f} GB/s of the arrays it must read once "
          f"({byts / s.n_rows:.0f} B per cell)")
    m.destroy()
