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


def _time(fn, reps=10):
    for _ in range(2):
        fn()
    api.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    api.synchronize()
    return (time.perf_counter() - t0) / reps


# Reconstruction operators on a non-orthogonal variant of the cube, arrays resident in HBM:
# least-squares gradient (imrgra 1), Green-Gauss with least-squares face values (imrgra 4, the
# reference's default) and cs_diffusion_potential reconstructed with it, boundary II' given.
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "tests", "golden"))
from make_golden_rhs import diipb_input, green_input, rc_input  # noqa: E402

for N in [int(a) for a in sys.argv[1:]] or [128, 256]:
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    m = api.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    m.set_boundary_iprime(diipb_input(N, len(bfc)))
    m.set_gradient_type(4)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in f.items()}
    t["grad"] = torch.zeros(3 * s.n_rows, dtype=torch.float64, device="cuda")
    t["div"] = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
    L = api.load_library()
    L.csb_gradient_scalar.argtypes = [api.C.c_void_p, api.C.c_int, api.C.c_int] + [api.C.c_void_p] * 4
    L.csb_diffusion_potential_rc.argtypes = [api.C.c_void_p, api.C.c_int, api.C.c_int, api.C.c_int] + [api.C.c_void_p] * 9
    ga = [api._ptr(t[k]) for k in ("pvar", "coefa", "coefb", "grad")]
    da = [api._ptr(t[k]) for k in ("pvar", "coefa", "coefb", "cofaf", "cofbf", "i_visc", "b_visc",
                                   "visel", "div")]
    nnz, n, nf = s.nnz, s.n_rows, s.n_faces
    adj = 4. * (n + 1) + nnz * (4 + 4 + 2)              # row_index, col_id, c2f, c2f_sgn
    b_lsq = adj + 8. * n * (3 + 1 + 3) + 4. * 6 * n      # + cell_cen, pvar, grad, cocg (float)
    b_green = b_lsq + adj + 8. * nf * (1 + 3 + 1 + 3) + 8. * n * (1 + 3 + 1 + 3 + 9 + 3)
    b_rc = b_green + adj + 8. * nf * (3 + 3 + 1 + 1 + 1) + 8. * n * (1 + 3 + 1 + 1)
    for name, fn, byts in (
            ("csb_gradient_scalar imrgra 1 (LSQ)", lambda: L.csb_gradient_scalar(m.handle, 1, 1, *ga), b_lsq),
            ("csb_gradient_scalar imrgra 4 (Green-Gauss, LSQ face values)",
             lambda: L.csb_gradient_scalar(m.handle, 4, 1, *ga), b_green),
            ("csb_diffusion_potential_rc imrgra 4", lambda: L.csb_diffusion_potential_rc(m.handle, 0, 1, 1, *da), b_rc)):
        dt = _time(fn)
        print(f"{name} {N}^3: {dt * 1e3:.3f} ms per call (one stream synchronize each); "
              f"{byts / dt / 1e9:.0f} GB/s of the arrays its kernels must read or write once "
              f"({byts / n:.0f} B per cell)", flush=True)
    m.destroy()
