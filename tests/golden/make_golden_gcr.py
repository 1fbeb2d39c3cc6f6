"""Golden fixtures for GCR (_gcr, cs_sles_it.cpp:3272-3480) and for the fallback sequence of
cs_sles_it_solve (_fallback, :4464-4528; call site :5383-5418) in its device form -- GCR takes
over from the first solver's vx, or from zero when that solver diverged.  Outputs of the
reference's own compiled CPU code (oracle/_ref/libcsref.so, OMP_NUM_THREADS=1).  The CPU build
of the reference falls back to GMRES (the choice is `on_device ? GCR : GMRES`), so the device
sequence is composed here from the reference's own pieces: its BiCGStab / block Jacobi run,
then its GCR run started as _fallback starts it.

    python tests/golden/make_golden_gcr.py   -> tests/golden/gcr.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import (cube_block3_convection_diffusion,  # noqa: E402
                                              cube_block_convection_diffusion, cube_poisson)
from oracle.pyoracle import RefLib  # noqa: E402

GCR = 6


def main():
    ref = RefLib(1)
    out = {}
    s = cube_poisson(12, "dirichlet")
    h = ref.system(s)
    r_norm = float(np.linalg.norm(s.rhs))
    vx0 = np.cos(0.1 * np.arange(s.n_rows))
    out["s_r_norm"] = r_norm
    out["s_vx0"] = vx0
    for tag, pd, v0, nmi, prec in (("jac_zero", 0, None, 200, 1e-10), ("jac_warm", 0, vx0, 200, 1e-10),
                                   ("npc_warm", -1, vx0, 200, 1e-10), ("jac_27", 0, vx0, 27, 1e-14)):
        r = ref.sles_it(h, GCR, s.rhs, nmi, prec, r_norm, v0, poly_degree=pd)
        for k in ("cvg", "n_iter", "residual", "initial_residual", "vx"):
            out[f"s_{tag}_{k}"] = r[k]
        print("scalar", tag, r["cvg"], r["n_iter"])
    ref.destroy(h)

    # block 3: BiCGStab stopped at 3 iterations (MAX_ITERATION), then GCR from its vx
    ri, ci, d, xv, rhs = cube_block3_convection_diffusion(8)
    n = len(ri) - 1
    h = ref.system_block(n, 3, ri, ci, d.ravel(), xv, False)
    r_norm = float(np.linalg.norm(rhs))
    out["b3_r_norm"] = r_norm
    r = ref.sles_it(h, GCR, rhs, 200, 1e-10, r_norm)
    for k in ("cvg", "n_iter", "residual", "vx"):
        out[f"b3_gcr_{k}"] = r[k]
    b = ref.sles_it(h, ref.BICGSTAB, rhs, 3, 1e-10, r_norm)
    f = ref.sles_it(h, GCR, rhs, 3, 1e-10, r_norm, vx0=b["vx"])       # fallback_n_max_iter = 3
    out["b3_fb_first_cvg"] = b["cvg"]
    out["b3_fb_first_n_iter"] = b["n_iter"]
    for k in ("cvg", "n_iter", "residual", "vx"):
        out[f"b3_fb_{k}"] = f[k]
    f = ref.sles_it(h, GCR, rhs, 200, 1e-10, r_norm, vx0=b["vx"])
    for k in ("cvg", "n_iter", "residual", "vx"):
        out[f"b3_fb200_{k}"] = f[k]
    print("block3: bicgstab", b["cvg"], b["n_iter"], "-> gcr", f["cvg"], f["n_iter"])
    ref.destroy(h)

    # block 4: block Jacobi diverges (LU factors applied as an inverse), GCR from zero
    ri, ci, d, xv, rhs = cube_block_convection_diffusion(5, 4)
    n = len(ri) - 1
    h = ref.system_block(n, 4, ri, ci, d.ravel(), xv, False)
    r_norm = float(np.linalg.norm(rhs))
    out["b4_r_norm"] = r_norm
    j = ref.sles_it(h, ref.JACOBI, rhs, 500, 1e-8, r_norm)
    f = ref.sles_it(h, GCR, rhs, 500, 1e-8, r_norm)
    out["b4_fb_first_cvg"] = j["cvg"]
    out["b4_fb_first_n_iter"] = j["n_iter"]
    for k in ("cvg", "n_iter", "residual", "vx"):
        out[f"b4_fb_{k}"] = f[k]
    print("block4: jacobi", j["cvg"], j["n_iter"], "-> gcr", f["cvg"], f["n_iter"])
    ref.destroy(h)
    path = os.path.join(ROOT, "tests", "golden", "gcr.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
