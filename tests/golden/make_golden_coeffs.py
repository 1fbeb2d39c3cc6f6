"""Golden fixtures for the matrix coefficient build (cs_matrix_compute_coeffs, scalar cases:
symmetric _sym_coeffs_msr, cs_matrix_building.cpp:136-206, and convection-diffusion
_coeffs_msr, :407-520; relaxation and diagonal shift :2243-2262): MSR coefficients
computed by the reference's own compiled CPU code (oracle/_ref/libcsref.so) from seeded face
viscosities / boundary coefficients on the N^3 cube:

    python tests/golden/make_golden_coeffs.py   -> tests/golden/coeffs_N{N}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_boundary_face_cells, cube_poisson  # noqa: E402
from oracle.pyoracle import RefLib  # noqa: E402

OPTIONS = {"default": dict(), "no_dirichlet": dict(ndircp=0),
           "theta_relax": dict(thetap=0.5, relaxp=0.7), "no_diffusion": dict(idiffp=0)}


CONV_OPTIONS = {"conv": dict(), "conv_theta_relax": dict(ndircp=0, thetap=0.5, relaxp=0.7),
                "conv_only": dict(idiffp=0), "conv_thermal": dict(thermal=True, thetap=0.5)}


V3_OPTIONS = {"v3_sym": dict(iconvp=0), "v3_sym_nofimp": dict(iconvp=0, fimp=False, ndircp=0, relaxp=0.5),
              "v3_conv": dict(iconvp=1), "v3_conv_theta": dict(iconvp=1, ndircp=0, thetap=0.5, relaxp=0.7),
              "v3_conv_only": dict(iconvp=1, idiffp=0, fimp=False)}


def inputs(N, s):
    rng = np.random.default_rng(1234)
    bfc = cube_boundary_face_cells(N)
    out = dict(b_face_cells=bfc, i_visc=0.5 + rng.random(s.n_faces),
               b_visc=0.5 + rng.random(len(bfc)), cofbfp=rng.random(len(bfc)),
               rovsdt=1e-3 * rng.random(s.n_rows))
    out.update(i_massflux=rng.standard_normal(s.n_faces), b_massflux=rng.standard_normal(len(bfc)),
               coefbp=rng.random(len(bfc)), xcpp=1000. + rng.random(s.n_rows))
    return out


def main():
    ref = RefLib(1)
    for N in (5, 7):
        s = cube_poisson(N, "neumann")
        h = ref.system(s)
        out = inputs(N, s)
        for name, kw in OPTIONS.items():
            d, x = ref.matrix_compute_coeffs_sym(h, s, out["b_face_cells"], out["cofbfp"],
                                                 out["rovsdt"], out["i_visc"], out["b_visc"], **kw)
            out[f"{name}_d"] = d
            out[f"{name}_x"] = x
        ref.destroy(h)
        h = ref.system(s, symmetric=False)
        for name, kw in CONV_OPTIONS.items():
            kw = dict(kw)
            xcpp = out["xcpp"] if kw.pop("thermal", False) else None
            d, x = ref.matrix_compute_coeffs(h, s, out["b_face_cells"], out["coefbp"],
                                             out["cofbfp"], out["rovsdt"], out["i_massflux"],
                                             out["b_massflux"], out["i_visc"], out["b_visc"],
                                             xcpp=xcpp, **kw)
            out[f"{name}_d"] = d
            out[f"{name}_x"] = x
        ref.destroy(h)
        # vector variable: 3x3 diagonal blocks (cs_matrix_compute_coeffs<3>)
        rng = np.random.default_rng(4321)
        nb = len(out["b_face_cells"])
        out.update(v3_coefbp=rng.random(nb * 9), v3_cofbfp=rng.random(nb * 9),
                   v3_fimp=rng.random(s.n_rows * 9))
        h = ref.system_block_with_mesh(s)
        for name, kw in V3_OPTIONS.items():
            kw = dict(kw)
            fimp = out["v3_fimp"] if kw.pop("fimp", True) else None
            d, x = ref.matrix_compute_coeffs_v3(h, s, out["b_face_cells"], out["v3_coefbp"],
                                                out["v3_cofbfp"], fimp, out["i_massflux"],
                                                out["b_massflux"], out["i_visc"], out["b_visc"],
                                                **kw)
            out[f"{name}_d"] = d
            out[f"{name}_x"] = x
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"coeffs_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
