"""Golden fixtures for cs_diffusion_potential without flux reconstruction
(src/alge/cs_convection_diffusion.cpp:8743-8990, eqp->ircflu = 0) on the cube mesh: outputs of
the reference's own compiled CPU code (oracle/_ref/libcsref.so, csref_diffusion_potential):

    python tests/golden/make_golden_rhs.py   -> tests/golden/rhs_N{N}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from code_saturne_b200.mesh_cartesian import cube_boundary_face_cells, cube_poisson  # noqa: E402

CASES = (5, 9)


def rhs_input(N, seed=3):
    """system, boundary faces, their geometry for the reference call, seeded fields"""
    s = cube_poisson(N, "dirichlet")
    h = 1.0 / N
    bfc = cube_boundary_face_cells(N)
    nb, per = len(bfc), N * N
    normals = np.zeros((nb, 3))
    for side, (ax, sg) in enumerate([(0, -1), (0, 1), (1, -1), (1, 1), (2, -1), (2, 1)]):
        normals[side * per:(side + 1) * per, ax] = sg
    cen = s.cell_cen.reshape(-1, 3)
    b_geo = np.hstack([normals, cen[bfc] + 0.5 * h * normals, np.full((nb, 1), 0.5 * h)])
    fc = s.face_cells
    i_face_cog = 0.5 * (cen[fc[:, 0]] + cen[fc[:, 1]])
    rng = np.random.default_rng(seed)
    f = dict(pvar=rng.standard_normal(s.n_rows), i_visc=rng.random(s.n_faces) + 0.5,
             b_visc=rng.random(nb) + 0.5, cofaf=rng.standard_normal(nb), cofbf=rng.random(nb),
             div0=rng.standard_normal(s.n_rows))
    return s, bfc, b_geo, i_face_cog, f


def main():
    from oracle.pyoracle import RefLib
    ref = RefLib(1)
    for N in CASES:
        s, bfc, b_geo, ifc, f = rhs_input(N)
        h = ref.system(s)
        out = {}
        for init, inc in ((1, 1), (0, 1), (1, 0)):
            out[f"div_init{init}_inc{inc}"] = ref.diffusion_potential(
                h, s, bfc, b_geo, ifc, f["pvar"], f["cofaf"], f["cofbf"], f["i_visc"], f["b_visc"],
                init=init, inc=inc, diverg=f["div0"])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
