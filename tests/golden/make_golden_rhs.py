"""Golden fixtures for cs_diffusion_potential without flux reconstruction
(src/alge/cs_convection_diffusion.cpp:8743-8990, eqp->ircflu = 0) on the cube mesh: outputs of
the reference's own compiled CPU code (oracle/_ref/libcsref.so, csref_diffusion_potential):

    python tests/golden/make_golden_rhs.py   -> tests/golden/rhs_N{N}.npz, rhs_rc_N{N}.npz,
                                                rhs_rcb_N{N}.npz, rhs_green_N{N}.npz,
                                                rhs_face_N{N}.npz, rhs_hyd_N{N}.npz,
                                                rhs_iter_N{N}.npz, rhs_clip_N{N}.npz

rhs_rc_*: the same operator WITH flux reconstruction (eqp->ircflu = 1, least-squares gradient,
csref_diffusion_potential_rc) on a non-orthogonal variant of the cube.
rhs_rcb_*: the same with II' of the boundary faces not zero (cs_mesh_quantities_t::diipb): the
boundary values go through cs_gradient_boundary_iprime_lsq_s.
rhs_green_*: eqp->imrgra = 4, the reference's default gradient (CS_GRADIENT_GREEN_LSQ: the
least-squares gradient, then _reconstruct_scalar_gradient), with and without boundary II' and
the warped-face correction matrices.
rhs_face_*: cs_face_diffusion_potential (the per-face fluxes; eqp->ircflu 0 and 1, imrgra 1 and
4, boundary II' given) and cs_divergence of those fluxes.
rhs_hyd_*: the hydrostatic pressure treatment (iphydp = 1 with a seeded exterior force: what
cs_pressure_correction hands over by default), imrgra 1 and 4, with and without boundary II'.
rhs_iter_*: imrgra = 0, the iterative Green-Gauss gradient (default nswrgr 100, epsrgr 1e-4),
with and without boundary II'.
rhs_clip_*: the gradient limiter (eqp->imligr 0 / 1 / 2 with factors climgr 0.3 and 1.5; the
values at I' of the boundary faces are limited too), gradient types 0 / 1 / 4, one case with
the hydrostatic treatment.
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


def rc_input(N, seed=5):
    """the same cube made non-orthogonal: perturbed cell centres, seeded II' / JJ' vectors and
    centre distances per interior face, boundary distances, BC coefficients a / b and a cell
    viscosity -- the reference uses these arrays as they are (no consistency check), so every
    term of the reconstruction is exercised"""
    s, bfc, b_geo, i_face_cog, f = rhs_input(N)
    rng = np.random.default_rng(seed)
    nf, nb, n, h = s.n_faces, len(bfc), s.n_rows, 1.0 / N
    i_geo = np.hstack([0.1 * h * rng.standard_normal((nf, 3)), 0.1 * h * rng.standard_normal((nf, 3)),
                       h * (1 + 0.2 * rng.random((nf, 1)))])
    s.cell_cen = (s.cell_cen.reshape(-1, 3) + 0.15 * h * rng.standard_normal((n, 3))).ravel().copy()
    f["coefa"] = rng.standard_normal(nb)
    f["coefb"] = rng.random(nb)
    f["visel"] = rng.random(n) + 0.5
    b_geo = b_geo.copy()
    b_geo[:, 6] *= (1 + 0.2 * rng.random(nb))
    return s, bfc, b_geo, i_face_cog, i_geo, f


def diipb_input(N, n_b_faces, seed=11):
    """seeded II' of the boundary faces; every 7th face keeps I' = I and every 11th (from 3) is
    below the reference's 1e-3 b_dist threshold, so both of its short cuts are taken"""
    rng = np.random.default_rng(seed)
    diipb = 0.2 / N * rng.standard_normal((n_b_faces, 3))
    diipb[::7] = 0.0
    diipb[3::11] *= 1e-4
    return diipb


def green_input(N, n_faces, n_b_faces, n_rows, seed=29):
    """seeded quantities of the Green-Gauss reconstruction: per interior face the weight and
    the OF vector (g_geo [n_faces][4]), boundary face surfaces, warped-face correction matrices
    (identity + 5 %)"""
    rng = np.random.default_rng(seed)
    h = 1.0 / N
    g_geo = np.hstack([0.3 + 0.4 * rng.random((n_faces, 1)),
                       0.1 * h * rng.standard_normal((n_faces, 3))])
    b_surf = h * h * (1 + 0.1 * rng.random(n_b_faces))
    cgl = (np.eye(3)[None] + 0.05 * rng.standard_normal((n_rows, 3, 3))).reshape(-1)
    return g_geo, b_surf, cgl


HYD_CASES = (("b", 1), ("b", 4), ("p", 4))        # (boundary II' given / plain, imrgra)


def hyd_input(n_rows, seed=43):
    """seeded exterior force per cell"""
    return np.random.default_rng(seed).standard_normal((n_rows, 3))


# (imligr, climgr, imrgra, hydrostatic)
CLIP_CASES = tuple((l, c, g, False) for l in (0, 1, 2) for c in (0.3, 1.5) for g in (0, 1, 4)) \
    + ((1, 1.0, 4, True), (2, 0.3, 1, True))


FACE_CASES = ((0, 1), (1, 1), (1, 4))            # (ircflu, imrgra)
FACE_INIT_INC = ((1, 1), (0, 1), (1, 0))


def face_input(s, n_b_faces, seed=41):
    """seeded previous content of (i_massflux, b_massflux) for the init = 0 calls"""
    rng = np.random.default_rng(seed)
    return rng.standard_normal(s.n_faces), rng.standard_normal(n_b_faces)


def main():
    from oracle.pyoracle import RefLib
    ref = RefLib(1)
    for N in CASES:
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        diipb = diipb_input(N, len(bfc))
        fx = hyd_input(s.n_rows)
        h = ref.system(s)
        out = {}
        for imligr, climgr, imrgra, hyd in CLIP_CASES:
            d, g = ref.diffusion_potential_rc(
                h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                f["cofbf"], f["i_visc"], f["b_visc"], f["visel"], irecmf=0, inc=1, diipb=diipb,
                imrgra=imrgra, g_geo=g_geo, b_face_surf=b_surf, corr_grad_lin=cgl, imligr=imligr,
                climgr=climgr, frcxt=(fx if hyd else None))
            tag = f"l{imligr}_c{climgr}_g{imrgra}_h{int(hyd)}"
            out["grad_" + tag], out["div_" + tag] = g, d
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_clip_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))
    for N in CASES:
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        h = ref.system(s)
        out = {}
        for bt in ("p", "b"):
            diipb = diipb_input(N, len(bfc)) if bt == "b" else None
            for irecmf in (0, 1):
                for inc in (1, 0):
                    d, g = ref.diffusion_potential_rc(
                        h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"], f["coefb"],
                        f["cofaf"], f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                        irecmf=irecmf, inc=inc, diipb=diipb, imrgra=0, g_geo=g_geo,
                        b_face_surf=b_surf, corr_grad_lin=cgl)
                    out[f"{bt}_div_irecmf{irecmf}_inc{inc}"] = d
                    out[f"{bt}_grad_inc{inc}"] = g
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_iter_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))
    for N in CASES:
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        fx = hyd_input(s.n_rows)
        mf0 = face_input(s, len(bfc))
        h = ref.system(s)
        out = {}
        for bt, imrgra in HYD_CASES:
            diipb = diipb_input(N, len(bfc)) if bt == "b" else None
            kw = dict(diipb=diipb, imrgra=imrgra, g_geo=g_geo, b_face_surf=b_surf,
                      corr_grad_lin=cgl, frcxt=fx)
            for irecmf in (0, 1):
                for inc in (1, 0):
                    d, g = ref.diffusion_potential_rc(
                        h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"], f["coefb"],
                        f["cofaf"], f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                        irecmf=irecmf, inc=inc, **kw)
                    out[f"{bt}{imrgra}_div_irecmf{irecmf}_inc{inc}"] = d
                    out[f"{bt}{imrgra}_grad_inc{inc}"] = g
            mf = (mf0[0].copy(), mf0[1].copy())
            ref.diffusion_potential_rc(h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"],
                                       f["coefb"], f["cofaf"], f["cofbf"], f["i_visc"],
                                       f["b_visc"], f["visel"], init=0, ircflu=1, massflux=mf, **kw)
            out[f"{bt}{imrgra}_i_massflux"], out[f"{bt}{imrgra}_b_massflux"] = mf
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_hyd_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))
    for N in CASES:
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        diipb = diipb_input(N, len(bfc))
        mf0 = face_input(s, len(bfc))
        h = ref.system(s)
        out = {}
        for ircflu, imrgra in FACE_CASES:
            for init, inc in FACE_INIT_INC:
                mf = (mf0[0].copy(), mf0[1].copy())
                ref.diffusion_potential_rc(h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"],
                                           f["coefb"], f["cofaf"], f["cofbf"], f["i_visc"],
                                           f["b_visc"], f["visel"], init=init, inc=inc,
                                           diipb=diipb, imrgra=imrgra, g_geo=g_geo,
                                           b_face_surf=b_surf, corr_grad_lin=cgl, ircflu=ircflu,
                                           massflux=mf)
                tag = f"rc{ircflu}_g{imrgra}_init{init}_inc{inc}"
                out["i_" + tag], out["b_" + tag] = mf
                out["div_" + tag] = ref.divergence(h, s, bfc, mf[0], mf[1], init=init,
                                                   diverg=f["div0"])
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_face_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))
    for N in CASES:
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
        h = ref.system(s)
        out = {}
        for tag, diipb, corr in (("plain", None, None), ("full", diipb_input(N, len(bfc)), cgl)):
            for irecmf in (0, 1):
                for inc in (1, 0):
                    d, g = ref.diffusion_potential_rc(
                        h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"], f["coefb"],
                        f["cofaf"], f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                        irecmf=irecmf, inc=inc, diipb=diipb, imrgra=4, g_geo=g_geo,
                        b_face_surf=b_surf, corr_grad_lin=corr)
                    out[f"{tag}_div_irecmf{irecmf}_inc{inc}"] = d
                    out[f"{tag}_grad_inc{inc}"] = g
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_green_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))
    for N in CASES:
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        diipb = diipb_input(N, len(bfc))
        h = ref.system(s)
        out = {}
        for irecmf in (0, 1):
            for inc in (1, 0):
                d, g = ref.diffusion_potential_rc(h, s, bfc, b_geo, ifc, i_geo, f["pvar"],
                                                  f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                                  f["i_visc"], f["b_visc"], f["visel"],
                                                  irecmf=irecmf, inc=inc, diipb=diipb)
                out[f"div_irecmf{irecmf}_inc{inc}"] = d
                out[f"grad_inc{inc}"] = g
        d0, _ = ref.diffusion_potential_rc(h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"],
                                           f["coefb"], f["cofaf"], f["cofbf"], f["i_visc"],
                                           f["b_visc"], f["visel"], irecmf=0)
        assert not np.array_equal(d0, out["div_irecmf0_inc1"])      # II' does change the result
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_rcb_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))
    for N in CASES:
        s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
        h = ref.system(s)
        out = {}
        for irecmf in (0, 1):
            d, g = ref.diffusion_potential_rc(h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"],
                                              f["coefb"], f["cofaf"], f["cofbf"], f["i_visc"],
                                              f["b_visc"], f["visel"], irecmf=irecmf)
            out[f"div_irecmf{irecmf}"] = d
            out["grad"] = g
        d, _ = ref.diffusion_potential_rc(h, s, bfc, b_geo, ifc, i_geo, f["pvar"], f["coefa"],
                                          f["coefb"], f["cofaf"], f["cofbf"], f["i_visc"],
                                          f["b_visc"], f["visel"], irecmf=0, init=0, inc=0,
                                          diverg=f["div0"])
        out["div_init0_inc0"] = d
        ref.destroy(h)
        path = os.path.join(ROOT, "tests", "golden", f"rhs_rc_N{N}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))
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
