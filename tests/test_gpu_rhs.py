"""cs_diffusion_potential without flux reconstruction (the
explicit right-hand-side operator of the pressure correction): csb_diffusion_potential.
Bar: bit-identical to the reference's one-thread result (same additions in the same order)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
from make_golden_rhs import CASES, rhs_input  # noqa: E402


@pytest.mark.parametrize("N", CASES)
@pytest.mark.parametrize("native", [False, True])
def test_diffusion_potential_matches_reference_golden(csb, N, native):
    g = np.load(os.path.join(GOLD, f"rhs_N{N}.npz"))
    s, bfc, _, _, f = rhs_input(N)
    m = csb.Matrix.from_system_native(s) if native else csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    for init, inc in ((1, 1), (0, 1), (1, 0)):
        d = m.diffusion_potential(f["pvar"], f["i_visc"], f["b_visc"], f["cofaf"], f["cofbf"],
                                  init=init, inc=inc, diverg=f["div0"])
        assert np.array_equal(d, g[f"div_init{init}_inc{inc}"]), (init, inc)
    m.destroy()


def test_diffusion_potential_device_arrays_vs_oracle(csb, oracle):
    """48^3, every array resident in HBM (torch tensors), result written in place"""
    import torch
    s, bfc, _, _, f = rhs_input(48, seed=21)
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    o = oracle.diffusion_potential(s, bfc, f["pvar"], f["cofaf"], f["cofbf"], f["i_visc"],
                                   f["b_visc"], init=0, diverg=f["div0"])
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in f.items()}
    L = csb.load_library()
    r = L.csb_diffusion_potential(m.handle, 0, 1, *[csb._ptr(t[k]) for k in
                                                    ("pvar", "cofaf", "cofbf", "i_visc", "b_visc")],
                                  csb._ptr(t["div0"]))
    assert r == 0
    csb.synchronize()
    assert np.array_equal(t["div0"].cpu().numpy(), o)
    # no boundary faces declared: interior part only
    m2 = csb.Matrix.from_system(s)
    d = m2.diffusion_potential(f["pvar"], f["i_visc"])
    zb = np.zeros(len(bfc))
    assert np.array_equal(d, oracle.diffusion_potential(s, bfc, f["pvar"], zb, zb, f["i_visc"], zb))
    m.destroy(); m2.destroy()


def test_diffusion_potential_errors(csb):
    s, bfc, _, _, f = rhs_input(5)
    m = csb.Matrix(s.n_rows, s.row_index, s.col_id, s.d_val, s.x_val)      # no mesh association
    with pytest.raises(csb.CsbError):
        m.diffusion_potential(f["pvar"], f["i_visc"])
    m.destroy()
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    with pytest.raises(csb.CsbError):
        m.diffusion_potential(f["pvar"], f["i_visc"])                      # boundary arrays missing
    with pytest.raises(csb.CsbError):
        m.diffusion_potential(f["pvar"], f["i_visc"], f["b_visc"], f["cofaf"], f["cofbf"], init=-1)
    m.destroy()


# ------------------- reconstruction on a non-orthogonal mesh: LSQ gradient + reconstructed flux

from make_golden_rhs import rc_input  # noqa: E402


def _rc_matrix(csb, s, bfc, b_geo, i_geo):
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b_geo[:, :3], b_geo[:, 6])
    return m


@pytest.mark.parametrize("N", CASES)
def test_gradient_lsq_and_reconstructed_flux_match_reference_golden(csb, N):
    """cs_gradient_scalar_synced_input (CS_GRADIENT_LSQ; single-precision cocg matrices inverted
    by the float Gauss-Jordan elimination, as in the reference) and cs_diffusion_potential with
    eqp->ircflu = 1, both reconstruction formulas: bit-identical"""
    g = np.load(os.path.join(GOLD, f"rhs_rc_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    assert np.array_equal(m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"]), g["grad"])
    for irecmf in (0, 1):
        d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                     f["i_visc"], f["b_visc"], f["visel"], irecmf=irecmf)
        assert np.array_equal(d, g[f"div_irecmf{irecmf}"]), irecmf
    d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                 f["i_visc"], f["b_visc"], f["visel"], irecmf=0, init=0, inc=0,
                                 diverg=f["div0"])
    assert np.array_equal(d, g["div_init0_inc0"])
    m.destroy()


def test_gradient_lsq_vs_oracle_larger_and_linear_field(csb, oracle):
    """40^3 non-orthogonal cube against the oracle (bit-exact), cocg cache reused across calls
    and rebuilt after a geometry change; a linear field's gradient is reproduced"""
    s, bfc, b_geo, _, i_geo, f = rc_input(40, seed=17)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    for k in range(2):
        p = f["pvar"] * (1 + k)
        go = oracle.gradient_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], p, f["coefa"], f["coefb"])
        assert np.array_equal(m.gradient_lsq(p, f["coefa"], f["coefb"]), go)
    do = oracle.diffusion_potential_rc(s, bfc, i_geo, go, p, f["cofaf"], f["cofbf"], f["i_visc"],
                                       f["b_visc"], f["visel"], irecmf=1)
    d = m.diffusion_potential_rc(p, f["coefa"], f["coefb"], f["cofaf"], f["cofbf"], f["i_visc"],
                                 f["b_visc"], f["visel"], irecmf=1)
    assert np.array_equal(d, do)
    cen = s.cell_cen.reshape(-1, 3)
    a = np.array([0.7, -1.3, 2.1])
    lin = cen @ a + 0.4
    vf = lin[bfc] + b_geo[:, 6] * (b_geo[:, :3] @ a)
    grad = m.gradient_lsq(lin, vf, np.zeros(len(bfc)))
    assert np.abs(grad - a).max() < 5e-6 * np.abs(a).max()
    # new boundary distances: the cached cocg matrices do not depend on them, the gradient does
    b2 = b_geo.copy(); b2[:, 6] *= 1.5
    m.set_face_geometry(i_geo[:, 6], i_geo[:, :3], i_geo[:, 3:6], b2[:, :3], b2[:, 6])
    go = oracle.gradient_lsq(s, bfc, b2[:, :3], b2[:, 6], f["pvar"], f["coefa"], f["coefb"])
    assert np.array_equal(m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"]), go)
    m.destroy()


@pytest.mark.parametrize("N", CASES)
def test_boundary_iprime_reconstruction_matches_reference_golden(csb, N):
    """II' of the boundary faces not zero (csb_matrix_set_boundary_iprime): the values at I' of
    cs_gradient_boundary_iprime_lsq_s feed the face values of the gradient and the boundary
    fluxes; gradient and divergence bit-identical to the reference (both formulas, inc 0 / 1);
    resetting II' gives the plain result back (cocg cache invalidated both ways)"""
    from make_golden_rhs import diipb_input
    g = np.load(os.path.join(GOLD, f"rhs_rcb_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    g0 = m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"])
    m.set_boundary_iprime(diipb_input(N, len(bfc)))
    for inc in (1, 0):
        gr = m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"], inc=inc)
        assert np.array_equal(gr, g[f"grad_inc{inc}"].reshape(-1, 3)), inc
        for irecmf in (0, 1):
            d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                         f["i_visc"], f["b_visc"], f["visel"], irecmf=irecmf,
                                         inc=inc)
            assert np.array_equal(d, g[f"div_irecmf{irecmf}_inc{inc}"]), (irecmf, inc)
    m.set_boundary_iprime(None)
    assert np.array_equal(m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"]), g0)
    m.destroy()


def test_boundary_iprime_vs_oracle_larger(csb, oracle):
    """40^3 non-orthogonal cube with boundary II' against the oracle, bit-exact"""
    from make_golden_rhs import diipb_input
    s, bfc, b_geo, _, i_geo, f = rc_input(40, seed=17)
    diipb = diipb_input(40, len(bfc), seed=23)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_boundary_iprime(diipb)
    vip = oracle.b_iprime_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], diipb, f["pvar"], f["coefa"],
                              f["coefb"])
    go = oracle.gradient_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], f["pvar"], f["coefa"],
                             f["coefb"], diipb=diipb)
    assert np.array_equal(m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"]), go)
    for irecmf in (0, 1):
        do = oracle.diffusion_potential_rc(s, bfc, i_geo, go, f["pvar"], f["cofaf"], f["cofbf"],
                                           f["i_visc"], f["b_visc"], f["visel"], irecmf=irecmf,
                                           val_ip=vip)
        d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                     f["i_visc"], f["b_visc"], f["visel"], irecmf=irecmf)
        assert np.array_equal(d, do), irecmf
    m.destroy()


@pytest.mark.parametrize("N", CASES)
def test_green_lsq_gradient_matches_reference_golden(csb, N):
    """imrgra = 4, the reference's default gradient (CS_GRADIENT_GREEN_LSQ: least-squares
    gradient, then _reconstruct_scalar_gradient) and cs_diffusion_potential reconstructed with
    it; plain, and with boundary II' + warped-face correction matrices: bit-identical"""
    from make_golden_rhs import diipb_input, green_input
    g = np.load(os.path.join(GOLD, f"rhs_green_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    with pytest.raises(csb.CsbError):
        m.set_gradient_type(4)                      # Green-Gauss quantities not given yet
    with pytest.raises(csb.CsbError):
        m.gradient(0, f["pvar"], f["coefa"], f["coefb"])      # iterative gradient: not built
    for tag, diipb, corr in (("plain", None, None), ("full", diipb_input(N, len(bfc)), cgl)):
        m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, corr)
        m.set_boundary_iprime(diipb)
        m.set_gradient_type(4)
        for inc in (1, 0):
            gr = m.gradient(4, f["pvar"], f["coefa"], f["coefb"], inc=inc)
            assert np.array_equal(gr, g[f"{tag}_grad_inc{inc}"].reshape(-1, 3)), (tag, inc)
            for irecmf in (0, 1):
                d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                             f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                             irecmf=irecmf, inc=inc)
                assert np.array_equal(d, g[f"{tag}_div_irecmf{irecmf}_inc{inc}"]), (tag, irecmf, inc)
    m.set_gradient_type(1)                          # back to least squares
    m.set_boundary_iprime(None)
    g1 = np.load(os.path.join(GOLD, f"rhs_rc_N{N}.npz"))
    d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                 f["i_visc"], f["b_visc"], f["visel"], irecmf=0)
    assert np.array_equal(d, g1["div_irecmf0"])
    m.destroy()


def test_green_lsq_gradient_vs_oracle_larger(csb, oracle):
    """40^3 non-orthogonal cube, imrgra = 4 with boundary II' and the warped-face correction,
    against the oracle: bit-exact"""
    from make_golden_rhs import diipb_input, green_input
    s, bfc, b_geo, _, i_geo, f = rc_input(40, seed=17)
    g_geo, b_surf, cgl = green_input(40, s.n_faces, len(bfc), s.n_rows, seed=31)
    diipb = diipb_input(40, len(bfc), seed=23)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    m.set_boundary_iprime(diipb)
    m.set_gradient_type(4)
    go = oracle.gradient_green_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], g_geo, b_surf, f["pvar"],
                                   f["coefa"], f["coefb"], diipb=diipb, corr_grad_lin=cgl)
    assert np.array_equal(m.gradient(4, f["pvar"], f["coefa"], f["coefb"]), go)
    vip = oracle.b_iprime_lsq(s, bfc, b_geo[:, :3], b_geo[:, 6], diipb, f["pvar"], f["coefa"],
                              f["coefb"])
    do = oracle.diffusion_potential_rc(s, bfc, i_geo, go, f["pvar"], f["cofaf"], f["cofbf"],
                                       f["i_visc"], f["b_visc"], f["visel"], irecmf=0, val_ip=vip)
    d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                 f["i_visc"], f["b_visc"], f["visel"], irecmf=0)
    assert np.array_equal(d, do)
    m.destroy()


@pytest.mark.parametrize("N", CASES)
def test_face_diffusion_potential_and_divergence_match_reference_golden(csb, N):
    """cs_face_diffusion_potential (per-face fluxes: ircflu 0 / 1, imrgra 1 / 4, boundary II',
    init 0 / 1, inc 0 / 1) and cs_divergence of them: bit-identical to the reference; the
    divergence of the fluxes equals csb_diffusion_potential_rc (irecmf 0)"""
    from make_golden_rhs import (FACE_CASES, FACE_INIT_INC, diipb_input, face_input, green_input)
    g = np.load(os.path.join(GOLD, f"rhs_face_N{N}.npz"))
    s, bfc, b_geo, _, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    mf0 = face_input(s, len(bfc))
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    m.set_boundary_iprime(diipb_input(N, len(bfc)))
    for ircflu, imrgra in FACE_CASES:
        m.set_gradient_type(imrgra)
        for init, inc in FACE_INIT_INC:
            im, bm = m.face_diffusion_potential(
                s.n_faces, len(bfc), f["pvar"], f["cofaf"], f["cofbf"], f["i_visc"], f["b_visc"],
                coefa=f["coefa"], coefb=f["coefb"], visel=f["visel"], ircflu=ircflu, init=init,
                inc=inc, massflux=mf0)
            tag = f"rc{ircflu}_g{imrgra}_init{init}_inc{inc}"
            assert np.array_equal(im, g["i_" + tag]) and np.array_equal(bm, g["b_" + tag]), tag
            d = m.divergence(im, bm, init=init, diverg=f["div0"])
            assert np.array_equal(d, g["div_" + tag]), tag
            if init == 1 and ircflu:
                d2 = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                              f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                              irecmf=0, inc=inc)
                assert np.array_equal(d, d2), tag
    with pytest.raises(csb.CsbError):
        m.divergence(g["i_rc0_g1_init1_inc1"], None)                 # boundary fluxes missing
    with pytest.raises(csb.CsbError):
        m.face_diffusion_potential(s.n_faces, len(bfc), f["pvar"], f["cofaf"], f["cofbf"],
                                   f["i_visc"], f["b_visc"], ircflu=1)   # coefa / coefb / visel
    m.destroy()


def test_face_fluxes_without_reconstruction_need_no_geometry(csb, oracle):
    """ircflu = 0 on the plain cube (mesh association and boundary faces only), device-resident
    arrays; 40^3 against the oracle, and divergence(fluxes) == csb_diffusion_potential"""
    import torch
    N = 40
    s, bfc, _, _, f = rhs_input(N)
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in f.items()}
    im, bm = m.face_diffusion_potential(s.n_faces, len(bfc), t["pvar"], t["cofaf"], t["cofbf"],
                                        t["i_visc"], t["b_visc"])
    imo, bmo = oracle.face_diffusion_potential(s, bfc, f["pvar"], f["cofaf"], f["cofbf"],
                                               f["i_visc"], f["b_visc"])
    assert np.array_equal(im, imo) and np.array_equal(bm, bmo)
    d = m.divergence(im, bm)
    assert np.array_equal(d, oracle.divergence(s, bfc, imo, bmo))
    assert np.array_equal(d, m.diffusion_potential(f["pvar"], f["i_visc"], f["b_visc"], f["cofaf"],
                                                   f["cofbf"]))
    m.destroy()


@pytest.mark.parametrize("N", CASES)
def test_hydrostatic_treatment_matches_reference_golden(csb, N):
    """iphydp = 1 with an exterior force (what cs_pressure_correction hands over by default):
    gradient (imrgra 1 / 4), reconstructed divergence (both formulas, inc 0 / 1) and face
    fluxes, with and without boundary II': bit-identical to the reference; flag 0 / 2 = off"""
    from make_golden_rhs import (HYD_CASES, diipb_input, face_input, green_input, hyd_input)
    g = np.load(os.path.join(GOLD, f"rhs_hyd_N{N}.npz"))
    s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    fx = hyd_input(s.n_rows)
    mf0 = face_input(s, len(bfc))
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    with pytest.raises(csb.CsbError):
        m.gradient_hyd(1, fx, f["pvar"], f["coefa"], f["coefb"])         # face centres missing
    m.set_face_cog(ifc, b_geo[:, 3:6])
    for bt, imrgra in HYD_CASES:
        m.set_boundary_iprime(diipb_input(N, len(bfc)) if bt == "b" else None)
        m.set_gradient_type(imrgra)
        for inc in (1, 0):
            gr = m.gradient_hyd(imrgra, fx, f["pvar"], f["coefa"], f["coefb"], inc=inc)
            assert np.array_equal(gr, g[f"{bt}{imrgra}_grad_inc{inc}"]), (bt, imrgra, inc)
            for irecmf in (0, 1):
                d = m.diffusion_potential_hyd(fx, f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                              f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                              irecmf=irecmf, inc=inc)
                assert np.array_equal(d, g[f"{bt}{imrgra}_div_irecmf{irecmf}_inc{inc}"]), \
                    (bt, imrgra, irecmf, inc)
        im, bm = m.face_diffusion_potential_hyd(
            s.n_faces, len(bfc), fx, f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
            f["i_visc"], f["b_visc"], f["visel"], init=0, massflux=mf0)
        assert np.array_equal(im, g[f"{bt}{imrgra}_i_massflux"]), (bt, imrgra)
        assert np.array_equal(bm, g[f"{bt}{imrgra}_b_massflux"]), (bt, imrgra)
    for flag in (0, 2):                              # off, like the reference reads 0 and 2
        assert np.array_equal(m.gradient_hyd(4, fx, f["pvar"], f["coefa"], f["coefb"],
                                             hyd_p_flag=flag),
                              m.gradient(4, f["pvar"], f["coefa"], f["coefb"]))
    m.destroy()


def test_hydrostatic_treatment_vs_oracle_larger(csb, oracle):
    """40^3, iphydp = 1, imrgra 4, boundary II', device-resident force: bit-exact vs the oracle"""
    import torch
    from make_golden_rhs import diipb_input, green_input, hyd_input
    s, bfc, b_geo, ifc, i_geo, f = rc_input(40, seed=17)
    g_geo, b_surf, cgl = green_input(40, s.n_faces, len(bfc), s.n_rows, seed=31)
    diipb = diipb_input(40, len(bfc), seed=23)
    fx = hyd_input(s.n_rows, seed=47)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    m.set_face_cog(ifc, b_geo[:, 3:6])
    m.set_boundary_iprime(diipb)
    m.set_gradient_type(4)
    go, vip = oracle.gradient_hyd(4, s, bfc, b_geo[:, :3], b_geo[:, 6], ifc, b_geo[:, 3:6], fx,
                                  f["pvar"], f["coefa"], f["coefb"], diipb=diipb, g_geo=g_geo,
                                  b_face_surf=b_surf, corr_grad_lin=cgl)
    fx_d = torch.from_numpy(np.ascontiguousarray(fx)).cuda()
    assert np.array_equal(m.gradient_hyd(4, fx_d, f["pvar"], f["coefa"], f["coefb"]), go)
    do = oracle.diffusion_potential_rc(s, bfc, i_geo, go, f["pvar"], f["cofaf"], f["cofbf"],
                                       f["i_visc"], f["b_visc"], f["visel"], irecmf=1, val_ip=vip)
    d = m.diffusion_potential_hyd(fx_d, f["pvar"], f["coefa"], f["coefb"], f["cofaf"], f["cofbf"],
                                  f["i_visc"], f["b_visc"], f["visel"], irecmf=1)
    assert np.array_equal(d, do)
    m.destroy()


@pytest.mark.parametrize("N", CASES)
def test_iterative_green_gauss_matches_reference_golden(csb, oracle, N):
    """imrgra = 0 (CS_GRADIENT_GREEN_ITER): gradient and reconstructed divergence bit-identical
    to the reference, same number of sweeps as the pinned oracle; sweep limits honoured; the
    hydrostatic form is refused"""
    from make_golden_rhs import diipb_input, green_input, hyd_input
    g = np.load(os.path.join(GOLD, f"rhs_iter_N{N}.npz"))
    s, bfc, b_geo, ifc, i_geo, f = rc_input(N)
    g_geo, b_surf, cgl = green_input(N, s.n_faces, len(bfc), s.n_rows)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    m.set_gradient_type(0)
    for bt in ("p", "b"):
        diipb = diipb_input(N, len(bfc)) if bt == "b" else None
        m.set_boundary_iprime(diipb)
        for inc in (1, 0):
            gr = m.gradient(0, f["pvar"], f["coefa"], f["coefb"], inc=inc)
            assert np.array_equal(gr, g[f"{bt}_grad_inc{inc}"]), (bt, inc)
            _, ns = oracle.gradient_green_iter(s, bfc, b_geo[:, :3], g_geo, b_surf, f["pvar"],
                                               f["coefa"], f["coefb"], inc=inc, diipb=diipb)
            assert m.gradient_sweeps() == ns, (bt, inc)
            for irecmf in (0, 1):
                d = m.diffusion_potential_rc(f["pvar"], f["coefa"], f["coefb"], f["cofaf"],
                                             f["cofbf"], f["i_visc"], f["b_visc"], f["visel"],
                                             irecmf=irecmf, inc=inc)
                assert np.array_equal(d, g[f"{bt}_div_irecmf{irecmf}_inc{inc}"]), (bt, irecmf, inc)
    for nsw in (1, 3):
        m.set_gradient_sweeps(nsw, 1e-4)
        go, ns = oracle.gradient_green_iter(s, bfc, b_geo[:, :3], g_geo, b_surf, f["pvar"],
                                            f["coefa"], f["coefb"], diipb=diipb, nswrgr=nsw)
        assert np.array_equal(m.gradient(0, f["pvar"], f["coefa"], f["coefb"]), go), nsw
        assert m.gradient_sweeps() == ns == nsw
    m.set_face_cog(ifc, b_geo[:, 3:6])
    with pytest.raises(csb.CsbError):
        m.gradient_hyd(0, hyd_input(s.n_rows), f["pvar"], f["coefa"], f["coefb"])
    m.destroy()


def test_iterative_green_gauss_vs_oracle_larger(csb, oracle):
    """40^3, imrgra = 0 with boundary II': bit-exact against the oracle, same sweep count"""
    from make_golden_rhs import diipb_input, green_input
    s, bfc, b_geo, _, i_geo, f = rc_input(40, seed=17)
    g_geo, b_surf, cgl = green_input(40, s.n_faces, len(bfc), s.n_rows, seed=31)
    diipb = diipb_input(40, len(bfc), seed=23)
    m = _rc_matrix(csb, s, bfc, b_geo, i_geo)
    m.set_green_geometry(g_geo[:, 0], g_geo[:, 1:4], b_surf, cgl)
    m.set_boundary_iprime(diipb)
    m.set_gradient_type(0)
    go, ns = oracle.gradient_green_iter(s, bfc, b_geo[:, :3], g_geo, b_surf, f["pvar"],
                                        f["coefa"], f["coefb"], diipb=diipb)
    assert np.array_equal(m.gradient(0, f["pvar"], f["coefa"], f["coefb"]), go)
    assert m.gradient_sweeps() == ns and 2 <= ns < 99
    m.destroy()


def test_reconstruction_needs_geometry(csb):
    s, bfc, _, _, _, f = rc_input(5)
    m = csb.Matrix.from_system(s)
    m.set_boundary_faces(bfc)
    with pytest.raises(csb.CsbError):
        m.gradient_lsq(f["pvar"], f["coefa"], f["coefb"])
    m.destroy()
