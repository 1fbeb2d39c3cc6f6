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
