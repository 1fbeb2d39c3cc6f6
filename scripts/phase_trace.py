#!/usr/bin/env python
"""Where the time of a small-level SpMV + dots phase goes (debug build with
-DCSB_PHASE_TRACE, selected with CSB_LIB=...libcsb200_trace.so)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson
    api.init(0)
    L = api.load_library()
    s = cube_poisson(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
    r_norm = float(np.linalg.norm(s.rhs))
    m = api.Matrix.from_system(s)
    mg = api.Multigrid()
    mg.setup(m)
    rhs_d = torch.from_numpy(s.rhs).cuda()
    vx_d = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
    mg.solve(m, rhs_d, 1e-8, r_norm, vx=vx_d)
    out = np.zeros(16)
    L.csb_debug_phase_trace(out.ctypes.data_as(C.c_void_p), 1)
    vx_d.zero_()
    r = mg.solve(m, rhs_d, 1e-8, r_norm, vx=vx_d)
    if L.csb_debug_phase_trace(out.ctypes.data_as(C.c_void_p), 0) != 1:
        raise SystemExit("library was built without -DCSB_PHASE_TRACE")
    n = max(out[15], 1)
    names = ["entry -> first stage full", "-> first gathers back", "-> own slabs consumed",
             "-> CTA joined", "reduction (all_sum)", "whole phase"]
    print(f"cycles {r['n_cycles']}, traced phases {int(out[15])}")
    for k, nm in enumerate(names):
        print(f"  {nm:28s} {out[k]/n/1e3:8.2f} us")


if __name__ == "__main__":
    main()
