#!/usr/bin/env python
"""One short multigrid solve (few cycles) for an ncu capture of k_mg_persistent.
Usage: python scripts/ncu_mg.py [--cube 256] [--precision 1e-2]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", type=int, default=256)
    ap.add_argument("--precision", type=float, default=1e-2)
    ap.add_argument("--solves", type=int, default=2)
    args = ap.parse_args()
    import torch
    from code_saturne_b200 import api
    from code_saturne_b200.mesh_cartesian import cube_poisson
    api.init(0)
    s = cube_poisson(args.cube)
    r_norm = float(np.linalg.norm(s.rhs))
    m = api.Matrix.from_system(s)
    mg = api.Multigrid()
    mg.setup(m)
    rhs_d = torch.from_numpy(s.rhs).cuda()
    vx_d = torch.zeros(s.n_rows, dtype=torch.float64, device="cuda")
    for _ in range(args.solves):
        vx_d.zero_()
        r = mg.solve(m, rhs_d, args.precision, r_norm, vx=vx_d)
    print("cycles", r["n_cycles"], "cvg", r["cvg"], "residual", r["residual"] / r_norm)


if __name__ == "__main__":
    main()
