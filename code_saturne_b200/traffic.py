"""Algorithmic HBM bytes of the hot path.

"Algorithmic" = every array an operation has to read or write once, nothing more:
MSR SpMV  12*nnz (col_id + x_val) + 4*(n+1) (row_index) + 8n (d_val) + 8n (x) + 8n (y).
The figures are per level of the hierarchy and per phase of one V-cycle with the
solver-mode defaults (PCG+Jacobi smoothers, 2 descent / 10 ascent sweeps); bench.py and
scripts/mg_phases.py divide them by device-side times.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

PHASES = ("descent", "residual", "restrict", "coarsest", "prolong", "ascent")


def spmv_bytes(n, nnz):
    return 12.0 * nnz + 4.0 * (n + 1) + 24.0 * n


def phase_bytes(n, nnz, nc, n_desc, n_asc, level, coarsest):
    """Bytes of each phase (PHASES order) of one cycle at one level."""
    M = 12.0 * nnz + 12.0 * n            # col_id + x_val + row_index + d_val

    def pcg(n_it, ini):
        b = (M + 32 * n) if ini else 40 * n          # residual / zero start
        b += M + 24 * n                               # zk = A.dk + 3 dots
        if n_it >= 1:
            b += 64 * n                               # first update
        b += max(n_it - 1, 0) * (M + 24 * n)          # sk = A.gk + 3 dots
        b += max(n_it - 2, 0) * 96 * n                # fused 5-vector update
        if n_it >= 2:
            b += 32 * n                               # last update of a smoother: vx only
        return b

    out = np.zeros(6)
    if level < coarsest:
        out[0] = pcg(n_desc, level == 0)
        out[1] = M + 24 * n
        out[2] = 12 * n + 12 * nc
        out[4] = 20 * n + 8 * nc
        if level > 0:
            out[5] = pcg(n_asc, True)
    return out


def level_sizes(mg):
    """[(n_rows, nnz)] of every level of an api.Multigrid after setup."""
    from . import api
    L = api.load_library()
    info = []
    for l in range(mg.n_levels()):
        a = np.zeros(3, dtype=np.int32)
        L.csb_multigrid_get_level_info(C.c_void_p(mg._h), l, a.ctypes.data_as(C.c_void_p))
        info.append((int(a[0]), int(a[2])))
    return info


def cycle_bytes(info, n_desc=2, n_asc=10):
    """Algorithmic bytes of one V-cycle over the hierarchy `info` (levels above the coarsest;
    the coarsest solve is latency-, not bandwidth-bound)."""
    coarsest = len(info) - 1
    tot = 0.0
    for l, (n, nnz) in enumerate(info):
        nc = info[l + 1][0] if l < coarsest else 0
        tot += phase_bytes(n, nnz, nc, n_desc, n_asc, l, coarsest).sum()
    return tot


# ---------------------------------------------------------------------------------------
# FCG preconditioned by one V-cycle with Jacobi-family smoothers (k_fcg_mg_persistent)
# ---------------------------------------------------------------------------------------

def sweep_bytes(n, nnz, mode="relaxed"):
    """One fused smoother sweep: E (col_id, x_val, row_index), x (read once: gather and
    staged copy are the same vector), rhs, 1/d, write; L1 adds d_val."""
    b = 12.0 * nnz + 4.0 * (n + 1) + 32.0 * n
    if mode == "jacobi":
        pass                      # same arrays, x only gathered
    elif mode == "l1":
        b += 8.0 * n
    return b


def fcg_mg_phase_bytes(n, nnz, nc, n_desc, n_asc, level, coarsest, mode="relaxed"):
    """Bytes of each phase (PHASES order) of one preconditioner V-cycle at one level.
    The first descent sweep (from zero) is fused into the phase that produces its right-hand
    side: 16 bytes per row (read 1/d, write) are counted with the restriction above (level
    > 0) or with the FCG update (level 0)."""
    out = np.zeros(6)
    if level < coarsest:
        out[0] = max(n_desc - 1, 0) * sweep_bytes(n, nnz, mode)
        out[1] = 12.0 * nnz + 4.0 * (n + 1) + 32.0 * n        # d, x, rhs, wr
        fused_first = 16.0 * nc if level + 1 < coarsest else 0.0
        out[2] = 12.0 * n + 12.0 * nc + fused_first
        out[4] = 20.0 * n + 8.0 * nc
        out[5] = n_asc * sweep_bytes(n, nnz, mode)
    return out


def fcg_mg_iter_bytes(info, n_desc=3, n_asc=3, mode="relaxed"):
    """Algorithmic bytes of ONE FCG iteration: the V-cycle over `info`, wk = A.vk with the 4
    dot products (vk, rk, qk read; wk written) and the fused 6-vector update + first sweep."""
    coarsest = len(info) - 1
    tot = 0.0
    for l, (n, nnz) in enumerate(info):
        nc = info[l + 1][0] if l < coarsest else 0
        tot += fcg_mg_phase_bytes(n, nnz, nc, n_desc, n_asc, l, coarsest, mode).sum()
    n0, nnz0 = info[0]
    tot += 12.0 * nnz0 + 4.0 * (n0 + 1) + 40.0 * n0           # SpMV + 4 dots
    tot += 96.0 * n0                                          # update (7 reads, 5 writes)
    return tot
