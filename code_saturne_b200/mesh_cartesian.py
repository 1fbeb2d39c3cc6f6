"""Synthetic cs_mesh_cartesian-style hex cubes and the pressure-Poisson system on them.

Host-side input generator (numpy only, no device work).  Follows the reference's
cartesian numbering so that matrices match what code_saturne would assemble:

* cell id = i + j*N + k*N^2            (src/mesh/cs_mesh_cartesian.cpp:224-226)
* interior faces ordered x-normal (k, j, i), then y-normal, then z-normal, each
  oriented from the lower to the higher cell id   (cs_mesh_cartesian.cpp:1313-1360)
* MSR structure: off-diagonal columns sorted ascending per row
  (src/alge/cs_matrix.cpp:579-581), diagonal stored apart
* cell -> interior face adjacency aligned with the MSR entries, sign +1 when the
  row is the face's first cell, -1 otherwise
  (src/mesh/cs_mesh_adjacencies.cpp:212-284)
* symmetric diffusion coefficients xa = -S/d = -h, da = sum(-xa) + Dirichlet
  terms + eps   (what _sym_coeffs_msr builds, src/alge/cs_matrix_building.cpp:136-222)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class CubeSystem:
    """MSR matrix + mesh association + right-hand side for an N^3 hex cube."""

    N: int
    n_rows: int
    n_faces: int
    row_index: np.ndarray      # int32 [n_rows+1]
    col_id: np.ndarray         # int32 [nnz]
    d_val: np.ndarray          # float64 [n_rows]
    x_val: np.ndarray          # float64 [nnz]
    face_cells: np.ndarray     # int32 [n_faces, 2]
    c2f: np.ndarray            # int32 [nnz]    interior face id of each entry
    c2f_sgn: np.ndarray        # int16 [nnz]
    cell_cen: np.ndarray       # float64 [n_rows, 3]
    cell_vol: np.ndarray       # float64 [n_rows]
    face_u_normal: np.ndarray  # float64 [n_faces, 3]
    face_surf: np.ndarray      # float64 [n_faces]
    rhs: np.ndarray            # float64 [n_rows]
    meta: dict = field(default_factory=dict)

    @property
    def nnz(self) -> int:
        return int(self.row_index[-1])


def cube_poisson(N: int, bc: str = "neumann", eps_factor: float = 1e-6,
                 with_mesh: bool = True) -> CubeSystem:
    """Pressure-Poisson system on the unit cube meshed with N^3 hexes.

    bc = "neumann":   all walls homogeneous Neumann, da += eps_factor*h (the
                      lid-driven-cavity pressure system, mildly regularised)
    bc = "dirichlet": every boundary face adds 2h to the diagonal (SPD,
                      well conditioned)
    """
    if N < 2:
        raise ValueError("N must be >= 2")
    n = N * N * N
    h = 1.0 / N
    idx = np.arange(n, dtype=np.int64)
    i = idx % N
    j = (idx // N) % N
    k = idx // (N * N)

    masks = [k > 0, j > 0, i > 0, i < N - 1, j < N - 1, k < N - 1]
    offs = [-N * N, -N, -1, 1, N, N * N]

    counts = np.zeros(n, dtype=np.int64)
    for m in masks:
        counts += m
    row_index = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=row_index[1:])
    nnz = int(row_index[-1])

    col_id = np.empty(nnz, dtype=np.int32)
    c2f = np.empty(nnz, dtype=np.int32) if with_mesh else None
    c2f_sgn = np.empty(nnz, dtype=np.int16) if with_mesh else None

    n_fx = (N - 1) * N * N
    n_fy = n_fx
    n_faces = 3 * n_fx

    def xface(ii, jj, kk):   # face between (ii-1,jj,kk) and (ii,jj,kk)
        return (ii - 1) + jj * (N - 1) + kk * (N - 1) * N

    def yface(ii, jj, kk):
        return n_fx + ii + (jj - 1) * N + kk * N * (N - 1)

    def zface(ii, jj, kk):
        return n_fx + n_fy + ii + jj * N + (kk - 1) * N * N

    face_of = [
        lambda s: zface(i[s], j[s], k[s]),
        lambda s: yface(i[s], j[s], k[s]),
        lambda s: xface(i[s], j[s], k[s]),
        lambda s: xface(i[s] + 1, j[s], k[s]),
        lambda s: yface(i[s], j[s] + 1, k[s]),
        lambda s: zface(i[s], j[s], k[s] + 1),
    ]
    sgns = [-1, -1, -1, 1, 1, 1]

    pos = row_index[:-1].copy()
    for d in range(6):
        s = np.nonzero(masks[d])[0]
        p = pos[s]
        col_id[p] = (s + offs[d]).astype(np.int32)
        if with_mesh:
            c2f[p] = face_of[d](s).astype(np.int32)
            c2f_sgn[p] = sgns[d]
        pos[s] += 1

    x_val = np.full(nnz, -h, dtype=np.float64)
    # da = sum over faces of (-xa) accumulated in entry order, like the
    # reference's face loop; all terms equal h so the order is immaterial.
    d_val = counts.astype(np.float64) * h
    n_bfaces = 6 - counts
    if bc == "dirichlet":
        d_val = d_val + n_bfaces.astype(np.float64) * (2.0 * h)
    elif bc == "neumann":
        d_val = d_val + eps_factor * h
    else:
        raise ValueError("bc must be 'neumann' or 'dirichlet'")

    cen = np.empty((n, 3), dtype=np.float64)
    cen[:, 0] = (i + 0.5) * h
    cen[:, 1] = (j + 0.5) * h
    cen[:, 2] = (k + 0.5) * h
    rhs = (np.sin(2 * np.pi * cen[:, 0]) * np.sin(2 * np.pi * cen[:, 1])
           * np.sin(2 * np.pi * cen[:, 2])) * (h ** 3)

    if with_mesh:
        face_cells = np.empty((n_faces, 2), dtype=np.int32)
        face_u_normal = np.zeros((n_faces, 3), dtype=np.float64)
        # x faces
        s = np.nonzero(masks[2])[0]
        f = xface(i[s], j[s], k[s])
        face_cells[f, 0] = s - 1
        face_cells[f, 1] = s
        face_u_normal[f, 0] = 1.0
        s = np.nonzero(masks[1])[0]
        f = yface(i[s], j[s], k[s])
        face_cells[f, 0] = s - N
        face_cells[f, 1] = s
        face_u_normal[f, 1] = 1.0
        s = np.nonzero(masks[0])[0]
        f = zface(i[s], j[s], k[s])
        face_cells[f, 0] = s - N * N
        face_cells[f, 1] = s
        face_u_normal[f, 2] = 1.0
        face_surf = np.full(n_faces, h * h, dtype=np.float64)
        cell_vol = np.full(n, h ** 3, dtype=np.float64)
    else:
        face_cells = np.zeros((0, 2), dtype=np.int32)
        face_u_normal = np.zeros((0, 3), dtype=np.float64)
        face_surf = np.zeros(0, dtype=np.float64)
        cell_vol = np.full(n, h ** 3, dtype=np.float64)
        c2f = np.zeros(0, dtype=np.int32)
        c2f_sgn = np.zeros(0, dtype=np.int16)

    return CubeSystem(
        N=N, n_rows=n, n_faces=n_faces if with_mesh else 0,
        row_index=row_index.astype(np.int32), col_id=col_id,
        d_val=d_val, x_val=x_val, face_cells=face_cells, c2f=c2f,
        c2f_sgn=c2f_sgn, cell_cen=cen, cell_vol=cell_vol,
        face_u_normal=face_u_normal, face_surf=face_surf, rhs=rhs,
        meta={"bc": bc, "eps_factor": eps_factor, "h": h},
    )


def matrix_check_coefficients(sys: CubeSystem, symmetric: bool = True):
    """Deterministic SpMV test coefficients in the style of the reference's
    _matrix_check (src/alge/cs_benchmark_matrix.cpp:868-896): values keyed on
    the global cell number.  Returns (d_val, x_val, x)."""
    n = sys.n_rows
    g = np.arange(1, n + 1, dtype=np.float64)
    d_val = 1.0 + np.cos(g * 0.1 + 0.3)
    ri = sys.row_index.astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ri))
    cols = sys.col_id.astype(np.int64)
    lo = np.minimum(rows, cols) + 1.0
    hi = np.maximum(rows, cols) + 1.0
    x_val = 0.5 * (0.45 + np.cos(lo * 0.13) + np.cos(hi * 0.07))
    if not symmetric:
        x_val = np.where(cols > rows, x_val, -x_val)
    x = np.sin(g * 0.37 + 0.2)
    return d_val, x_val, x
