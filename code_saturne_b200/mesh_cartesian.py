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


# --------------------------------------------------------------------------- partitioning

@dataclass
class HaloDesc:
    """cs_halo_t content for one rank (src/base/cs_halo.h:74-171, CS_HALO_STANDARD only,
    no periodicity): neighbour ranks ascending, one ghost section and one send section
    per neighbour.  Ghost g of section i is column n_local_elts + index[i] + g."""

    n_local_elts: int
    c_domain_rank: np.ndarray   # int32 [n_c_domains]
    index: np.ndarray           # int32 [n_c_domains + 1]  ghost sections
    send_index: np.ndarray      # int32 [n_c_domains + 1]
    send_list: np.ndarray       # int32 [n_send]  local ids

    @property
    def n_c_domains(self) -> int:
        return len(self.c_domain_rank)

    @property
    def n_ghost(self) -> int:
        return int(self.index[-1])


def partition_dims(n_ranks: int):
    """Bricks per direction (px, py, pz).  Powers of two split z, then y, then x, which
    is what the reference's default Morton-box partitioning (src/mesh/cs_partition.cpp:
    2947-2948) yields on a cube: halves, quarter bricks, octants."""
    p = [1, 1, 1]
    n, d = n_ranks, 2
    while n > 1:
        f = 2 if n % 2 == 0 else n
        p[d] *= f
        n //= f
        d = (d - 1) % 3
    return tuple(p)


def _bounds(N, p):
    return [(N * r) // p for r in range(p + 1)]


def cube_boundary_face_cells(N: int) -> np.ndarray:
    """Adjacent cell of every boundary face of the N^3 cube: the six sides in the order
    x-min, x-max, y-min, y-max, z-min, z-max, cells in increasing id inside a side (a corner
    cell appears three times)."""
    ids = np.arange(N * N * N, dtype=np.int32).reshape(N, N, N)      # [k, j, i]
    sides = [ids[:, :, 0], ids[:, :, N - 1], ids[:, 0, :], ids[:, N - 1, :], ids[0, :, :],
             ids[N - 1, :, :]]
    return np.concatenate([x.ravel() for x in sides]).astype(np.int32)


def cube_poisson_local(N: int, n_ranks: int, rank: int, bc: str = "neumann",
                       eps_factor: float = 1e-6):
    """The rows of cube_poisson(N, bc) owned by `rank` out of `n_ranks` bricks, in the
    layout an MPI run of the reference would hold them: local cells numbered in global
    order, ghost cells appended (grouped by owner rank ascending, global order inside a
    section), local MSR column ids, local interior-face numbering (faces on a parallel
    boundary included, numbered in global face order).  Returns (CubeSystem, HaloDesc,
    global_ids) where global_ids[n_cols_ext] maps local columns to global cell ids."""
    px, py, pz = partition_dims(n_ranks)
    bx, by, bz = _bounds(N, px), _bounds(N, py), _bounds(N, pz)
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    x0, x1, y0, y1, z0, z1 = bx[rx], bx[rx + 1], by[ry], by[ry + 1], bz[rz], bz[rz + 1]
    nx, ny, nz = x1 - x0, y1 - y0, z1 - z0
    n = nx * ny * nz
    h = 1.0 / N

    lid = np.arange(n, dtype=np.int64)
    I = x0 + lid % nx
    J = y0 + (lid // nx) % ny
    K = z0 + lid // (nx * ny)
    gid = I + J * N + K * N * N

    def owner_of(ii, jj, kk):
        ox = np.searchsorted(bx, ii, side="right") - 1
        oy = np.searchsorted(by, jj, side="right") - 1
        oz = np.searchsorted(bz, kk, side="right") - 1
        return ox + oy * px + oz * px * py

    n_fx = (N - 1) * N * N
    n_fy = n_fx
    dirs = [(0, 0, -1), (0, -1, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    rows, cols_g, own, fgid, sgn = [], [], [], [], []
    for (dx, dy, dz) in dirs:
        ii, jj, kk = I + dx, J + dy, K + dz
        ok = (ii >= 0) & (ii < N) & (jj >= 0) & (jj < N) & (kk >= 0) & (kk < N)
        s = np.nonzero(ok)[0]
        ii, jj, kk = ii[s], jj[s], kk[s]
        rows.append(s)
        cols_g.append(ii + jj * N + kk * N * N)
        own.append(owner_of(ii, jj, kk))
        # global interior face id between the two cells (cs_mesh_cartesian ordering)
        hi_i, hi_j, hi_k = np.maximum(ii, I[s]), np.maximum(jj, J[s]), np.maximum(kk, K[s])
        if dx != 0:
            f = (hi_i - 1) + hi_j * (N - 1) + hi_k * (N - 1) * N
        elif dy != 0:
            f = n_fx + hi_i + (hi_j - 1) * N + hi_k * N * (N - 1)
        else:
            f = n_fx + n_fy + hi_i + hi_j * N + (hi_k - 1) * N * N
        fgid.append(f)
        sgn.append(np.full(len(s), 1 if (dx + dy + dz) > 0 else -1, dtype=np.int16))
    rows = np.concatenate(rows)
    cols_g = np.concatenate(cols_g)
    own = np.concatenate(own)
    fgid = np.concatenate(fgid)
    sgn = np.concatenate(sgn)

    # ghosts: unique (owner, global id), owner ascending then global id ascending
    is_ghost = own != rank
    gkey = own[is_ghost] * (N ** 3) + cols_g[is_ghost]
    ukey = np.unique(gkey)
    g_owner = (ukey // (N ** 3)).astype(np.int32)
    g_gid = ukey % (N ** 3)
    n_ghost = len(ukey)
    c_domain_rank = np.unique(g_owner).astype(np.int32)
    index = np.concatenate([[0], np.cumsum([np.count_nonzero(g_owner == r)
                                            for r in c_domain_rank])]).astype(np.int32)

    # local column ids
    col_l = np.empty(len(rows), dtype=np.int64)
    loc = ~is_ghost
    ci, cj, ck = cols_g[loc] % N, (cols_g[loc] // N) % N, cols_g[loc] // (N * N)
    col_l[loc] = (ci - x0) + (cj - y0) * nx + (ck - z0) * nx * ny
    col_l[is_ghost] = n + np.searchsorted(ukey, gkey)

    order = np.lexsort((col_l, rows))
    rows, col_l, fgid, sgn = rows[order], col_l[order], fgid[order], sgn[order]
    counts = np.bincount(rows, minlength=n)
    row_index = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=row_index[1:])

    # send lists: my cells that are ghosts of neighbour d, in global (= local) order
    send_list, send_index = [], [0]
    entry_own = own[order]
    for d in c_domain_rank:
        cells = np.unique(rows[entry_own == d])
        send_list.append(cells)
        send_index.append(send_index[-1] + len(cells))
    send_list = (np.concatenate(send_list) if send_list else np.zeros(0)).astype(np.int32)

    # local faces in global face order
    ufaces, c2f = np.unique(fgid, return_inverse=True)
    n_faces = len(ufaces)
    face_u_normal = np.zeros((n_faces, 3), dtype=np.float64)
    face_u_normal[ufaces < n_fx, 0] = 1.0
    face_u_normal[(ufaces >= n_fx) & (ufaces < n_fx + n_fy), 1] = 1.0
    face_u_normal[ufaces >= n_fx + n_fy, 2] = 1.0
    face_cells = np.zeros((n_faces, 2), dtype=np.int32)
    face_cells[c2f[sgn > 0], 0] = rows[sgn > 0]
    face_cells[c2f[sgn > 0], 1] = col_l[sgn > 0]
    face_cells[c2f[sgn < 0], 1] = rows[sgn < 0]
    face_cells[c2f[sgn < 0], 0] = col_l[sgn < 0]

    x_val = np.full(len(rows), -h, dtype=np.float64)
    d_val = counts.astype(np.float64) * h
    if bc == "dirichlet":
        d_val = d_val + (6 - counts).astype(np.float64) * (2.0 * h)
    elif bc == "neumann":
        d_val = d_val + eps_factor * h
    else:
        raise ValueError("bc must be 'neumann' or 'dirichlet'")

    gids = np.concatenate([gid, g_gid])
    cen = np.empty((n + n_ghost, 3), dtype=np.float64)
    cen[:, 0] = (gids % N + 0.5) * h
    cen[:, 1] = ((gids // N) % N + 0.5) * h
    cen[:, 2] = (gids // (N * N) + 0.5) * h
    rhs = (np.sin(2 * np.pi * cen[:n, 0]) * np.sin(2 * np.pi * cen[:n, 1])
           * np.sin(2 * np.pi * cen[:n, 2])) * (h ** 3)

    s = CubeSystem(
        N=N, n_rows=n, n_faces=n_faces, row_index=row_index.astype(np.int32),
        col_id=col_l.astype(np.int32), d_val=d_val, x_val=x_val, face_cells=face_cells,
        c2f=c2f.astype(np.int32), c2f_sgn=sgn.astype(np.int16), cell_cen=cen,
        cell_vol=np.full(n + n_ghost, h ** 3, dtype=np.float64),
        face_u_normal=face_u_normal, face_surf=np.full(n_faces, h * h, dtype=np.float64),
        rhs=rhs, meta={"bc": bc, "eps_factor": eps_factor, "h": h, "rank": rank,
                       "n_ranks": n_ranks, "n_cols_ext": n + n_ghost})
    halo = HaloDesc(n_local_elts=n, c_domain_rank=c_domain_rank, index=index,
                    send_index=np.array(send_index, dtype=np.int32), send_list=send_list)
    return s, halo, gids


# --------------------------------------------------------------------------- config 4

def cube_block3_convection_diffusion(N: int, seed: int = 1234):
    """Non-symmetric 3x3 block-diagonal system on the hex cube (BASELINE config 4 in the
    small: velocity convection-diffusion).  Extra-diagonal entries are scalars
    `xa = -h +- 0.5*phi_f` (upwind-like, two values per face => non-symmetric), diagonal
    blocks are `(sum of -xa + eps) I + small dense coupling` (diagonally dominant).
    Returns (row_index, col_id, d_val[n,3,3], x_val, rhs[3n])."""
    s = cube_poisson(N, "dirichlet", with_mesh=False)
    n = s.n_rows
    ri = s.row_index.astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ri))
    cols = s.col_id.astype(np.int64)
    h = 1.0 / N
    lo, hi = np.minimum(rows, cols), np.maximum(rows, cols)
    phi = 0.6 * h * np.cos(0.37 * lo + 0.11 * hi)          # face mass flux, same for both sides
    x_val = -h + np.where(cols > rows, 0.5, -0.5) * phi
    row_sum = np.zeros(n)
    np.add.at(row_sum, rows, -x_val)
    rng = np.random.default_rng(seed)
    d_val = np.zeros((n, 3, 3))
    for k in range(3):
        d_val[:, k, k] = row_sum + (0.05 + 0.02 * k) * h + (6 - np.diff(ri)) * 2.0 * h
    d_val += 0.02 * h * rng.standard_normal((n, 3, 3))
    g = np.arange(3 * n, dtype=np.float64)
    rhs = np.sin(0.013 * g + 0.4) * h ** 3
    return s.row_index, s.col_id, d_val, x_val, rhs


def cube_block_convection_diffusion(N: int, db: int, seed: int = 4321):
    """The system of cube_block3_convection_diffusion with db x db diagonal blocks (any block
    size of CS_MATRIX_BLOCK_D: 6 for symmetric tensors, others for the generic code paths).
    Returns (row_index, col_id, d_val[n,db,db], x_val, rhs[db*n])."""
    s = cube_poisson(N, "dirichlet", with_mesh=False)
    n = s.n_rows
    ri = s.row_index.astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ri))
    cols = s.col_id.astype(np.int64)
    h = 1.0 / N
    lo, hi = np.minimum(rows, cols), np.maximum(rows, cols)
    phi = 0.6 * h * np.cos(0.37 * lo + 0.11 * hi)
    x_val = -h + np.where(cols > rows, 0.5, -0.5) * phi
    row_sum = np.zeros(n)
    np.add.at(row_sum, rows, -x_val)
    rng = np.random.default_rng(seed + db)
    d_val = np.zeros((n, db, db))
    for k in range(db):
        d_val[:, k, k] = row_sum + (0.05 + 0.02 * k) * h + (6 - np.diff(ri)) * 2.0 * h
    d_val += 0.02 * h * rng.standard_normal((n, db, db))
    g = np.arange(db * n, dtype=np.float64)
    rhs = np.sin(0.013 * g + 0.4) * h ** 3
    return s.row_index, s.col_id, d_val, x_val, rhs
