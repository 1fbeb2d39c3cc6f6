"""CPU tests of the product's host-side logic (no device work): mesh generator
and the host part of the hierarchy construction exported by libcsb200.so
(csb_host_aggregate_dx_msr / csb_host_coarse_structure), against the golden
fixtures and the oracle.  Index data: bit-exact."""
import numpy as np
import pytest

from code_saturne_b200 import api
from code_saturne_b200.mesh_cartesian import cube_poisson
from conftest import golden_cases


def test_cube_generator_structure():
    s = cube_poisson(5)
    assert s.n_rows == 125 and s.n_faces == 3 * 4 * 25 and s.nnz == 2 * s.n_faces
    ri = s.row_index
    rows = np.repeat(np.arange(s.n_rows), np.diff(ri))
    # sorted columns, no diagonal (src/alge/cs_matrix.cpp:579-581)
    for r in range(s.n_rows):
        c = s.col_id[ri[r]:ri[r + 1]]
        assert np.all(np.diff(c) > 0) and r not in c
    # cell -> face adjacency consistent with face_cells and its sign convention
    fc = s.face_cells[s.c2f]
    ok = np.where(s.c2f_sgn > 0, (fc[:, 0] == rows) & (fc[:, 1] == s.col_id),
                  (fc[:, 1] == rows) & (fc[:, 0] == s.col_id))
    assert ok.all()
    # faces ordered x, y, z and oriented low -> high id
    assert np.all(s.face_cells[:, 0] < s.face_cells[:, 1])
    assert np.all(s.face_u_normal[:100, 0] == 1.0)
    # symmetric, weakly diagonally dominant
    assert np.allclose(s.d_val, -np.add.reduceat(s.x_val, ri[:-1]) + 1e-6 / 5)
    assert abs(s.rhs.sum()) < 1e-12


def test_cube_generator_dirichlet_is_dominant():
    s = cube_poisson(4, "dirichlet")
    offd = -np.add.reduceat(s.x_val, s.row_index[:-1])
    assert np.all(s.d_val >= offd)
    assert np.any(s.d_val > offd)


@pytest.mark.parametrize("N,bc,path", golden_cases())
def test_host_hierarchy_structure_matches_golden(N, bc, path):
    """Walk the golden hierarchy level by level with the product's host code."""
    g = np.load(path)
    s = cube_poisson(N, bc)
    ri, ci, d, xv = s.row_index, s.col_id, s.d_val, s.x_val
    for l in range(1, int(g["mg_n_levels"])):
        nc, f_c_row = api.host_aggregate(l - 1, ri, ci, d, xv, 3, 1)
        assert np.array_equal(f_c_row, g[f"mg_l{l}_coarse_row"])
        assert nc == len(g[f"mg_l{l}_row_index"]) - 1
        cs = api.host_coarse_structure(ri, ci, f_c_row, nc)
        assert np.array_equal(cs["row_index"], g[f"mg_l{l}_row_index"])
        assert np.array_equal(cs["col_id"], g[f"mg_l{l}_col_id"])
        # upper entries number the coarse faces in order; lower ones mirror them
        up = cs["cell_face_sgn"] > 0
        assert np.array_equal(cs["cell_face"][up], np.arange(cs["n_faces"]))
        assert cs["n_faces"] * 2 == len(cs["col_id"])
        ri, ci = g[f"mg_l{l}_row_index"], g[f"mg_l{l}_col_id"]
        d, xv = g[f"mg_l{l}_d_val"], g[f"mg_l{l}_x_val"]


@pytest.mark.parametrize("N,n_blocks", [(12, 1), (24, 4), (33, 3), (40, 8)])
def test_host_aggregation_matches_oracle_with_blocks(oracle, N, n_blocks):
    s = cube_poisson(N)
    nc, f_c_row = api.host_aggregate(0, s.row_index, s.col_id, s.d_val, s.x_val, 3, n_blocks)
    nco, f_c_row_o = oracle.aggregation(0, s.row_index, s.col_id, s.d_val, s.x_val, 3, n_blocks)
    assert nc == nco
    assert np.array_equal(f_c_row, f_c_row_o)
    # every row aggregated, aggregate sizes bounded by limit + 1
    assert f_c_row.min() >= 0
    assert np.bincount(f_c_row).max() <= 4


def test_host_aggregation_penalized_and_isolated_rows(oracle):
    """Rows with a huge diagonal are left out (-1); isolated rows become singletons."""
    s = cube_poisson(8)
    d = s.d_val.copy()
    d[[3, 77, 300]] *= 1e6
    nc, f_c_row = api.host_aggregate(0, s.row_index, s.col_id, d, s.x_val, 3, 1)
    nco, ref = oracle.aggregation(0, s.row_index, s.col_id, d, s.x_val, 3, 1)
    assert np.array_equal(f_c_row, ref)
    assert list(f_c_row[[3, 77, 300]]) == [-1, -1, -1]
    # diagonal matrix: every row its own aggregate
    n = 300
    ri = np.zeros(n + 1, dtype=np.int32)
    nc, f = api.host_aggregate(1, ri, np.zeros(0, dtype=np.int32), np.ones(n), np.zeros(0), 3, 1)
    assert nc == n and np.array_equal(f, np.arange(n))


def test_host_aggregation_empty():
    ri = np.zeros(1, dtype=np.int32)
    nc, f = api.host_aggregate(0, ri, np.zeros(0, dtype=np.int32), np.zeros(0), np.zeros(0))
    assert nc == 0 and len(f) == 0


@pytest.mark.parametrize("N,bc", [(8, "neumann"), (12, "dirichlet")])
def test_host_pairwise_coarsening_matches_reference_golden(N, bc):
    """csb_host_coarsen_pairwise (the host recursion csb_multigrid_setup uses for
    CS_GRID_COARSENING_SPD_PW: _pairwise_msr + Galerkin values) level by level against the
    reference-built fixtures: maps, structures and values bit-exact.  No GPU needed."""
    import os
    from code_saturne_b200 import api
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"pairwise_N{N}_{bc}.npz"))
    s = cube_poisson(N, bc)
    for lim in (2, 4, 8):
        f = dict(row_index=s.row_index, col_id=s.col_id, d_val=s.d_val, x_val=s.x_val)
        for l in range(1, int(g[f"lim{lim}_n_levels"])):
            c = api.host_coarsen_pairwise(l - 1, f["row_index"], f["col_id"], f["d_val"],
                                          f["x_val"], lim)
            for k in ("coarse_row", "row_index", "col_id", "d_val", "x_val"):
                assert np.array_equal(c[k], g[f"lim{lim}_l{l}_{k}"]), (lim, l, k)
            f = c


def test_host_pairwise_edge_cases():
    from code_saturne_b200 import api
    # a diagonal matrix: every row is left alone (level > 0) ...
    ri = np.arange(5, dtype=np.int32) * 0
    c = api.host_coarsen_pairwise(1, ri, np.zeros(0, np.int32), np.ones(4), np.zeros(0), 4)
    assert c["n_rows"] == 4 and list(c["coarse_row"]) == [0, 1, 2, 3]
    # ... and at level 0 its rows are strongly diagonally dominant: no coarse rows at all
    c = api.host_coarsen_pairwise(0, ri, np.zeros(0, np.int32), np.ones(4), np.zeros(0), 4)
    assert c["n_rows"] == 0 and list(c["coarse_row"]) == [-1, -1, -1, -1]
    # two coupled rows form one pair
    ri = np.array([0, 1, 2], dtype=np.int32)
    ci = np.array([1, 0], dtype=np.int32)
    c = api.host_coarsen_pairwise(1, ri, ci, np.array([2., 2.]), np.array([-1., -1.]), 2)
    assert c["n_rows"] == 1 and list(c["coarse_row"]) == [0, 0]
    assert c["d_val"][0] == 2. and len(c["x_val"]) == 0
