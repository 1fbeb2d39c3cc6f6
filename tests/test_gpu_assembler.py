"""Matrix assembler: cs_matrix_assembler_t on one rank with an
MSR matrix -- csb_matrix_assembler_* (csrc/assembler.cu).  Bar: structure bit-exact, every
entry the reference's arrival-order sum bit for bit (one fold kernel, no atomics); the
assembled matrix multiplies and solves like any other."""
import os
import sys

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
from make_golden_assembler import CASES, assembler_input  # noqa: E402


def _assemble(csb, n, l0, rows, cols, batches):
    ma = csb.MatrixAssembler((l0, l0 + n))
    ma.add_g_ids(rows[:100], cols[:100])
    ma.add_g_ids(rows[100:], cols[100:])
    ma.compute()
    m = ma.create_matrix()
    ma.values_init(m)
    for b in batches:
        ma.values_add_g(*b)
    ma.values_finalize()
    return ma, m


@pytest.mark.parametrize("N,case", sorted(CASES.items()))
def test_assembler_matches_reference_golden(csb, N, case):
    l0, seed = case
    g = np.load(os.path.join(GOLD, f"assembler_N{N}.npz"))
    n, rows, cols, vals, batches = assembler_input(N, l0, seed)
    ma, m = _assemble(csb, n, l0, rows, cols, batches)
    ri, ci = ma.structure()
    assert np.array_equal(ri, g["row_index"]) and np.array_equal(ci, g["col_id"])
    d, x = m.get_coefficients()
    assert np.array_equal(d, g["d_val"]) and np.array_equal(x, g["x_val"])
    xv = np.sin(np.arange(n) * 0.37)
    assert np.array_equal(m.vector_multiply(xv.copy()), g["y"])
    # the declared sequence as one kernel over the plan kept in HBM; host, then device values
    for dev in (False, True):
        ma.values_init(m)
        if dev:
            import torch
            ma.values_add_declared(torch.from_numpy(vals).cuda())
        else:
            ma.values_add_declared(vals)
        ma.values_finalize()
        d, x = m.get_coefficients()
        assert np.array_equal(d, g["d_val_declared"]) and np.array_equal(x, g["x_val_declared"])
    m.destroy(); ma.destroy()


def test_assembler_vs_oracle_larger_and_solve(csb, oracle):
    """a 32^3 cube's faces plus long-range couplings; values_done between batches keeps the
    arrival order; PCG on the (symmetric positive definite) result matches the oracle's"""
    n, rows, cols, vals, batches = assembler_input(32, 12345, 5)
    # SPD: one full pass of the sequence (the face part and the long-range part are symmetric
    # by construction, the partial second pass of the fixture is not) plus a diagonal shift
    batches = batches[:4]
    shift = (np.arange(n, dtype=np.int64) + 12345, np.arange(n, dtype=np.int64) + 12345,
             np.full(n, 0.05))
    ma = csb.MatrixAssembler((12345, 12345 + n))
    ma.add_g_ids(rows, cols)
    ma.add_g_ids(shift[0], shift[1])
    ma.compute()
    m = ma.create_matrix()
    ma.values_init(m)
    ma.values_add_g(*batches[0]); ma.values_add_g(*batches[1])
    ma.values_done()
    for b in batches[2:] + [shift]:
        ma.values_add_g(*b)
    ma.values_finalize()
    ri, ci, d, x = oracle.assembler(n, np.concatenate([rows, shift[0]]),
                                    np.concatenate([cols, shift[1]]), batches + [shift], 12345)
    sri, sci = ma.structure()
    assert np.array_equal(sri, ri) and np.array_equal(sci, ci)
    dd, xx = m.get_coefficients()
    assert np.array_equal(dd, d) and np.array_equal(xx, x)
    rhs = np.cos(np.arange(n) * 0.013)
    r_norm = float(np.linalg.norm(rhs))
    c = csb.SlesIt(csb.SlesItType.PCG, 2000)
    r = c.solve(m, rhs, 1e-10, r_norm)
    o = oracle.pcg(ri, ci, d, x, rhs, 2000, 1e-10, r_norm)
    assert int(r["cvg"]) == o["cvg"] == 1 and abs(r["n_iter"] - o["n_iter"]) <= 2
    assert rel_l2(r["vx"][:n], o["vx"]) <= 1e-8
    c.destroy(); m.destroy(); ma.destroy()


def test_assembler_edge_cases(csb):
    """empty assembler, diagonal-only rows, coefficients outside the structure or the range"""
    ma = csb.MatrixAssembler((10, 14))
    ma.add_g_ids([10, 11, 13], [10, 11, 13])          # diagonal only: no extra-diagonal entry
    ma.compute()
    ri, ci = ma.structure()
    assert np.array_equal(ri, np.zeros(5, dtype=np.int32)) and ci.size == 0
    m = ma.create_matrix()
    ma.values_init(m)
    ma.values_add_g([10, 13, 10], [10, 13, 10], [1.0, 2.0, 0.5])
    with pytest.raises(csb.CsbError):
        ma.values_add_g([10], [11], [1.0])            # not declared
    ma.values_finalize()
    d, _ = m.get_coefficients()
    assert np.array_equal(d, [1.5, 0., 0., 2.0])
    assert np.array_equal(m.vector_multiply(np.ones(4)), [1.5, 0., 0., 2.0])
    m.destroy()
    with pytest.raises(csb.CsbError):
        ma.add_g_ids([10], [12])                      # after compute
    ma.destroy()
    ma = csb.MatrixAssembler((0, 4))
    with pytest.raises(csb.CsbError):
        ma.add_g_ids([0], [4])                        # column of another rank
    ma.compute()                                      # nothing declared
    m = ma.create_matrix()
    assert np.array_equal(m.vector_multiply(np.ones(4)), np.zeros(4))
    m.destroy(); ma.destroy()
