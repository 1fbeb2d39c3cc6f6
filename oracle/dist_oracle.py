"""Multi-rank emulation of the reference's distributed solve -- TEST INFRASTRUCTURE ONLY.

All ranks of a partitioned system live in ONE process and advance in lockstep; a halo
exchange is a numpy copy from the owner's send list to the neighbour's ghost section and
a global sum adds the per-rank partial sums in rank order.  Per-rank arithmetic is done
by the pinned C restatement (oracle/cs_oracle.c); what this file adds is the restatement
of the rank-coupling logic of the reference:

* cs_halo_sync_var               src/base/cs_halo.cpp:1739-2220 (pack, exchange, ghosts
                                 stored after the local values, one section per neighbour)
* _coarsen_halo                  src/alge/cs_grid.cpp:924-1329  (coarse send lists and
                                 coarse ghost numbering)
* _coarse_row_count_and_halo     src/alge/cs_grid.cpp:1346-1389 (global row count)
* coarse cell_cen / cell_vol sync  src/alge/cs_grid.cpp:6561-6581
* _setup_hierarchy stop criteria on GLOBAL row counts  src/alge/cs_multigrid.cpp:2518-2529
* _dot_products_* + MPI_Allreduce  src/alge/cs_sles_it.cpp:341-690

Parity status: the reference has no MPI in this image, so _coarsen_halo cannot be run to
pin this restatement ("parity unpinned" for that function); everything per-rank goes
through the pinned oracle, and tests check the distributed solution against the
single-rank reference solution.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from oracle.pyoracle import OraGrid, Oracle, _dp, _f64, _i32, _ip, _np_from, _sp


@dataclass
class Halo:
    n_local: int
    ranks: np.ndarray        # neighbour ranks, ascending
    index: np.ndarray        # ghost sections [n_dom+1]
    send_index: np.ndarray   # [n_dom+1]
    send_list: np.ndarray    # local ids

    @property
    def n_ghost(self):
        return int(self.index[-1])


def halo_from_desc(h) -> Halo:
    return Halo(int(h.n_local_elts), np.asarray(h.c_domain_rank, dtype=np.int32).copy(),
                np.asarray(h.index, dtype=np.int32).copy(),
                np.asarray(h.send_index, dtype=np.int32).copy(),
                np.asarray(h.send_list, dtype=np.int32).copy())


def halo_sync(halos, arrays, stride=1):
    """cs_halo_sync_var on every rank at once: arrays[r] has (n_local + n_ghost)*stride
    entries; ghost section i of rank r <- send section of rank ranks[i] aimed at r."""
    for r, h in enumerate(halos):
        a = arrays[r].reshape(-1, stride) if stride > 1 else arrays[r]
        for i, q in enumerate(h.ranks):
            hq = halos[q]
            j = int(np.nonzero(hq.ranks == r)[0][0])
            src = arrays[q].reshape(-1, stride) if stride > 1 else arrays[q]
            ids = hq.send_list[hq.send_index[j]:hq.send_index[j + 1]]
            a[h.n_local + h.index[i]:h.n_local + h.index[i + 1]] = src[ids]


def coarsen_halos(halos, f_c_rows, c_n_rows):
    """_coarsen_halo for every rank: fills the ghost part of f_c_rows[r] (size
    n_cols_ext) with coarse ghost ids and returns the coarse halos."""
    R = len(halos)
    coarse_send = []
    c_send_lists, c_send_index = [], []
    for r, h in enumerate(halos):
        f_c_rows[r][h.n_local:] = -1                           # :960-962
        cs = np.full(len(h.send_list), -1, dtype=np.int32)
        lists, sidx = [], [0]
        for i in range(len(h.ranks)):                          # :988-1061, :1152-1209
            sub = {}
            lst = []
            for ii in range(h.send_index[i], h.send_index[i + 1]):
                jj = int(f_c_rows[r][h.send_list[ii]])
                if jj < 0:        # penalised row: sub_num[0] == -2 (:974, :1027-1032)
                    cs[ii] = -2
                    continue
                if jj not in sub:
                    sub[jj] = len(sub)
                    lst.append(jj)
                cs[ii] = sub[jj]
            lists.append(np.array(lst, dtype=np.int32))
            sidx.append(sidx[-1] + len(lst))
        coarse_send.append(cs)
        c_send_lists.append(lists)
        c_send_index.append(sidx)
    # _exchange_halo_coarsening (:802-910)
    for r, h in enumerate(halos):
        for i, q in enumerate(h.ranks):
            hq = halos[q]
            j = int(np.nonzero(hq.ranks == r)[0][0])
            f_c_rows[r][h.n_local + h.index[i]:h.n_local + h.index[i + 1]] \
                = coarse_send[q][hq.send_index[j]:hq.send_index[j + 1]]
    out = []
    for r, h in enumerate(halos):
        n_elts = 0
        index = [0]
        for i in range(len(h.ranks)):                          # :1076-1140
            jj = -1
            for ii in range(h.n_local + h.index[i], h.n_local + h.index[i + 1]):
                if f_c_rows[r][ii] > -1:
                    if f_c_rows[r][ii] > jj:
                        jj += 1
                    f_c_rows[r][ii] += c_n_rows[r] + n_elts
            n_elts += jj + 1
            index.append(n_elts)
        keep = [i for i in range(len(h.ranks))                 # prune (:1218-1304)
                if (index[i + 1] - index[i]) + (c_send_index[r][i + 1] - c_send_index[r][i]) > 0]
        ranks = h.ranks[keep] if keep else np.zeros(0, dtype=np.int32)
        idx = np.array([0] + [index[i + 1] for i in keep], dtype=np.int32)
        sidx = np.array([0] + [c_send_index[r][i + 1] for i in keep], dtype=np.int32)
        sl = (np.concatenate([c_send_lists[r][i] for i in keep]) if keep
              else np.zeros(0, dtype=np.int32))
        out.append(Halo(int(c_n_rows[r]), ranks.astype(np.int32), idx, sidx, sl.astype(np.int32)))
    return out


class DistOracle:
    """Hierarchy + V-cycle / PCG solves over all ranks of a partitioned system."""

    def __init__(self, systems, halos, oracle: Oracle | None = None):
        self.ora = oracle or Oracle()
        L = self.ora.lib
        L.ora_grid_create_fine_ext.restype = C.POINTER(OraGrid)
        L.ora_grid_coarsen_begin.restype = C.POINTER(OraGrid)
        L.ora_grid_coarsen_begin.argtypes = [C.POINTER(OraGrid), C.POINTER(C.c_int), C.c_int, C.c_int]
        L.ora_grid_coarsen_finish.argtypes = [C.POINTER(OraGrid), C.c_double]
        L.ora_grid_free.argtypes = [C.POINTER(OraGrid)]
        self.R = len(systems)
        self.systems = systems
        self.levels = []      # levels[l] = dict(grids=[...], halos=[...], coarse_row=[...])
        grids = []
        for s in systems:
            nce = int(s.meta.get("n_cols_ext", s.n_rows))
            ri, ci = _i32(s.row_index), _i32(s.col_id)
            d, xv = _f64(s.d_val), _f64(s.x_val)
            c2f, sgn = _i32(s.c2f), np.ascontiguousarray(s.c2f_sgn, dtype=np.int16)
            cen, vol = _f64(s.cell_cen), _f64(s.cell_vol)
            fun, fs = _f64(s.face_u_normal), _f64(s.face_surf)
            grids.append(L.ora_grid_create_fine_ext(
                s.n_rows, nce, _ip(ri), _ip(ci), _dp(d), _dp(xv), s.n_faces, _ip(c2f), _sp(sgn),
                _dp(cen), _dp(vol), _dp(fun), _dp(fs)))
        self.levels.append(dict(grids=grids, halos=[halo_from_desc(h) for h in halos],
                                coarse_row=None))

    # ------------------------------------------------------------------ views
    @staticmethod
    def _mat(g):
        g = g.contents
        return (_np_from(g.row_index, g.n_rows + 1, np.int32), _np_from(g.col_id, g.nnz, np.int32),
                _np_from(g.d_val, g.n_rows, np.float64), _np_from(g.x_val, g.nnz, np.float64))

    def level_arrays(self, l, r):
        g = self.levels[l]["grids"][r].contents
        ri, ci, d, xv = self._mat(self.levels[l]["grids"][r])
        out = dict(n_rows=g.n_rows, n_cols_ext=g.n_cols_ext, nnz=g.nnz, row_index=ri, col_id=ci,
                   d_val=d, x_val=xv)
        if l > 0:
            pf = self.levels[l - 1]["grids"][r].contents
            out["coarse_row"] = _np_from(g.coarse_row, max(pf.n_cols_ext, pf.n_rows), np.int32)
        h = self.levels[l]["halos"][r]
        out["halo"] = h
        return out

    # ------------------------------------------------------------------ setup
    def mg_setup(self, aggregation_limit=3, n_levels_max=25, n_g_rows_min=30, p0p1_relax=0.95):
        L = self.ora.lib
        n_g_rows = sum(g.contents.n_rows for g in self.levels[0]["grids"])
        while True:
            n_g_prev = n_g_rows
            if len(self.levels) >= n_levels_max:
                break
            f = self.levels[-1]
            f_level = len(self.levels) - 1
            f_c_rows, c_n_rows = [], []
            for r in range(self.R):
                g = f["grids"][r].contents
                ri, ci, d, xv = self._mat(f["grids"][r])
                nc, fc = self.ora.aggregation(f_level, ri, ci, d, xv, aggregation_limit, 1)
                ext = np.full(max(g.n_cols_ext, g.n_rows), -1, dtype=np.int32)
                ext[:g.n_rows] = fc
                f_c_rows.append(ext)
                c_n_rows.append(nc)
            n_g_rows = int(sum(c_n_rows))
            add = True
            if n_g_rows == 0:
                add = False
            elif len(self.levels) > 1:
                if n_g_rows < n_g_rows_min or n_g_rows > 0.8 * n_g_prev:
                    add = False
            if not add:
                break
            c_halos = coarsen_halos(f["halos"], f_c_rows, c_n_rows)
            c_grids = []
            for r in range(self.R):
                nce = c_n_rows[r] + c_halos[r].n_ghost
                c_grids.append(L.ora_grid_coarsen_begin(f["grids"][r], _ip(f_c_rows[r]),
                                                        c_n_rows[r], nce))
            # ghost cell centres and volumes (cs_grid.cpp:6561-6581)
            cens, vols = [], []
            for r in range(self.R):
                g = c_grids[r].contents
                cens.append(np.ctypeslib.as_array(g.cell_cen, shape=(3 * g.n_cols_ext,)))
                vols.append(np.ctypeslib.as_array(g.cell_vol, shape=(g.n_cols_ext,)))
            halo_sync(c_halos, cens, 3)
            halo_sync(c_halos, vols, 1)
            for r in range(self.R):
                L.ora_grid_coarsen_finish(c_grids[r], p0p1_relax)
            self.levels.append(dict(grids=c_grids, halos=c_halos, coarse_row=f_c_rows))
        return len(self.levels)

    # ------------------------------------------------------------------ kernels
    def _spmv(self, l, xs):
        """y_r = A_r . x_r after a halo sync of x (x_r has n_cols_ext entries)."""
        lv = self.levels[l]
        halo_sync(lv["halos"], xs)
        ys = []
        for r in range(self.R):
            ri, ci, d, xv = self._cached(l, r)
            ys.append(self.ora.spmv(ri, ci, d, xv, xs[r]))
        return ys

    def _cached(self, l, r):
        key = (l, r)
        if not hasattr(self, "_cache"):
            self._cache = {}
        if key not in self._cache:
            self._cache[key] = self._mat(self.levels[l]["grids"][r])
        return self._cache[key]

    def _gsum(self, parts):
        s = 0.0
        for p in parts:       # rank order
            s += p
        return s

    def _dot3(self, x, y, z):
        p = [self.ora.dot_xx_xy_yz(x[r], y[r], z[r]) for r in range(self.R)]
        return tuple(self._gsum([q[k] for q in p]) for k in range(3))

    # ------------------------------------------------------------------ PCG
    def pcg(self, l, rhs, vx, n_max_iter, precision, r_norm, use_vx_ini, history=None):
        """_conjugate_gradient_sr_p0 (cs_sles_it.cpp:1441-1645) over all ranks; vx[r] has
        n_cols_ext entries and is updated in place.  Returns (cvg, n_iter, residual,
        initial_residual)."""
        R = self.R
        lv = self.levels[l]
        n = [g.contents.n_rows for g in lv["grids"]]
        nce = [max(g.contents.n_cols_ext, g.contents.n_rows) for g in lv["grids"]]
        ad_inv = [1.0 / self._cached(l, r)[2] for r in range(R)]

        def ext(r):
            return np.zeros(nce[r])

        def test(n_iter, residual, state):
            state["n_iter"], state["residual"] = n_iter, residual
            if history is not None:
                history.append(residual)
            if precision > 0 and residual < precision * r_norm:
                return 1
            if n_iter >= n_max_iter:
                return -1
            if residual > state["initial"] * 10000.0 and residual > 100.:
                return -3
            if math.isnan(residual) or math.isinf(residual):
                return -3
            return 0

        st = dict(n_iter=0, residual=-1., initial=-1.)
        rk, dk, gk, zk, sk = ([ext(r) for r in range(R)] for _ in range(5))
        if use_vx_ini:
            y = self._spmv(l, vx)
            for r in range(R):
                rk[r][:n[r]] = y[r] - rhs[r][:n[r]]
                dk[r][:n[r]] = rk[r][:n[r]] * ad_inv[r]
        else:
            for r in range(R):
                vx[r][:n[r]] = 0
                rk[r][:n[r]] = -rhs[r][:n[r]]
                dk[r][:n[r]] = rk[r][:n[r]] * ad_inv[r]
        y = self._spmv(l, dk)
        for r in range(R):
            zk[r][:n[r]] = y[r]
        loc = lambda v: [v[r][:n[r]] for r in range(R)]
        rr, ro_0, ro_1 = self._dot3(loc(rk), loc(dk), loc(zk))
        residual = math.sqrt(rr)
        st["initial"] = residual
        cvg = test(0, residual, st)
        n_iter = 0
        rk_gkm1 = 0.
        if cvg == 0:
            n_iter = 1
            d_ro_1 = 1. / ro_1 if abs(ro_1) > 2.2250738585072014e-308 else 0.
            alpha = -ro_0 * d_ro_1
            rk_gkm1 = ro_0
            for r in range(R):
                m = n[r]
                vx[r][:m] += alpha * dk[r][:m]
                rk[r][:m] += alpha * zk[r][:m]
                gk[r][:m] = rk[r][:m] * ad_inv[r]
            residual = math.sqrt(self._gsum([self.ora.dot_xx(rk[r][:n[r]]) for r in range(R)]))
            cvg = test(n_iter, residual, st)
        while cvg == 0:
            y = self._spmv(l, gk)
            for r in range(R):
                sk[r][:n[r]] = y[r]
            rr, rk_gk, gk_sk = self._dot3(loc(rk), loc(gk), loc(sk))
            residual = math.sqrt(rr)
            if n_iter > 1:
                cvg = test(n_iter, residual, st)
            if cvg != 0:
                break
            n_iter += 1
            beta = rk_gk / rk_gkm1 if abs(rk_gkm1) > 2.2250738585072014e-308 else 0.
            rk_gkm1 = rk_gk
            ro_1 = gk_sk - beta * beta * ro_1
            ro_0 = rk_gk
            d_ro_1 = 1. / ro_1 if abs(ro_1) > 2.2250738585072014e-308 else 0.
            alpha = -ro_0 * d_ro_1
            for r in range(R):
                m = n[r]
                dk[r][:m] = gk[r][:m] + beta * dk[r][:m]
                vx[r][:m] += alpha * dk[r][:m]
                zk[r][:m] = sk[r][:m] + beta * zk[r][:m]
                rk[r][:m] += alpha * zk[r][:m]
                gk[r][:m] = rk[r][:m] * ad_inv[r]
        return cvg, st["n_iter"], st["residual"], st["initial"]

    def pcg_solve(self, rhs, n_max_iter, precision, r_norm, vx0=None):
        lv = self.levels[0]
        nce = [max(g.contents.n_cols_ext, g.contents.n_rows) for g in lv["grids"]]
        vx = [np.zeros(nce[r]) for r in range(self.R)]
        if vx0 is not None:
            for r in range(self.R):
                vx[r][:len(vx0[r])] = vx0[r]
        hist = []
        cvg, n_iter, res, ini = self.pcg(0, rhs, vx, n_max_iter, precision, r_norm,
                                         vx0 is not None, hist)
        return dict(cvg=cvg, n_iter=n_iter, residual=res, initial_residual=ini,
                    history=np.array(hist), vx=vx)

    # ------------------------------------------------------------------ V-cycle
    def mg_solve(self, rhs, precision, r_norm, n_max_cycles=100, n_max_iter=(2, 10, 500),
                 precision_mult=(1., 1., 1.)):
        """cs_multigrid_solve / _multigrid_v_cycle (cs_multigrid.cpp:5362, :3127) over all
        ranks.  n_equiv_iter uses GLOBAL row counts (n_g_rows, :3341)."""
        R, nl = self.R, len(self.levels)
        coarsest = nl - 1
        n = [[g.contents.n_rows for g in lv["grids"]] for lv in self.levels]
        nce = [[max(g.contents.n_cols_ext, g.contents.n_rows) for g in lv["grids"]]
               for lv in self.levels]
        n_g = [sum(x) for x in n]
        denom = 1.0 / n_g[0]
        rhs_l = [None] + [[np.zeros(nce[l][r]) for r in range(R)] for l in range(1, nl)]
        vx_l = [[np.zeros(nce[l][r]) for r in range(R)] for l in range(nl)]
        rhs_l[0] = rhs
        cvg, n_equiv, n_cycles = 0, 0, 0
        initial_residual = residual = -1.
        cyc = []
        while cvg == 0:
            cycle_id = n_cycles + 1
            end_cycle = False
            for level in range(coarsest):
                ini = (n_cycles > 0) if level == 0 else False
                c_cvg, n_it, _, ir = self.pcg(level, rhs_l[level], vx_l[level], n_max_iter[0],
                                              precision * precision_mult[0], r_norm, ini)
                if level == 0 and cycle_id == 1:
                    initial_residual = ir
                if c_cvg < -2:
                    end_cycle = True
                    break
                y = self._spmv(level, vx_l[level])
                wr = [rhs_l[level][r][:n[level][r]] - y[r] for r in range(R)]
                if level == 0:
                    residual = math.sqrt(self._gsum([self.ora.dot_xx(w) for w in wr]))
                    ir0 = residual if cycle_id == 1 else initial_residual
                    cyc.append(residual)
                    if residual < precision * r_norm:
                        cvg = 1
                    elif cycle_id > n_max_cycles:
                        cvg = -1
                    elif residual > ir0 * 10000.0 and residual > 100.:
                        cvg = -3
                    elif math.isnan(residual) or math.isinf(residual):
                        cvg = -3
                    if cvg != 0:
                        end_cycle = True
                        n_equiv = int(n_equiv + n_it * n_g[level] * denom)
                        break
                n_equiv = int(n_equiv + n_it * n_g[level] * denom)
                cr = self.levels[level + 1]["coarse_row"]
                for r in range(R):
                    c = np.zeros(n[level + 1][r])
                    self.ora.lib.ora_restrict(n[level][r], n[level + 1][r], _ip(cr[r]),
                                              _dp(_f64(wr[r])), _dp(c))
                    rhs_l[level + 1][r][:n[level + 1][r]] = c
            if not end_cycle:
                c_cvg, n_it, _, _ = self.pcg(coarsest, rhs_l[coarsest], vx_l[coarsest],
                                             n_max_iter[2], precision * precision_mult[2],
                                             r_norm, False)
                n_equiv = int(n_equiv + n_it * n_g[coarsest] * denom)
                if c_cvg < -2:
                    end_cycle = True
            if not end_cycle:
                for level in range(coarsest - 1, -1, -1):
                    cr = self.levels[level + 1]["coarse_row"]
                    for r in range(R):
                        f = np.ascontiguousarray(vx_l[level][r][:n[level][r]])
                        self.ora.lib.ora_prolong_inc(n[level][r], _ip(cr[r]),
                                                     _dp(_f64(vx_l[level + 1][r])), _dp(f))
                        vx_l[level][r][:n[level][r]] = f
                    if level > 0:
                        c_cvg, n_it, _, _ = self.pcg(level, rhs_l[level], vx_l[level],
                                                     n_max_iter[1],
                                                     precision * precision_mult[1], r_norm, True)
                        n_equiv = int(n_equiv + n_it * n_g[level] * denom)
                        if c_cvg < -2:
                            break
            n_cycles += 1
        return dict(cvg=cvg, n_iter=n_equiv, residual=residual, n_cycles=n_cycles,
                    cycle_residuals=np.array(cyc), vx=[v[:n[0][r]] for r, v in enumerate(vx_l[0])])

    # ------------------------------------------------- FCG + V-cycle preconditioner
    SMOOTHERS = {3: ("jacobi", (1.0,)), 14: ("relaxed", (2. / 3.,)),
                 15: ("relaxed", (1.7319, 0.5695)), 16: ("relaxed", (2.2473, 0.8573, 0.5296))}

    def _spmv_e(self, l, xs):
        """E_r . x_r (diagonal excluded) after a halo sync of x."""
        lv = self.levels[l]
        halo_sync(lv["halos"], xs)
        out = []
        for r in range(self.R):
            ri, ci, d, xv = self._cached(l, r)
            out.append(self.ora.spmv(ri, ci, d, xv, xs[r], exclude_diag=True))
        return out

    def smoother(self, l, stype, n_max_iter, rhs, vx, use_vx_ini):
        """cs_multigrid_smoother _jacobi / _relaxed_jacobi (cs_multigrid_smoother.cpp:143-373)
        over all ranks, in place in vx[r] (n_cols_ext entries)."""
        mode, w = self.SMOOTHERS[stype]
        n_iter = n_max_iter
        if mode == "relaxed" and n_iter % len(w):
            n_iter += len(w) - n_iter % len(w)        # whole schedules (:1888-1898)
        R = self.R
        n = [g.contents.n_rows for g in self.levels[l]["grids"]]
        ad_inv = [1.0 / self._cached(l, r)[2] for r in range(R)]
        it0 = 0
        if not use_vx_ini:
            for r in range(R):
                b = rhs[r][:n[r]]
                vx[r][:n[r]] = (w[0] * b * ad_inv[r]) if mode == "relaxed" else (b * ad_inv[r])
            it0 = 1
        for it in range(it0, n_iter):
            wk = w[it % len(w)]
            ex = self._spmv_e(l, vx)
            for r in range(R):
                b, x = rhs[r][:n[r]], vx[r][:n[r]]
                t = b - ex[r]
                if mode == "relaxed":
                    vx[r][:n[r]] = (1.0 - wk) * x + wk * t * ad_inv[r]
                else:
                    vx[r][:n[r]] = t * ad_inv[r]
        return n_iter

    def vcycle_pc(self, rhs, vx, precision, r_norm, types=(16, 16, 0), n_max_iter=(3, 3, 500),
                  precision_mult=(-1., -1., 1.)):
        """_multigrid_v_cycle_pc (cs_multigrid.cpp:3569-4020) from a zero start over all
        ranks: vx[r] (n_cols_ext of level 0) receives M^-1 rhs."""
        R, nl = self.R, len(self.levels)
        coarsest = nl - 1
        n = [[g.contents.n_rows for g in lv["grids"]] for lv in self.levels]
        nce = [[max(g.contents.n_cols_ext, g.contents.n_rows) for g in lv["grids"]]
               for lv in self.levels]
        rhs_l = [rhs] + [[np.zeros(nce[l][r]) for r in range(R)] for l in range(1, nl)]
        vx_l = [vx] + [[np.zeros(nce[l][r]) for r in range(R)] for l in range(1, nl)]
        for level in range(coarsest):
            self.smoother(level, types[0], n_max_iter[0], rhs_l[level], vx_l[level], False)
            y = self._spmv(level, vx_l[level])
            wr = [rhs_l[level][r][:n[level][r]] - y[r] for r in range(R)]
            cr = self.levels[level + 1]["coarse_row"]
            for r in range(R):
                c = np.zeros(n[level + 1][r])
                self.ora.lib.ora_restrict(n[level][r], n[level + 1][r], _ip(cr[r]),
                                          _dp(_f64(wr[r])), _dp(c))
                rhs_l[level + 1][r][:n[level + 1][r]] = c
        self.pcg(coarsest, rhs_l[coarsest], vx_l[coarsest], n_max_iter[2],
                 precision * precision_mult[2], r_norm, False)
        for level in range(coarsest - 1, -1, -1):
            cr = self.levels[level + 1]["coarse_row"]
            for r in range(R):
                f = np.ascontiguousarray(vx_l[level][r][:n[level][r]])
                self.ora.lib.ora_prolong_inc(n[level][r], _ip(cr[r]),
                                             _dp(_f64(vx_l[level + 1][r])), _dp(f))
                vx_l[level][r][:n[level][r]] = f
            self.smoother(level, types[1], n_max_iter[1], rhs_l[level], vx_l[level], True)

    def fcg_mgpc(self, rhs, n_max_iter, precision, r_norm, types=(16, 16, 0),
                 mg_n_max_iter=(3, 3, 500), precision_mult=(-1., -1., 1.), vx0=None):
        """_flexible_conjugate_gradient (cs_sles_it.cpp:832-985) preconditioned by vcycle_pc,
        over all ranks; dot products summed per rank, then in rank order."""
        R = self.R
        lv = self.levels[0]
        n = [g.contents.n_rows for g in lv["grids"]]
        nce = [max(g.contents.n_cols_ext, g.contents.n_rows) for g in lv["grids"]]
        ext = lambda: [np.zeros(nce[r]) for r in range(R)]
        vx, rk, vk, wk, dk, qk = (ext() for _ in range(6))
        if vx0 is not None:
            for r in range(R):
                vx[r][:n[r]] = vx0[r][:n[r]]
            y = self._spmv(0, vx)
            for r in range(R):
                rk[r][:n[r]] = rhs[r][:n[r]] - y[r]
        else:
            for r in range(R):
                rk[r][:n[r]] = rhs[r][:n[r]]
        hist = []
        cvg, n_iter, residual, initial = 0, 0, -1., -1.
        rho_km1 = 0.
        tiny = 2.2250738585072014e-308
        loc = lambda v, r: v[r][:n[r]]
        while cvg == 0:
            for r in range(R):
                vk[r][:] = 0.
            self.vcycle_pc(rk, vk, precision, r_norm, types, mg_n_max_iter, precision_mult)
            y = self._spmv(0, vk)
            for r in range(R):
                wk[r][:n[r]] = y[r]
            a = b = g = rr = 0.0
            parts = [(self.ora.dot(loc(vk, r), loc(rk, r)), self.ora.dot(loc(vk, r), loc(wk, r)),
                      self.ora.dot(loc(vk, r), loc(qk, r)), self.ora.dot_xx(loc(rk, r)))
                     for r in range(R)]
            for p4 in parts:
                a += p4[0]; b += p4[1]; g += p4[2]; rr += p4[3]
            residual = math.sqrt(rr)
            hist.append(residual)
            if n_iter > 0:
                if precision > 0 and residual < precision * r_norm:
                    cvg = 1
                elif n_iter >= n_max_iter:
                    cvg = -1
                elif residual > initial * 10000.0 and residual > 100.:
                    cvg = -3
            else:
                initial = residual
            if cvg != 0:
                break
            if n_iter > 0:
                gk_rk1 = g / rho_km1 if abs(rho_km1) > tiny else 0.
                rho_k = b - g * gk_rk1
            else:
                gk_rk1 = 0.
                rho_k = b
            ak_rk = a / rho_k if abs(rho_k) > tiny else 0.
            for r in range(R):
                m = n[r]
                if n_iter > 0:
                    dk[r][:m] = vk[r][:m] - gk_rk1 * dk[r][:m]
                    qk[r][:m] = wk[r][:m] - gk_rk1 * qk[r][:m]
                else:
                    dk[r][:m] = vk[r][:m]
                    qk[r][:m] = wk[r][:m]
                vx[r][:m] = vx[r][:m] + ak_rk * dk[r][:m]
                rk[r][:m] = rk[r][:m] - ak_rk * qk[r][:m]
            rho_km1 = rho_k
            n_iter += 1
        return dict(cvg=cvg, n_iter=n_iter, residual=residual, initial_residual=initial,
                    history=np.array(hist), vx=[v[:n[r]] for r, v in enumerate(vx)])

    def free(self):
        for l, lv in enumerate(self.levels):
            for g in lv["grids"]:
                self.ora.lib.ora_grid_free(g)
        self.levels = []
