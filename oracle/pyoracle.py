"""ctypes bindings for the CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``  : oracle/libcs_oracle.so, the plain-C restatement (oracle/cs_oracle.c)
* ``RefLib``  : oracle/_ref/libcsref.so, the reference's own compiled CPU code
                driven by oracle/ref_harness.cpp (built by oracle/Makefile.ref)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

c_int_p = C.POINTER(C.c_int)
c_dbl_p = C.POINTER(C.c_double)
c_short_p = C.POINTER(C.c_short)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dbl_p)


def _sp(a):
    return None if a is None else a.ctypes.data_as(c_short_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build_oracle(force: bool = False) -> str:
    """Compile oracle/cs_oracle.c into oracle/libcs_oracle.so (gcc, seconds)."""
    so = os.path.join(_HERE, "libcs_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cs_oracle.c", "cs_oracle_native.c", "cs_oracle_assembler.c", "cs_oracle_rhs.c", "cs_oracle.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-f", "oracle/Makefile"], cwd=_ROOT,
                              stdout=subprocess.DEVNULL)
    return so


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libcsref.so"))


class OraGrid(C.Structure):
    _fields_ = [("level", C.c_int), ("n_rows", C.c_int), ("nnz", C.c_int),
                ("n_faces", C.c_int),
                ("row_index", c_int_p), ("col_id", c_int_p),
                ("d_val", c_dbl_p), ("x_val", c_dbl_p),
                ("coarse_row", c_int_p), ("cell_face", c_int_p),
                ("cell_face_sgn", c_short_p),
                ("cell_cen", c_dbl_p), ("cell_vol", c_dbl_p),
                ("face_normal", c_dbl_p), ("face_u_normal", c_dbl_p),
                ("face_surf", c_dbl_p), ("xa0", c_dbl_p), ("xa0ij", c_dbl_p),
                ("n_cols_ext", C.c_int), ("face_cell", c_int_p), ("xa", c_dbl_p)]


class OraMg(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("grids", C.POINTER(C.POINTER(OraGrid)))]


class OraMgOpt(C.Structure):
    _fields_ = [("n_max_cycles", C.c_int), ("n_max_iter", C.c_int * 3),
                ("precision_mult", C.c_double * 3)]


class OraMgPcOpt(C.Structure):
    _fields_ = [("type", C.c_int * 3), ("n_max_iter", C.c_int * 3),
                ("precision_mult", C.c_double * 3)]


# cs_sles_it_type_t values (src/alge/cs_sles_it.h:55-87)
SLES_PCG, SLES_FCG, SLES_IPCG, SLES_JACOBI, SLES_BICGSTAB = 0, 1, 2, 3, 4
SLES_L1_JACOBI, SLES_R_JACOBI, SLES_RJ2, SLES_RJ3 = 13, 14, 15, 16


def _np_from(ptr, n, dtype):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def native_arrays(s):
    """(da, xa) of a symmetric CubeSystem in native form: xa[f] = the value of the MSR entry
    (face_cells[f][0], face_cells[f][1]) (sign +1 side of the mesh association)."""
    xa = np.zeros(max(s.n_faces, 1))
    up = np.asarray(s.c2f_sgn) > 0
    xa[np.asarray(s.c2f)[up]] = np.asarray(s.x_val)[up]
    return np.array(s.d_val, dtype=np.float64), xa[:s.n_faces]


class Oracle:
    """The C restatement (oracle/cs_oracle.c)."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.ora_dot.restype = C.c_double
        L.ora_dot_xx.restype = C.c_double
        L.ora_grid_create_fine.restype = C.POINTER(OraGrid)
        L.ora_grid_coarsen.restype = C.POINTER(OraGrid)
        L.ora_mg_setup.restype = C.POINTER(OraMg)
        L.ora_grid_coarsen.argtypes = [C.POINTER(OraGrid), C.c_int, C.c_double, C.c_int]
        L.ora_mg_setup.argtypes = [C.POINTER(OraGrid), C.c_int, C.c_int, C.c_int,
                                   C.c_double, C.c_int]
        L.ora_mg_solve.argtypes = [C.POINTER(OraMg), C.POINTER(OraMgOpt), C.c_double,
                                   C.c_double, c_dbl_p, c_dbl_p, C.c_int, c_dbl_p, c_dbl_p]

    # -- kernels ----------------------------------------------------------
    def spmv(self, ri, ci, d, x_val, x, exclude_diag=False):
        ri, ci, d, x_val, x = _i32(ri), _i32(ci), _f64(d), _f64(x_val), _f64(x)
        y = np.empty(len(ri) - 1)
        self.lib.ora_spmv_msr(C.c_int(len(ri) - 1), _ip(ri), _ip(ci), _dp(d), _dp(x_val),
                              C.c_int(int(exclude_diag)), _dp(x), _dp(y))
        return y

    def spmv_b3(self, ri, ci, d, x_val, x, exclude_diag=False):
        ri, ci, d, x_val, x = _i32(ri), _i32(ci), _f64(d), _f64(x_val), _f64(x)
        n = len(ri) - 1
        y = np.empty(3 * n)
        self.lib.ora_spmv_msr_b3(C.c_int(n), _ip(ri), _ip(ci), _dp(d), _dp(x_val),
                                 C.c_int(int(exclude_diag)), _dp(x), _dp(y))
        return y

    def bicgstab_b3(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        ri, ci, d, xv, rhs = _i32(ri), _i32(ci), _f64(d), _f64(xv), _f64(rhs)
        n = len(ri) - 1
        vx = np.zeros(3 * n) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        cvg = self.lib.ora_bicgstab_b(C.c_int(n), C.c_int(3), _ip(ri), _ip(ci), _dp(d), _dp(xv),
                                      C.c_int(n_max_iter), C.c_double(precision),
                                      C.c_double(r_norm), _dp(rhs), _dp(vx),
                                      C.c_int(0 if vx0 is None else 1), _dp(out), _dp(hist),
                                      C.c_int(len(hist)))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)

    def block3_jacobi(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        ri, ci, d, xv, rhs = _i32(ri), _i32(ci), _f64(d), _f64(xv), _f64(rhs)
        n = len(ri) - 1
        vx = np.zeros(3 * n) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        cvg = self.lib.ora_block3_jacobi(C.c_int(n), _ip(ri), _ip(ci), _dp(d), _dp(xv),
                                         C.c_int(n_max_iter), C.c_double(precision),
                                         C.c_double(r_norm), _dp(rhs), _dp(vx),
                                         C.c_int(0 if vx0 is None else 1), _dp(out), _dp(hist),
                                         C.c_int(len(hist)))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)

    def spmv_bn(self, db, ri, ci, d, x_val, x, exclude_diag=False):
        """_b_mat_vec_p_l_msr_generic / _6 / _3 by block size."""
        ri, ci, d, x_val, x = _i32(ri), _i32(ci), _f64(d), _f64(x_val), _f64(x)
        n = len(ri) - 1
        y = np.empty(db * n)
        self.lib.ora_spmv_msr_bn(C.c_int(n), C.c_int(db), _ip(ri), _ip(ci), _dp(d), _dp(x_val),
                                 C.c_int(int(exclude_diag)), _dp(x), _dp(y))
        return y

    def _solve_b(self, fn, db, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0):
        ri, ci, d, xv, rhs = _i32(ri), _i32(ci), _f64(d), _f64(xv), _f64(rhs)
        n = len(ri) - 1
        vx = np.zeros(db * n) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        cvg = fn(C.c_int(n), C.c_int(db), _ip(ri), _ip(ci), _dp(d), _dp(xv),
                 C.c_int(n_max_iter), C.c_double(precision), C.c_double(r_norm), _dp(rhs),
                 _dp(vx), C.c_int(0 if vx0 is None else 1), _dp(out), _dp(hist),
                 C.c_int(len(hist)))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)

    def bicgstab_bn(self, db, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        return self._solve_b(self.lib.ora_bicgstab_b, db, ri, ci, d, xv, rhs, n_max_iter,
                             precision, r_norm, vx0)

    def gcr(self, db, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None, jacobi=True,
            restart=20):
        """_gcr (cs_sles_it.cpp:3272), the fallback solver on a device."""
        import functools

        def fn(n, dbc, *rest):
            return self.lib.ora_gcr(n, dbc, *rest[:4], C.c_int(int(jacobi)), C.c_int(restart),
                                    *rest[4:])
        return self._solve_b(fn, db, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0)

    def smoother_block(self, stype, db, ri, ci, d, xv, rhs, n_max_iter, vx0=None):
        """cs_multigrid_smoother on a block-diagonal matrix (block Jacobi / block relaxed Jacobi)."""
        ri, ci, d, xv, rhs = _i32(ri), _i32(ci), _f64(d), _f64(xv), _f64(rhs)
        n = len(ri) - 1
        vx = np.zeros(db * n) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(2)
        cvg = self.lib.ora_smoother_block(C.c_int(stype), C.c_int(n), C.c_int(db), _ip(ri), _ip(ci),
                                          _dp(d), _dp(xv), C.c_int(n_max_iter), _dp(rhs), _dp(vx),
                                          C.c_int(0 if vx0 is None else 1), _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), vx=vx)

    def block_jacobi(self, db, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        """_block_jacobi (cs_sles_it.cpp:2460) / _block_3_jacobi by block size."""
        return self._solve_b(self.lib.ora_block_jacobi, db, ri, ci, d, xv, rhs, n_max_iter,
                             precision, r_norm, vx0)

    def spmv_csr(self, ri, ci, val, x, exclude_diag=False):
        ri, ci, val, x = _i32(ri), _i32(ci), _f64(val), _f64(x)
        y = np.empty(len(ri) - 1)
        self.lib.ora_spmv_csr(C.c_int(len(ri) - 1), _ip(ri), _ip(ci), _dp(val),
                              C.c_int(int(exclude_diag)), _dp(x), _dp(y))
        return y

    def dot(self, x, y):
        x, y = _f64(x), _f64(y)
        return self.lib.ora_dot(C.c_int(len(x)), _dp(x), _dp(y))

    def dot_xx(self, x):
        x = _f64(x)
        return self.lib.ora_dot_xx(C.c_int(len(x)), _dp(x))

    def dot_xx_xy_yz(self, x, y, z):
        x, y, z = _f64(x), _f64(y), _f64(z)
        r = (C.c_double * 3)()
        self.lib.ora_dot_xx_xy_yz(C.c_int(len(x)), _dp(x), _dp(y), _dp(z),
                                  C.byref(r, 0), C.byref(r, 8), C.byref(r, 16))
        return tuple(r)

    def _solve(self, fn, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0):
        ri, ci, d, xv, rhs = _i32(ri), _i32(ci), _f64(d), _f64(xv), _f64(rhs)
        n = len(ri) - 1
        vx = np.zeros(n) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        cvg = fn(C.c_int(n), _ip(ri), _ip(ci), _dp(d), _dp(xv), C.c_int(n_max_iter),
                 C.c_double(precision), C.c_double(r_norm), _dp(rhs), _dp(vx),
                 C.c_int(0 if vx0 is None else 1), _dp(out), _dp(hist), C.c_int(len(hist)))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)

    def pcg(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        return self._solve(self.lib.ora_pcg_sr_p0, ri, ci, d, xv, rhs, n_max_iter,
                           precision, r_norm, vx0)

    def fcg(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        return self._solve(self.lib.ora_fcg, ri, ci, d, xv, rhs, n_max_iter,
                           precision, r_norm, vx0)

    def ipcg(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        return self._solve(self.lib.ora_ipcg, ri, ci, d, xv, rhs, n_max_iter,
                           precision, r_norm, vx0)

    def pcg_npc(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        return self._solve(self.lib.ora_pcg_sr_npc, ri, ci, d, xv, rhs, n_max_iter,
                           precision, r_norm, vx0)

    def jacobi(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        return self._solve(self.lib.ora_jacobi, ri, ci, d, xv, rhs, n_max_iter,
                           precision, r_norm, vx0)

    def gauss_seidel(self, sym, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        """_p_gauss_seidel_msr (sym False) / _p_sym_gauss_seidel_msr (True), with convergence test."""
        import functools
        return self._solve(functools.partial(self.lib.ora_gauss_seidel, C.c_int(1 if sym else 0)),
                           ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0)

    def bicgstab(self, ri, ci, d, xv, rhs, n_max_iter, precision, r_norm, vx0=None):
        return self._solve(self.lib.ora_bicgstab, ri, ci, d, xv, rhs, n_max_iter,
                           precision, r_norm, vx0)

    def aggregation(self, level, ri, ci, d, xv, max_aggregation=3, n_threads=1):
        ri, ci, d, xv = _i32(ri), _i32(ci), _f64(d), _f64(xv)
        n = len(ri) - 1
        f_c_row = np.empty(n, dtype=np.int32)
        nc = self.lib.ora_aggregation_dx_msr(C.c_int(level), C.c_int(n), _ip(ri), _ip(ci),
                                             _dp(d), _dp(xv), C.c_int(max_aggregation),
                                             C.c_int(n_threads), _ip(f_c_row))
        return nc, f_c_row

    # -- hierarchy ---------------------------------------------------------
    def fine_grid(self, s):
        ri, ci = _i32(s.row_index), _i32(s.col_id)
        args = [_f64(s.d_val), _f64(s.x_val)]
        c2f, sgn = _i32(s.c2f), np.ascontiguousarray(s.c2f_sgn, dtype=np.int16)
        cen, vol = _f64(s.cell_cen), _f64(s.cell_vol)
        fun, fs = _f64(s.face_u_normal), _f64(s.face_surf)
        return self.lib.ora_grid_create_fine(
            C.c_int(s.n_rows), _ip(ri), _ip(ci), _dp(args[0]), _dp(args[1]),
            C.c_int(s.n_faces), _ip(c2f), _sp(sgn), _dp(cen), _dp(vol), _dp(fun), _dp(fs))

    def mg_setup(self, s, aggregation_limit=3, n_levels_max=25, n_g_rows_min=30,
                 p0p1_relax=0.95, n_threads=1):
        g = self.fine_grid(s)
        return self.lib.ora_mg_setup(g, aggregation_limit, n_levels_max, n_g_rows_min,
                                     p0p1_relax, n_threads)

    def mg_setup_native(self, s, aggregation_limit=3, n_levels_max=25, n_g_rows_min=30,
                        p0p1_relax=0.95, n_threads=1, da=None, xa=None):
        """Hierarchy of the matrix mapped from native arrays (cs_oracle_native.c)."""
        if da is None or xa is None:
            da, xa = native_arrays(s)
        self.lib.ora_grid_create_fine_native.restype = C.POINTER(OraGrid)
        self.lib.ora_mg_setup_native.restype = C.POINTER(OraMg)
        g = self.lib.ora_grid_create_fine_native(
            C.c_int(s.n_rows), C.c_int(s.n_faces), _ip(_i32(s.face_cells)), _dp(_f64(da)),
            _dp(_f64(xa)), _dp(_f64(s.cell_cen)), _dp(_f64(s.cell_vol)),
            _dp(_f64(s.face_u_normal)), _dp(_f64(s.face_surf)))
        return self.lib.ora_mg_setup_native(g, C.c_int(aggregation_limit), C.c_int(n_levels_max),
                                            C.c_int(n_g_rows_min), C.c_double(p0p1_relax),
                                            C.c_int(n_threads))

    def assembler(self, n_rows, rows, cols, batches, l0=0):
        """Structure from the declared (rows, cols) pairs, then the values of `batches`
        (list of (rows, cols, vals)) added in order; returns (row_index, col_id, d_val, x_val)."""
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_longlong))
        rows, cols = i64(rows), i64(cols)
        ri = np.zeros(n_rows + 1, dtype=np.int32)
        ci = np.zeros(max(len(rows), 1), dtype=np.int32)
        nnz = self.lib.ora_assembler_structure(C.c_int(n_rows), C.c_int(len(rows)), lp(rows),
                                               lp(cols), C.c_longlong(l0), _ip(ri), _ip(ci))
        ci = ci[:nnz].copy()
        d, x = np.zeros(n_rows), np.zeros(max(nnz, 1))
        for br, bc, bv in batches:
            br, bc, bv = i64(br), i64(bc), _f64(bv)
            miss = self.lib.ora_assembler_values_add(C.c_int(n_rows), _ip(ri), _ip(ci),
                                                     C.c_int(len(br)), lp(br), lp(bc),
                                                     C.c_longlong(l0), _dp(bv), _dp(d), _dp(x))
            assert miss == 0
        return ri, ci, d, x[:nnz]

    def diffusion_potential(self, s, b_face_cells, pvar, cofaf, cofbf, i_visc, b_visc, init=1,
                            inc=1, diverg=None):
        div = np.zeros(s.n_rows) if diverg is None else _f64(diverg).copy()
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        self.lib.ora_diffusion_potential.restype = None
        self.lib.ora_diffusion_potential(
            C.c_int(s.n_rows), C.c_int(len(fc)), _ip(fc), C.c_int(len(bfc)), _ip(bfc),
            C.c_int(init), C.c_int(inc), _dp(_f64(pvar)), _dp(_f64(cofaf)), _dp(_f64(cofbf)),
            _dp(_f64(i_visc)), _dp(_f64(b_visc)), _dp(div))
        return div

    def b_iprime_lsq(self, s, b_face_cells, b_un, b_dist, diipb, pvar, coefa, coefb,
                     clip_coeff=-1.):
        """value of pvar at I' of every boundary face (cs_gradient_boundary_iprime_lsq_s)"""
        bfc = _i32(b_face_cells)
        val_ip = np.zeros(max(len(bfc), 1))
        self.lib.ora_b_iprime_lsq.restype = None
        self.lib.ora_b_iprime_lsq(
            C.c_int(s.n_rows), _ip(_i32(s.row_index)), _ip(_i32(s.col_id)), C.c_int(len(bfc)),
            _ip(bfc), _dp(_f64(s.cell_cen)), _dp(_f64(b_un)), _dp(_f64(b_dist)),
            _dp(_f64(diipb)), C.c_double(clip_coeff), _dp(_f64(pvar)), _dp(_f64(coefa)),
            _dp(_f64(coefb)), _dp(val_ip))
        return val_ip[:len(bfc)]

    def gradient_clip(self, s, b_face_cells, i_geo, imligr, climgr, pvar, val_f, grad):
        """_scalar_gradient_clipping applied to a copy of grad; val_f: boundary face values"""
        bfc = _i32(b_face_cells)
        g = _f64(grad).reshape(-1).copy()
        self.lib.ora_gradient_clip.restype = None
        self.lib.ora_gradient_clip(
            C.c_int(s.n_rows), _ip(_i32(s.row_index)), _ip(_i32(s.col_id)), _ip(_i32(s.c2f)),
            np.ascontiguousarray(s.c2f_sgn, dtype=np.int16).ctypes.data_as(C.c_void_p),
            C.c_int(len(bfc)), _ip(bfc), _dp(_f64(s.cell_cen)), _dp(_f64(i_geo)),
            C.c_int(imligr), C.c_double(climgr), _dp(_f64(pvar)), _dp(_f64(val_f)), _dp(g))
        return g.reshape(-1, 3)

    def gradient_lsq(self, s, b_face_cells, b_un, b_dist, pvar, coefa, coefb, inc=1, diipb=None,
                     clip_coeff=-1.):
        """diipb (II' of the boundary faces) given: face values from the values at I'"""
        grad = np.zeros(3 * s.n_rows)
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        val_ip = dip = None
        if diipb is not None:
            dip = _f64(diipb)
            val_ip = self.b_iprime_lsq(s, bfc, b_un, b_dist, dip, pvar, coefa, coefb,
                                       clip_coeff=clip_coeff)
        self.lib.ora_gradient_lsq_b.restype = None
        self.lib.ora_gradient_lsq_b(
            C.c_int(s.n_rows), _ip(_i32(s.row_index)), _ip(_i32(s.col_id)), C.c_int(len(fc)),
            _ip(fc), C.c_int(len(bfc)), _ip(bfc), _dp(_f64(s.cell_cen)), _dp(_f64(b_un)),
            _dp(_f64(b_dist)), (None if dip is None else _dp(dip)),
            (None if val_ip is None else _dp(val_ip)), C.c_int(inc), _dp(_f64(pvar)),
            _dp(_f64(coefa)), _dp(_f64(coefb)), _dp(grad))
        return grad.reshape(-1, 3)

    def gradient_green_lsq(self, s, b_face_cells, b_un, b_dist, g_geo, b_face_surf, pvar, coefa,
                           coefb, inc=1, diipb=None, corr_grad_lin=None, clip_coeff=-1.):
        """imrgra = 4: the least-squares gradient, then the Green-Gauss reconstruction pass;
        g_geo: weight [1] and OF vector [3] per interior face"""
        r_grad = self.gradient_lsq(s, b_face_cells, b_un, b_dist, pvar, coefa, coefb, inc=inc,
                                   diipb=diipb, clip_coeff=clip_coeff)
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        vip = None
        if diipb is not None:
            vip = self.b_iprime_lsq(s, bfc, b_un, b_dist, diipb, pvar, coefa, coefb,
                                    clip_coeff=clip_coeff)
        gg = _f64(g_geo).reshape(-1, 4)
        grad = np.zeros(3 * s.n_rows)
        self.lib.ora_gradient_green_reconstruct.restype = None
        self.lib.ora_gradient_green_reconstruct(
            C.c_int(s.n_rows), C.c_int(len(fc)), _ip(fc), C.c_int(len(bfc)), _ip(bfc),
            _dp(_f64(gg[:, 0])), _dp(_f64(gg[:, 1:4])), _dp(_f64(s.face_surf)),
            _dp(_f64(s.face_u_normal)), _dp(_f64(b_face_surf)), _dp(_f64(b_un)),
            _dp(_f64(s.cell_vol)), (None if vip is None else _dp(_f64(vip))), C.c_int(inc),
            _dp(_f64(pvar)), _dp(_f64(coefa)), _dp(_f64(coefb)), _dp(_f64(r_grad)),
            (None if corr_grad_lin is None else _dp(_f64(corr_grad_lin))), _dp(grad))
        return grad.reshape(-1, 3)

    def diffusion_potential_rc(self, s, b_face_cells, i_geo, grad, pvar, cofaf, cofbf, i_visc,
                               b_visc, visel, irecmf=0, init=1, inc=1, diverg=None, val_ip=None):
        """val_ip: values at I' of the boundary faces (b_iprime_lsq) when II' is not zero"""
        div = np.zeros(s.n_rows) if diverg is None else _f64(diverg).copy()
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        vip = None if val_ip is None else _f64(val_ip)
        self.lib.ora_diffusion_potential_rc_b.restype = None
        self.lib.ora_diffusion_potential_rc_b(
            C.c_int(s.n_rows), C.c_int(len(fc)), _ip(fc), C.c_int(len(bfc)), _ip(bfc),
            _dp(_f64(i_geo)), _dp(_f64(s.face_surf)), (None if vip is None else _dp(vip)),
            C.c_int(irecmf), C.c_int(init), C.c_int(inc),
            _dp(_f64(pvar)), _dp(_f64(grad)), _dp(_f64(cofaf)), _dp(_f64(cofbf)),
            _dp(_f64(i_visc)), _dp(_f64(b_visc)), _dp(_f64(visel)), _dp(div))
        return div

    def gradient_green_iter(self, s, b_face_cells, b_un, g_geo, b_face_surf, pvar, coefa, coefb,
                            inc=1, diipb=None, nswrgr=100, epsrgr=1e-4):
        """imrgra = 0: iterative Green-Gauss gradient; returns (grad, n_sweeps)"""
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        gg = _f64(g_geo).reshape(-1, 4)
        grad = np.zeros(3 * s.n_rows)
        self.lib.ora_gradient_green_iter.restype = C.c_int
        n = self.lib.ora_gradient_green_iter(
            C.c_int(s.n_rows), C.c_int(len(fc)), _ip(fc), C.c_int(len(bfc)), _ip(bfc),
            _dp(_f64(gg[:, 0])), _dp(_f64(gg[:, 1:4])), _dp(_f64(s.face_surf)),
            _dp(_f64(s.face_u_normal)), _dp(_f64(b_face_surf)), _dp(_f64(b_un)),
            _dp(_f64(s.cell_vol)), (None if diipb is None else _dp(_f64(diipb))),
            C.c_int(nswrgr), C.c_double(epsrgr), C.c_int(inc), _dp(_f64(pvar)),
            _dp(_f64(coefa)), _dp(_f64(coefb)), _dp(grad))
        return grad.reshape(-1, 3), int(n)

    def gradient_hyd(self, imrgra, s, b_face_cells, b_un, b_dist, i_face_cog, b_face_cog, f_ext,
                     pvar, coefa, coefb, inc=1, diipb=None, g_geo=None, b_face_surf=None,
                     corr_grad_lin=None, clip_coeff=-1., with_val_f=False):
        """hyd_p_flag = 1 (exterior force f_ext): returns (grad, val_ip[, val_f]); imrgra 1 or 4"""
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        nb = len(bfc)
        dip = np.zeros(3 * max(nb, 1)) if diipb is None else _f64(diipb)
        vip, vf = np.zeros(max(nb, 1)), np.zeros(max(nb, 1))
        ri, ci = _i32(s.row_index), _i32(s.col_id)
        cen, fx = _f64(s.cell_cen), _f64(f_ext)
        icog, bcog, bun, bd = _f64(i_face_cog), _f64(b_face_cog), _f64(b_un), _f64(b_dist)
        p = _f64(pvar)
        self.lib.ora_b_face_values_hyd.restype = None
        self.lib.ora_b_face_values_hyd(
            C.c_int(s.n_rows), _ip(ri), _ip(ci), _ip(_i32(s.c2f)), C.c_int(nb), _ip(bfc), _dp(cen),
            _dp(bun), _dp(bd), _dp(dip), _dp(icog), _dp(bcog), _dp(fx), C.c_double(clip_coeff),
            C.c_int(inc), _dp(p),
            _dp(_f64(coefa)), _dp(_f64(coefb)), _dp(vip), _dp(vf))
        r_grad = np.zeros(3 * s.n_rows)
        self.lib.ora_gradient_lsq_hyd.restype = None
        self.lib.ora_gradient_lsq_hyd(
            C.c_int(s.n_rows), _ip(ri), _ip(ci), C.c_int(len(fc)), _ip(fc), C.c_int(nb), _ip(bfc),
            _dp(cen), _dp(bun), _dp(bd), (None if diipb is None else _dp(dip)), _dp(icog),
            _dp(bcog), _dp(fx), _dp(vf), _dp(p), _dp(r_grad))
        if imrgra != 4:
            if with_val_f:
                return r_grad.reshape(-1, 3), vip[:nb], vf[:nb]
            return r_grad.reshape(-1, 3), vip[:nb]
        gg = _f64(g_geo).reshape(-1, 4)
        grad = np.zeros(3 * s.n_rows)
        self.lib.ora_gradient_green_reconstruct_hyd.restype = None
        self.lib.ora_gradient_green_reconstruct_hyd(
            C.c_int(s.n_rows), C.c_int(len(fc)), _ip(fc), C.c_int(nb), _ip(bfc),
            _dp(_f64(gg[:, 0])), _dp(_f64(gg[:, 1:4])), _dp(_f64(s.face_surf)),
            _dp(_f64(s.face_u_normal)), _dp(_f64(b_face_surf)), _dp(bun), _dp(_f64(s.cell_vol)),
            _dp(cen), _dp(icog), _dp(fx), _dp(vf), _dp(p), _dp(r_grad),
            (None if corr_grad_lin is None else _dp(_f64(corr_grad_lin))), _dp(grad))
        if with_val_f:
            return grad.reshape(-1, 3), vip[:nb], vf[:nb]
        return grad.reshape(-1, 3), vip[:nb]

    def divergence(self, s, b_face_cells, i_massflux, b_massflux, init=1, diverg=None):
        div = np.zeros(s.n_rows) if diverg is None else _f64(diverg).copy()
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        self.lib.ora_divergence.restype = None
        self.lib.ora_divergence(C.c_int(s.n_rows), C.c_int(len(fc)), _ip(fc), C.c_int(len(bfc)),
                                _ip(bfc), C.c_int(init), _dp(_f64(i_massflux)),
                                _dp(_f64(b_massflux)), _dp(div))
        return div

    def face_diffusion_potential(self, s, b_face_cells, pvar, cofaf, cofbf, i_visc, b_visc,
                                 i_geo=None, grad=None, visel=None, val_ip=None, init=1, inc=1,
                                 massflux=None):
        """grad None: eqp->ircflu = 0; returns (i_massflux, b_massflux)"""
        fc, bfc = _i32(s.face_cells), _i32(b_face_cells)
        im = np.zeros(max(len(fc), 1)) if massflux is None else _f64(massflux[0]).copy()
        bm = np.zeros(max(len(bfc), 1)) if massflux is None else _f64(massflux[1]).copy()
        opt = lambda a: None if a is None else _dp(_f64(a))
        self.lib.ora_face_diffusion_potential.restype = None
        self.lib.ora_face_diffusion_potential(
            C.c_int(s.n_rows), C.c_int(len(fc)), _ip(fc), C.c_int(len(bfc)), _ip(bfc), opt(i_geo),
            _dp(_f64(s.face_surf)), opt(val_ip), C.c_int(init), C.c_int(inc), _dp(_f64(pvar)),
            opt(grad), _dp(_f64(cofaf)), _dp(_f64(cofbf)), _dp(_f64(i_visc)), _dp(_f64(b_visc)),
            opt(visel), _dp(im), _dp(bm))
        return im[:len(fc)], bm[:len(bfc)]

    def aggregation_fc(self, level, n_rows, face_cells, da, xa, max_aggregation=3):
        fc = _i32(face_cells)
        f_c = np.empty(max(n_rows, 1), dtype=np.int32)
        nc = self.lib.ora_aggregation_fc(C.c_int(level), C.c_int(n_rows), C.c_int(n_rows),
                                         C.c_int(len(fc)), _ip(fc), _dp(_f64(da)), _dp(_f64(xa)),
                                         C.c_int(max_aggregation), _ip(f_c))
        return nc, f_c[:n_rows]

    def mg_setup_pw(self, s, aggregation_limit=4, n_levels_max=25, n_g_rows_min=30):
        """CS_GRID_COARSENING_SPD_PW (pairwise aggregation) with Galerkin coarse matrices."""
        g = self.fine_grid(s)
        self.lib.ora_mg_setup_pw.restype = C.POINTER(OraMg)
        self.lib.ora_mg_setup_pw.argtypes = [C.POINTER(OraGrid), C.c_int, C.c_int, C.c_int]
        return self.lib.ora_mg_setup_pw(g, aggregation_limit, n_levels_max, n_g_rows_min)

    def matrix_compute_coeffs_sym(self, ri, c2f, b_face_cells, cofbfp, rovsdt, i_visc, b_visc,
                                  idiffp=1, ndircp=1, thetap=1., relaxp=1.):
        ri, c2f, bfc = _i32(ri), _i32(c2f), _i32(b_face_cells)
        n = len(ri) - 1
        da, ea = np.zeros(max(n, 1)), np.zeros(max(int(ri[-1]), 1))
        self.lib.ora_matrix_compute_coeffs_sym(
            C.c_int(n), _ip(ri), _ip(c2f), C.c_int(len(bfc)), _ip(bfc), C.c_int(idiffp),
            C.c_int(ndircp), C.c_double(thetap), C.c_double(relaxp), _dp(_f64(cofbfp)),
            _dp(_f64(rovsdt)), _dp(_f64(i_visc)), _dp(_f64(b_visc)), _dp(da), _dp(ea))
        return da[:n], ea[:int(ri[-1])]

    def matrix_compute_coeffs(self, s, b_face_cells, coefbp, cofbfp, rovsdt, i_massflux,
                              b_massflux, i_visc, b_visc, xcpp=None, iconvp=1, idiffp=1,
                              ndircp=1, thetap=1., relaxp=1.):
        ri, ci, c2f, bfc = _i32(s.row_index), _i32(s.col_id), _i32(s.c2f), _i32(b_face_cells)
        sgn = np.ascontiguousarray(s.c2f_sgn, dtype=np.int16)
        n = len(ri) - 1
        da, ea = np.zeros(max(n, 1)), np.zeros(max(int(ri[-1]), 1))
        f = lambda a: None if a is None else _f64(a)
        arrs = [f(a) for a in (coefbp, cofbfp, rovsdt, i_massflux, b_massflux, i_visc, b_visc, xcpp)]
        self.lib.ora_matrix_compute_coeffs(
            C.c_int(n), _ip(ri), _ip(ci), _ip(c2f), _sp(sgn), C.c_int(len(bfc)), _ip(bfc),
            C.c_int(iconvp), C.c_int(idiffp), C.c_int(ndircp), C.c_double(thetap),
            C.c_double(relaxp), *[_dp(a) for a in arrs], _dp(da), _dp(ea))
        return da[:n], ea[:int(ri[-1])]

    def matrix_compute_coeffs_v3(self, s, b_face_cells, coefbp, cofbfp, fimp, i_massflux,
                                 b_massflux, i_visc, b_visc, iconvp=1, idiffp=1, ndircp=1,
                                 thetap=1., relaxp=1.):
        ri, c2f, bfc = _i32(s.row_index), _i32(s.c2f), _i32(b_face_cells)
        sgn = np.ascontiguousarray(s.c2f_sgn, dtype=np.int16)
        n = len(ri) - 1
        da, ea = np.zeros(9 * max(n, 1)), np.zeros(max(int(ri[-1]), 1))
        f = lambda a: None if a is None else _f64(a)
        arrs = [f(a) for a in (coefbp, cofbfp, fimp, i_massflux, b_massflux, i_visc, b_visc)]
        self.lib.ora_matrix_compute_coeffs_v3(
            C.c_int(n), _ip(ri), _ip(c2f), _sp(sgn), C.c_int(len(bfc)), _ip(bfc),
            C.c_int(iconvp), C.c_int(idiffp), C.c_int(ndircp), C.c_double(thetap),
            C.c_double(relaxp), *[_dp(a) for a in arrs], _dp(da), _dp(ea))
        return da[:9 * n], ea[:int(ri[-1])]

    def pairwise(self, level, ri, ci, d, xv):
        ri, ci, d, xv = _i32(ri), _i32(ci), _f64(d), _f64(xv)
        n = len(ri) - 1
        f_c_row = np.empty(max(n, 1), dtype=np.int32)
        nc = self.lib.ora_pairwise_msr(C.c_int(level), C.c_int(n), _ip(ri), _ip(ci), _dp(d),
                                       _dp(xv), _ip(f_c_row))
        return nc, f_c_row[:n]

    @staticmethod
    def mg_levels(mg):
        """Copy every level out as numpy arrays."""
        out = []
        for l in range(mg.contents.n_levels):
            g = mg.contents.grids[l].contents
            parent_n = mg.contents.grids[l - 1].contents.n_rows if l > 0 else 0
            out.append(dict(
                n_rows=g.n_rows, nnz=g.nnz, n_faces=g.n_faces,
                row_index=_np_from(g.row_index, g.n_rows + 1, np.int32),
                col_id=_np_from(g.col_id, g.nnz, np.int32),
                d_val=_np_from(g.d_val, g.n_rows, np.float64),
                x_val=_np_from(g.x_val, g.nnz, np.float64),
                coarse_row=_np_from(g.coarse_row, parent_n, np.int32),
                cell_cen=_np_from(g.cell_cen, 3 * g.n_rows, np.float64),
                cell_vol=_np_from(g.cell_vol, g.n_rows, np.float64)))
        return out

    def mg_solve(self, mg, rhs, precision, r_norm, n_max_cycles=100, n_max_iter=(2, 10, 500),
                 precision_mult=(1., 1., 1.), vx0=None):
        rhs = _f64(rhs)
        n = len(rhs)
        vx = np.zeros(n) if vx0 is None else _f64(vx0).copy()
        opt = OraMgOpt()
        opt.n_max_cycles = n_max_cycles
        for i in range(3):
            opt.n_max_iter[i] = n_max_iter[i]
            opt.precision_mult[i] = precision_mult[i]
        out = np.zeros(4)
        cyc = np.zeros(n_max_cycles + 2)
        cvg = self.lib.ora_mg_solve(mg, C.byref(opt), precision, r_norm, _dp(rhs), _dp(vx),
                                    0 if vx0 is None else 1, _dp(out), _dp(cyc))
        nc = int(out[2])
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], n_cycles=nc,
                    cycle_residuals=cyc[:nc].copy(), vx=vx)

    def mg_free(self, mg):
        self.lib.ora_mg_free(mg, 1)

    # -- smoothers, single-cycle (preconditioner) V-cycle, FCG + multigrid pc --------
    def smoother(self, stype, ri, ci, d, xv, rhs, n_max_iter, vx0=None):
        ri, ci, d, xv, rhs = _i32(ri), _i32(ci), _f64(d), _f64(xv), _f64(rhs)
        n = len(ri) - 1
        vx = np.zeros(n) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(4)
        cvg = self.lib.ora_smoother(C.c_int(stype), C.c_int(n), _ip(ri), _ip(ci), _dp(d),
                                    _dp(xv), C.c_int(n_max_iter), _dp(rhs), _dp(vx),
                                    C.c_int(0 if vx0 is None else 1), _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), vx=vx)

    @staticmethod
    def _pc_opt(types, n_max_iter, precision_mult):
        opt = OraMgPcOpt()
        for i in range(3):
            opt.type[i] = types[i]
            opt.n_max_iter[i] = n_max_iter[i]
            opt.precision_mult[i] = precision_mult[i]
        return opt

    def mg_cycle_pc(self, mg, rhs, precision, r_norm, types=(SLES_RJ3, SLES_RJ3, SLES_PCG),
                    n_max_iter=(3, 3, 500), precision_mult=(-1., -1., 1.), vx0=None):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        opt = self._pc_opt(types, n_max_iter, precision_mult)
        out = np.zeros(4)
        self.lib.ora_mg_cycle_pc.argtypes = [C.POINTER(OraMg), C.POINTER(OraMgPcOpt),
                                             C.c_double, C.c_double, c_dbl_p, c_dbl_p,
                                             C.c_int, c_dbl_p]
        cvg = self.lib.ora_mg_cycle_pc(mg, C.byref(opt), precision, r_norm, _dp(rhs),
                                       _dp(vx), 0 if vx0 is None else 1, _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), vx=vx)

    def it_poly(self, solver, ri, ci, d, xv, rhs, poly_degree, n_max_iter, precision, r_norm,
                vx0=None):
        """FCG (solver 1) / single-reduction PCG (0) with the polynomial preconditioner
        (poly_degree >= 1) or none (< 0, FCG)."""
        ri, ci, d, xv, rhs = _i32(ri), _i32(ci), _f64(d), _f64(xv), _f64(rhs)
        n = len(ri) - 1
        vx = np.zeros(n) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        cvg = self.lib.ora_it_poly(C.c_int(solver), C.c_int(n), _ip(ri), _ip(ci), _dp(d), _dp(xv),
                                   C.c_int(poly_degree), C.c_int(n_max_iter),
                                   C.c_double(precision), C.c_double(r_norm), _dp(rhs), _dp(vx),
                                   C.c_int(0 if vx0 is None else 1), _dp(out), _dp(hist),
                                   C.c_int(len(hist)))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)

    def it_mgpc(self, solver, mg, rhs, n_max_iter, precision, r_norm,
                types=(SLES_RJ3, SLES_RJ3, SLES_PCG), mg_n_max_iter=(3, 3, 500),
                precision_mult=(-1., -1., 1.), vx0=None):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        opt = self._pc_opt(types, mg_n_max_iter, precision_mult)
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        self.lib.ora_it_mgpc.argtypes = [C.c_int, C.POINTER(OraMg), C.POINTER(OraMgPcOpt),
                                         C.c_int, C.c_double, C.c_double, c_dbl_p, c_dbl_p,
                                         C.c_int, c_dbl_p, c_dbl_p, C.c_int]
        cvg = self.lib.ora_it_mgpc(solver, mg, C.byref(opt), n_max_iter, precision, r_norm,
                                   _dp(rhs), _dp(vx), 0 if vx0 is None else 1, _dp(out),
                                   _dp(hist), len(hist))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)

    def mg_solve_sm(self, mg, rhs, precision, r_norm, types=(SLES_RJ3, SLES_RJ3, SLES_PCG),
                    n_max_iter=(3, 3, 500), precision_mult=(-1., -1., 1.), n_max_cycles=100,
                    vx0=None):
        """V-cycle in solver mode with cs_multigrid_smoother smoothers."""
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        opt = self._pc_opt(types, n_max_iter, precision_mult)
        out = np.zeros(4)
        cyc = np.zeros(n_max_cycles + 2)
        self.lib.ora_mg_solve_sm.argtypes = [C.POINTER(OraMg), C.POINTER(OraMgPcOpt), C.c_int,
                                             C.c_double, C.c_double, c_dbl_p, c_dbl_p, C.c_int,
                                             c_dbl_p, c_dbl_p]
        cvg = self.lib.ora_mg_solve_sm(mg, C.byref(opt), n_max_cycles, precision, r_norm,
                                       _dp(rhs), _dp(vx), 0 if vx0 is None else 1, _dp(out),
                                       _dp(cyc))
        nc = int(out[2])
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], n_cycles=nc,
                    cycle_residuals=cyc[:nc].copy(), vx=vx)

    def mg_k_cycle(self, mg, rhs, precision, r_norm, types=(SLES_JACOBI, SLES_JACOBI, SLES_PCG),
                   n_max_iter=(2, 2, 500), precision_mult=(-1., -1., 10.), trsh=0.25):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs))
        opt = self._pc_opt(types, n_max_iter, precision_mult)
        out = np.zeros(4)
        self.lib.ora_mg_k_cycle.argtypes = [C.POINTER(OraMg), C.POINTER(OraMgPcOpt), C.c_double,
                                            C.c_double, C.c_double, c_dbl_p, c_dbl_p, c_dbl_p]
        cvg = self.lib.ora_mg_k_cycle(mg, C.byref(opt), trsh, precision, r_norm, _dp(rhs),
                                      _dp(vx), _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), vx=vx)

    def it_kpc(self, solver, mg, rhs, n_max_iter, precision, r_norm,
               types=(SLES_JACOBI, SLES_JACOBI, SLES_PCG), mg_n_max_iter=(2, 2, 500),
               precision_mult=(-1., -1., 10.), trsh=0.25, vx0=None):
        """FCG (1) / PCG (0) preconditioned by one K-cycle."""
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        opt = self._pc_opt(types, mg_n_max_iter, precision_mult)
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        self.lib.ora_it_kpc.argtypes = [C.c_int, C.POINTER(OraMg), C.POINTER(OraMgPcOpt),
                                        C.c_double, C.c_int, C.c_double, C.c_double, c_dbl_p,
                                        c_dbl_p, C.c_int, c_dbl_p, c_dbl_p, C.c_int]
        cvg = self.lib.ora_it_kpc(solver, mg, C.byref(opt), trsh, n_max_iter, precision, r_norm,
                                  _dp(rhs), _dp(vx), 0 if vx0 is None else 1, _dp(out),
                                  _dp(hist), len(hist))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)

    def fcg_mgpc(self, mg, rhs, n_max_iter, precision, r_norm,
                 types=(SLES_RJ3, SLES_RJ3, SLES_PCG), mg_n_max_iter=(3, 3, 500),
                 precision_mult=(-1., -1., 1.), vx0=None):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        opt = self._pc_opt(types, mg_n_max_iter, precision_mult)
        out = np.zeros(4)
        hist = np.zeros(n_max_iter + 8)
        self.lib.ora_fcg_mgpc.argtypes = [C.POINTER(OraMg), C.POINTER(OraMgPcOpt), C.c_int,
                                          C.c_double, C.c_double, c_dbl_p, c_dbl_p, C.c_int,
                                          c_dbl_p, c_dbl_p, C.c_int]
        cvg = self.lib.ora_fcg_mgpc(mg, C.byref(opt), n_max_iter, precision, r_norm,
                                    _dp(rhs), _dp(vx), 0 if vx0 is None else 1, _dp(out),
                                    _dp(hist), len(hist))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    history=hist[:int(out[3])].copy(), vx=vx)


class RefLib:
    """The reference's own CPU code (oracle/_ref/libcsref.so)."""

    PCG, FCG, IPCG, JACOBI, BICGSTAB = 0, 1, 2, 3, 4

    LIBNAME = "libcsref.so"

    def __init__(self, n_threads=1, quiet=True):
        path = os.path.join(_HERE, "_ref", self.LIBNAME)
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -f oracle/Makefile.ref -j8` where "
                "/root/reference is mounted")
        self.lib = C.CDLL(path)
        L = self.lib
        L.csref_system_create.restype = C.c_void_p
        L.csref_hierarchy_build.restype = C.c_void_p
        L.csref_dot.restype = C.c_double
        L.csref_dot_xx.restype = C.c_double
        self.n_threads = L.csref_init(C.c_int(n_threads), C.c_int(int(quiet)))

    def set_threads(self, n):
        self.n_threads = self.lib.csref_init(C.c_int(n), C.c_int(1))

    def system(self, s, d_val=None, x_val=None, symmetric=True, with_mesh=True):
        ri, ci = _i32(s.row_index), _i32(s.col_id)
        d = _f64(s.d_val if d_val is None else d_val)
        xv = _f64(s.x_val if x_val is None else x_val)
        h = C.c_void_p(self.lib.csref_system_create(C.c_int(s.n_rows), _ip(ri), _ip(ci),
                                                    _dp(d), _dp(xv), C.c_int(int(symmetric))))
        if with_mesh and s.n_faces > 0:
            fc, c2f = _i32(s.face_cells), _i32(s.c2f)
            sgn = np.ascontiguousarray(s.c2f_sgn, dtype=np.int16)
            cen, vol = _f64(s.cell_cen), _f64(s.cell_vol)
            fun, fs = _f64(s.face_u_normal), _f64(s.face_surf)
            self.lib.csref_system_set_mesh(h, C.c_int(s.n_faces), _ip(fc), _ip(c2f), _sp(sgn),
                                           _dp(cen), _dp(vol), _dp(fun), _dp(fs))
        return h

    def system_native(self, s, da=None, xa=None):
        """MSR matrix mapped from native (face-based) arrays: cs_matrix_set_coefficients with
        edges = s.face_cells, da, xa[n_faces] (symmetric); geometry attached."""
        fc = _i32(s.face_cells)
        if da is None or xa is None:
            da, xa = native_arrays(s)
        self.lib.csref_system_create_native.restype = C.c_void_p
        return C.c_void_p(self.lib.csref_system_create_native(
            C.c_int(s.n_rows), C.c_int(s.n_faces), _ip(fc), _dp(_f64(da)), _dp(_f64(xa)),
            _dp(_f64(s.cell_cen)), _dp(_f64(s.cell_vol)), _dp(_f64(s.face_u_normal)),
            _dp(_f64(s.face_surf))))

    def assembler(self, n_rows, rows, cols, batches, l0=0):
        """The reference's cs_matrix_assembler on one rank: structure from the declared pairs,
        MSR matrix, values of `batches` added with cs_matrix_assembler_values_add_g in order.
        Returns (handle, row_index, col_id, d_val, x_val); the handle is a system usable with
        spmv / sles_it / destroy."""
        u64 = lambda a: np.ascontiguousarray(a, dtype=np.uint64)
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_ulonglong))
        rows, cols = u64(rows), u64(cols)
        self.lib.csref_assembler_system_create.restype = C.c_void_p
        h = C.c_void_p(self.lib.csref_assembler_system_create(
            C.c_ulonglong(l0), C.c_int(n_rows), C.c_int(len(rows)), up(rows), up(cols)))
        for br, bc, bv in batches:
            br, bc, bv = u64(br), u64(bc), _f64(bv)
            self.lib.csref_assembler_values_add(h, C.c_int(len(br)), up(br), up(bc), _dp(bv))
        nnz = self.lib.csref_assembler_values_finalize(h)
        ri = np.zeros(n_rows + 1, dtype=np.int32)
        ci = np.zeros(max(nnz, 1), dtype=np.int32)
        d, x = np.zeros(n_rows), np.zeros(max(nnz, 1))
        self.lib.csref_get_msr_arrays(h, _ip(ri), _ip(ci), _dp(d), _dp(x))
        return h, ri, ci[:nnz], d, x[:nnz]

    def diffusion_potential(self, h, s, b_face_cells, b_geo, i_face_cog, pvar, cofaf, cofbf,
                            i_visc, b_visc, init=1, inc=1, diverg=None):
        """cs_diffusion_potential without flux reconstruction on the system's mesh."""
        div = np.zeros(s.n_rows) if diverg is None else _f64(diverg).copy()
        bfc = _i32(b_face_cells)
        r = self.lib.csref_diffusion_potential(
            h, _ip(_i32(s.face_cells)), C.c_int(len(bfc)), _ip(bfc), _dp(_f64(b_geo)),
            _dp(_f64(i_face_cog)), C.c_int(init), C.c_int(inc), _dp(_f64(pvar)), _dp(_f64(cofaf)),
            _dp(_f64(cofbf)), _dp(_f64(i_visc)), _dp(_f64(b_visc)), _dp(div))
        assert r == 0
        return div

    def diffusion_potential_rc(self, h, s, b_face_cells, b_geo, i_face_cog, i_geo, pvar, coefa,
                               coefb, cofaf, cofbf, i_visc, b_visc, visel, irecmf=0, init=1,
                               inc=1, diverg=None, diipb=None, imrgra=1, g_geo=None,
                               b_face_surf=None, corr_grad_lin=None, ircflu=1, massflux=None,
                               frcxt=None, imligr=-1, climgr=-1.):
        """cs_diffusion_potential WITH flux reconstruction (LSQ gradient) on the geometry given;
        returns (diverg, grad)."""
        div = np.zeros(s.n_rows) if diverg is None else _f64(diverg).copy()
        grad = np.zeros(3 * s.n_rows)
        bfc = _i32(b_face_cells)
        keep = [_f64(x) for x in (b_geo, i_face_cog, i_geo, pvar, coefa, coefb, cofaf, cofbf,
                                  i_visc, b_visc, visel)]
        r = self.lib.csref_diffusion_potential_rc(
            h, _ip(_i32(s.face_cells)), C.c_int(len(bfc)), _ip(bfc), _dp(keep[0]), _dp(keep[1]),
            _dp(keep[2]), (None if diipb is None else _dp(_f64(diipb))), C.c_int(imrgra),
            (None if g_geo is None else _dp(_f64(g_geo))),
            (None if b_face_surf is None else _dp(_f64(b_face_surf))),
            (None if corr_grad_lin is None else _dp(_f64(corr_grad_lin))), C.c_int(irecmf),
            C.c_int(init), C.c_int(inc), _dp(keep[3]), _dp(keep[4]),
            _dp(keep[5]), _dp(keep[6]), _dp(keep[7]), _dp(keep[8]), _dp(keep[9]), _dp(keep[10]),
            _dp(div), _dp(grad), C.c_int(ircflu),
            (None if massflux is None else _dp(massflux[0])),
            (None if massflux is None else _dp(massflux[1])),
            (None if frcxt is None else _dp(_f64(frcxt))), C.c_int(imligr), C.c_double(climgr))
        assert r == 0
        return div, grad.reshape(-1, 3)

    def divergence(self, h, s, b_face_cells, i_massflux, b_massflux, init=1, diverg=None):
        """cs_divergence on the system's faces"""
        div = np.zeros(s.n_rows) if diverg is None else _f64(diverg).copy()
        bfc = _i32(b_face_cells)
        r = self.lib.csref_divergence(h, _ip(_i32(s.face_cells)), C.c_int(len(bfc)), _ip(bfc),
                                      C.c_int(init), _dp(_f64(i_massflux)), _dp(_f64(b_massflux)),
                                      _dp(div))
        assert r == 0
        return div

    def system_block(self, n_rows, db, ri, ci, d_val, x_val, symmetric=False):
        """MSR matrix with db x db diagonal blocks (d_val: n_rows*db*db) built by the reference."""
        ri, ci, d, xv = _i32(ri), _i32(ci), _f64(d_val), _f64(x_val)
        self.lib.csref_system_create_block.restype = C.c_void_p
        return C.c_void_p(self.lib.csref_system_create_block(
            C.c_int(n_rows), C.c_int(db), _ip(ri), _ip(ci), _dp(d), _dp(xv),
            C.c_int(int(symmetric))))

    def system_csr(self, n_rows, edges, da, xa, symmetric=True):
        """CS_MATRIX_CSR built by the reference from an edge list; returns (handle,
        row_index, col_id, val) with the arrays the reference holds."""
        edges = np.ascontiguousarray(edges, dtype=np.int32)
        da, xa = _f64(da), _f64(xa)
        self.lib.csref_system_create_csr.restype = C.c_void_p
        h = C.c_void_p(self.lib.csref_system_create_csr(C.c_int(n_rows), C.c_int(len(edges)),
                                                        _ip(edges), _dp(da), _dp(xa),
                                                        C.c_int(int(symmetric))))
        nnz = self.lib.csref_get_csr_arrays(h, None, None, None)
        ri = np.empty(n_rows + 1, dtype=np.int32)
        ci = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz)
        self.lib.csref_get_csr_arrays(h, _ip(ri), _ip(ci), _dp(val))
        return h, ri, ci, val

    def destroy(self, h):
        self.lib.csref_system_destroy(h)

    def spmv(self, h, x, n_rows, exclude_diag=False):
        x = _f64(x).copy()
        y = np.empty(n_rows)
        self.lib.csref_spmv(h, C.c_int(int(exclude_diag)), _dp(x), _dp(y))
        return y

    def dot(self, x, y):
        x, y = _f64(x), _f64(y)
        return self.lib.csref_dot(C.c_int(len(x)), _dp(x), _dp(y))

    def dot_xx(self, x):
        x = _f64(x)
        return self.lib.csref_dot_xx(C.c_int(len(x)), _dp(x))

    def dot_xx_xy_yz(self, x, y, z):
        x, y, z = _f64(x), _f64(y), _f64(z)
        r = np.zeros(3)
        self.lib.csref_dot_xx_xy_yz(C.c_int(len(x)), _dp(x), _dp(y), _dp(z), _dp(r))
        return tuple(r)

    def sles_it(self, h, solver, rhs, n_max_iter, precision, r_norm, vx0=None, poly_degree=0):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(5)
        cvg = self.lib.csref_sles_it_solve(h, C.c_int(solver), C.c_int(poly_degree),
                                           C.c_int(n_max_iter), C.c_double(precision),
                                           C.c_double(r_norm), _dp(rhs), _dp(vx),
                                           C.c_int(0 if vx0 is None else 1), _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    t_setup=out[3], t_solve=out[4], vx=vx)

    def multigrid(self, h, rhs, precision, r_norm, n_max_cycles=100, n_max_iter=(2, 10, 500),
                  types=(0, 0, 0), poly_degree=(0, 0, 0), precision_mult=(1., 1., 1.),
                  aggregation_limit=3, coarsening_type=0, n_levels_max=25, n_g_rows_min=30,
                  p0p1_relax=0.95, vx0=None, verbosity=0, mg_type=0):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        iopt = np.array(list(types) + [n_max_cycles] + list(n_max_iter) + list(poly_degree)
                        + [aggregation_limit, coarsening_type, n_levels_max, n_g_rows_min,
                           mg_type], dtype=np.int32)
        dopt = np.array(list(precision_mult) + [p0p1_relax], dtype=np.float64)
        out = np.zeros(4)
        cvg = self.lib.csref_multigrid_solve(h, _ip(iopt), _dp(dopt), C.c_double(precision),
                                             C.c_double(r_norm), _dp(rhs), _dp(vx),
                                             C.c_int(0 if vx0 is None else 1),
                                             C.c_int(verbosity), _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], t_setup=out[2],
                    t_solve=out[3], vx=vx)

    def matrix_compute_coeffs_sym(self, h, s, b_face_cells, cofbfp, rovsdt, i_visc, b_visc,
                                  idiffp=1, ndircp=1, thetap=1., relaxp=1.):
        """cs_matrix_compute_coeffs (iconvp = 0) on the system's matrix; returns (d_val, x_val)."""
        bfc = _i32(b_face_cells)
        d, x = np.zeros(max(s.n_rows, 1)), np.zeros(max(s.nnz, 1))
        rc = self.lib.csref_matrix_compute_coeffs_sym(
            h, C.c_int(len(bfc)), _ip(bfc), C.c_int(idiffp), C.c_int(ndircp), C.c_double(thetap),
            C.c_double(relaxp), _dp(_f64(cofbfp)), _dp(_f64(rovsdt)), _dp(_f64(i_visc)),
            _dp(_f64(b_visc)), _dp(d), _dp(x))
        if rc != 0:
            raise RuntimeError("csref_matrix_compute_coeffs_sym: system has no mesh")
        return d[:s.n_rows], x[:s.nnz]

    def matrix_compute_coeffs(self, h, s, b_face_cells, coefbp, cofbfp, rovsdt, i_massflux,
                              b_massflux, i_visc, b_visc, xcpp=None, iconvp=1, idiffp=1,
                              ndircp=1, thetap=1., relaxp=1.):
        """cs_matrix_compute_coeffs, scalar cases, on the system's matrix (created with
        symmetric=False when iconvp = 1); returns (d_val, x_val)."""
        bfc = _i32(b_face_cells)
        d, x = np.zeros(max(s.n_rows, 1)), np.zeros(max(s.nnz, 1))
        f = lambda a: None if a is None else _f64(a)
        arrs = [f(a) for a in (coefbp, cofbfp, rovsdt, i_massflux, b_massflux, i_visc, b_visc, xcpp)]
        rc = self.lib.csref_matrix_compute_coeffs(
            h, C.c_int(len(bfc)), _ip(bfc), C.c_int(iconvp), C.c_int(idiffp), C.c_int(ndircp),
            C.c_double(thetap), C.c_double(relaxp), *[_dp(a) for a in arrs], _dp(d), _dp(x))
        if rc != 0:
            raise RuntimeError("csref_matrix_compute_coeffs: system has no mesh")
        return d[:s.n_rows], x[:s.nnz]

    def system_block_with_mesh(self, s, db=3):
        """Block-diagonal MSR system on the structure and mesh association of CubeSystem s
        (coefficients zero: to be filled by matrix_compute_coeffs_v3)."""
        h = self.system_block(s.n_rows, db, s.row_index, s.col_id,
                              np.tile(np.eye(db).ravel(), s.n_rows), np.zeros(s.nnz), False)
        fc, c2f = _i32(s.face_cells), _i32(s.c2f)
        sgn = np.ascontiguousarray(s.c2f_sgn, dtype=np.int16)
        self.lib.csref_system_set_mesh(h, C.c_int(s.n_faces), _ip(fc), _ip(c2f), _sp(sgn),
                                       _dp(_f64(s.cell_cen)), _dp(_f64(s.cell_vol)),
                                       _dp(_f64(s.face_u_normal)), _dp(_f64(s.face_surf)))
        return h

    def matrix_compute_coeffs_v3(self, h, s, b_face_cells, coefbp, cofbfp, fimp, i_massflux,
                                 b_massflux, i_visc, b_visc, iconvp=1, idiffp=1, ndircp=1,
                                 thetap=1., relaxp=1.):
        bfc = _i32(b_face_cells)
        d, x = np.zeros(9 * max(s.n_rows, 1)), np.zeros(max(s.nnz, 1))
        f = lambda a: None if a is None else _f64(a)
        arrs = [f(a) for a in (coefbp, cofbfp, fimp, i_massflux, b_massflux, i_visc, b_visc)]
        rc = self.lib.csref_matrix_compute_coeffs_v3(
            h, C.c_int(len(bfc)), _ip(bfc), C.c_int(iconvp), C.c_int(idiffp), C.c_int(ndircp),
            C.c_double(thetap), C.c_double(relaxp), *[_dp(a) for a in arrs], _dp(d), _dp(x))
        if rc != 0:
            raise RuntimeError("csref_matrix_compute_coeffs_v3: system has no mesh")
        return d[:9 * s.n_rows], x[:s.nnz]

    def smoother(self, h, stype, rhs, n_max_iter, vx0=None):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(4)
        cvg = self.lib.csref_smoother_solve(h, C.c_int(stype), C.c_int(n_max_iter), _dp(rhs),
                                            _dp(vx), C.c_int(0 if vx0 is None else 1),
                                            _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), vx=vx)

    def fcg_mgpc(self, h, rhs, n_max_iter, precision, r_norm, types=(16, 16, 0),
                 mg_n_max_iter=(3, 3, 500), **kw):
        return self.it_mgpc(h, 1, rhs, n_max_iter, precision, r_norm, types, mg_n_max_iter, **kw)

    def it_mgpc(self, h, solver, rhs, n_max_iter, precision, r_norm, types=(16, 16, 0),
                mg_n_max_iter=(3, 3, 500), poly_degree=(0, 0, 0),
                precision_mult=(-1., -1., 1.), aggregation_limit=3, coarsening_type=0,
                n_levels_max=25, n_g_rows_min=30, p0p1_relax=0.95, vx0=None, mg_type=0):
        rhs = _f64(rhs)
        vx = np.zeros(len(rhs)) if vx0 is None else _f64(vx0).copy()
        iopt = np.array(list(types) + [1] + list(mg_n_max_iter) + list(poly_degree)
                        + [aggregation_limit, coarsening_type, n_levels_max, n_g_rows_min,
                           mg_type], dtype=np.int32)
        dopt = np.array(list(precision_mult) + [p0p1_relax], dtype=np.float64)
        out = np.zeros(5)
        cvg = self.lib.csref_it_mgpc_solve(h, C.c_int(solver), _ip(iopt), _dp(dopt), C.c_int(n_max_iter),
                                            C.c_double(precision), C.c_double(r_norm),
                                            _dp(rhs), _dp(vx),
                                            C.c_int(0 if vx0 is None else 1), _dp(out))
        return dict(cvg=cvg, n_iter=int(out[0]), residual=out[1], initial_residual=out[2],
                    t_setup=out[3], t_solve=out[4], vx=vx)

    def hierarchy(self, h, aggregation_limit=3, coarsening_type=0, n_levels_max=25,
                  n_g_rows_min=30, p0p1_relax=0.95, verbosity=0):
        hh = C.c_void_p(self.lib.csref_hierarchy_build(h, C.c_int(aggregation_limit),
                                                       C.c_int(coarsening_type),
                                                       C.c_int(n_levels_max),
                                                       C.c_int(n_g_rows_min),
                                                       C.c_double(p0p1_relax),
                                                       C.c_int(verbosity)))
        n_levels = self.lib.csref_hierarchy_n_levels(hh)
        levels = []
        parent_n = 0
        for l in range(n_levels):
            info = np.zeros(3, dtype=np.int32)
            self.lib.csref_hierarchy_level_info(hh, C.c_int(l), _ip(info))
            n, nnz = int(info[0]), int(info[1])
            cr = np.zeros(max(parent_n, 1), dtype=np.int32)
            ri = np.zeros(n + 1, dtype=np.int32)
            ci = np.zeros(max(nnz, 1), dtype=np.int32)
            d = np.zeros(max(n, 1))
            xv = np.zeros(max(nnz, 1))
            cen = np.zeros(3 * max(n, 1))
            vol = np.zeros(max(n, 1))
            self.lib.csref_hierarchy_level_copy(hh, C.c_int(l), _ip(cr), _ip(ri), _ip(ci),
                                                _dp(d), _dp(xv), _dp(cen), _dp(vol))
            levels.append(dict(n_rows=n, nnz=nnz, n_faces=int(info[2]),
                               row_index=ri, col_id=ci[:nnz], d_val=d[:n], x_val=xv[:nnz],
                               coarse_row=cr[:parent_n], cell_cen=cen[:3 * n],
                               cell_vol=vol[:n]))
            parent_n = n
        self.lib.csref_hierarchy_destroy(hh)
        return levels


class RefLibB200(RefLib):
    """oracle/_ref/libcsref_b200.so: the same reference build plus integration/cs_sles_b200.cpp
    (the reference-side binding of libcsb200.so) and a driver that solves through the
    reference's cs_sles registry.  Needs a GPU at solve time."""

    LIBNAME = "libcsref_b200.so"
    MULTIGRID, FCG_MG, PCG_JACOBI, FCG_JACOBI = 0, 1, 2, 3

    def b200_sles_solve(self, h, kind, rhs, n_max_iter, precision, r_norm, vx0=None,
                        n_solves=1, free_between=False, n_cols=None):
        rhs = _f64(rhs)
        vx = np.zeros(n_cols or len(rhs)) if vx0 is None else _f64(vx0).copy()
        out = np.zeros(8)
        state = self.lib.csref_b200_sles_solve(
            h, C.c_int(kind), C.c_int(n_max_iter), C.c_double(precision), C.c_double(r_norm),
            _dp(rhs), _dp(vx), C.c_int(0 if vx0 is None else 1), C.c_int(n_solves),
            C.c_int(int(free_between)), _dp(out))
        return dict(cvg=state, n_iter=int(out[0]), residual=out[1], uploads=int(out[3]),
                    value_updates=int(out[4]), vx=vx)

    def finalize(self):
        self.lib.csref_b200_finalize()


def ref_b200_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libcsref_b200.so"))
