/*============================================================================
 * oracle/cs_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C, single-threaded CPU restatement of the reference (code_saturne)
 * src/alge linear-solve hot path, used ONLY as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product
 * (code_saturne_b200/csrc) never includes, links or calls this file.
 *
 * Parity status: PINNED.  Every function here is checked against the
 * reference's own compiled code (oracle/_ref/libcsref.so, built from
 * /root/reference by oracle/Makefile.ref) in tests/test_oracle.py,
 * and / or against the committed fixtures in tests/golden/ (generated from that
 * same reference build by the tests/golden/make_golden*.py scripts).  The reference tree
 * itself holds no golden vectors for this path.
 *
 * Arithmetic: plain IEEE double mul/add in the reference's evaluation order
 * (compile with -ffp-contract=off; the reference build has no FMA either).
 * All of this restates the reference behaviour with OMP_NUM_THREADS=1, except
 * where a "n_threads" argument restates cs_parall_thread_range blocking.
 *============================================================================*/

#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "cs_oracle.h"

#define ORA_CL_SIZE 64            /* CS_CL_SIZE, src/base/cs_defs.h:491 */
#define ORA_SBLOCK 60             /* CS_SBLOCK_BLOCK_SIZE, src/alge/cs_blas.cpp:81 */

/*----------------------------------------------------------------------------
 * cs_parall_thread_range (src/base/cs_parall.h:571-590) for thread t_id of n_t
 *----------------------------------------------------------------------------*/

static int _align(int i, int m) { return ((i + m - 1) / m) * m; }

void
ora_thread_range(int n, int type_size, int t_id, int n_t, int *s_id, int *e_id)
{
  const int t_n = (n + n_t - 1) / n_t;
  const int cl_m = ORA_CL_SIZE / type_size;
  *s_id = t_id * t_n;
  *e_id = (t_id + 1) * t_n;
  *s_id = _align(*s_id, cl_m);
  *e_id = _align(*e_id, cl_m);
  if (*e_id > n) *e_id = n;
  if (*s_id > n) *s_id = n;
}

/*----------------------------------------------------------------------------
 * y = A.x, scalar MSR: _mat_vec_p_l_msr (src/alge/cs_matrix_spmv.cpp:2103-2165)
 * off-diagonal sum in CSR order first, diagonal term added last.
 *----------------------------------------------------------------------------*/

void
ora_spmv_msr(int n_rows, const int *row_index, const int *col_id,
             const double *d_val, const double *x_val, int exclude_diag,
             const double *x, double *y)
{
  for (int ii = 0; ii < n_rows; ii++) {
    double sii = 0.0;
    for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++)
      sii += (x_val[jj] * x[col_id[jj]]);
    if (!exclude_diag && d_val != NULL)
      y[ii] = sii + d_val[ii] * x[ii];
    else
      y[ii] = sii;
  }
}

/*----------------------------------------------------------------------------
 * Superblock dot products: _cs_dot_*_superblock (src/alge/cs_blas.cpp:173-640)
 * blocks of 60, sqrt(n_blocks) superblocks (_sbloc_sizes :173-184).
 *----------------------------------------------------------------------------*/

static void
_sbloc_sizes(int n, int *n_sblocks, int *blocks_in_sblocks)
{
  int n_blocks = (n + ORA_SBLOCK - 1) / ORA_SBLOCK;
  *n_sblocks = (n_blocks > 1) ? (int)sqrt((double)n_blocks) : 1;
  int n_b = ORA_SBLOCK * *n_sblocks;
  *blocks_in_sblocks = (n + n_b - 1) / n_b;
}

/* generic up-to-3-way: s[0]=x.x (or x.y when mode says so) ... see callers */
#define ORA_SB_LOOP(BODY_DECL, BODY_ACC, BODY_SB, BODY_TOT)                  \
  int n_sblocks, blocks_in_sblocks;                                          \
  _sbloc_sizes(n, &n_sblocks, &blocks_in_sblocks);                           \
  for (int sid = 0; sid < n_sblocks; sid++) {                                \
    BODY_DECL(s)                                                             \
    for (int bid = 0; bid < blocks_in_sblocks; bid++) {                      \
      int start_id = ORA_SBLOCK * (blocks_in_sblocks*sid + bid);             \
      int end_id = ORA_SBLOCK * (blocks_in_sblocks*sid + bid + 1);           \
      if (end_id > n) end_id = n;                                            \
      BODY_DECL(c)                                                           \
      for (int i = start_id; i < end_id; i++) { BODY_ACC }                   \
      BODY_SB                                                                \
    }                                                                        \
    BODY_TOT                                                                 \
  }

double
ora_dot(int n, const double *x, const double *y)
{
  double dot = 0.0;
#define D1(p) double p##0 = 0.0;
  ORA_SB_LOOP(D1, c0 += x[i]*y[i];, s0 += c0;, dot += s0;)
  return dot;
}

double
ora_dot_xx(int n, const double *x)
{
  double dot = 0.0;
  ORA_SB_LOOP(D1, c0 += x[i]*x[i];, s0 += c0;, dot += s0;)
  return dot;
}

void
ora_dot_xx_xy(int n, const double *x, const double *y, double *xx, double *xy)
{
  double d0 = 0.0, d1 = 0.0;
#define D2(p) double p##0 = 0.0, p##1 = 0.0;
  ORA_SB_LOOP(D2, c0 += x[i]*x[i]; c1 += x[i]*y[i];,
              s0 += c0; s1 += c1;, d0 += s0; d1 += s1;)
  *xx = d0; *xy = d1;
}

void
ora_dot_xy_yz(int n, const double *x, const double *y, const double *z,
              double *xy, double *yz)
{
  double d0 = 0.0, d1 = 0.0;
  ORA_SB_LOOP(D2, c0 += x[i]*y[i]; c1 += y[i]*z[i];,
              s0 += c0; s1 += c1;, d0 += s0; d1 += s1;)
  *xy = d0; *yz = d1;
}

void
ora_dot_xx_xy_yz(int n, const double *x, const double *y, const double *z,
                 double *xx, double *xy, double *yz)
{
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
#define D3(p) double p##0 = 0.0, p##1 = 0.0, p##2 = 0.0;
  ORA_SB_LOOP(D3, c0 += x[i]*x[i]; c1 += x[i]*y[i]; c2 += y[i]*z[i];,
              s0 += c0; s1 += c1; s2 += c2;, d0 += s0; d1 += s1; d2 += s2;)
  *xx = d0; *xy = d1; *yz = d2;
}

/*----------------------------------------------------------------------------
 * CSR SpMV: _mat_vec_p_l_csr (src/alge/cs_matrix_spmv.cpp:1890-1955).  The
 * diagonal is an ordinary entry; exclude_diag skips col_id == row.
 *----------------------------------------------------------------------------*/

void
ora_spmv_csr(int n_rows, const int *row_index, const int *col_id, const double *val,
             int exclude_diag, const double *x, double *y)
{
  for (int ii = 0; ii < n_rows; ii++) {
    const int *restrict c = col_id + row_index[ii];
    const double *restrict m_row = val + row_index[ii];
    const int n_cols = row_index[ii+1] - row_index[ii];
    double sii = 0.0;
    if (!exclude_diag) {
      for (int jj = 0; jj < n_cols; jj++) sii += (m_row[jj]*x[c[jj]]);
    }
    else {
      for (int jj = 0; jj < n_cols; jj++)
        if (c[jj] != ii) sii += (m_row[jj]*x[c[jj]]);
    }
    y[ii] = sii;
  }
}

/*----------------------------------------------------------------------------
 * Convergence test: _convergence_test (src/alge/cs_sles_it.cpp:243-325)
 *----------------------------------------------------------------------------*/

typedef struct {
  double precision, r_norm, initial_residual;
  int    n_iterations_max;
  int    n_iterations;
  double residual;
  double *history;   /* optional: residual per test */
  int     history_size, history_n;
} ora_cvg_t;

static int
_cvg_test(ora_cvg_t *c, int n_iter, double residual)
{
  c->n_iterations = n_iter;
  c->residual = residual;
  if (c->history != NULL && c->history_n < c->history_size)
    c->history[c->history_n++] = residual;

  if (c->precision > 0 && residual < c->precision * c->r_norm)
    return ORA_CONVERGED;
  else if (n_iter >= c->n_iterations_max)
    return ORA_MAX_ITERATION;
  else if (residual > c->initial_residual * 10000.0 && residual > 100.)
    return ORA_DIVERGED;
  if (isnan(residual) || isinf(residual))
    return ORA_DIVERGED;
  return ORA_ITERATING;
}

/*----------------------------------------------------------------------------
 * Flexible conjugate gradient with Jacobi preconditioning:
 * _flexible_conjugate_gradient (src/alge/cs_sles_it.cpp:832-985), the default
 * outer solver of the pressure system.  Residual sign r = b - A.x.
 * Dots as in _dot_products_vr_vw_vq_rr (:693-724): cs_dot_xy_yz(w, v, q) then
 * cs_dot_xx_xy(r, v).  pc_apply = ad_inv scaling (cs_sles_pc.cpp:502-535).
 * out[0]=n_iter out[1]=residual out[2]=initial_residual
 *----------------------------------------------------------------------------*/

int
ora_fcg(int n_rows, const int *row_index, const int *col_id,
        const double *d_val, const double *x_val,
        int n_max_iter, double precision, double r_norm,
        const double *rhs, double *vx, int use_vx_ini,
        double *out, double *history, int history_size)
{
  const int n = n_rows;
  double *rk = malloc(sizeof(double)*n*6);
  double *vk = rk + n, *wk = rk + 2*n, *dk = rk + 3*n, *qk = rk + 4*n;
  double *ad_inv = rk + 5*n;
  for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[i];

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  double residual = -1;

  if (use_vx_ini) {
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, vx, rk);
    for (int ii = 0; ii < n; ii++) rk[ii] = rhs[ii] - rk[ii];
  }
  else {
    for (int ii = 0; ii < n; ii++) { vx[ii] = 0; rk[ii] = rhs[ii]; }
  }
  for (int ii = 0; ii < n; ii++) qk[ii] = 0;

  double rho_km1 = 0;

  while (cvg == ORA_ITERATING) {
    for (int ii = 0; ii < n; ii++) vk[ii] = rk[ii] * ad_inv[ii];
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, vk, wk);

    double alpha_k, beta_k, gamma_k;
    ora_dot_xy_yz(n, wk, vk, qk, &beta_k, &gamma_k);
    ora_dot_xx_xy(n, rk, vk, &residual, &alpha_k);
    residual = sqrt(residual);

    if (n_iter > 0)
      cvg = _cvg_test(&cv, n_iter, residual);
    else {
      cv.initial_residual = residual;
      if (history != NULL && cv.history_n < history_size)
        history[cv.history_n++] = residual;
    }
    if (cvg != ORA_ITERATING) break;

    if (n_iter > 0) {
      const double gk_rk1 = (fabs(rho_km1) > DBL_MIN) ? gamma_k / rho_km1 : 0.;
      const double rho_k = beta_k - gamma_k * gk_rk1;
      const double ak_rk = (fabs(rho_k) > DBL_MIN) ? alpha_k / rho_k : 0.;
      for (int ii = 0; ii < n; ii++) {
        dk[ii] = vk[ii] - gk_rk1 * dk[ii];
        vx[ii] = vx[ii] + ak_rk * dk[ii];
      }
      for (int ii = 0; ii < n; ii++) {
        qk[ii] = wk[ii] - gk_rk1 * qk[ii];
        rk[ii] = rk[ii] - ak_rk * qk[ii];
      }
      rho_km1 = rho_k;
    }
    else {
      const double rho_k = beta_k;
      const double ak_rk = (fabs(rho_k) > DBL_MIN) ? alpha_k / rho_k : 0.;
      for (int ii = 0; ii < n; ii++) {
        dk[ii] = vk[ii];
        vx[ii] = vx[ii] + ak_rk * vk[ii];
      }
      for (int ii = 0; ii < n; ii++) {
        qk[ii] = wk[ii];
        rk[ii] = rk[ii] - ak_rk * wk[ii];
      }
      rho_km1 = rho_k;
    }
    n_iter += 1;
  }

  out[0] = n_iter; out[1] = residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Inexact-preconditioning CG with Jacobi preconditioning:
 * _conjugate_gradient_ip (src/alge/cs_sles_it.cpp:1016-1203).
 * Residual sign r = A.x - b; beta = (rk.gk - rkm1.gk) / rkm1.gkm1.
 * out[0]=n_iter out[1]=residual out[2]=initial_residual out[3]=history count
 *----------------------------------------------------------------------------*/

int
ora_ipcg(int n_rows, const int *row_index, const int *col_id,
         const double *d_val, const double *x_val,
         int n_max_iter, double precision, double r_norm,
         const double *rhs, double *vx, int use_vx_ini,
         double *out, double *history, int history_size)
{
  const int n = n_rows;
  double *rk = malloc(sizeof(double)*n*6);
  double *rkm1 = rk + n, *dk = rk + 2*n, *gk = rk + 3*n, *zk = rk + 4*n;
  double *ad_inv = rk + 5*n;
  for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[i];

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg;
  int n_iter = 0;
  double ro_0, ro_1, alpha, rk_gk_m1, rkm1_gk, rk_gk, beta, residual;

  if (use_vx_ini) {
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, vx, rk);
    for (int ii = 0; ii < n; ii++) rk[ii] -= rhs[ii];
  }
  else {
    for (int ii = 0; ii < n; ii++) { vx[ii] = 0; rk[ii] = -rhs[ii]; }
  }
  for (int ii = 0; ii < n; ii++) gk[ii] = rk[ii] * ad_inv[ii];
  for (int ii = 0; ii < n; ii++) dk[ii] = gk[ii];

  ora_dot_xx_xy(n, rk, gk, &residual, &rk_gk_m1);
  residual = sqrt(residual);
  cv.initial_residual = residual;
  cvg = _cvg_test(&cv, n_iter, residual);

  if (cvg == ORA_ITERATING) {
    n_iter = 1;
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, dk, zk);
    ora_dot_xy_yz(n, rk, dk, zk, &ro_0, &ro_1);
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    for (int ii = 0; ii < n; ii++) vx[ii] += (alpha * dk[ii]);
    for (int ii = 0; ii < n; ii++) { rkm1[ii] = rk[ii]; rk[ii] += (alpha * zk[ii]); }
    residual = sqrt(ora_dot_xx(n, rk));
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  while (cvg == ORA_ITERATING) {
    for (int ii = 0; ii < n; ii++) gk[ii] = rk[ii] * ad_inv[ii];
    ora_dot_xx_xy_yz(n, rk, gk, rkm1, &residual, &rk_gk, &rkm1_gk);
    residual = sqrt(residual);
    if (n_iter > 1)
      cvg = _cvg_test(&cv, n_iter, residual);
    if (cvg != ORA_ITERATING) break;
    n_iter += 1;
    beta = (fabs(rk_gk_m1) > DBL_MIN) ? (rk_gk - rkm1_gk) / rk_gk_m1 : 0.;
    rk_gk_m1 = rk_gk;
    for (int ii = 0; ii < n; ii++) dk[ii] = gk[ii] + (beta * dk[ii]);
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, dk, zk);
    ora_dot_xy_yz(n, rk, dk, zk, &ro_0, &ro_1);
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    for (int ii = 0; ii < n; ii++) vx[ii] += (alpha * dk[ii]);
    for (int ii = 0; ii < n; ii++) { rkm1[ii] = rk[ii]; rk[ii] += (alpha * zk[ii]); }
  }

  out[0] = cv.n_iterations; out[1] = cv.residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Unpreconditioned CG, single reduction: _conjugate_gradient_sr_npc
 * (src/alge/cs_sles_it.cpp:1673-1860; selected when poly_degree < 0).
 * out[0]=n_iter out[1]=residual out[2]=initial_residual out[3]=history count
 *----------------------------------------------------------------------------*/

int
ora_pcg_sr_npc(int n_rows, const int *row_index, const int *col_id,
               const double *d_val, const double *x_val,
               int n_max_iter, double precision, double r_norm,
               const double *rhs, double *vx, int use_vx_ini,
               double *out, double *history, int history_size)
{
  const int n = n_rows;
  double *rk = malloc(sizeof(double)*n*4);
  double *dk = rk + n, *zk = rk + 2*n, *sk = rk + 3*n;

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;

  if (use_vx_ini) {
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, vx, rk);
    for (int ii = 0; ii < n; ii++) { rk[ii] = rk[ii] - rhs[ii]; dk[ii] = rk[ii]; }
  }
  else {
    for (int ii = 0; ii < n; ii++) { vx[ii] = 0; rk[ii] = - rhs[ii]; dk[ii] = rk[ii]; }
  }
  ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, dk, zk);

  double ro_0 = 0, ro_1 = 0;
  double rk_rkm1 = 0, residual = -1, rk_rk = 0, rk_sk = 0;
  ora_dot_xx_xy_yz(n, rk, dk, zk, &residual, &ro_0, &ro_1);
  residual = sqrt(residual);
  cv.initial_residual = residual;
  cvg = _cvg_test(&cv, n_iter, residual);

  if (cvg == ORA_ITERATING) {
    n_iter = 1;
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    double alpha = - ro_0 * d_ro_1;
    rk_rkm1 = ro_0;
    for (int ii = 0; ii < n; ii++) {
      vx[ii] += (alpha * dk[ii]);
      rk[ii] += (alpha * zk[ii]);
    }
    residual = sqrt(ora_dot_xx(n, rk));
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  while (cvg == ORA_ITERATING) {
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, rk, sk);
    ora_dot_xx_xy(n, rk, sk, &residual, &rk_sk);
    rk_rk = residual;
    residual = sqrt(residual);
    if (n_iter > 1)
      cvg = _cvg_test(&cv, n_iter, residual);
    if (cvg != ORA_ITERATING) break;
    n_iter += 1;
    double beta = (fabs(rk_rkm1) > DBL_MIN) ? rk_rk / rk_rkm1 : 0.;
    rk_rkm1 = rk_rk;
    ro_1 = rk_sk - beta*beta*ro_1;
    ro_0 = rk_rk;
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    double alpha = - ro_0 * d_ro_1;
    for (int ii = 0; ii < n; ii++) {
      dk[ii] = rk[ii] + (beta * dk[ii]);
      vx[ii] += alpha * dk[ii];
      zk[ii] = sk[ii] + (beta * zk[ii]);
      rk[ii] += alpha * zk[ii];
    }
  }

  out[0] = cv.n_iterations; out[1] = cv.residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/*----------------------------------------------------------------------------
 * PCG + Jacobi, single reduction: _conjugate_gradient_sr_p0
 * (src/alge/cs_sles_it.cpp:1441-1645).  Residual sign r = A.x - b.
 * out[0]=n_iter out[1]=residual out[2]=initial_residual
 *----------------------------------------------------------------------------*/

int
ora_pcg_sr_p0(int n_rows, const int *row_index, const int *col_id,
              const double *d_val, const double *x_val,
              int n_max_iter, double precision, double r_norm,
              const double *rhs, double *vx, int use_vx_ini,
              double *out, double *history, int history_size)
{
  const int n = n_rows;
  double *rk = malloc(sizeof(double)*n*6);
  double *dk = rk + n, *gk = rk + 2*n, *zk = rk + 3*n, *sk = rk + 4*n;
  double *ad_inv = rk + 5*n;
  for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[i]; /* cs_sles_pc.cpp:376 */

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  const int update_residual = (precision > 0.);

  if (use_vx_ini) {
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, vx, rk);
    for (int ii = 0; ii < n; ii++) {
      rk[ii] -= rhs[ii];
      dk[ii] = rk[ii] * ad_inv[ii];
    }
  }
  else {
    for (int ii = 0; ii < n; ii++) {
      vx[ii] = 0;
      rk[ii] = -rhs[ii];
      dk[ii] = rk[ii] * ad_inv[ii];
    }
  }

  ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, dk, zk);

  double ro_0 = 0, ro_1 = 0;
  double rk_gkm1 = 0, residual = -1, rk_gk = 0, gk_sk = 0;

  if (update_residual) {
    ora_dot_xx_xy_yz(n, rk, dk, zk, &residual, &ro_0, &ro_1);
    residual = sqrt(residual);
  }
  else
    ora_dot_xy_yz(n, rk, dk, zk, &ro_0, &ro_1);

  cv.initial_residual = residual;
  cvg = _cvg_test(&cv, n_iter, residual);

  if (cvg == ORA_ITERATING) {
    n_iter = 1;
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    double alpha = - ro_0 * d_ro_1;
    rk_gkm1 = ro_0;

    /* parallel_for_reduce_sum: same superblock order (cs_dispatch.h:369-420) */
    for (int ii = 0; ii < n; ii++) {
      vx[ii] += (alpha * dk[ii]);
      rk[ii] += (alpha * zk[ii]);
      gk[ii] = rk[ii] * ad_inv[ii];
    }
    if (update_residual)
      residual = sqrt(ora_dot_xx(n, rk));

    cvg = _cvg_test(&cv, n_iter, residual);
  }

  while (cvg == ORA_ITERATING) {
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, gk, sk);

    if (update_residual) {
      ora_dot_xx_xy_yz(n, rk, gk, sk, &residual, &rk_gk, &gk_sk);
      residual = sqrt(residual);
    }
    else
      ora_dot_xy_yz(n, rk, gk, sk, &rk_gk, &gk_sk);

    if (n_iter > 1)
      cvg = _cvg_test(&cv, n_iter, residual);
    if (cvg != ORA_ITERATING)
      break;

    n_iter += 1;

    double beta = (fabs(rk_gkm1) > DBL_MIN) ? rk_gk / rk_gkm1 : 0.;
    rk_gkm1 = rk_gk;
    ro_1 = gk_sk - beta*beta*ro_1;
    ro_0 = rk_gk;
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    double alpha = - ro_0 * d_ro_1;

    for (int ii = 0; ii < n; ii++) {
      dk[ii] = gk[ii] + (beta * dk[ii]);
      vx[ii] += alpha * dk[ii];
      zk[ii] = sk[ii] + (beta * zk[ii]);
      rk[ii] += alpha * zk[ii];
      gk[ii] = rk[ii] * ad_inv[ii];
    }
  }

  out[0] = cv.n_iterations;
  out[1] = cv.residual;
  out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Jacobi: _jacobi (src/alge/cs_sles_it.cpp:2098-2270)
 *----------------------------------------------------------------------------*/

int
ora_jacobi(int n_rows, const int *row_index, const int *col_id,
           const double *d_val, const double *x_val,
           int n_max_iter, double precision, double r_norm,
           const double *rhs, double *vx, int use_vx_ini,
           double *out, double *history, int history_size)
{
  const int n = n_rows;
  double *rk = malloc(sizeof(double)*n*2);
  double *ad_inv = rk + n;
  for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[i];

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  double residual = -1;

  if (!use_vx_ini) {
    n_iter += 1;
    double res2 = 0.0;
    for (int ii = 0; ii < n; ii++) {
      vx[ii] = rhs[ii]*ad_inv[ii];
      double r = rhs[ii];
      res2 += (r*r);
      rk[ii] = vx[ii];
    }
    if (precision > 0.)
      residual = sqrt(res2);
    if (n_iter == 1) cv.initial_residual = residual;
    cvg = _cvg_test(&cv, n_iter, residual);
  }
  else {
    for (int ii = 0; ii < n; ii++) rk[ii] = vx[ii];
  }

  while (cvg == ORA_ITERATING) {
    n_iter += 1;
    ora_spmv_msr(n, row_index, col_id, d_val, x_val, 1, rk, vx);
    double res2 = 0.0;
    for (int ii = 0; ii < n; ii++) {
      vx[ii] = (rhs[ii]-vx[ii])*ad_inv[ii];
      double r = d_val[ii] * (vx[ii]-rk[ii]);
      res2 += (r*r);
      rk[ii] = vx[ii];
    }
    if (precision > 0.)
      residual = sqrt(res2);
    if (n_iter == 1) cv.initial_residual = residual;
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  out[0] = cv.n_iterations;
  out[1] = cv.residual;
  out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Process-local Gauss-Seidel solvers with convergence test, one thread:
 *   sym == 0: _p_gauss_seidel_msr      (src/alge/cs_sles_it.cpp:3940-4110)
 *   sym != 0: _p_sym_gauss_seidel_msr  (src/alge/cs_sles_it.cpp:4124-4360)
 * The residual of an iteration is sqrt(sum (d_i (vx_i_new - vx_i_old))^2) over
 * the forward sweep (GS) or over the backward sweep (symmetric GS), rows summed
 * in sweep order.  With precision <= 0 no residual is computed and the loop
 * ends at n_max_iter (:4098-4101, :4349-4352).
 * out as ora_jacobi.
 *----------------------------------------------------------------------------*/

int
ora_gauss_seidel(int sym, int n_rows, const int *row_index, const int *col_id,
                 const double *d_val, const double *x_val,
                 int n_max_iter, double precision, double r_norm,
                 const double *rhs, double *vx, int use_vx_ini,
                 double *out, double *history, int history_size)
{
  const int n = n_rows;
  double *ad_inv = malloc(sizeof(double)*(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[i];

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  double residual = -1;

  while (cvg == ORA_ITERATING) {
    if (n_iter == 0 && !use_vx_ini)
      for (int ii = 0; ii < n; ii++) vx[ii] = 0;
    n_iter += 1;
    double res2 = 0.0;
    if (sym) {
      for (int ii = 0; ii < n; ii++) {
        double vx0 = rhs[ii];
        for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++)
          vx0 -= (x_val[jj]*vx[col_id[jj]]);
        vx[ii] = vx0 * ad_inv[ii];
      }
      for (int ii = n - 1; ii > -1; ii--) {
        double vxm1 = vx[ii];
        double vx0 = rhs[ii];
        for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++)
          vx0 -= (x_val[jj]*vx[col_id[jj]]);
        vx0 *= ad_inv[ii];
        double r = d_val[ii] * (vx0 - vxm1);
        res2 += (r*r);
        vx[ii] = vx0;
      }
    }
    else {
      for (int ii = 0; ii < n; ii++) {
        double vxm1 = vx[ii];
        double vx0 = rhs[ii];
        for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++)
          vx0 -= (x_val[jj]*vx[col_id[jj]]);
        vx0 *= ad_inv[ii];
        double r = d_val[ii] * (vx0 - vxm1);
        res2 += (r*r);
        vx[ii] = vx0;
      }
    }
    if (precision > 0.) {
      residual = sqrt(res2);   /* of the previous iterate */
      if (n_iter == 1) cv.initial_residual = residual;
      cvg = _cvg_test(&cv, n_iter, residual);
    }
    else if (n_iter >= n_max_iter) {
      cv.n_iterations = n_iter;
      cvg = ORA_MAX_ITERATION;
    }
  }

  out[0] = cv.n_iterations;
  out[1] = cv.residual;
  out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(ad_inv);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Block-diagonal MSR SpMV, 3x3 blocks: _b_mat_vec_p_l_msr_3
 * (src/alge/cs_matrix_spmv.cpp:2394-2455) with _dense_3_3_ax (:590-607):
 * y_i = D_i.x_i first, then += e_ij * x_j in CSR order, per component.
 *----------------------------------------------------------------------------*/

void
ora_spmv_msr_b3(int n_rows, const int *row_index, const int *col_id,
                const double *d_val, const double *x_val, int exclude_diag,
                const double *x, double *y)
{
  for (int ii = 0; ii < n_rows; ii++) {
    const int *restrict c = col_id + row_index[ii];
    const double *restrict m_row = x_val + row_index[ii];
    const int n_cols = row_index[ii+1] - row_index[ii];
    if (!exclude_diag) {
      const double *a = d_val;
      y[ii*3]     =   a[ii*9]         * x[ii*3]
                    + a[ii*9 + 1]     * x[ii*3 + 1]
                    + a[ii*9 + 2]     * x[ii*3 + 2];
      y[ii*3 + 1] =   a[ii*9 + 3]     * x[ii*3]
                    + a[ii*9 + 3 + 1] * x[ii*3 + 1]
                    + a[ii*9 + 3 + 2] * x[ii*3 + 2];
      y[ii*3 + 2] =   a[ii*9 + 6]     * x[ii*3]
                    + a[ii*9 + 6 + 1] * x[ii*3 + 1]
                    + a[ii*9 + 6 + 2] * x[ii*3 + 2];
    }
    else {
      for (int kk = 0; kk < 3; kk++) y[ii*3 + kk] = 0.;
    }
    for (int jj = 0; jj < n_cols; jj++)
      for (int kk = 0; kk < 3; kk++)
        y[ii*3 + kk] += (m_row[jj]*x[c[jj]*3 + kk]);
  }
}

/*----------------------------------------------------------------------------
 * Block-diagonal MSR SpMV, any block size: _b_mat_vec_p_l_msr_generic
 * (src/alge/cs_matrix_spmv.cpp:2308-2380) with _dense_b_ax (:560-576, sums from 0.), and for
 * 6x6 blocks _b_mat_vec_p_l_msr_6 (:2478-2545) with _dense_6_6_ax (:622-670, sums from the
 * first product); selected by block size like _b_mat_vec_p_l_msr (:2562-2580).
 *----------------------------------------------------------------------------*/

void
ora_spmv_msr_bn(int n_rows, int db, const int *row_index, const int *col_id,
                const double *d_val, const double *x_val, int exclude_diag,
                const double *x, double *y)
{
  if (db == 3) {
    ora_spmv_msr_b3(n_rows, row_index, col_id, d_val, x_val, exclude_diag, x, y);
    return;
  }
  const int db2 = db*db;
  for (int ii = 0; ii < n_rows; ii++) {
    const int *restrict c = col_id + row_index[ii];
    const double *restrict m_row = x_val + row_index[ii];
    const int n_cols = row_index[ii+1] - row_index[ii];
    if (!exclude_diag) {
      for (int kk = 0; kk < db; kk++) {
        const double *a = d_val + (size_t)ii*db2 + kk*db;
        if (db == 6)
          y[ii*6 + kk] =   a[0]*x[ii*6]     + a[1]*x[ii*6 + 1] + a[2]*x[ii*6 + 2]
                         + a[3]*x[ii*6 + 3] + a[4]*x[ii*6 + 4] + a[5]*x[ii*6 + 5];
        else {
          y[ii*db + kk] = 0.;
          for (int jj = 0; jj < db; jj++)
            y[ii*db + kk] += a[jj] * x[ii*db + jj];
        }
      }
    }
    else {
      for (int kk = 0; kk < db; kk++) y[ii*db + kk] = 0.;
    }
    for (int jj = 0; jj < n_cols; jj++)
      for (int kk = 0; kk < db; kk++)
        y[ii*db + kk] += (m_row[jj]*x[c[jj]*db + kk]);
  }
}

/*----------------------------------------------------------------------------
 * BiCGStab + Jacobi preconditioner: _bi_cgstab (src/alge/cs_sles_it.cpp:2693-2880)
 * with pc_apply = x*ad_inv (src/alge/cs_sles_pc.cpp:502-535).
 * db = 1 (scalar) or 3 (block-diagonal matrix: vectors of n_rows*3, point
 * Jacobi ad_inv[i] = 1/ad[i*db + i%db], cs_sles_pc.cpp:383-388).
 *----------------------------------------------------------------------------*/

int
ora_bicgstab_b(int n_rows, int db, const int *row_index, const int *col_id,
               const double *d_val, const double *x_val,
               int n_max_iter, double precision, double r_norm,
               const double *rhs, double *vx, int use_vx_ini,
               double *out, double *history, int history_size)
{
  const int n = n_rows*db;
  const double epzero = 1.e-30;   /* _epzero, cs_sles_it.cpp:169 */
  double *res0 = malloc(sizeof(double)*n*7);
  double *rk = res0 + n, *pk = res0 + 2*n, *zk = res0 + 3*n;
  double *uk = res0 + 4*n, *vk = res0 + 5*n, *ad_inv = res0 + 6*n;
  for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[(size_t)i*db + (i % db)];

#define ORA_SPMV(xx, yy) \
  do { if (db == 1) ora_spmv_msr(n_rows, row_index, col_id, d_val, x_val, 0, xx, yy); \
       else ora_spmv_msr_bn(n_rows, db, row_index, col_id, d_val, x_val, 0, xx, yy); } while (0)

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  double ro_0, ro_1, alpha, beta = 0, betam1, gamma, omega, ukres0;
  double residual = -1;

  for (int ii = 0; ii < n; ii++) { pk[ii] = 0.0; uk[ii] = 0.0; }

  if (use_vx_ini) {
    ORA_SPMV(vx, res0);
    for (int ii = 0; ii < n; ii++) {
      res0[ii] = -res0[ii] + rhs[ii];
      rk[ii] = res0[ii];
    }
  }
  else {
    for (int ii = 0; ii < n; ii++) {
      vx[ii] = 0; res0[ii] = rhs[ii]; rk[ii] = rhs[ii];
    }
  }

  alpha = 1.0; betam1 = 1.0; gamma = 1.0;

  while (cvg == ORA_ITERATING) {
    if (n_iter == 0) {
      beta = ora_dot_xx(n, rk);
      residual = sqrt(beta);
      cv.initial_residual = residual;
    }
    else {
      ora_dot_xx_xy(n, rk, res0, &residual, &beta);
      residual = sqrt(residual);
    }
    cvg = _cvg_test(&cv, n_iter, residual);
    if (cvg != ORA_ITERATING) break;

    n_iter += 1;

    if (fabs(beta) < epzero) { cvg = ORA_BREAKDOWN; break; }
    if (fabs(alpha) < epzero) { cvg = ORA_BREAKDOWN; break; }

    omega = beta*gamma / (alpha*betam1);
    betam1 = beta;

    for (int ii = 0; ii < n; ii++)
      pk[ii] = rk[ii] + omega*(pk[ii] - alpha*uk[ii]);
    for (int ii = 0; ii < n; ii++)
      zk[ii] = pk[ii] * ad_inv[ii];
    ORA_SPMV(zk, uk);
    ukres0 = ora_dot(n, uk, res0);
    gamma = beta / ukres0;
    for (int ii = 0; ii < n; ii++) {
      vx[ii] += (gamma * zk[ii]);
      rk[ii] -= (gamma * uk[ii]);
    }
    for (int ii = 0; ii < n; ii++)
      zk[ii] = rk[ii] * ad_inv[ii];
    ORA_SPMV(zk, vk);
    ora_dot_xx_xy(n, vk, rk, &ro_1, &ro_0);
    if (fabs(ro_1) < epzero) { cvg = ORA_BREAKDOWN; break; }
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = ro_0 * d_ro_1;
    for (int ii = 0; ii < n; ii++) {
      vx[ii] += (alpha * zk[ii]);
      rk[ii] -= (alpha * vk[ii]);
    }
  }
#undef ORA_SPMV

  out[0] = cv.n_iterations;
  out[1] = cv.residual;
  out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(res0);
  return cvg;
}

int
ora_bicgstab(int n_rows, const int *row_index, const int *col_id,
             const double *d_val, const double *x_val,
             int n_max_iter, double precision, double r_norm,
             const double *rhs, double *vx, int use_vx_ini,
             double *out, double *history, int history_size)
{
  return ora_bicgstab_b(n_rows, 1, row_index, col_id, d_val, x_val, n_max_iter, precision,
                        r_norm, rhs, vx, use_vx_ini, out, history, history_size);
}

/*============================================================================
 * Grid hierarchy (cs_grid.cpp)
 *============================================================================*/

/* _reorder_f_c_row (src/alge/cs_grid.cpp:1638-1664) */
static void
_reorder_f_c_row(int f_n_rows, int c_n_rows, int *c_o2n, int *f_c_row)
{
  for (int i = 0; i < c_n_rows; i++) c_o2n[i] = -1;
  int c_count = 0;
  for (int i = 0; i < f_n_rows; i++) {
    int c_row = f_c_row[i];
    if (c_row > -1) {
      if (c_o2n[c_row] < 0) {
        c_o2n[c_row] = c_count;
        f_c_row[i] = c_count;
        c_count++;
      }
      else
        f_c_row[i] = c_o2n[c_row];
    }
  }
}

/*----------------------------------------------------------------------------
 * _automatic_aggregation_dx_msr (src/alge/cs_grid.cpp:4282-4767), symmetric
 * scalar case (isym = 1, CS_GRID_COARSENING_SPD_DX).  n_threads restates the
 * per-thread row blocks (aggregation never crosses a block, :4374-4376,
 * :4492) followed by _scan_f_c_row (:1686-1716).
 * f_c_row (size n_rows) is written; returns the number of coarse rows.
 *----------------------------------------------------------------------------*/

int
ora_aggregation_dx_msr(int level, int f_n_rows, const int *row_index,
                       const int *col_id, const double *d_val,
                       const double *x_val, int max_aggregation,
                       int n_threads, int *f_c_row)
{
  const int npass_max = 10;
  const double epsilon = 1.e-6;
  const double penalization_threshold = 1e4;  /* cs_grid.cpp:165 */
  const int f_nnz = row_index[f_n_rows];

  int n_loc_threads = n_threads;
  { /* cs_parall_n_threads(f_n_rows, CS_THR_MIN) (cs_parall.h:549-565) */
    int n_t_l = f_n_rows / 128;
    if (n_t_l < n_loc_threads) n_loc_threads = n_t_l;
    if (n_loc_threads < 1) n_loc_threads = 1;
  }

  for (int ii = 0; ii < f_n_rows; ii++) f_c_row[ii] = -2;  /* _coarse_init :520 */

  short *c_cardinality = malloc(sizeof(short)*(f_n_rows > 0 ? f_n_rows : 1));
  short *c_aggr_count = malloc(sizeof(short)*(f_n_rows > 0 ? f_n_rows : 1));
  const int ag_queue_stride = 2;
  int ag_work_size = f_nnz*ag_queue_stride;
  if (ag_work_size < f_n_rows) ag_work_size = f_n_rows;
  int *ag_work = malloc(sizeof(int)*(ag_work_size > 0 ? ag_work_size : 1));
  int *ag_queue = ag_work;
  int *t_c_scan = calloc(n_loc_threads, sizeof(int));
  int *t_range = malloc(sizeof(int)*2*n_loc_threads);
  int c_n_rows = 0;

  for (int t_id = 0; t_id < n_loc_threads; t_id++) {
    int t_s_id, t_e_id;
    ora_thread_range(f_n_rows, 8, t_id, n_loc_threads, &t_s_id, &t_e_id);
    t_range[2*t_id] = t_s_id; t_range[2*t_id+1] = t_e_id;

    int ncoarse = 8, inc_nei = 0;
    int _max_aggregation = 1;

    if (level == 0) {
      for (int ii = t_s_id; ii < t_e_id; ii++) {
        c_cardinality[ii] = row_index[ii+1] - row_index[ii];
        c_aggr_count[ii] = 1;
        double sum = 0.0;
        for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++) {
          const double xv = x_val[jj];
          if (xv < 0.) sum -= xv; else sum += xv;
        }
        if (d_val[ii] > penalization_threshold * sum) {
          f_c_row[ii] = -1;
          c_aggr_count[ii] = 0;
        }
      }
    }
    else {
      for (int ii = t_s_id; ii < t_e_id; ii++) {
        c_cardinality[ii] = row_index[ii+1] - row_index[ii];
        c_aggr_count[ii] = 1;
      }
    }

    const int t_r_s_id = row_index[t_s_id] * ag_queue_stride;
    int t_r_e_id = t_r_s_id;
    int t_c_n_rows = 0;
    int r_n_faces = 0;

    /* First pass */
    {
      if (_max_aggregation < max_aggregation) _max_aggregation++;
      const double ag_threshold = 1. - epsilon;

      for (int ii = t_s_id; ii < t_e_id; ii++) {
        for (int ix0 = row_index[ii]; ix0 < row_index[ii+1]; ix0++) {
          int jj = col_id[ix0];
          if (jj < ii || jj >= t_e_id) continue;
          int ii_c = f_c_row[ii];
          int jj_c = f_c_row[jj];
          if (ii_c == -1 || jj_c == -1 || (ii_c >= 0 && jj_c >= 0)) continue;
          int ix1 = ix0;

          double f_da0_da1_inv = (c_cardinality[ii]*c_cardinality[jj])
                                 / (d_val[ii] * d_val[jj]);
          double f_xa0_xa1 = x_val[ix0] * x_val[ix1];
          double aggr_crit = f_da0_da1_inv * f_xa0_xa1;

          int count = 0;
          if (aggr_crit > ag_threshold) {
            if (ii_c > -1 && jj_c < -1) {
              if (c_aggr_count[t_s_id + ii_c] < _max_aggregation + 1) {
                f_c_row[jj] = ii_c;
                c_aggr_count[t_s_id + ii_c] += 1;
                count++;
              }
            }
            else if (ii_c < -1 && jj_c > -1) {
              if (c_aggr_count[t_s_id + jj_c] < _max_aggregation + 1) {
                f_c_row[ii] = jj_c;
                c_aggr_count[t_s_id + jj_c] += 1;
                count++;
              }
            }
            else if (ii_c < -1 && jj_c < -1) {
              f_c_row[ii] = t_c_n_rows;
              f_c_row[jj] = t_c_n_rows;
              c_aggr_count[t_s_id + t_c_n_rows] += 1;
              t_c_n_rows++;
              count++;
            }
          }
          if (count == 0) {
            ag_queue[t_r_e_id] = ii;
            ag_queue[t_r_e_id+1] = ix0;
            t_r_e_id += ag_queue_stride;
            r_n_faces++;
          }
        }
      }
    }

    /* Other passes */
    for (int npass = 2; npass <= npass_max; npass++) {
      int aggr_count = 0;
      for (int ii = t_s_id; ii < t_e_id; ii++)
        if (f_c_row[ii] < -1) aggr_count++;

      if (   aggr_count == 0
          || (t_c_n_rows + aggr_count)*ncoarse < (t_e_id - t_s_id)
          || r_n_faces == 0)
        break;

      if (_max_aggregation < max_aggregation) _max_aggregation++;
      r_n_faces = 0;
      inc_nei += 1;
      int r_idx_n = t_r_s_id;

      for (int r_idx = t_r_s_id; r_idx < t_r_e_id; r_idx += ag_queue_stride) {
        const double ag_threshold = 1. - epsilon;
        int ii = ag_queue[r_idx];
        int ix0 = ag_queue[r_idx+1];
        int ix1 = ix0;
        int jj = col_id[ix0];
        int ii_c = f_c_row[ii];
        int jj_c = f_c_row[jj];
        if (ii_c >= 0 && jj_c >= 0) continue;
        int count = 0;

        double f_da0_da1_inv = (  (c_cardinality[ii]+inc_nei)
                                * (c_cardinality[jj]+inc_nei))
                               / (d_val[ii] * d_val[jj]);
        double f_xa0_xa1 = x_val[ix0] * x_val[ix1];
        double aggr_crit = f_da0_da1_inv * f_xa0_xa1;

        if (aggr_crit > ag_threshold) {
          if (ii_c > -1 && jj_c < -1) {
            if (c_aggr_count[t_s_id + ii_c] < _max_aggregation + 1) {
              f_c_row[jj] = ii_c;
              c_aggr_count[t_s_id + ii_c] += 1;
              count++;
            }
          }
          else if (ii_c < -1 && jj_c > -1) {
            if (c_aggr_count[t_s_id + jj_c] < _max_aggregation + 1) {
              f_c_row[ii] = jj_c;
              c_aggr_count[t_s_id + jj_c] += 1;
              count++;
            }
          }
          else if (ii_c < -1 && jj_c < -1) {
            f_c_row[ii] = t_c_n_rows;
            f_c_row[jj] = t_c_n_rows;
            c_aggr_count[t_s_id + t_c_n_rows] += 1;
            t_c_n_rows++;
            count++;
          }
        }
        if (count == 0) {
          ag_queue[r_idx_n] = ii;
          ag_queue[r_idx_n+1] = ix0;
          r_idx_n += ag_queue_stride;
          r_n_faces++;
        }
      }
      t_r_e_id = r_idx_n;
    }

    for (int ii = t_s_id; ii < t_e_id; ii++) {
      if (f_c_row[ii] < -1) {
        f_c_row[ii] = t_c_n_rows;
        t_c_n_rows++;
      }
    }
    t_c_scan[t_id] = t_c_n_rows;
  }

  /* after the barrier (:4706): per-block reordering using ag_work */
  for (int t_id = 0; t_id < n_loc_threads; t_id++) {
    int t_s_id = t_range[2*t_id], t_e_id = t_range[2*t_id+1];
    _reorder_f_c_row(t_e_id - t_s_id, t_c_scan[t_id],
                     ag_work + t_s_id, f_c_row + t_s_id);
  }

  /* _scan_f_c_row */
  for (int t_id = 0; t_id < n_loc_threads; t_id++) {
    int n_t_rows = t_c_scan[t_id];
    t_c_scan[t_id] = c_n_rows;
    c_n_rows += n_t_rows;
  }
  for (int t_id = 1; t_id < n_loc_threads; t_id++) {
    int r_shift = t_c_scan[t_id];
    if (r_shift != 0)
      for (int ii = t_range[2*t_id]; ii < t_range[2*t_id+1]; ii++)
        if (f_c_row[ii] > -1) f_c_row[ii] += r_shift;
  }

  free(ag_work); free(c_cardinality); free(c_aggr_count);
  free(t_c_scan); free(t_range);
  return c_n_rows;
}

/* cs_sort_lnum restated only as "ascending order" (result is order-unique) */
static int _cmp_int(const void *a, const void *b)
{ int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }

/* _l_id_binary_search (src/alge/cs_grid.cpp:277-300) */
static int
_bsearch_id(int size, int id, const int *a)
{
  if (size < 1) return -1;
  int start_id = 0, end_id = size - 1;
  int mid_id = (end_id - start_id) / 2;
  while (start_id < end_id) {
    if (a[mid_id] < id) start_id = mid_id + 1;
    else if (a[mid_id] > id) end_id = mid_id - 1;
    else break;
    mid_id = start_id + ((end_id - start_id) / 2);
  }
  if (a[mid_id] != id) mid_id = -1;
  return mid_id;
}

void
ora_grid_free(ora_grid_t *g)
{
  if (g == NULL) return;
  free(g->row_index); free(g->col_id); free(g->d_val); free(g->x_val);
  free(g->coarse_row); free(g->cell_face); free(g->cell_face_sgn);
  free(g->cell_cen); free(g->cell_vol); free(g->face_normal);
  free(g->face_u_normal); free(g->face_surf); free(g->xa0); free(g->xa0ij);
  free(g->face_cell); free(g->xa);
  free(g);
}

static void *_dup(const void *p, size_t s)
{ if (p == NULL) return NULL; void *q = malloc(s > 0 ? s : 1); memcpy(q, p, s); return q; }

/* Level-0 grid from MSR arrays + mesh association
   (cs_grid_create_from_shared, src/alge/cs_grid.cpp:7425-7562). */
ora_grid_t *
ora_grid_create_fine(int n_rows, const int *row_index, const int *col_id,
                     const double *d_val, const double *x_val,
                     int n_faces, const int *c2f, const short *c2f_sgn,
                     const double *cell_cen, const double *cell_vol,
                     const double *face_u_normal, const double *face_surf)
{
  ora_grid_t *g = calloc(1, sizeof(ora_grid_t));
  const int nnz = row_index[n_rows];
  g->level = 0; g->n_rows = n_rows; g->nnz = nnz; g->n_faces = n_faces;
  g->row_index = _dup(row_index, sizeof(int)*(n_rows+1));
  g->col_id = _dup(col_id, sizeof(int)*nnz);
  g->d_val = _dup(d_val, sizeof(double)*n_rows);
  g->x_val = _dup(x_val, sizeof(double)*nnz);
  g->cell_face = _dup(c2f, sizeof(int)*nnz);
  g->cell_face_sgn = _dup(c2f_sgn, sizeof(short)*nnz);
  g->cell_cen = _dup(cell_cen, sizeof(double)*3*n_rows);
  g->cell_vol = _dup(cell_vol, sizeof(double)*n_rows);
  g->face_u_normal = _dup(face_u_normal, sizeof(double)*3*n_faces);
  g->face_surf = _dup(face_surf, sizeof(double)*n_faces);
  g->n_cols_ext = n_rows;
  return g;
}

/* Same for one rank of a distributed matrix: columns >= n_rows are ghost rows
   (cs_grid_create_from_shared reads n_cols_ext from the matrix, :7470-7480);
   cell_cen / cell_vol hold n_cols_ext entries (ghost values already synced). */
ora_grid_t *
ora_grid_create_fine_ext(int n_rows, int n_cols_ext, const int *row_index, const int *col_id,
                         const double *d_val, const double *x_val,
                         int n_faces, const int *c2f, const short *c2f_sgn,
                         const double *cell_cen, const double *cell_vol,
                         const double *face_u_normal, const double *face_surf)
{
  ora_grid_t *g = ora_grid_create_fine(n_rows, row_index, col_id, d_val, x_val, n_faces,
                                       c2f, c2f_sgn, cell_cen, cell_vol, face_u_normal,
                                       face_surf);
  free(g->cell_cen); free(g->cell_vol);
  g->cell_cen = _dup(cell_cen, sizeof(double)*3*n_cols_ext);
  g->cell_vol = _dup(cell_vol, sizeof(double)*n_cols_ext);
  g->n_cols_ext = n_cols_ext;
  return g;
}

/*----------------------------------------------------------------------------
 * One coarsening step on the live path for a symmetric MSR matrix with mesh
 * association ("msr_gather", src/alge/cs_grid.cpp:8019-8022):
 *   aggregation      _automatic_aggregation_dx_msr            (:4282)
 *   adjacency        _coarse_to_fine_adjacency_msr(_serial)   (:2860)
 *   structure        _coarse_msr_struct                       (:3020)
 *   coarse faces     _msr_face_adjacency                      (:3751)
 *   quantities       _coarse_quantities_msr_with_faces_stage_1 (:6226)
 *   values           _compute_coarse_quantities_msr_with_faces (:6841-6897)
 *----------------------------------------------------------------------------*/

/* Structure + stage 1 for a given fine -> coarse map.  f_c_row has f->n_cols_ext
   entries (ghost entries = coarse ghost ids from _coarsen_halo, or absent when
   n_cols_ext == n_rows); it is copied.  Coarse cell_cen / cell_vol get
   c_n_cols_ext entries; the ghost part is to be filled by the caller (halo sync,
   cs_grid.cpp:6561-6581) before ora_grid_coarsen_finish. */
ora_grid_t *
ora_grid_coarsen_begin(const ora_grid_t *f, const int *f_c_row_in, int c_n_rows,
                       int c_n_cols_ext)
{
  const int f_n_rows = f->n_rows;
  const int f_n_cols_ext = (f->n_cols_ext > f_n_rows) ? f->n_cols_ext : f_n_rows;
  const int *f_row_index = f->row_index, *f_col_id = f->col_id;
  const double *f_d_val = f->d_val, *f_x_val = f->x_val;

  ora_grid_t *c = calloc(1, sizeof(ora_grid_t));
  c->level = f->level + 1;
  c->coarse_row = malloc(sizeof(int)*(f_n_cols_ext > 0 ? f_n_cols_ext : 1));
  memcpy(c->coarse_row, f_c_row_in, sizeof(int)*f_n_cols_ext);
  int *f_c_row = c->coarse_row;
  c->n_rows = c_n_rows;
  c->n_cols_ext = c_n_cols_ext;

  /* coarse -> fine adjacency, ascending fine row order (serial variant) */
  int *c_f_row_index = calloc(c_n_rows + 1, sizeof(int));
  int *c_f_row_ids = malloc(sizeof(int)*(f_n_rows > 0 ? f_n_rows : 1));
  for (int ii = 0; ii < f_n_rows; ii++)
    if (f_c_row[ii] > -1) c_f_row_index[f_c_row[ii] + 1] += 1;
  for (int i = 0; i < c_n_rows; i++) c_f_row_index[i+1] += c_f_row_index[i];
  {
    int *shift = calloc(c_n_rows > 0 ? c_n_rows : 1, sizeof(int));
    for (int ii = 0; ii < f_n_rows; ii++) {
      int jj = f_c_row[ii];
      if (jj > -1) c_f_row_ids[c_f_row_index[jj] + shift[jj]++] = ii;
    }
    free(shift);
  }

  /* coarse MSR structure: sorted unique coarse columns per coarse row */
  int *c_row_index = calloc(c_n_rows + 1, sizeof(int));
  int c_nnz_0 = 0;
  for (int ic = 0; ic < c_n_rows; ic++)
    for (int r = c_f_row_index[ic]; r < c_f_row_index[ic+1]; r++) {
      int ii = c_f_row_ids[r];
      c_nnz_0 += f_row_index[ii+1] - f_row_index[ii];
    }
  int *c_col_id = malloc(sizeof(int)*(c_nnz_0 > 0 ? c_nnz_0 : 1));
  int c_nnz = 0;
  for (int ic = 0; ic < c_n_rows; ic++) {
    int *l_buf = c_col_id + c_nnz;
    int b_e = 0;
    for (int r = c_f_row_index[ic]; r < c_f_row_index[ic+1]; r++) {
      int ii = c_f_row_ids[r];
      for (int jj = f_row_index[ii]; jj < f_row_index[ii+1]; jj++) {
        int kk = f_c_row[f_col_id[jj]];
        if (kk > -1 && kk != ic) l_buf[b_e++] = kk;
      }
    }
    qsort(l_buf, b_e, sizeof(int), _cmp_int);
    int r_count = 0;
    if (b_e > 0) {
      r_count = 1;
      for (int jj = 1; jj < b_e; jj++)
        if (l_buf[jj] > l_buf[jj-1]) l_buf[r_count++] = l_buf[jj];
    }
    c_nnz += r_count;
    c_row_index[ic+1] = c_nnz;
  }
  c->nnz = c_nnz;

  /* coarse faces: upper-triangle entries in (row, col) order */
  int *c_cell_face = malloc(sizeof(int)*(c_nnz > 0 ? c_nnz : 1));
  short *c_cell_face_sgn = malloc(sizeof(short)*(c_nnz > 0 ? c_nnz : 1));
  int c_n_faces = 0;
  for (int ic = 0; ic < c_n_rows; ic++)
    for (int idx = c_row_index[ic]; idx < c_row_index[ic+1]; idx++) {
      if (c_col_id[idx] < ic) {
        c_cell_face[idx] = -1; c_cell_face_sgn[idx] = -1;
      }
      else {
        c_cell_face[idx] = c_n_faces++; c_cell_face_sgn[idx] = 1;
      }
    }
  for (int ic = 0; ic < c_n_rows; ic++)
    for (int idx = c_row_index[ic]; idx < c_row_index[ic+1]; idx++)
      if (c_cell_face[idx] == -1) {
        int ia = c_col_id[idx];
        int k = _bsearch_id(c_row_index[ia+1] - c_row_index[ia], ic,
                            c_col_id + c_row_index[ia]);
        c_cell_face[idx] = c_cell_face[c_row_index[ia] + k];
      }
  c->n_faces = c_n_faces;

  /* stage 1 */
  double *c_d_val = malloc(sizeof(double)*(c_n_rows > 0 ? c_n_rows : 1));
  double *c_x_val = malloc(sizeof(double)*(c_nnz > 0 ? c_nnz : 1));
  double *c_cell_cen = calloc(3*(c_n_cols_ext > 0 ? c_n_cols_ext : 1), sizeof(double));
  double *c_cell_vol = calloc(c_n_cols_ext > 0 ? c_n_cols_ext : 1, sizeof(double));
  double *c_face_normal = calloc(3*(c_n_faces > 0 ? c_n_faces : 1), sizeof(double));
  double *c_xa0 = calloc(c_n_faces > 0 ? c_n_faces : 1, sizeof(double));
  double *c_xa0ij = calloc(3*(c_n_faces > 0 ? c_n_faces : 1), sizeof(double));

  const double *f_cell_cen = f->cell_cen, *f_cell_vol = f->cell_vol;
  const int *f_cell_face = f->cell_face;
  const short *f_cell_face_sgn = f->cell_face_sgn;

  for (int ii_c = 0; ii_c < c_n_rows; ii_c++) {
    const int c_s_idx = c_row_index[ii_c];
    const int n_c_cols = c_row_index[ii_c+1] - c_s_idx;
    double *_cen = c_cell_cen + 3*ii_c;
    _cen[0] = 0.; _cen[1] = 0.; _cen[2] = 0.;
    c_cell_vol[ii_c] = 0;
    c_d_val[ii_c] = 0;

    for (int r_idx = c_f_row_index[ii_c]; r_idx < c_f_row_index[ii_c+1]; r_idx++) {
      int ii = c_f_row_ids[r_idx];
      for (int k = 0; k < 3; k++)
        _cen[k] += f_cell_cen[ii*3+k]*f_cell_vol[ii];
      c_cell_vol[ii_c] += f_cell_vol[ii];
      c_d_val[ii_c] += f_d_val[ii];

      for (int f_idx = f_row_index[ii]; f_idx < f_row_index[ii+1]; f_idx++) {
        int jj = f_col_id[f_idx];
        int jj_c = f_c_row[jj];
        c_d_val[ii_c] += f_x_val[f_idx];

        if (jj_c > -1 && jj_c != ii_c) {
          int k = _bsearch_id(n_c_cols, jj_c, c_col_id + c_s_idx);
          int c_idx = c_s_idx + k;
          if (jj_c > ii_c) {
            int sgn = f_cell_face_sgn[f_idx];
            int f_face_id = f_cell_face[f_idx];
            int c_face_id = c_cell_face[c_idx];
            if (f->level == 0) {
              double sgn_surf = sgn * f->face_surf[f_face_id];
              for (int q = 0; q < 3; q++)
                c_face_normal[c_face_id*3+q]
                  += sgn_surf * f->face_u_normal[f_face_id*3+q];
              const double f_xa0_f = f_x_val[f_idx];
              c_xa0[c_face_id] += f_xa0_f;
              for (int q = 0; q < 3; q++)
                c_xa0ij[c_face_id*3+q]
                  += f_xa0_f * (f_cell_cen[jj*3+q] - f_cell_cen[ii*3+q]);
            }
            else {
              for (int q = 0; q < 3; q++)
                c_face_normal[c_face_id*3+q]
                  += sgn * f->face_normal[f_face_id*3+q];
              c_xa0[c_face_id] += f->xa0[f_face_id];
              for (int q = 0; q < 3; q++)
                c_xa0ij[c_face_id*3+q] += sgn * f->xa0ij[f_face_id*3+q];
            }
          }
        }
      }
    }
    for (int k = 0; k < 3; k++) _cen[k] /= c_cell_vol[ii_c];
  }

  free(c_f_row_index); free(c_f_row_ids);

  c->row_index = c_row_index;
  c->col_id = realloc(c_col_id, sizeof(int)*(c_nnz > 0 ? c_nnz : 1));
  c->d_val = c_d_val; c->x_val = c_x_val;
  c->cell_face = c_cell_face; c->cell_face_sgn = c_cell_face_sgn;
  c->cell_cen = c_cell_cen; c->cell_vol = c_cell_vol;
  c->face_normal = c_face_normal; c->xa0 = c_xa0; c->xa0ij = c_xa0ij;
  return c;
}

/* Stage 2: coarse extra-diagonal values and final diagonal. */
void
ora_grid_coarsen_finish(ora_grid_t *c, double relax_param)
{
  const int c_n_rows = c->n_rows;
  const int *c_row_index = c->row_index, *c_col_id = c->col_id;
  const int *c_cell_face = c->cell_face;
  const short *c_cell_face_sgn = c->cell_face_sgn;
  const double *c_cell_cen = c->cell_cen, *c_face_normal = c->face_normal;
  const double *c_xa0 = c->xa0, *c_xa0ij = c->xa0ij;
  double *c_d_val = c->d_val, *c_x_val = c->x_val;

  /* values: P0 restriction relaxed with P1 (:6848-6895) */
  for (int ic = 0; ic < c_n_rows; ic++) {
    const double *cen_ic = c_cell_cen + 3*ic;
    for (int r_idx = c_row_index[ic]; r_idx < c_row_index[ic+1]; r_idx++) {
      const int jc = c_col_id[r_idx];
      const int c_face_id = c_cell_face[r_idx];
      const double sgn = (double)c_cell_face_sgn[r_idx];
      const double *cen_jc = c_cell_cen + 3*jc;
      const double *nrm = c_face_normal + 3*c_face_id;
      const double *xij = c_xa0ij + 3*c_face_id;

      const double dsigjg
        = sgn * (  (cen_jc[0] - cen_ic[0])*nrm[0]
                 + (cen_jc[1] - cen_ic[1])*nrm[1]
                 + (cen_jc[2] - cen_ic[2])*nrm[2]);

      double c_x_val_c = c_xa0[c_face_id];

      if (fabs(dsigjg) > 1e-12) {
        const double dsxaij = xij[0]*nrm[0] + xij[1]*nrm[1] + xij[2]*nrm[2];
        const double agij = dsxaij/dsigjg;
        if (agij >= c_x_val_c && agij <= 0.)
          c_x_val_c = agij;
      }
      c_x_val_c =         relax_param *c_x_val_c
                  + (1. - relax_param)*c_xa0[c_face_id];
      c_x_val[r_idx] = c_x_val_c;
      c_d_val[ic] -= c_x_val_c;
    }
  }

}

/* Galerkin coarse matrix (relaxation parameter <= 0): _compute_coarse_quantities_msr, scalar
   case (src/alge/cs_grid.cpp:5969-6090) -- d_c = sum of the fine diagonals and of the
   fine extra-diagonal terms interior to the aggregate, x_c[k] += fine terms between two
   aggregates (binary search of the sorted coarse column), fine rows in ascending order.
   No mesh quantities are needed or produced. */
static ora_grid_t *
_grid_coarsen_galerkin(const ora_grid_t *f, const int *f_c_row_in, int c_n_rows)
{
  const int f_n_rows = f->n_rows;
  const int *f_row_index = f->row_index, *f_col_id = f->col_id;
  const double *f_d_val = f->d_val, *f_x_val = f->x_val;

  ora_grid_t *c = calloc(1, sizeof(ora_grid_t));
  c->level = f->level + 1;
  c->coarse_row = malloc(sizeof(int)*(f_n_rows > 0 ? f_n_rows : 1));
  memcpy(c->coarse_row, f_c_row_in, sizeof(int)*f_n_rows);
  const int *f_c_row = c->coarse_row;
  c->n_rows = c_n_rows;
  c->n_cols_ext = c_n_rows;

  int *c_f_row_index = calloc(c_n_rows + 1, sizeof(int));
  int *c_f_row_ids = malloc(sizeof(int)*(f_n_rows > 0 ? f_n_rows : 1));
  for (int ii = 0; ii < f_n_rows; ii++)
    if (f_c_row[ii] > -1) c_f_row_index[f_c_row[ii] + 1] += 1;
  for (int i = 0; i < c_n_rows; i++) c_f_row_index[i+1] += c_f_row_index[i];
  {
    int *shift = calloc(c_n_rows > 0 ? c_n_rows : 1, sizeof(int));
    for (int ii = 0; ii < f_n_rows; ii++) {
      int jj = f_c_row[ii];
      if (jj > -1) c_f_row_ids[c_f_row_index[jj] + shift[jj]++] = ii;
    }
    free(shift);
  }

  /* coarse MSR structure (_coarse_msr_struct, :3020-3200) */
  int *c_row_index = calloc(c_n_rows + 1, sizeof(int));
  int c_nnz_0 = 0;
  for (int ic = 0; ic < c_n_rows; ic++)
    for (int r = c_f_row_index[ic]; r < c_f_row_index[ic+1]; r++) {
      int ii = c_f_row_ids[r];
      c_nnz_0 += f_row_index[ii+1] - f_row_index[ii];
    }
  int *c_col_id = malloc(sizeof(int)*(c_nnz_0 > 0 ? c_nnz_0 : 1));
  int c_nnz = 0;
  for (int ic = 0; ic < c_n_rows; ic++) {
    int *l_buf = c_col_id + c_nnz;
    int b_e = 0;
    for (int r = c_f_row_index[ic]; r < c_f_row_index[ic+1]; r++) {
      int ii = c_f_row_ids[r];
      for (int jj = f_row_index[ii]; jj < f_row_index[ii+1]; jj++) {
        int kk = f_c_row[f_col_id[jj]];
        if (kk > -1 && kk != ic) l_buf[b_e++] = kk;
      }
    }
    qsort(l_buf, b_e, sizeof(int), _cmp_int);
    int r_count = 0;
    if (b_e > 0) {
      r_count = 1;
      for (int jj = 1; jj < b_e; jj++)
        if (l_buf[jj] > l_buf[jj-1]) l_buf[r_count++] = l_buf[jj];
    }
    c_nnz += r_count;
    c_row_index[ic+1] = c_nnz;
  }

  double *c_d_val = calloc(c_n_rows > 0 ? c_n_rows : 1, sizeof(double));
  double *c_x_val = calloc(c_nnz > 0 ? c_nnz : 1, sizeof(double));
  for (int ic = 0; ic < c_n_rows; ic++) {
    const int s_id = c_row_index[ic];
    const int n_cols = c_row_index[ic+1] - s_id;
    double *row_x_vals = c_x_val + s_id;
    c_d_val[ic] = 0;
    for (int r_idx = c_f_row_index[ic]; r_idx < c_f_row_index[ic+1]; r_idx++) {
      const int ii = c_f_row_ids[r_idx];
      c_d_val[ic] += f_d_val[ii];
      for (int jj = f_row_index[ii]; jj < f_row_index[ii+1]; jj++) {
        const int jc = f_c_row[f_col_id[jj]];
        if (jc > -1) {
          if (ic != jc) {
            int lo = 0, hi = n_cols - 1, k = -1;
            while (lo <= hi) {
              int mid = (lo + hi) / 2;
              if (c_col_id[s_id + mid] < jc) lo = mid + 1;
              else if (c_col_id[s_id + mid] > jc) hi = mid - 1;
              else { k = mid; break; }
            }
            row_x_vals[k] += f_x_val[jj];
          }
          else
            c_d_val[ic] += f_x_val[jj];
        }
      }
    }
  }

  free(c_f_row_index); free(c_f_row_ids);
  c->nnz = c_nnz;
  c->row_index = c_row_index;
  c->col_id = c_col_id;
  c->d_val = c_d_val; c->x_val = c_x_val;
  return c;
}

/* the same for a given fine -> coarse map (cs_oracle_native.c: level 0 of a matrix mapped
   from native arrays is aggregated over its face list even when the values are Galerkin) */
ora_grid_t *
ora_grid_coarsen_galerkin_map(const ora_grid_t *f, const int *f_c_row, int c_n_rows)
{
  return _grid_coarsen_galerkin(f, f_c_row, c_n_rows);
}

/*----------------------------------------------------------------------------
 * Pairwise aggregation (Notay), CS_GRID_COARSENING_SPD_PW:
 *   _pairwise_msr (src/alge/cs_grid.cpp:1744-2020), single-thread form, with the bucket
 *   lists _graph_m_ptr_remove_m / _insert_m (:1544-1620); called with beta = 0.25 and
 *   dd_threshold = 5 at level 0, -1 below (_automatic_aggregation_pw_msr :2041-2110).
 * Rows are taken by increasing number of remaining (strong) neighbours a_m; row i is paired
 * with the free neighbour of most negative coefficient if it is below -beta * max_k(-a_ik).
 *----------------------------------------------------------------------------*/

typedef struct {
  int  m_min, m_max;
  int *m_head, *next, *prev;
} ora_graph_m_ptr_t;

static void
_gm_remove(ora_graph_m_ptr_t *s, int _m, int elt_id)
{
  int s_p = s->prev[elt_id];
  int s_n = s->next[elt_id];
  if (s_p > -1) {
    s->next[s_p] = s_n;
    if (s_n > -1) s->prev[s_n] = s_p;
  }
  else {
    s->m_head[_m] = s_n;
    if (s_n > -1) s->prev[s_n] = -1;
    if (s->m_head[_m] < 0) {
      if (_m == s->m_min) {
        while (s->m_min < s->m_max && s->m_head[s->m_min] < 0) s->m_min++;
      }
      else if (_m == s->m_max) {
        s->m_max -= 1;
        while (s->m_max > s->m_min) {
          if (s->m_head[s->m_max] > -1) break;
          else s->m_max--;
        }
      }
    }
  }
  s->prev[elt_id] = -1;
  s->next[elt_id] = -1;
}

static void
_gm_insert(ora_graph_m_ptr_t *s, int _m, int elt_id)
{
  int s_p = s->m_head[_m];
  if (s_p > -1) {
    s->prev[s_p] = elt_id;
    s->next[elt_id] = s_p;
  }
  else {
    if (_m < s->m_min) s->m_min = _m;
    else if (_m > s->m_max) {
      s->m_max = _m;
      if (s->m_head[s->m_min] < 0) s->m_min = _m;
    }
  }
  s->m_head[_m] = elt_id;
}

int
ora_pairwise_msr(int level, int f_n_rows, const int *row_index, const int *col_id,
                 const double *d_val, const double *x_val, int *f_c_row)
{
  const double beta = 0.25;
  const double dd_threshold = (level == 0) ? 5. : -1.;
  const int t_s_id = 0, t_e_id = f_n_rows;
  int c_n_rows = 0;

  short *a_m = malloc(sizeof(short)*(f_n_rows > 0 ? f_n_rows : 1));
  double *a_max = malloc(sizeof(double)*(f_n_rows > 0 ? f_n_rows : 1));
  int *s_next = malloc(sizeof(int)*(2*(size_t)f_n_rows + 1));

  int t_c_n_rows = 0;
  for (int ii = t_s_id; ii < t_e_id; ii++) f_c_row[ii] = -2;

  int m_max = -1;
  if (dd_threshold > 0) {
    for (int ii = t_s_id; ii < t_e_id; ii++) {
      double sum = 0.0;
      const int s_id = row_index[ii], e_id = row_index[ii+1];
      a_max[ii] = 0.0;
      for (int jj = s_id; jj < e_id; jj++) {
        double xv = x_val[jj];
        sum += fabs(xv);
        if (xv < 0) a_max[ii] = (a_max[ii] > -xv) ? a_max[ii] : -xv;
      }
      if (d_val[ii] > dd_threshold * sum) {
        a_m[ii] = -1;
        f_c_row[ii] = -1;
      }
      else {
        a_m[ii] = 0;
        for (int jj = e_id-1; jj >= s_id; jj--)
          if (col_id[jj] < f_n_rows && x_val[jj] < beta*sum) a_m[ii] += 1;
      }
      if (m_max < a_m[ii]) m_max = a_m[ii];
    }
  }
  else {
    for (int ii = t_s_id; ii < t_e_id; ii++) {
      const int s_id = row_index[ii], e_id = row_index[ii+1];
      a_max[ii] = 0.0;
      for (int jj = s_id; jj < e_id; jj++) {
        double xv = x_val[jj];
        if (xv < 0) a_max[ii] = (a_max[ii] > -xv) ? a_max[ii] : -xv;
      }
      a_m[ii] = 0;
      for (int jj = e_id-1; jj >= s_id; jj--)
        if (col_id[jj] < f_n_rows) a_m[ii] += 1;
      if (m_max < a_m[ii]) m_max = a_m[ii];
    }
  }

  if (m_max >= 0) {
    ora_graph_m_ptr_t s;
    s.m_min = 0;
    s.m_max = m_max;
    s.m_head = malloc(sizeof(int)*(m_max + 1));
    s.next = s_next;
    s.prev = s.next + f_n_rows;
    for (int ii = 0; ii < s.m_max+1; ii++) s.m_head[ii] = -1;
    for (int ii = t_s_id; ii < t_e_id; ii++) { s.next[ii] = -1; s.prev[ii] = -1; }
    for (int ii = t_e_id-1; ii >= t_s_id; ii--) {
      short _m = a_m[ii];
      if (_m >= 0) {
        int prev_head = s.m_head[_m];
        s.m_head[_m] = ii;
        s.next[ii] = prev_head;
        if (prev_head > -1) s.prev[prev_head] = ii;
      }
    }
    while (s.m_min < s.m_max) {
      if (s.m_head[s.m_min] < 0) s.m_min++;
      else break;
    }

    while (s.m_min > -1) {
      int ii = s.m_head[s.m_min];
      int gg[2] = {ii, -1};
      const int s_id = row_index[ii], e_id = row_index[ii+1];
      f_c_row[ii] = t_c_n_rows;
      if (e_id > s_id) {
        int jj = -1;
        double _a_min = HUGE_VAL;
        for (int kk_idx = s_id; kk_idx < e_id; kk_idx++) {
          int kk = col_id[kk_idx];
          if (kk >= t_s_id && kk < t_e_id && f_c_row[kk] == -2) {
            double xv = x_val[kk_idx];
            if (xv < _a_min) { _a_min = xv; jj = kk; }
          }
        }
        if (_a_min >= -beta*a_max[ii])
          jj = -1;
        else {
          f_c_row[jj] = t_c_n_rows;
          gg[1] = jj;
        }
      }
      t_c_n_rows++;

      for (int ip = 0; ip < 2; ip++) {
        int i = gg[ip];
        if (i >= t_s_id && i < t_e_id) {
          int _m = a_m[i];
          if (_m < 0) continue;
          _gm_remove(&s, _m, i);
          a_m[i] = -1;
          for (int k = row_index[i+1]-1; k >= row_index[i]; k--) {
            int j = col_id[k];
            if (j < t_s_id || j >= t_e_id) continue;
            _m = a_m[j];
            if (_m >= 0) {
              _gm_remove(&s, _m, j);
              if (_m > 0) _gm_insert(&s, _m-1, j);
              a_m[j] = _m - 1;
            }
          }
        }
      }

      if (s.m_min >= s.m_max) {
        if (s.m_head[s.m_min] < 0) { s.m_min = -1; s.m_max = -1; }
      }
    }

    for (int ii = t_s_id; ii < t_e_id; ii++)
      if (f_c_row[ii] < -1) f_c_row[ii] = t_c_n_rows++;

    free(s.m_head);
    c_n_rows = t_c_n_rows;
  }

  free(s_next); free(a_max); free(a_m);
  return c_n_rows;
}

/* cs_grid_coarsen with CS_GRID_COARSENING_SPD_PW and relaxation <= 0
   (src/alge/cs_grid.cpp:8143-8150, :8222-8223, :8305-8346): pairs, Galerkin values, and for an
   aggregation limit > 2 a recursion on the intermediate grid (limit / 2) whose map is
   projected back to the fine rows; only the coarsest grid is kept. */
static ora_grid_t *
_grid_coarsen_pw(const ora_grid_t *f, int aggregation_limit)
{
  int *f_c_row = malloc(sizeof(int)*(f->n_rows > 0 ? f->n_rows : 1));
  const int c_n_rows = ora_pairwise_msr(f->level, f->n_rows, f->row_index, f->col_id,
                                        f->d_val, f->x_val, f_c_row);
  ora_grid_t *c = _grid_coarsen_galerkin(f, f_c_row, c_n_rows);
  free(f_c_row);
  if (aggregation_limit > 2) {
    ora_grid_t *cc = _grid_coarsen_pw(c, aggregation_limit / 2);
    for (int ii = 0; ii < f->n_rows; ii++) {
      int ic = c->coarse_row[ii];
      if (ic >= 0) c->coarse_row[ii] = cc->coarse_row[ic];
    }
    free(cc->coarse_row);
    cc->coarse_row = c->coarse_row;
    c->coarse_row = NULL;
    cc->level -= 1;
    ora_grid_free(c);
    return cc;
  }
  return c;
}

ora_mg_t *
ora_mg_setup_pw(ora_grid_t *fine, int aggregation_limit, int n_levels_max, int n_g_rows_min)
{
  ora_mg_t *mg = calloc(1, sizeof(ora_mg_t));
  mg->grids = calloc(n_levels_max > 0 ? n_levels_max : 1, sizeof(ora_grid_t *));
  mg->grids[0] = fine;
  mg->n_levels = 1;
  ora_grid_t *g = fine;
  long n_g_rows = g->n_rows, n_g_rows_prev;
  while (mg->n_levels < n_levels_max) {
    n_g_rows_prev = n_g_rows;
    ora_grid_t *c = _grid_coarsen_pw(g, aggregation_limit);
    n_g_rows = c->n_rows;
    int add = 1;
    if (n_g_rows == 0) add = 0;
    else if (mg->n_levels > 1) {
      if (n_g_rows < n_g_rows_min || n_g_rows > 0.8*n_g_rows_prev) add = 0;
    }
    if (!add) { ora_grid_free(c); break; }
    mg->grids[mg->n_levels++] = c;
    g = c;
  }
  return mg;
}

ora_grid_t *
ora_grid_coarsen(const ora_grid_t *f, int max_aggregation, double relax_param,
                 int n_threads)
{
  const int f_n_rows = f->n_rows;
  const int f_n_cols_ext = (f->n_cols_ext > f_n_rows) ? f->n_cols_ext : f_n_rows;
  int *f_c_row = malloc(sizeof(int)*(f_n_cols_ext > 0 ? f_n_cols_ext : 1));
  for (int i = f_n_rows; i < f_n_cols_ext; i++) f_c_row[i] = -1;
  const int c_n_rows
    = ora_aggregation_dx_msr(f->level, f_n_rows, f->row_index, f->col_id,
                             f->d_val, f->x_val, max_aggregation, n_threads,
                             f_c_row);
  if (relax_param <= 0) {        /* cs_grid_coarsen, src/alge/cs_grid.cpp:8222-8223 */
    ora_grid_t *cg = _grid_coarsen_galerkin(f, f_c_row, c_n_rows);
    free(f_c_row);
    return cg;
  }
  ora_grid_t *c = ora_grid_coarsen_begin(f, f_c_row, c_n_rows, c_n_rows);
  free(f_c_row);
  ora_grid_coarsen_finish(c, relax_param);
  return c;
}

/*----------------------------------------------------------------------------
 * Hierarchy: _setup_hierarchy stop criteria (src/alge/cs_multigrid.cpp:2454-2584)
 *----------------------------------------------------------------------------*/

ora_mg_t *
ora_mg_setup(ora_grid_t *fine, int aggregation_limit, int n_levels_max,
             int n_g_rows_min, double p0p1_relax, int n_threads)
{
  ora_mg_t *mg = calloc(1, sizeof(ora_mg_t));
  mg->grids = calloc(n_levels_max > 0 ? n_levels_max : 1, sizeof(ora_grid_t *));
  mg->grids[0] = fine;
  mg->n_levels = 1;
  ora_grid_t *g = fine;
  long n_g_rows = g->n_rows, n_g_rows_prev;

  while (1) {
    n_g_rows_prev = n_g_rows;
    if (mg->n_levels >= n_levels_max) break;
    ora_grid_t *c = ora_grid_coarsen(g, aggregation_limit, p0p1_relax, n_threads);
    n_g_rows = c->n_rows;
    int add = 1;
    if (n_g_rows == 0) add = 0;
    else if (mg->n_levels > 1) {
      if (n_g_rows < n_g_rows_min || n_g_rows > (0.8 * n_g_rows_prev)) add = 0;
    }
    if (!add) { ora_grid_free(c); break; }
    mg->grids[mg->n_levels++] = c;
    g = c;
  }
  return mg;
}

void
ora_mg_free(ora_mg_t *mg, int free_fine)
{
  if (mg == NULL) return;
  for (int i = (free_fine ? 0 : 1); i < mg->n_levels; i++)
    ora_grid_free(mg->grids[i]);
  free(mg->grids);
  free(mg);
}

/* cs_grid_restrict_row_var, CS_DISPATCH_SUM_SIMPLE (src/alge/cs_grid.cpp:8987-9040) */
void
ora_restrict(int f_n_rows, int c_n_rows, const int *coarse_row,
             const double *f_var, double *c_var)
{
  for (int ii = 0; ii < c_n_rows; ii++) c_var[ii] = 0.;
  for (int ii = 0; ii < f_n_rows; ii++) {
    int i = coarse_row[ii];
    if (i >= 0) c_var[i] += f_var[ii];
  }
}

/* cs_grid_prolong_row_var, increment mode (src/alge/cs_grid.cpp:9156-9166) */
void
ora_prolong_inc(int f_n_rows, const int *coarse_row,
                const double *c_var, double *f_var)
{
  for (int ii = 0; ii < f_n_rows; ii++) {
    int ic = coarse_row[ii];
    if (ic >= 0) f_var[ii] += c_var[ic];
  }
}

/*----------------------------------------------------------------------------
 * Multigrid solve, V-cycle in solver mode with cs_sles_it PCG(+Jacobi)
 * smoothers and coarse solver:
 *   cs_multigrid_solve  (src/alge/cs_multigrid.cpp:5362-5588)
 *   _multigrid_v_cycle  (:3127-3542),  _convergence_test (:2927-3035)
 * opt: n_max_cycles, n_max_iter[3] (descent, ascent, coarse),
 *      precision_mult[3].
 * out[0]=n_equiv_iter out[1]=residual out[2]=n_cycles
 * cycle_res (size >= n_max_cycles+1): fine residual norm at each cycle start.
 *----------------------------------------------------------------------------*/

int
ora_mg_solve(const ora_mg_t *mg, const ora_mg_opt_t *opt,
             double precision, double r_norm,
             const double *rhs, double *vx, int use_vx_ini,
             double *out, double *cycle_res)
{
  const int n_levels = mg->n_levels;
  const int coarsest_level = n_levels - 1;
  const ora_grid_t *g0 = mg->grids[0];
  const double denom_n_g_rows_0 = 1.0 / g0->n_rows;

  /* rhs_vx work arrays (cs_multigrid.cpp:1888-1935) */
  double **rhs_l = calloc(n_levels, sizeof(double *));
  double **vx_l = calloc(n_levels, sizeof(double *));
  int wr_size = g0->n_rows;
  for (int l = 1; l < n_levels; l++) {
    int n = mg->grids[l]->n_rows;
    rhs_l[l] = calloc(n > 0 ? n : 1, sizeof(double));
    vx_l[l] = calloc(n > 0 ? n : 1, sizeof(double));
    if (n > wr_size) wr_size = n;
  }
  vx_l[0] = vx;
  double *wr = malloc(sizeof(double)*(wr_size > 0 ? wr_size : 1));

  int cvg = ORA_ITERATING;
  int n_equiv_iter = 0;
  double initial_residual = -1, residual = -1;
  int n_cycles = 0;
  double sout[4];

  for (n_cycles = 0; cvg == ORA_ITERATING; n_cycles++) {
    const int cycle_id = n_cycles + 1;
    const int vx_ini_nonzero = (n_cycles > 0) ? 1 : use_vx_ini;
    int end_cycle = 0;
    int c_cvg = ORA_ITERATING;
    int level;

    /* Descent */
    for (level = 0; level < coarsest_level; level++) {
      const ora_grid_t *f = mg->grids[level];
      const ora_grid_t *c = mg->grids[level+1];
      const double *rhs_lv = (level == 0) ? rhs : rhs_l[level];
      double *vx_lv = vx_l[level];
      const int ini = (level == 0) ? vx_ini_nonzero : 0;

      c_cvg = ora_pcg_sr_p0(f->n_rows, f->row_index, f->col_id, f->d_val,
                            f->x_val, opt->n_max_iter[0],
                            precision*opt->precision_mult[0], r_norm,
                            rhs_lv, vx_lv, ini, sout, NULL, 0);
      int n_iter = (int)sout[0];
      if (level == 0 && cycle_id == 1) initial_residual = sout[2];
      if (c_cvg < ORA_BREAKDOWN) { end_cycle = 1; break; }

      ora_spmv_msr(f->n_rows, f->row_index, f->col_id, f->d_val, f->x_val,
                   0, vx_lv, wr);
      for (int ii = 0; ii < f->n_rows; ii++) wr[ii] = rhs_lv[ii] - wr[ii];

      if (level == 0) {
        double r2 = ora_dot_xx(f->n_rows, wr);
        residual = sqrt(r2);
        /* initial_residual is passed by value to _convergence_test (:2937,
           :2967-2968): cycle 1 compares against itself, later cycles against
           the level-0 smoother's initial residual of cycle 1 (:3278-3279) */
        const double ir = (cycle_id == 1) ? residual : initial_residual;
        if (cycle_res != NULL) cycle_res[cycle_id - 1] = residual;
        if (residual < precision*r_norm) cvg = ORA_CONVERGED;
        else if (cycle_id > opt->n_max_cycles) cvg = ORA_MAX_ITERATION;
        else if (residual > ir * 10000.0 && residual > 100.)
          cvg = ORA_DIVERGED;
        else if (isnan(residual) || isinf(residual)) cvg = ORA_DIVERGED;
        if (cvg != 0) {
          c_cvg = cvg; end_cycle = 1;
          n_equiv_iter += n_iter * f->n_rows * denom_n_g_rows_0;
          break;
        }
      }
      n_equiv_iter += n_iter * f->n_rows * denom_n_g_rows_0;

      ora_restrict(f->n_rows, c->n_rows, c->coarse_row, wr, rhs_l[level+1]);
    }

    if (!end_cycle) {
      const ora_grid_t *c = mg->grids[coarsest_level];
      const double *rhs_lv = (coarsest_level == 0) ? rhs : rhs_l[coarsest_level];
      c_cvg = ora_pcg_sr_p0(c->n_rows, c->row_index, c->col_id, c->d_val,
                            c->x_val, opt->n_max_iter[2],
                            precision*opt->precision_mult[2], r_norm,
                            rhs_lv, vx_l[coarsest_level], 0, sout, NULL, 0);
      n_equiv_iter += (int)sout[0] * c->n_rows * denom_n_g_rows_0;
      if (c_cvg < ORA_BREAKDOWN) end_cycle = 1;
    }

    if (!end_cycle) {
      for (level = coarsest_level - 1; level > -1; level--) {
        const ora_grid_t *f = mg->grids[level];
        const ora_grid_t *c = mg->grids[level+1];
        ora_prolong_inc(f->n_rows, c->coarse_row, vx_l[level+1], vx_l[level]);
        if (level > 0) {
          c_cvg = ora_pcg_sr_p0(f->n_rows, f->row_index, f->col_id, f->d_val,
                                f->x_val, opt->n_max_iter[1],
                                precision*opt->precision_mult[1], r_norm,
                                rhs_l[level], vx_l[level], 1, sout, NULL, 0);
          n_equiv_iter += (int)sout[0] * f->n_rows * denom_n_g_rows_0;
          if (c_cvg < ORA_BREAKDOWN) break;
        }
      }
    }
  }

  out[0] = n_equiv_iter;
  out[1] = residual;
  out[2] = n_cycles;
  for (int l = 1; l < n_levels; l++) { free(rhs_l[l]); free(vx_l[l]); }
  free(rhs_l); free(vx_l); free(wr);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Multigrid smoothers without convergence test (fixed sweep counts):
 *   _jacobi          (src/alge/cs_multigrid_smoother.cpp:143-226)
 *   _relaxed_jacobi  (:244-373): w = 2/3; RJ2 {1.7319, 0.5695};
 *                    RJ3 {2.2473, 0.8573, 0.5296}, w_k = w[n_iter % M]
 *   _l1_jacobi       (:804-905) with l1_inv = 1/(d_i + sum_j |x_ij|)
 *                    (_setup_ad_l1_inv, src/alge/cs_sles_it_priv.cpp:231-262)
 * Sweep count rounded up to a multiple of M for the scheduled variants
 * (cs_multigrid_smoother_setup, :1888-1898).
 * type: cs_sles_it_type_t value.  out[0] = number of sweeps done.
 * Returns ORA_MAX_ITERATION like the reference, -100 for an unknown type.
 *----------------------------------------------------------------------------*/

int
ora_smoother(int type, int n_rows, const int *row_index, const int *col_id,
             const double *d_val, const double *x_val, int n_max_iter,
             const double *rhs, double *vx, int use_vx_ini, double *out)
{
  const int n = n_rows;
  const int gs = (   type == ORA_SLES_P_GAUSS_SEIDEL || type == ORA_SLES_P_SYM_GAUSS_SEIDEL
                  || type == ORA_SLES_TS_F_GAUSS_SEIDEL || type == ORA_SLES_TS_B_GAUSS_SEIDEL);
  if (   type != ORA_SLES_JACOBI && type != ORA_SLES_L1_JACOBI
      && type != ORA_SLES_R_JACOBI && type != ORA_SLES_RJ2 && type != ORA_SLES_RJ3 && !gs)
    return -100;

  if (gs) {
    /* Process-local Gauss-Seidel family, ONE thread (the reference's OpenMP loops over the
       rows update vx in place, so with several threads the result depends on their timing;
       the sequential sweep is the defined one):
         _p_gauss_seidel_msr      cs_multigrid_smoother.cpp:1053-1160  forward sweeps
         _p_sym_gauss_seidel_msr  :1179-1370   forward then backward sweep per iteration
         _ts_f_gauss_seidel_msr   :1389-1500   ONE forward sweep over the lower part
         _ts_b_gauss_seidel_msr   :1519-1630   ONE backward sweep over the upper part */
    int n_it = 0;
    if (type == ORA_SLES_TS_F_GAUSS_SEIDEL || type == ORA_SLES_TS_B_GAUSS_SEIDEL) {
      if (!use_vx_ini) for (int ii = 0; ii < n; ii++) vx[ii] = 0.0;
      if (type == ORA_SLES_TS_F_GAUSS_SEIDEL)
        for (int ii = 0; ii < n; ii++) {
          double vx0 = rhs[ii];
          for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++) {
            if (col_id[jj] > ii) break;
            vx0 -= (x_val[jj]*vx[col_id[jj]]);
          }
          vx0 *= 1.0 / d_val[ii];
          vx[ii] = vx0;
        }
      else
        for (int ii = n - 1; ii > -1; ii--) {
          double vx0 = rhs[ii];
          for (int jj = row_index[ii+1] - 1; jj > row_index[ii] - 1; jj--) {
            if (col_id[jj] < ii) break;
            vx0 -= (x_val[jj]*vx[col_id[jj]]);
          }
          vx0 *= 1.0 / d_val[ii];
          vx[ii] = vx0;
        }
      n_it = 1;
    }
    else {
      for (n_it = 0; n_it < n_max_iter; n_it++) {
        if (n_it == 0 && !use_vx_ini)
          for (int ii = 0; ii < n; ii++) vx[ii] = 0;
        for (int ii = 0; ii < n; ii++) {
          double vx0 = rhs[ii];
          for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++)
            vx0 -= (x_val[jj]*vx[col_id[jj]]);
          vx[ii] = vx0 * (1.0 / d_val[ii]);
        }
        if (type == ORA_SLES_P_SYM_GAUSS_SEIDEL)
          for (int ii = n - 1; ii > -1; ii--) {
            double vx0 = rhs[ii];
            for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++)
              vx0 -= (x_val[jj]*vx[col_id[jj]]);
            vx0 *= 1.0 / d_val[ii];
            vx[ii] = vx0;
          }
      }
    }
    if (out != NULL) out[0] = n_it;
    return ORA_MAX_ITERATION;
  }

  double *tmp = malloc(sizeof(double)*(2*(size_t)n + 2));
  double *ad_inv = tmp + n;

  if (type == ORA_SLES_L1_JACOBI) {
    for (int i = 0; i < n; i++) {
      double s1 = d_val[i];
      for (int j = row_index[i]; j < row_index[i+1]; j++)
        s1 += fabs(x_val[j]);
      ad_inv[i] = 1.0 / s1;
    }
  }
  else
    for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[i];

  int n_iter = 0;

  if (type == ORA_SLES_JACOBI) {
    double *rk = tmp;
    int iter_ini = 0;
    if (!use_vx_ini) {
      for (int ii = 0; ii < n; ii++) vx[ii] = rhs[ii]*ad_inv[ii];
      iter_ini = 1;
    }
    for (n_iter = iter_ini; n_iter < n_max_iter; n_iter++) {
      for (int ii = 0; ii < n; ii++) rk[ii] = vx[ii];
      ora_spmv_msr(n, row_index, col_id, d_val, x_val, 1, rk, vx);
      for (int ii = 0; ii < n; ii++) vx[ii] = (rhs[ii]-vx[ii])*ad_inv[ii];
    }
  }

  else if (type == ORA_SLES_L1_JACOBI) {
    double *vx_n = vx, *vx_np = tmp;
    int iter_ini = 0;
    if (!use_vx_ini) {
      for (int ii = 0; ii < n; ii++) vx_n[ii] = rhs[ii]*ad_inv[ii];
      iter_ini = 1;
    }
    for (n_iter = iter_ini; n_iter < n_max_iter; n_iter++) {
      double *t = vx_n; vx_n = vx_np; vx_np = t;
      ora_spmv_msr(n, row_index, col_id, d_val, x_val, 0, vx_np, vx_n);
      for (int ii = 0; ii < n; ii++)
        vx_n[ii] = vx_np[ii] + (rhs[ii]-vx_n[ii]) * ad_inv[ii];
    }
    if (vx_n != vx)
      for (int ii = 0; ii < n; ii++) vx[ii] = vx_n[ii];
  }

  else {
    double w[3] = {2./3., 1.0, 1.0};
    int wk_m = 1;
    if (type == ORA_SLES_RJ2) { w[0] = 1.7319; w[1] = 0.5695; wk_m = 2; }
    if (type == ORA_SLES_RJ3) { w[0] = 2.2473; w[1] = 0.8573; w[2] = 0.5296; wk_m = 3; }
    if (n_max_iter % wk_m) n_max_iter += wk_m - (n_max_iter % wk_m);

    double *vx_n = vx, *vx_np = tmp;
    int iter_ini = 0;
    if (!use_vx_ini) {
      const double wk = w[0];
      for (int ii = 0; ii < n; ii++) vx_n[ii] = wk * rhs[ii]*ad_inv[ii];
      iter_ini = 1;
    }
    for (n_iter = iter_ini; n_iter < n_max_iter; n_iter++) {
      const double wk = w[n_iter % wk_m];
      const double o_m_wk = 1.0 - wk;
      double *t = vx_n; vx_n = vx_np; vx_np = t;
      ora_spmv_msr(n, row_index, col_id, d_val, x_val, 1, vx_np, vx_n);
      for (int ii = 0; ii < n; ii++)
        vx_n[ii] = o_m_wk*vx_np[ii] + wk*(rhs[ii]-vx_n[ii])*ad_inv[ii];
    }
    if (vx_n != vx)
      for (int ii = 0; ii < n; ii++) vx[ii] = vx_n[ii];
  }

  if (out != NULL) out[0] = n_iter;
  free(tmp);
  return ORA_MAX_ITERATION;
}

/* One level's smoother / solver as _multigrid_setup_sles builds it
   (src/alge/cs_multigrid.cpp:2150-2175): precision_mult < 0 and a smoother-capable
   type -> cs_multigrid_smoother_*; otherwise cs_sles_it (PCG + Jacobi here). */

static int
_mg_level_solve(int type, double precision_mult, const ora_grid_t *g, int n_max_iter,
                double precision, double r_norm, const double *rhs, double *vx,
                int use_vx_ini, double *sout)
{
  if (precision_mult < 0 && type != ORA_SLES_PCG) {
    sout[2] = HUGE_VAL;
    return ora_smoother(type, g->n_rows, g->row_index, g->col_id, g->d_val, g->x_val,
                        n_max_iter, rhs, vx, use_vx_ini, sout);
  }
  if (type == ORA_SLES_PCG)
    return ora_pcg_sr_p0(g->n_rows, g->row_index, g->col_id, g->d_val, g->x_val,
                         n_max_iter, precision*precision_mult, r_norm,
                         rhs, vx, use_vx_ini, sout, NULL, 0);
  return -100;
}

/*----------------------------------------------------------------------------
 * Single V-cycle without convergence tests, the form used when the multigrid
 * is a preconditioner (n_max_cycles == 1):
 *   _multigrid_v_cycle_pc (src/alge/cs_multigrid.cpp:3569-4000), reached from
 *   cs_multigrid_solve (:5484-5499) and _multigrid_pc_apply (:1500-1600).
 * Descent: smoother from 0 (or vx at level 0 when use_vx_ini), residual,
 * restriction; coarsest solve from 0; ascent: prolongation (increment), then
 * the ascent smoother on EVERY level including 0.
 * out[0] = n_equiv_iter.  Returns ORA_ITERATING like the reference, or the
 * failing level solver's state (< ORA_BREAKDOWN).
 *----------------------------------------------------------------------------*/

int
ora_mg_cycle_pc(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt,
                double precision, double r_norm,
                const double *rhs, double *vx, int use_vx_ini, double *out)
{
  const int n_levels = mg->n_levels;
  const int coarsest_level = n_levels - 1;
  const ora_grid_t *g0 = mg->grids[0];
  const double denom_n_g_rows_0 = 1.0 / g0->n_rows;

  double **rhs_l = calloc(n_levels, sizeof(double *));
  double **vx_l = calloc(n_levels, sizeof(double *));
  int wr_size = g0->n_rows;
  for (int l = 1; l < n_levels; l++) {
    int n = mg->grids[l]->n_rows;
    rhs_l[l] = calloc(n > 0 ? n : 1, sizeof(double));
    vx_l[l] = calloc(n > 0 ? n : 1, sizeof(double));
    if (n > wr_size) wr_size = n;
  }
  vx_l[0] = vx;
  double *wr = malloc(sizeof(double)*(wr_size > 0 ? wr_size : 1));

  int n_equiv_iter = 0;
  int c_cvg = ORA_ITERATING;
  int end_cycle = 0;
  int ret = ORA_ITERATING;
  double sout[4];
  int level;

  for (level = 0; level < coarsest_level; level++) {
    const ora_grid_t *f = mg->grids[level];
    const ora_grid_t *c = mg->grids[level+1];
    const double *rhs_lv = (level == 0) ? rhs : rhs_l[level];
    double *vx_lv = vx_l[level];
    const int ini = (level == 0) ? use_vx_ini : 0;

    c_cvg = _mg_level_solve(opt->type[0], opt->precision_mult[0], f, opt->n_max_iter[0],
                            precision, r_norm, rhs_lv, vx_lv, ini, sout);
    if (c_cvg < ORA_BREAKDOWN) { end_cycle = 1; ret = c_cvg; break; }

    ora_spmv_msr(f->n_rows, f->row_index, f->col_id, f->d_val, f->x_val, 0, vx_lv, wr);
    for (int ii = 0; ii < f->n_rows; ii++) wr[ii] = rhs_lv[ii] - wr[ii];

    n_equiv_iter += (int)sout[0] * f->n_rows * denom_n_g_rows_0;
    ora_restrict(f->n_rows, c->n_rows, c->coarse_row, wr, rhs_l[level+1]);
  }

  if (!end_cycle) {
    const ora_grid_t *c = mg->grids[coarsest_level];
    const double *rhs_lv = (coarsest_level == 0) ? rhs : rhs_l[coarsest_level];
    c_cvg = _mg_level_solve(opt->type[2], opt->precision_mult[2], c, opt->n_max_iter[2],
                            precision, r_norm, rhs_lv, vx_l[coarsest_level], 0, sout);
    n_equiv_iter += (int)sout[0] * c->n_rows * denom_n_g_rows_0;
    if (c_cvg < ORA_BREAKDOWN) { end_cycle = 1; ret = c_cvg; }
  }

  if (!end_cycle) {
    for (level = coarsest_level - 1; level > -1; level--) {
      const ora_grid_t *f = mg->grids[level];
      const ora_grid_t *c = mg->grids[level+1];
      const double *rhs_lv = (level == 0) ? rhs : rhs_l[level];
      ora_prolong_inc(f->n_rows, c->coarse_row, vx_l[level+1], vx_l[level]);
      c_cvg = _mg_level_solve(opt->type[1], opt->precision_mult[1], f, opt->n_max_iter[1],
                              precision, r_norm, rhs_lv, vx_l[level], 1, sout);
      n_equiv_iter += (int)sout[0] * f->n_rows * denom_n_g_rows_0;
      if (c_cvg < ORA_BREAKDOWN) { ret = c_cvg; break; }
    }
  }

  if (out != NULL) out[0] = n_equiv_iter;
  for (int l = 1; l < n_levels; l++) { free(rhs_l[l]); free(vx_l[l]); }
  free(rhs_l); free(vx_l); free(wr);
  return ret;
}

/*----------------------------------------------------------------------------
 * General preconditioners of cs_sles_it (cs_sles_pc_t apply functions):
 *   none   cs_sles_pc_none            (x_out = x_in)
 *   poly   _sles_pc_poly_apply_poly   (src/alge/cs_sles_pc.cpp:549-610): x = r/d, then
 *          degree times x = (r - E.x)/d; cs_sles_it_create maps poly_degree 1 -> 1,
 *          >= 2 -> 2 (cs_sles_it.cpp:4683-4690)
 *   mg     _multigrid_pc_apply        (src/alge/cs_multigrid.cpp:1500-1600): one V-cycle,
 *          identity for a one-level hierarchy; tolerance = the outer solver's
 *          (_multigrid_pc_tolerance_t :1473-1481)
 *----------------------------------------------------------------------------*/

static int
_k_cycle(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt, double trsh, int level,
         double precision, double r_norm, const double *rhs, double *vx, int *n_equiv_iter);

typedef struct {
  int     kind;            /* 0 none, 1 poly, 2 multigrid V-cycle, 3 multigrid K-cycle */
  double  trsh;            /* K-cycle threshold */
  int     n_rows;
  const int *row_index, *col_id;
  const double *d_val, *x_val;
  int     poly_degree;
  double *ad_inv, *w;      /* poly */
  const ora_mg_t *mg;
  const ora_mg_pc_opt_t *opt;
  double  precision, r_norm;
} ora_pc_t;

static void
_pc_init_matrix(ora_pc_t *pc, int n, const int *ri, const int *ci, const double *d,
                const double *xv)
{
  pc->n_rows = n; pc->row_index = ri; pc->col_id = ci; pc->d_val = d; pc->x_val = xv;
  pc->ad_inv = NULL; pc->w = NULL;
  if (pc->kind == 1) {
    pc->ad_inv = malloc(sizeof(double)*(2*(size_t)n + 1));
    pc->w = pc->ad_inv + n;
    for (int i = 0; i < n; i++) pc->ad_inv[i] = 1.0 / d[i];
  }
}

static void
_pc_apply(const ora_pc_t *pc, const double *x_in, double *x_out)
{
  const int n = pc->n_rows;
  if (pc->kind == 0 || (pc->kind >= 2 && pc->mg->n_levels == 1)) {
    for (int ii = 0; ii < n; ii++) x_out[ii] = x_in[ii];
  }
  else if (pc->kind == 1) {
    for (int ii = 0; ii < n; ii++) x_out[ii] = x_in[ii] * pc->ad_inv[ii];
    for (int deg_id = 1; deg_id <= pc->poly_degree; deg_id++) {
      ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 1, x_out, pc->w);
      for (int ii = 0; ii < n; ii++)
        x_out[ii] = (x_in[ii] - pc->w[ii]) * pc->ad_inv[ii];
    }
  }
  else if (pc->kind == 3) {
    int n_equiv = 0;
    _k_cycle(pc->mg, pc->opt, pc->trsh, 0, pc->precision, pc->r_norm, x_in, x_out, &n_equiv);
  }
  else
    ora_mg_cycle_pc(pc->mg, pc->opt, pc->precision, pc->r_norm, x_in, x_out, 0, NULL);
}

/* _flexible_conjugate_gradient (src/alge/cs_sles_it.cpp:832-985) with pc_apply = pc.
   out[0]=n_iter out[1]=residual out[2]=initial_residual out[3]=history count */

static int
_fcg_pc(const ora_pc_t *pc, int n_max_iter, double precision, double r_norm,
        const double *rhs, double *vx, int use_vx_ini,
        double *out, double *history, int history_size)
{
  const int n = pc->n_rows;
  double *rk = malloc(sizeof(double)*((size_t)n*5 + 1));
  double *vk = rk + n, *wk = rk + 2*n, *dk = rk + 3*n, *qk = rk + 4*n;

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  double residual = -1;

  if (use_vx_ini) {
    ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, vx, rk);
    for (int ii = 0; ii < n; ii++) rk[ii] = rhs[ii] - rk[ii];
  }
  else {
    for (int ii = 0; ii < n; ii++) { vx[ii] = 0; rk[ii] = rhs[ii]; }
  }
  for (int ii = 0; ii < n; ii++) qk[ii] = 0;

  double rho_km1 = 0;

  while (cvg == ORA_ITERATING) {
    _pc_apply(pc, rk, vk);
    ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, vk, wk);

    double alpha_k, beta_k, gamma_k;
    ora_dot_xy_yz(n, wk, vk, qk, &beta_k, &gamma_k);
    ora_dot_xx_xy(n, rk, vk, &residual, &alpha_k);
    residual = sqrt(residual);

    if (n_iter > 0)
      cvg = _cvg_test(&cv, n_iter, residual);
    else {
      cv.initial_residual = residual;
      if (history != NULL && cv.history_n < history_size)
        history[cv.history_n++] = residual;
    }
    if (cvg != ORA_ITERATING) break;

    if (n_iter > 0) {
      const double gk_rk1 = (fabs(rho_km1) > DBL_MIN) ? gamma_k / rho_km1 : 0.;
      const double rho_k = beta_k - gamma_k * gk_rk1;
      const double ak_rk = (fabs(rho_k) > DBL_MIN) ? alpha_k / rho_k : 0.;
      for (int ii = 0; ii < n; ii++) {
        dk[ii] = vk[ii] - gk_rk1 * dk[ii];
        vx[ii] = vx[ii] + ak_rk * dk[ii];
      }
      for (int ii = 0; ii < n; ii++) {
        qk[ii] = wk[ii] - gk_rk1 * qk[ii];
        rk[ii] = rk[ii] - ak_rk * qk[ii];
      }
      rho_km1 = rho_k;
    }
    else {
      const double rho_k = beta_k;
      const double ak_rk = (fabs(rho_k) > DBL_MIN) ? alpha_k / rho_k : 0.;
      for (int ii = 0; ii < n; ii++) {
        dk[ii] = vk[ii];
        vx[ii] = vx[ii] + ak_rk * vk[ii];
      }
      for (int ii = 0; ii < n; ii++) {
        qk[ii] = wk[ii];
        rk[ii] = rk[ii] - ak_rk * wk[ii];
      }
      rho_km1 = rho_k;
    }
    n_iter += 1;
  }

  out[0] = n_iter; out[1] = residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/* _conjugate_gradient_sr (src/alge/cs_sles_it.cpp:1229-1415): single-reduction PCG with a
   general preconditioner; residual sign r = A.x - b.  The first iteration has its own
   residual norm and convergence test (:1330-1345).
   out[0]=n_iter out[1]=residual out[2]=initial_residual out[3]=history count */

static int
_pcg_sr_pc(const ora_pc_t *pc, int n_max_iter, double precision, double r_norm,
           const double *rhs, double *vx, int use_vx_ini,
           double *out, double *history, int history_size)
{
  const int n = pc->n_rows;
  double *rk = malloc(sizeof(double)*((size_t)n*5 + 1));
  double *dk = rk + n, *gk = rk + 2*n, *zk = rk + 3*n, *sk = rk + 4*n;

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  const int update_residual = (precision > 0.) ? 1 : 0;

  if (use_vx_ini) {
    ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, vx, rk);
    for (int ii = 0; ii < n; ii++) rk[ii] -= rhs[ii];
  }
  else {
    for (int ii = 0; ii < n; ii++) { vx[ii] = 0; rk[ii] = -rhs[ii]; }
  }

  _pc_apply(pc, rk, dk);
  ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, dk, zk);

  double ro_0 = 0, ro_1 = 0;
  double rk_gkm1 = 0, residual = -1, rk_gk = 0, gk_sk = 0;

  if (update_residual) {
    ora_dot_xx_xy_yz(n, rk, dk, zk, &residual, &ro_0, &ro_1);
    residual = sqrt(residual);
  }
  else
    ora_dot_xy_yz(n, rk, dk, zk, &ro_0, &ro_1);

  cv.initial_residual = residual;
  cvg = _cvg_test(&cv, n_iter, residual);

  if (cvg == ORA_ITERATING) {
    n_iter = 1;
    const double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    const double alpha = - ro_0 * d_ro_1;
    rk_gkm1 = ro_0;
    for (int ii = 0; ii < n; ii++) {
      vx[ii] += (alpha * dk[ii]);
      rk[ii] += (alpha * zk[ii]);
    }
    if (update_residual)
      residual = sqrt(ora_dot_xx(n, rk));
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  while (cvg == ORA_ITERATING) {
    _pc_apply(pc, rk, gk);
    ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, gk, sk);

    if (update_residual) {
      ora_dot_xx_xy_yz(n, rk, gk, sk, &residual, &rk_gk, &gk_sk);
      residual = sqrt(residual);
    }
    else
      ora_dot_xy_yz(n, rk, gk, sk, &rk_gk, &gk_sk);

    if (n_iter > 1)
      cvg = _cvg_test(&cv, n_iter, residual);
    if (cvg != ORA_ITERATING)
      break;

    n_iter += 1;

    const double beta = (fabs(rk_gkm1) > DBL_MIN) ? rk_gk / rk_gkm1 : 0.;
    rk_gkm1 = rk_gk;
    ro_1 = gk_sk - beta*beta*ro_1;
    ro_0 = rk_gk;
    const double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    const double alpha = - ro_0 * d_ro_1;

    for (int ii = 0; ii < n; ii++) {
      dk[ii] = gk[ii] + (beta * dk[ii]);
      vx[ii] += alpha * dk[ii];
      zk[ii] = sk[ii] + (beta * zk[ii]);
      rk[ii] += alpha * zk[ii];
    }
  }

  out[0] = cv.n_iterations; out[1] = cv.residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

static int
_ipcg_pc(const ora_pc_t *pc, int n_max_iter, double precision, double r_norm,
         const double *rhs, double *vx, int use_vx_ini,
         double *out, double *history, int history_size);

/* FCG (solver 1), IPCG (2) or single-reduction PCG (0) preconditioned by one multigrid
   V-cycle -- FCG + cs_multigrid_pc is the reference's production pressure solver. */

int
ora_it_mgpc(int solver, const ora_mg_t *mg, const ora_mg_pc_opt_t *opt,
            int n_max_iter, double precision, double r_norm,
            const double *rhs, double *vx, int use_vx_ini,
            double *out, double *history, int history_size)
{
  const ora_grid_t *g = mg->grids[0];
  ora_pc_t pc;
  memset(&pc, 0, sizeof(pc));
  pc.kind = 2; pc.mg = mg; pc.opt = opt; pc.precision = precision; pc.r_norm = r_norm;
  _pc_init_matrix(&pc, g->n_rows, g->row_index, g->col_id, g->d_val, g->x_val);
  if (solver == 2)
    return _ipcg_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                    out, history, history_size);
  if (solver == 1)
    return _fcg_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                   out, history, history_size);
  return _pcg_sr_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                    out, history, history_size);
}

int
ora_fcg_mgpc(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt,
             int n_max_iter, double precision, double r_norm,
             const double *rhs, double *vx, int use_vx_ini,
             double *out, double *history, int history_size)
{
  return ora_it_mgpc(1, mg, opt, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                     out, history, history_size);
}

/* FCG (solver 1) or single-reduction PCG (solver 0) with the polynomial preconditioner
   (poly_degree >= 1), or with none (poly_degree < 0; FCG only -- unpreconditioned PCG is
   ora_pcg_sr_npc, a different function in the reference). */

int
ora_it_poly(int solver, int n_rows, const int *row_index, const int *col_id,
            const double *d_val, const double *x_val, int poly_degree,
            int n_max_iter, double precision, double r_norm,
            const double *rhs, double *vx, int use_vx_ini,
            double *out, double *history, int history_size)
{
  ora_pc_t pc;
  memset(&pc, 0, sizeof(pc));
  pc.kind = (poly_degree < 0) ? 0 : 1;
  pc.poly_degree = (poly_degree > 2) ? 2 : poly_degree;
  _pc_init_matrix(&pc, n_rows, row_index, col_id, d_val, x_val);
  int cvg;
  if (solver == 2)
    cvg = _ipcg_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                   out, history, history_size);
  else if (solver == 1)
    cvg = _fcg_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                  out, history, history_size);
  else
    cvg = _pcg_sr_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                     out, history, history_size);
  free(pc.ad_inv);
  return cvg;
}

/*----------------------------------------------------------------------------
 * 3x3 block Jacobi: _block_3_jacobi (src/alge/cs_sles_it.cpp:2284-2440) with
 *   ad_inv = cs_math_matrix_gauss_inverse<3> of each diagonal block (Gauss-Jordan with
 *            row pivoting by pointer swap, src/base/cs_math.h:540-586; called from
 *            _setup_ad_block_nn_inv, src/alge/cs_sles_it_priv.cpp:174-222)
 *   x = ad_inv.(rhs - E.x)  (_mat_c_m_b_33, src/alge/cs_sles_it_priv.h:272-289)
 *   residual^2 = sum_rows (D.(x_new - x_old))^2, tested every sweep.
 *----------------------------------------------------------------------------*/

void
ora_gauss_inverse_33(const double a[9], double b[9])
{
  double t[3][6];
  double *p_a[3];
  for (int i = 0; i < 3; i++) {
    p_a[i] = t[i];
    for (int j = 0; j < 3; j++) {
      p_a[i][j] = a[i*3 + j];
      p_a[i][j+3] = 0;
    }
    p_a[i][i+3] = 1;
  }
  for (int i = 0; i < 3; i++) {
    for (int j = i+1; j < 3; j++) {
      if (fabs(p_a[i][i]) < fabs(p_a[j][i])) {
        double *tmp = p_a[j]; p_a[j] = p_a[i]; p_a[i] = tmp;
      }
    }
    for (int j = 0; j < 3; j++) {
      if (j != i) {
        double s = p_a[j][i] / p_a[i][i];
        for (int k = 0; k < 6; k++)
          p_a[j][k] -= p_a[i][k] * s;
      }
    }
  }
  for (int i = 0; i < 3; i++) {
    double s = 1. / p_a[i][i];
    for (int j = 0; j < 3; j++)
      b[i*3 + j] = p_a[i][3+j] * s;
  }
}

static void
_ora_mat_c_m_b_33(const double *mat, double *x, const double *b, const double *c)
{
  double aux[3];
  for (int ii = 0; ii < 3; ii++) aux[ii] = (c[ii] - b[ii]);
  for (int ii = 0; ii < 3; ii++)
    x[ii] = mat[3*ii + 0]*aux[0] + mat[3*ii + 1]*aux[1] + mat[3*ii + 2]*aux[2];
}

int
ora_block3_jacobi(int n_rows, const int *row_index, const int *col_id,
                  const double *d_val, const double *x_val,
                  int n_max_iter, double precision, double r_norm,
                  const double *rhs, double *vx, int use_vx_ini,
                  double *out, double *history, int history_size)
{
  const int n_blocks = n_rows;
  const size_t n = (size_t)n_rows*3;
  double *rk = malloc(sizeof(double)*(2*n + 9*(size_t)n_blocks + 1));
  double *vxx = rk + n, *ad_inv = rk + 2*n;
  const double *ad = d_val;
  for (int i = 0; i < n_blocks; i++)
    ora_gauss_inverse_33(d_val + 9*(size_t)i, ad_inv + 9*(size_t)i);

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  double residual = -1, res2;

  if (!use_vx_ini) {
    n_iter += 1;
    res2 = 0.0;
    for (int ii = 0; ii < n_blocks; ii++) {
      for (int jj = 0; jj < 3; jj++) { rk[3*ii + jj] = 0; vxx[3*ii + jj] = 0; }
      _ora_mat_c_m_b_33(ad_inv + 9*ii, vx + 3*ii, vxx + 3*ii, rhs + 3*ii);
      for (int jj = 0; jj < 3; jj++) {
        double r = 0.0;
        for (int kk = 0; kk < 3; kk++)
          r += ad[ii*9 + jj*3 + kk] * (vx[ii*3 + kk] - rk[ii*3 + kk]);
        res2 += (r*r);
      }
    }
    residual = sqrt(res2);
    if (n_iter == 1) cv.initial_residual = residual;
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  while (cvg == ORA_ITERATING) {
    n_iter += 1;
    ora_spmv_msr_b3(n_rows, row_index, col_id, d_val, x_val, 1, vx, vxx);
    res2 = 0.0;
    for (int ii = 0; ii < n_blocks; ii++) {
      for (int kk = 0; kk < 3; kk++) rk[ii*3 + kk] = vx[ii*3 + kk];
      _ora_mat_c_m_b_33(ad_inv + 9*ii, vx + 3*ii, vxx + 3*ii, rhs + 3*ii);
      for (int jj = 0; jj < 3; jj++) {
        double r = 0.0;
        for (int kk = 0; kk < 3; kk++)
          r += ad[ii*9 + jj*3 + kk] * (vx[ii*3 + kk] - rk[ii*3 + kk]);
        res2 += (r*r);
      }
    }
    residual = sqrt(res2);
    if (n_iter == 1) cv.initial_residual = residual;
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  out[0] = cv.n_iterations; out[1] = cv.residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Preconditioned GCR with restart: _gcr (src/alge/cs_sles_it.cpp:3272-3480), Jacobi
 * preconditioner (jacobi != 0: pc_apply = x*ad_inv, point Jacobi for block matrices) or none.
 * It is also the solver cs_sles_it_solve falls back to on a device (:5383-5385).
 * restart = c->restart_interval (20 by default, :4699).
 *----------------------------------------------------------------------------*/

int
ora_gcr(int n_rows, int db, const int *row_index, const int *col_id,
        const double *d_val, const double *x_val, int jacobi, int restart,
        int n_max_iter, double precision, double r_norm,
        const double *rhs, double *vx, int use_vx_ini,
        double *out, double *history, int history_size)
{
  const int n = n_rows*db;
  const size_t wa = (size_t)n;
  const int n_k = restart;
  double *rk = malloc(sizeof(double)*(wa*(2 + 2*(size_t)n_k) + 1));
  double *zk = rk + wa, *ck = rk + wa*(1 + (size_t)n_k), *ad_inv = rk + wa*(1 + 2*(size_t)n_k);
  double *alpha = malloc(sizeof(double)*n_k);
  double *gkj = malloc(sizeof(double)*((n_k + 1)*n_k/2));
  double *gkj_inv = malloc(sizeof(double)*((n_k + 1)*n_k/2));
  for (int i = 0; i < n; i++) ad_inv[i] = 1.0 / d_val[(size_t)i*db + (i % db)];

#define ORA_SPMV_G(xx, yy) \
  do { if (db == 1) ora_spmv_msr(n_rows, row_index, col_id, d_val, x_val, 0, xx, yy); \
       else ora_spmv_msr_bn(n_rows, db, row_index, col_id, d_val, x_val, 0, xx, yy); } while (0)

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1, history, history_size, 0};
  int cvg = ORA_ITERATING;
  unsigned n_restart = 0;
  int have_vx = use_vx_ini;

  while (cvg == ORA_ITERATING) {
    int n_c_iter = 0;
    if (have_vx) {
      ORA_SPMV_G(vx, rk);
      for (int ii = 0; ii < n; ii++) rk[ii] -= rhs[ii];
    }
    else {
      for (int ii = 0; ii < n; ii++) { vx[ii] = 0; rk[ii] = -rhs[ii]; }
      have_vx = 1;
    }
    double residual = sqrt(ora_dot_xx(n, rk));
    if (n_restart == 0) cv.initial_residual = residual;

    while (cvg == ORA_ITERATING && n_c_iter < n_k) {
      double *zk_n = zk + n_c_iter*wa, *ck_n = ck + n_c_iter*wa;
      if (jacobi) for (int ii = 0; ii < n; ii++) zk_n[ii] = rk[ii]*ad_inv[ii];
      else for (int ii = 0; ii < n; ii++) zk_n[ii] = rk[ii];
      ORA_SPMV_G(zk_n, ck_n);
      for (int jj = 0; jj < n_c_iter; jj++) {
        double *ck_j = ck + jj*wa;
        int ii_jn = (n_c_iter + 1)*n_c_iter/2 + jj;
        gkj[ii_jn] = ora_dot(n, ck_j, ck_n);
        for (int ii = 0; ii < n; ii++) ck_n[ii] += - gkj[ii_jn] * ck_j[ii];
      }
      const int iter_shift = (n_c_iter+1)*n_c_iter/2 + n_c_iter;
      gkj[iter_shift] = sqrt(ora_dot(n, ck_n, ck_n));
      if (fabs(gkj[iter_shift]) > 0) {
        double scale = 1. / gkj[iter_shift];
        for (int ii = 0; ii < n; ii++) ck_n[ii] *= scale;
        alpha[n_c_iter] = ora_dot(n, ck_n, rk);
      }
      else
        alpha[n_c_iter] = 0.;
      for (int ii = 0; ii < n; ii++) rk[ii] += - alpha[n_c_iter] * ck_n[ii];
      residual = sqrt(ora_dot_xx(n, rk));
      n_c_iter += 1;
      cvg = _cvg_test(&cv, (int)(n_restart*n_k) + n_c_iter, residual);
      if (cvg != ORA_ITERATING) break;
    }

    if (n_c_iter == 1 && !(fabs(alpha[0]) > 0))
      gkj_inv[0] = 1.0;
    else {
      int n_g_inv = (n_k + 1)*n_k/2;
      for (int jj = 0; jj < n_g_inv; jj++) gkj_inv[jj] = 0.0;
      for (int kk = 0; kk < n_c_iter; kk++) {
        for (int ii = 0; ii < kk; ii++)
          for (int jj = 0; jj < kk; jj++)
            gkj_inv[(kk + 1)*kk/2 + ii]
              +=   ((ii <= jj) ? gkj_inv[(jj + 1)*jj/2 + ii] : 0.0)
                 * gkj[(kk + 1)*kk/2 + jj];
        for (int jj = 0; jj < kk; jj++)
          gkj_inv[(kk + 1)*kk/2 + jj] /= - gkj[(kk + 1)*kk/2 + kk];
        gkj_inv[(kk + 1)*kk/2 + kk] = 1.0 / gkj[(kk + 1)*kk/2 + kk];
      }
    }
    for (int kk = 0; kk < n_c_iter; kk++)
      for (int jj = 0; jj <= kk; jj++) {
        const double *zk_j = zk + jj*wa;
        for (int ii = 0; ii < n; ii++)
          vx[ii] -= alpha[kk] * zk_j[ii] * gkj_inv[(kk + 1)*kk/2 + jj];
      }
    n_restart += 1;
  }
#undef ORA_SPMV_G

  out[0] = cv.n_iterations; out[1] = cv.residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk); free(alpha); free(gkj); free(gkj_inv);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Block Jacobi for block sizes other than 3: _block_jacobi (src/alge/cs_sles_it.cpp:2460-2620).
 * Setup (_setup_ad_block_nn_inv, cs_sles_it_priv.cpp:178-222): 6x6 blocks get their
 * Gauss-Jordan inverse (cs_math_matrix_gauss_inverse<6>, cs_math.h:540-586); every other size
 * gets the LU factors of _fact_lu (:118-160) -- which the sweep then applies with _mat_c_m_b
 * (cs_sles_it_priv.h:302-322) as it does an inverse.  Restated as it is.
 *----------------------------------------------------------------------------*/

static void
_ora_gauss_inverse_n(int n, const double *a, double *b)
{
  double t[9][18];
  double *p_a[9];
  for (int i = 0; i < n; i++) {
    p_a[i] = t[i];
    for (int j = 0; j < n; j++) { p_a[i][j] = a[i*n + j]; p_a[i][j+n] = 0; }
    p_a[i][i+n] = 1;
  }
  for (int i = 0; i < n; i++) {
    for (int j = i+1; j < n; j++) {
      if (fabs(p_a[i][i]) < fabs(p_a[j][i])) {
        double *tmp = p_a[j]; p_a[j] = p_a[i]; p_a[i] = tmp;
      }
    }
    for (int j = 0; j < n; j++) {
      if (j != i) {
        double s = p_a[j][i] / p_a[i][i];
        for (int k = 0; k < 2*n; k++) p_a[j][k] -= p_a[i][k] * s;
      }
    }
  }
  for (int i = 0; i < n; i++) {
    double s = 1. / p_a[i][i];
    for (int j = 0; j < n; j++) b[i*n + j] = p_a[i][n+j] * s;
  }
}

static void
_ora_fact_lu(int db_size, const double *_ad, double *_ad_inv)
{
  _ad_inv[0] = _ad[0];
  for (int ii = 1; ii < db_size; ii++) {
    _ad_inv[ii] = _ad[ii];
    _ad_inv[ii*db_size] = _ad[ii*db_size]/_ad[0];
  }
  for (int ii = 1; ii < db_size - 1; ii++) {
    _ad_inv[ii + ii*db_size] = _ad[ii + ii*db_size];
    for (int kk = 0; kk < ii; kk++)
      _ad_inv[ii + ii*db_size] -= _ad_inv[ii*db_size + kk]*_ad_inv[kk*db_size + ii];
    for (int jj = ii + 1; jj < db_size; jj++) {
      _ad_inv[ii*db_size + jj] = _ad[ii*db_size + jj];
      _ad_inv[jj*db_size + ii] = _ad[jj*db_size + ii] / _ad_inv[ii*db_size + ii];
      for (int kk = 0; kk < ii; kk++) {
        _ad_inv[ii*db_size + jj] -= _ad_inv[ii*db_size + kk]*_ad_inv[kk*db_size + jj];
        _ad_inv[jj*db_size + ii] -=  _ad_inv[jj*db_size + kk]*_ad_inv[kk*db_size + ii]
                                    /_ad_inv[ii*db_size + ii];
      }
    }
  }
  _ad_inv[db_size*db_size - 1] = _ad[db_size*db_size - 1];
  for (int kk = 0; kk < db_size - 1; kk++)
    _ad_inv[db_size*db_size - 1] -=  _ad_inv[(db_size-1)*db_size + kk]
                                    *_ad_inv[kk*db_size + db_size - 1];
}

static void
_ora_mat_c_m_b(const double *mat, int db_size, double *x, const double *b, const double *c)
{
  double aux[9];
  for (int ii = 0; ii < db_size; ii++) aux[ii] = (c[ii] - b[ii]);
  for (int ii = 0; ii < db_size; ii++) {
    x[ii] = 0;
    for (int jj = 0; jj < db_size; jj++) x[ii] += aux[jj]*mat[ii*db_size + jj];
  }
}

int
ora_block_jacobi(int n_rows, int db, const int *row_index, const int *col_id,
                 const double *d_val, const double *x_val,
                 int n_max_iter, double precision, double r_norm,
                 const double *rhs, double *vx, int use_vx_ini,
                 double *out, double *history, int history_size)
{
  if (db == 3)
    return ora_block3_jacobi(n_rows, row_index, col_id, d_val, x_val, n_max_iter, precision,
                             r_norm, rhs, vx, use_vx_ini, out, history, history_size);
  const int n_blocks = n_rows, db2 = db*db;
  const size_t n = (size_t)n_rows*db;
  double *rk = malloc(sizeof(double)*(2*n + db2*(size_t)n_blocks + 1));
  double *vxx = rk + n, *ad_inv = rk + 2*n;
  const double *ad = d_val;
  for (int i = 0; i < n_blocks; i++) {
    if (db == 6) _ora_gauss_inverse_n(6, d_val + db2*(size_t)i, ad_inv + db2*(size_t)i);
    else _ora_fact_lu(db, d_val + db2*(size_t)i, ad_inv + db2*(size_t)i);
  }

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg = ORA_ITERATING;
  int n_iter = 0;
  double residual = -1, res2;

  if (!use_vx_ini) {
    n_iter += 1;
    res2 = 0.0;
    for (int ii = 0; ii < n_blocks; ii++) {
      double _vxx[9];
      for (int jj = 0; jj < db; jj++) { rk[db*ii + jj] = 0; _vxx[jj] = 0; }
      _ora_mat_c_m_b(ad_inv + db2*ii, db, vx + db*ii, _vxx, rhs + db*ii);
      for (int jj = 0; jj < db; jj++) {
        double r = 0.0;
        for (int kk = 0; kk < db; kk++)
          r += ad[ii*db2 + jj*db + kk] * (vx[ii*db + kk] - rk[ii*db + kk]);
        res2 += (r*r);
      }
    }
    residual = sqrt(res2);
    if (n_iter == 1) cv.initial_residual = residual;
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  while (cvg == ORA_ITERATING) {
    n_iter += 1;
    ora_spmv_msr_bn(n_rows, db, row_index, col_id, d_val, x_val, 1, vx, vxx);
    res2 = 0.0;
    for (int ii = 0; ii < n_blocks; ii++) {
      for (int kk = 0; kk < db; kk++) rk[ii*db + kk] = vx[ii*db + kk];
      _ora_mat_c_m_b(ad_inv + db2*ii, db, vx + db*ii, vxx + db*ii, rhs + db*ii);
      for (int jj = 0; jj < db; jj++) {
        double r = 0.0;
        for (int kk = 0; kk < db; kk++)
          r += ad[ii*db2 + jj*db + kk] * (vx[ii*db + kk] - rk[ii*db + kk]);
        res2 += (r*r);
      }
    }
    residual = sqrt(res2);
    if (n_iter == 1) cv.initial_residual = residual;
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  out[0] = cv.n_iterations; out[1] = cv.residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}

/*----------------------------------------------------------------------------
 * cs_multigrid_smoother on block-diagonal matrices (fixed sweeps, no residual):
 *   type 3 (Jacobi): _block_3_jacobi (cs_multigrid_smoother.cpp:390-500) for 3x3 blocks,
 *                    _block_jacobi (:507-620) otherwise;
 *   types 14 / 15 / 16 (R_JACOBI, RJ2, RJ3): _block_relaxed_jacobi (:626-790);
 * ad_inv = the block_nn_inverse of cs_multigrid_smoother_setup (:1844-1866 ->
 * _setup_ad_block_nn_inv): Gauss-Jordan inverse for 3x3 and 6x6, LU factors otherwise.
 * out[0] = n_iter.
 *----------------------------------------------------------------------------*/

int
ora_smoother_block(int type, int n_rows, int db, const int *row_index, const int *col_id,
                   const double *d_val, const double *x_val, int n_max_iter,
                   const double *rhs, double *vx, int use_vx_ini, double *out)
{
  const int n_blocks = n_rows, db2 = db*db;
  const size_t n = (size_t)n_rows*db;
  if (type != ORA_SLES_JACOBI && type != ORA_SLES_R_JACOBI && type != ORA_SLES_RJ2
      && type != ORA_SLES_RJ3)
    return -100;
  double *tmp = malloc(sizeof(double)*(2*n + db2*(size_t)n_blocks + 1));
  double *vxx = tmp + n, *ad_inv = tmp + 2*n;
  for (int i = 0; i < n_blocks; i++) {
    if (db == 3) ora_gauss_inverse_33(d_val + 9*(size_t)i, ad_inv + 9*(size_t)i);
    else if (db == 6) _ora_gauss_inverse_n(6, d_val + db2*(size_t)i, ad_inv + db2*(size_t)i);
    else _ora_fact_lu(db, d_val + db2*(size_t)i, ad_inv + db2*(size_t)i);
  }
  int n_iter = 0;

  if (type == ORA_SLES_JACOBI) {
    int iter_ini = 0;
    if (!use_vx_ini) {
      for (int ii = 0; ii < n_blocks; ii++) {
        double _vxx[9];
        for (int jj = 0; jj < db; jj++) _vxx[jj] = 0;
        if (db == 3) _ora_mat_c_m_b_33(ad_inv + 9*ii, vx + 3*ii, _vxx, rhs + 3*ii);
        else _ora_mat_c_m_b(ad_inv + db2*ii, db, vx + db*ii, _vxx, rhs + db*ii);
      }
      iter_ini = 1;
    }
    for (n_iter = iter_ini; n_iter < n_max_iter; n_iter++) {
      ora_spmv_msr_bn(n_rows, db, row_index, col_id, d_val, x_val, 1, vx, vxx);
      for (int ii = 0; ii < n_blocks; ii++) {
        if (db == 3) _ora_mat_c_m_b_33(ad_inv + 9*ii, vx + 3*ii, vxx + 3*ii, rhs + 3*ii);
        else _ora_mat_c_m_b(ad_inv + db2*ii, db, vx + db*ii, vxx + db*ii, rhs + db*ii);
      }
    }
  }
  else {
    int wk_m = 1;
    double w[3] = {2./3., 1.0, 1.0};
    if (type == ORA_SLES_RJ2) { w[0] = 1.7319; w[1] = 0.5695; wk_m = 2; }
    if (type == ORA_SLES_RJ3) { w[0] = 2.2473; w[1] = 0.8573; w[2] = 0.5296; wk_m = 3; }
    if (n_max_iter % wk_m) n_max_iter += wk_m - (n_max_iter % wk_m);   /* setup, :1893-1898 */
    double *vx_n = vx, *vx_np = tmp;
    int iter_ini = 0;
    if (!use_vx_ini) {
      const double wk = w[0];
      for (size_t r = 0; r < n; r++) {
        const size_t ii = r / db, jj = r % db;
        const double *_ad_inv = ad_inv + db2*ii + db*jj, *_rhs = rhs + db*ii;
        double _vx = 0;
        for (int kk = 0; kk < db; kk++) _vx += _rhs[kk] * _ad_inv[kk];
        vx_n[r] = wk * _vx;
      }
      iter_ini = 1;
    }
    for (n_iter = iter_ini; n_iter < n_max_iter; n_iter++) {
      const double wk = w[n_iter % wk_m];
      const double o_m_wk = 1.0 - wk;
      double *t = vx_n; vx_n = vx_np; vx_np = t;
      ora_spmv_msr_bn(n_rows, db, row_index, col_id, d_val, x_val, 1, vx_np, vxx);
      for (size_t r = 0; r < n; r++) {
        const size_t ii = r / db, jj = r % db;
        const double *_ad_inv = ad_inv + db2*ii + db*jj, *_rhs = rhs + db*ii,
                     *_vxx = vxx + db*ii;
        double _vx = 0;
        for (int kk = 0; kk < db; kk++) _vx += _ad_inv[kk] * (_rhs[kk] - _vxx[kk]);
        vx_n[r] = o_m_wk*vx_np[r] + wk*_vx;
      }
    }
    if (vx_n != vx)
      for (size_t r = 0; r < n; r++) vx[r] = vx_n[r];
  }
  if (out != NULL) out[0] = n_iter;
  free(tmp);
  return ORA_MAX_ITERATION;
}

/*----------------------------------------------------------------------------
 * Matrix coefficients of a symmetric scalar diffusion system, straight into MSR:
 *   cs_matrix_compute_coeffs (src/alge/cs_matrix_building.cpp:1948-2283) with iconvp = 0 ->
 *   _sym_coeffs_msr (:136-206): per cell, in the order of its MSR row,
 *     x_k = -thetap * i_visc[face of entry k],  d = rovsdt - sum_k x_k
 *   then the boundary faces in face order, d[cell] += thetap * b_visc * cofbfp
 *   (sequential "simple" sum on the CPU), the relaxation d *= 1/relaxp (:2243-2249) and the
 *   (1 + 1e-7) shift of the diagonal when no Dirichlet condition exists (:2256-2262).
 * cell_i_faces = interior face id of every MSR entry (the matrix's mesh association).
 *----------------------------------------------------------------------------*/

void
ora_matrix_compute_coeffs_sym(int n_rows, const int *row_index, const int *cell_i_faces,
                              int n_b_faces, const int *b_face_cells,
                              int idiffp, int ndircp, double thetap, double relaxp,
                              const double *cofbfp, const double *rovsdt,
                              const double *i_visc, const double *b_visc,
                              double *da, double *ea)
{
  if (idiffp) {
    for (int c_id = 0; c_id < n_rows; c_id++) {
      double _da = rovsdt[c_id];
      for (int cidx = row_index[c_id]; cidx < row_index[c_id+1]; cidx++) {
        const int f_id = cell_i_faces[cidx];
        double _xa = -thetap*i_visc[f_id];
        ea[cidx] = _xa;
        _da -= _xa;
      }
      da[c_id] = _da;
    }
    for (int f_id = 0; f_id < n_b_faces; f_id++) {
      int ii = b_face_cells[f_id];
      da[ii] += thetap*b_visc[f_id]*cofbfp[f_id];
    }
  }
  else {
    for (int c_id = 0; c_id < n_rows; c_id++) {
      da[c_id] = rovsdt[c_id];
      for (int cidx = row_index[c_id]; cidx < row_index[c_id+1]; cidx++)
        ea[cidx] = 0.;
    }
  }
  if (relaxp < 1) {
    double rf = 1. / relaxp;
    for (int i = 0; i < n_rows; i++) da[i] *= rf;
  }
  if (ndircp <= 0) {
    const double epsi_p1 = 1. + 1.e-7;
    for (int c_id = 0; c_id < n_rows; c_id++) da[c_id] *= epsi_p1;
  }
}

/*----------------------------------------------------------------------------
 * Same for a convection-diffusion scalar (iconvp = 1, non-symmetric matrix):
 *   _coeffs_msr<is_thermal> (src/alge/cs_matrix_building.cpp:407-520): upwind convective
 *   fluxes from the face mass flux, sign of the face relative to the cell from the mesh
 *   association; boundary faces add the convective / diffusive boundary terms in face order.
 *   xcpp != NULL (imucpp != 0): convective part multiplied by Cp (temperature).
 * cell_i_faces_sgn: +1 when the cell is the first cell of the face.
 *----------------------------------------------------------------------------*/

void
ora_matrix_compute_coeffs(int n_rows, const int *row_index, const int *col_id,
                          const int *cell_i_faces, const short *cell_i_faces_sgn,
                          int n_b_faces, const int *b_face_cells,
                          int iconvp, int idiffp, int ndircp, double thetap, double relaxp,
                          const double *coefbp, const double *cofbfp, const double *rovsdt,
                          const double *i_massflux, const double *b_massflux,
                          const double *i_visc, const double *b_visc, const double *xcpp,
                          double *da, double *ea)
{
  if (!iconvp) {
    ora_matrix_compute_coeffs_sym(n_rows, row_index, cell_i_faces, n_b_faces, b_face_cells,
                                  idiffp, ndircp, thetap, relaxp, cofbfp, rovsdt, i_visc,
                                  b_visc, da, ea);
    return;
  }
  const int is_thermal = (xcpp != NULL);
  for (int c_id = 0; c_id < n_rows; c_id++) {
    double _da = rovsdt[c_id];
    for (int cidx = row_index[c_id]; cidx < row_index[c_id+1]; cidx++) {
      const int f_id = cell_i_faces[cidx];
      double _i_massflux = i_massflux[f_id];
      if (cell_i_faces_sgn[cidx] > 0) {
        double flui = 0.5*(_i_massflux - fabs(_i_massflux));
        double cpi = (is_thermal) ? xcpp[c_id] : 1.0;
        double xai = thetap*(cpi*flui - idiffp*i_visc[f_id]);
        ea[cidx] = xai;
        _da -= xai + (1.-thetap) * cpi * _i_massflux;
      }
      else {
        double fluj = -0.5*(_i_massflux + fabs(_i_massflux));
        double cpj = (is_thermal) ? xcpp[col_id[cidx]] : 1.0;
        double xaj = thetap*(cpj*fluj - idiffp*i_visc[f_id]);
        ea[cidx] = xaj;
        _da -= xaj - (1.-thetap) * cpj * _i_massflux;
      }
    }
    da[c_id] = _da;
  }
  for (int f_id = 0; f_id < n_b_faces; f_id++) {
    int ii = b_face_cells[f_id];
    double _b_massflux = b_massflux[f_id];
    double flui = 0.5*(_b_massflux - fabs(_b_massflux));
    double cpi = (is_thermal) ? xcpp[ii] : 1.0;
    double bfac =          cpi *(  flui * thetap * (coefbp[f_id] - 1.)
                                 - (1. - thetap) * _b_massflux)
                  + idiffp * thetap * b_visc[f_id] * cofbfp[f_id];
    da[ii] += bfac;
  }
  if (relaxp < 1) {
    double rf = 1. / relaxp;
    for (int i = 0; i < n_rows; i++) da[i] *= rf;
  }
  if (ndircp <= 0) {
    const double epsi_p1 = 1. + 1.e-7;
    for (int c_id = 0; c_id < n_rows; c_id++) da[c_id] *= epsi_p1;
  }
}

/*----------------------------------------------------------------------------
 * K-cycle (Notay): _multigrid_k_cycle (src/alge/cs_multigrid.cpp:4051-4480), the single
 * cycle form used as a preconditioner (n_max_cycles == 1: no fine-level convergence test).
 * At each level: pre-smoothing from zero, residual, restriction; the coarser problem is
 * solved by the coarse solver (next level coarsest) or by one or two recursive K-cycles
 * combined by a 2-step Krylov (FCG-like) acceleration -- the second step is skipped when the
 * first reduced the residual norm below trsh (k_cycle_threshold, 0.25); then prolongation
 * (overwrite form, cs_grid.cpp:9180-9190), new residual, post-smoothing from zero, sum.
 * Dot products as _dot_xy_yz / _dot_xx_yy / _dot_xu_xv_xw (:2792-2901).
 *----------------------------------------------------------------------------*/

static int
_k_cycle(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt, double trsh, int level,
         double precision, double r_norm, const double *rhs, double *vx, int *n_equiv_iter)
{
  const int coarsest_level = mg->n_levels - 1;
  const ora_grid_t *f = mg->grids[level];
  const ora_grid_t *c = mg->grids[level+1];
  const int f_n = f->n_rows, c_n = c->n_rows;
  const double denom_n_g_rows_0 = 1.0 / mg->grids[0]->n_rows;
  double sout[4];
  int cvg = ORA_ITERATING;
  int n_iter = 0;

  double *wf = malloc(sizeof(double)*(4*(size_t)f_n + 1));
  double *rt_lv = wf, *rb_lv = wf + f_n, *z1_lv = wf + 2*(size_t)f_n, *z2_lv = wf + 3*(size_t)f_n;
  double *wc = malloc(sizeof(double)*(6*(size_t)c_n + 1));
  double *vx_lv1 = wc, *rhs_lv1 = wc + c_n, *v_lv1 = wc + 2*(size_t)c_n,
         *rt_lv1 = wc + 3*(size_t)c_n, *vx2_lv1 = wc + 4*(size_t)c_n, *w_lv1 = wc + 5*(size_t)c_n;

  /* Pre-smoothing */
  int c_cvg = _mg_level_solve(opt->type[0], opt->precision_mult[0], f, opt->n_max_iter[0],
                              precision, r_norm, rhs, vx, 0, sout);
  n_iter = (int)sout[0];
  *n_equiv_iter += n_iter * f_n * denom_n_g_rows_0;

  ora_spmv_msr(f_n, f->row_index, f->col_id, f->d_val, f->x_val, 0, vx, rt_lv);
  for (int ii = 0; ii < f_n; ii++) rt_lv[ii] = rhs[ii] - rt_lv[ii];

  if (c_cvg < ORA_BREAKDOWN) { free(wf); free(wc); return c_cvg; }

  ora_restrict(f_n, c_n, c->coarse_row, rt_lv, rhs_lv1);

  if (level == coarsest_level - 1) {
    cvg = _mg_level_solve(opt->type[2], opt->precision_mult[2], c, opt->n_max_iter[2],
                          precision, r_norm, rhs_lv1, vx_lv1, 0, sout);
    n_iter = (int)sout[0];
    *n_equiv_iter += n_iter * c_n * denom_n_g_rows_0;
  }
  else {
    _k_cycle(mg, opt, trsh, level + 1, precision, r_norm, rhs_lv1, vx_lv1, &n_iter);
    ora_spmv_msr(c_n, c->row_index, c->col_id, c->d_val, c->x_val, 0, vx_lv1, v_lv1);

    double rho1 = 0, alpha1 = 0;
    ora_dot_xy_yz(c_n, v_lv1, vx_lv1, rhs_lv1, &rho1, &alpha1);
    const double ar1 = alpha1 / rho1;
    for (int i = 0; i < c_n; i++) rt_lv1[i] = rhs_lv1[i] - ar1 * v_lv1[i];

    double rt_lv1_norm = 1.0, r_lv1_norm = 0.0;
    if (trsh > 0) {
      rt_lv1_norm = ora_dot_xx(c_n, rt_lv1);
      r_lv1_norm = ora_dot_xx(c_n, rhs_lv1);
    }

    if (rt_lv1_norm < trsh * trsh * r_lv1_norm) {
      for (int i = 0; i < c_n; i++) vx_lv1[i] = ar1 * vx_lv1[i];
    }
    else {
      for (int i = 0; i < c_n; i++) w_lv1[i] = 0.0;
      _k_cycle(mg, opt, trsh, level + 1, precision, r_norm, rt_lv1, vx2_lv1, &n_iter);
      ora_spmv_msr(c_n, c->row_index, c->col_id, c->d_val, c->x_val, 0, vx2_lv1, w_lv1);

      double gamma = 0, beta = 0, alpha2 = 0;
      ora_dot_xy_yz(c_n, v_lv1, vx2_lv1, w_lv1, &gamma, &beta);
      alpha2 = ora_dot(c_n, vx2_lv1, rt_lv1);
      const double rho2 = beta - (gamma * gamma) / rho1;
      const double ar2 = alpha2 / rho2;
      const double ar1_ar2 = ar1 - (gamma / rho1) * ar2;
      for (int i = 0; i < c_n; i++)
        vx_lv1[i] = ar1_ar2 * vx_lv1[i] + ar2 * vx2_lv1[i];
    }
  }

  /* Ascent */
  for (int ii = 0; ii < f_n; ii++) z2_lv[ii] = 0.0;
  for (int ii = 0; ii < f_n; ii++) {
    int ic = c->coarse_row[ii];
    z1_lv[ii] = (ic >= 0) ? vx_lv1[ic] : 0;
  }
  ora_spmv_msr(f_n, f->row_index, f->col_id, f->d_val, f->x_val, 0, z1_lv, rb_lv);
  for (int ii = 0; ii < f_n; ii++) rb_lv[ii] = rt_lv[ii] - rb_lv[ii];

  cvg = _mg_level_solve(opt->type[1], opt->precision_mult[1], f, opt->n_max_iter[1],
                        precision, r_norm, rb_lv, z2_lv, 1, sout);
  n_iter = (int)sout[0];
  *n_equiv_iter += n_iter * f_n * denom_n_g_rows_0;

  for (int ii = 0; ii < f_n; ii++) vx[ii] += z1_lv[ii] + z2_lv[ii];

  free(wf); free(wc);
  return cvg;
}

/* One K-cycle from rhs (cs_multigrid_solve, CS_MULTIGRID_K_CYCLE, n_max_cycles 1: :5513-5533).
   A one-level hierarchy cannot cycle: returns -100.  out[0] = n_equiv_iter. */
int
ora_mg_k_cycle(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt, double k_cycle_threshold,
               double precision, double r_norm, const double *rhs, double *vx, double *out)
{
  if (mg->n_levels < 2) return -100;
  int n_equiv_iter = 0;
  int cvg = _k_cycle(mg, opt, k_cycle_threshold, 0, precision, r_norm, rhs, vx, &n_equiv_iter);
  if (out != NULL) out[0] = n_equiv_iter;
  return cvg;
}

/* FCG (solver 1) or single-reduction PCG (0) preconditioned by one K-cycle
   (cs_multigrid_pc_create(CS_MULTIGRID_K_CYCLE), src/alge/cs_multigrid.cpp:1642-1655). */
int
ora_it_kpc(int solver, const ora_mg_t *mg, const ora_mg_pc_opt_t *opt, double k_cycle_threshold,
           int n_max_iter, double precision, double r_norm,
           const double *rhs, double *vx, int use_vx_ini,
           double *out, double *history, int history_size)
{
  const ora_grid_t *g = mg->grids[0];
  ora_pc_t pc;
  memset(&pc, 0, sizeof(pc));
  pc.kind = 3; pc.trsh = k_cycle_threshold;
  pc.mg = mg; pc.opt = opt; pc.precision = precision; pc.r_norm = r_norm;
  _pc_init_matrix(&pc, g->n_rows, g->row_index, g->col_id, g->d_val, g->x_val);
  if (solver == 2)
    return _ipcg_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                    out, history, history_size);
  if (solver == 1)
    return _fcg_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                   out, history, history_size);
  return _pcg_sr_pc(&pc, n_max_iter, precision, r_norm, rhs, vx, use_vx_ini,
                    out, history, history_size);
}

/*----------------------------------------------------------------------------
 * Same for a vector variable (3x3 diagonal blocks, scalar extra-diagonal: CS_MATRIX_BLOCK_D):
 *   cs_matrix_compute_coeffs<3> (src/alge/cs_matrix_building.cpp:2284-2600) with scalar
 *   diffusion (tensorial_diffusion 1, eb_size 1): _sym_coeffs_msr<3> (:250-353) when
 *   iconvp = 0, _coeffs_msr<3> (:560-692) when iconvp = 1.  fimp (n x 3 x 3) may be NULL;
 *   coefbp / cofbfp are 3x3 per boundary face.
 *----------------------------------------------------------------------------*/

void
ora_matrix_compute_coeffs_v3(int n_rows, const int *row_index,
                             const int *cell_i_faces, const short *cell_i_faces_sgn,
                             int n_b_faces, const int *b_face_cells,
                             int iconvp, int idiffp, int ndircp, double thetap, double relaxp,
                             const double *coefbp, const double *cofbfp, const double *fimp,
                             const double *i_massflux, const double *b_massflux,
                             const double *i_visc, const double *b_visc,
                             double *da, double *ea)
{
  const int stride = 3;
  for (int c_id = 0; c_id < n_rows; c_id++) {
    double *_da = da + 9*(size_t)c_id;
    for (int q = 0; q < 9; q++) _da[q] = (fimp != NULL) ? fimp[9*(size_t)c_id + q] : 0.;
    for (int cidx = row_index[c_id]; cidx < row_index[c_id+1]; cidx++) {
      const int f_id = cell_i_faces[cidx];
      if (!iconvp) {
        if (idiffp) {
          double _xa = -thetap*i_visc[f_id];
          ea[cidx] = _xa;
          for (int i = 0; i < stride; i++) _da[i*3+i] -= _xa;
        }
        else
          ea[cidx] = 0.;
      }
      else {
        double _i_massflux = i_massflux[f_id];
        if (cell_i_faces_sgn[cidx] > 0) {
          double flui = 0.5*(_i_massflux - fabs(_i_massflux));
          double xai = thetap*(flui - idiffp*i_visc[f_id]);
          ea[cidx] = xai;
          double ifac = xai + (1.-thetap) * _i_massflux;
          for (int i = 0; i < stride; i++) _da[i*3+i] -= ifac;
        }
        else {
          double fluj = -0.5*(_i_massflux + fabs(_i_massflux));
          double xaj = thetap*(fluj - idiffp*i_visc[f_id]);
          ea[cidx] = xaj;
          double jfac = xaj - (1.-thetap) * _i_massflux;
          for (int i = 0; i < stride; i++) _da[i*3+i] -= jfac;
        }
      }
    }
  }
  if (!iconvp) {
    if (idiffp)
      for (int f_id = 0; f_id < n_b_faces; f_id++) {
        double *_da = da + 9*(size_t)b_face_cells[f_id];
        for (int q = 0; q < 9; q++)
          _da[q] += thetap * b_visc[f_id] * cofbfp[9*(size_t)f_id + q];
      }
  }
  else {
    for (int f_id = 0; f_id < n_b_faces; f_id++) {
      double _b_massflux = b_massflux[f_id];
      double *_da = da + 9*(size_t)b_face_cells[f_id];
      double flui = 0.5*(_b_massflux - fabs(_b_massflux));
      for (int i = 0; i < stride; i++)
        for (int j = 0; j < stride; j++) {
          double d_ij = ((i == j) ? 1. : 0.);
          double diag = d_ij * (thetap * flui + (1. - thetap) * b_massflux[f_id]);
          _da[3*i+j] += - diag + thetap * (  flui * coefbp[9*(size_t)f_id + 3*i + j]
                                           + idiffp * b_visc[f_id] * cofbfp[9*(size_t)f_id + 3*i + j]);
        }
    }
  }
  if (relaxp < 1) {
    double rf = 1. / relaxp;
    for (size_t i = 0; i < 9*(size_t)n_rows; i++) da[i] *= rf;
  }
  if (ndircp <= 0) {
    const double epsi_p1 = 1. + 1.e-7;
    for (int c_id = 0; c_id < n_rows; c_id++)
      for (int i = 0; i < stride; i++)
        da[9*(size_t)c_id + 3*i + i] = epsi_p1 * da[9*(size_t)c_id + 3*i + i];
  }
}

/*----------------------------------------------------------------------------
 * ora_mg_solve with any of the level solvers of _mg_level_solve: the V-cycle in solver
 * mode (n_max_cycles > 1, _multigrid_v_cycle) with cs_multigrid_smoother smoothers
 * (precision_mult < 0).  A smoother reports no initial residual: HUGE_VAL
 * (cs_multigrid.cpp:3270-3276), so the divergence test of the cycle never fires.
 *----------------------------------------------------------------------------*/

int
ora_mg_solve_sm(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt, int n_max_cycles,
             double precision, double r_norm,
             const double *rhs, double *vx, int use_vx_ini,
             double *out, double *cycle_res)
{
  const int n_levels = mg->n_levels;
  const int coarsest_level = n_levels - 1;
  const ora_grid_t *g0 = mg->grids[0];
  const double denom_n_g_rows_0 = 1.0 / g0->n_rows;

  /* rhs_vx work arrays (cs_multigrid.cpp:1888-1935) */
  double **rhs_l = calloc(n_levels, sizeof(double *));
  double **vx_l = calloc(n_levels, sizeof(double *));
  int wr_size = g0->n_rows;
  for (int l = 1; l < n_levels; l++) {
    int n = mg->grids[l]->n_rows;
    rhs_l[l] = calloc(n > 0 ? n : 1, sizeof(double));
    vx_l[l] = calloc(n > 0 ? n : 1, sizeof(double));
    if (n > wr_size) wr_size = n;
  }
  vx_l[0] = vx;
  double *wr = malloc(sizeof(double)*(wr_size > 0 ? wr_size : 1));

  int cvg = ORA_ITERATING;
  int n_equiv_iter = 0;
  double initial_residual = -1, residual = -1;
  int n_cycles = 0;
  double sout[4];

  for (n_cycles = 0; cvg == ORA_ITERATING; n_cycles++) {
    const int cycle_id = n_cycles + 1;
    const int vx_ini_nonzero = (n_cycles > 0) ? 1 : use_vx_ini;
    int end_cycle = 0;
    int c_cvg = ORA_ITERATING;
    int level;

    /* Descent */
    for (level = 0; level < coarsest_level; level++) {
      const ora_grid_t *f = mg->grids[level];
      const ora_grid_t *c = mg->grids[level+1];
      const double *rhs_lv = (level == 0) ? rhs : rhs_l[level];
      double *vx_lv = vx_l[level];
      const int ini = (level == 0) ? vx_ini_nonzero : 0;

      c_cvg = _mg_level_solve(opt->type[0], opt->precision_mult[0], f, opt->n_max_iter[0],
                              precision, r_norm, rhs_lv, vx_lv, ini, sout);
      int n_iter = (int)sout[0];
      if (level == 0 && cycle_id == 1) initial_residual = sout[2];
      if (c_cvg < ORA_BREAKDOWN) { end_cycle = 1; break; }

      ora_spmv_msr(f->n_rows, f->row_index, f->col_id, f->d_val, f->x_val,
                   0, vx_lv, wr);
      for (int ii = 0; ii < f->n_rows; ii++) wr[ii] = rhs_lv[ii] - wr[ii];

      if (level == 0) {
        double r2 = ora_dot_xx(f->n_rows, wr);
        residual = sqrt(r2);
        /* initial_residual is passed by value to _convergence_test (:2937,
           :2967-2968): cycle 1 compares against itself, later cycles against
           the level-0 smoother's initial residual of cycle 1 (:3278-3279) */
        const double ir = (cycle_id == 1) ? residual : initial_residual;
        if (cycle_res != NULL) cycle_res[cycle_id - 1] = residual;
        if (residual < precision*r_norm) cvg = ORA_CONVERGED;
        else if (cycle_id > n_max_cycles) cvg = ORA_MAX_ITERATION;
        else if (residual > ir * 10000.0 && residual > 100.)
          cvg = ORA_DIVERGED;
        else if (isnan(residual) || isinf(residual)) cvg = ORA_DIVERGED;
        if (cvg != 0) {
          c_cvg = cvg; end_cycle = 1;
          n_equiv_iter += n_iter * f->n_rows * denom_n_g_rows_0;
          break;
        }
      }
      n_equiv_iter += n_iter * f->n_rows * denom_n_g_rows_0;

      ora_restrict(f->n_rows, c->n_rows, c->coarse_row, wr, rhs_l[level+1]);
    }

    if (!end_cycle) {
      const ora_grid_t *c = mg->grids[coarsest_level];
      const double *rhs_lv = (coarsest_level == 0) ? rhs : rhs_l[coarsest_level];
      c_cvg = _mg_level_solve(opt->type[2], opt->precision_mult[2], c, opt->n_max_iter[2],
                              precision, r_norm, rhs_lv, vx_l[coarsest_level], 0, sout);
      n_equiv_iter += (int)sout[0] * c->n_rows * denom_n_g_rows_0;
      if (c_cvg < ORA_BREAKDOWN) end_cycle = 1;
    }

    if (!end_cycle) {
      for (level = coarsest_level - 1; level > -1; level--) {
        const ora_grid_t *f = mg->grids[level];
        const ora_grid_t *c = mg->grids[level+1];
        ora_prolong_inc(f->n_rows, c->coarse_row, vx_l[level+1], vx_l[level]);
        if (level > 0) {
          c_cvg = _mg_level_solve(opt->type[1], opt->precision_mult[1], f, opt->n_max_iter[1],
                                  precision, r_norm, rhs_l[level], vx_l[level], 1, sout);
          n_equiv_iter += (int)sout[0] * f->n_rows * denom_n_g_rows_0;
          if (c_cvg < ORA_BREAKDOWN) break;
        }
      }
    }
  }

  out[0] = n_equiv_iter;
  out[1] = residual;
  out[2] = n_cycles;
  for (int l = 1; l < n_levels; l++) { free(rhs_l[l]); free(vx_l[l]); }
  free(rhs_l); free(vx_l); free(wr);
  return cvg;
}

/*----------------------------------------------------------------------------
 * Inexact-preconditioning CG with a general preconditioner: _conjugate_gradient_ip
 * (src/alge/cs_sles_it.cpp:1016-1203) with pc_apply = pc -- the solver the reference pairs
 * with the K-cycle ("has to be used with IPCG (or FCG)", cs_multigrid.cpp:4030-4035).
 *----------------------------------------------------------------------------*/

static int
_ipcg_pc(const ora_pc_t *pc, int n_max_iter, double precision, double r_norm,
         const double *rhs, double *vx, int use_vx_ini,
         double *out, double *history, int history_size)
{
  const int n = pc->n_rows;
  double *rk = malloc(sizeof(double)*((size_t)n*5 + 1));
  double *rkm1 = rk + n, *dk = rk + 2*n, *gk = rk + 3*n, *zk = rk + 4*n;

  ora_cvg_t cv = {precision, r_norm, -1, n_max_iter, 0, -1,
                  history, history_size, 0};
  int cvg;
  int n_iter = 0;
  double ro_0, ro_1, alpha, rk_gk_m1, rkm1_gk, rk_gk, beta, residual;

  if (use_vx_ini) {
    ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, vx, rk);
    for (int ii = 0; ii < n; ii++) rk[ii] -= rhs[ii];
  }
  else {
    for (int ii = 0; ii < n; ii++) { vx[ii] = 0; rk[ii] = -rhs[ii]; }
  }
  _pc_apply(pc, rk, gk);
  for (int ii = 0; ii < n; ii++) dk[ii] = gk[ii];

  ora_dot_xx_xy(n, rk, gk, &residual, &rk_gk_m1);
  residual = sqrt(residual);
  cv.initial_residual = residual;
  cvg = _cvg_test(&cv, n_iter, residual);

  if (cvg == ORA_ITERATING) {
    n_iter = 1;
    ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, dk, zk);
    ora_dot_xy_yz(n, rk, dk, zk, &ro_0, &ro_1);
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    for (int ii = 0; ii < n; ii++) vx[ii] += (alpha * dk[ii]);
    for (int ii = 0; ii < n; ii++) { rkm1[ii] = rk[ii]; rk[ii] += (alpha * zk[ii]); }
    residual = sqrt(ora_dot_xx(n, rk));
    cvg = _cvg_test(&cv, n_iter, residual);
  }

  while (cvg == ORA_ITERATING) {
    _pc_apply(pc, rk, gk);
    ora_dot_xx_xy_yz(n, rk, gk, rkm1, &residual, &rk_gk, &rkm1_gk);
    residual = sqrt(residual);
    if (n_iter > 1)
      cvg = _cvg_test(&cv, n_iter, residual);
    if (cvg != ORA_ITERATING) break;
    n_iter += 1;
    beta = (fabs(rk_gk_m1) > DBL_MIN) ? (rk_gk - rkm1_gk) / rk_gk_m1 : 0.;
    rk_gk_m1 = rk_gk;
    for (int ii = 0; ii < n; ii++) dk[ii] = gk[ii] + (beta * dk[ii]);
    ora_spmv_msr(n, pc->row_index, pc->col_id, pc->d_val, pc->x_val, 0, dk, zk);
    ora_dot_xy_yz(n, rk, dk, zk, &ro_0, &ro_1);
    double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    for (int ii = 0; ii < n; ii++) vx[ii] += (alpha * dk[ii]);
    for (int ii = 0; ii < n; ii++) { rkm1[ii] = rk[ii]; rk[ii] += (alpha * zk[ii]); }
  }

  out[0] = cv.n_iterations; out[1] = cv.residual; out[2] = cv.initial_residual;
  out[3] = cv.history_n;
  free(rk);
  return cvg;
}
