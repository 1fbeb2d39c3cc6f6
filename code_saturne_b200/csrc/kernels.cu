/*============================================================================
 * kernels.cu -- hand-written sm_100a kernels of the linear-solve hot path.
 *
 * All arithmetic is FP64, compiled with -fmad=false so that every per-entry
 * operation (products, in-row sums, vector updates, coarse-matrix assembly)
 * is the same IEEE mul/add sequence as the reference's CPU path; only the
 * order of the global reductions differs (fixed-shape 2-stage trees, so they
 * are deterministic run to run).
 *
 * Everything here is HBM-bandwidth bound: no tensor cores.
 *   - SpMV: "CSR-stream": a CTA streams the contiguous off-diagonal entries
 *     of 256 consecutive rows with fully coalesced loads, stages the products
 *     x_val[k]*x[col_id[k]] in shared memory, then each thread sums its own
 *     row in CSR order (same order as _mat_vec_p_l_msr,
 *     src/alge/cs_matrix_spmv.cpp:2130-2142) and runs a fused epilogue
 *     (residual, PCG initialisation, the 3 PCG dot products ...).
 *   - dots: block tree + "last CTA reduces the partials" (one launch, result
 *     and all Krylov scalars stay on the device: no host sync per iteration).
 *============================================================================*/

#include "csb_internal.h"
#include "persistent.h"

#include <cfloat>
#include <climits>
#include <cstring>
#include <cmath>

namespace csb {

namespace {

constexpr int TPB = 256;           /* threads per CTA for streaming kernels */
constexpr int SB_TPB = 1024;       /* single-CTA solver */

inline int vec_grid(int n)
{
  int g = (n + TPB*4 - 1) / (TPB*4);
  int gmax = ctx().sm_count * 8;
  if (g > gmax) g = gmax;
  if (g > RED_MAX_BLOCKS) g = RED_MAX_BLOCKS;
  if (g < 1) g = 1;
  return g;
}

inline bool post_launch(const char *name)
{
  ctx().launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel %s launch failed: %s", name, cudaGetErrorString(e));
    return false;
  }
  return true;
}

/* Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to
   the (function, device) pair and ranks may live on different devices in one process (one
   host thread each, thread_local Context): what has been granted is remembered per thread,
   never process-wide. */
inline void smem_optin(const void *func, size_t bytes)
{
  if (bytes <= 48*1024) return;
  struct Granted { const void *f; size_t bytes; int device; };
  static thread_local std::vector<Granted> granted;
  const int dev = ctx().device;
  for (Granted &g : granted)
    if (g.f == func && g.device == dev) {
      if (g.bytes >= bytes) return;
      cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      g.bytes = bytes;
      return;
    }
  cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  granted.push_back({func, bytes, dev});
}

/*--------------------------------------------------------------------------
 * Reductions
 *--------------------------------------------------------------------------*/

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

/* Sum Q values over the CTA; result valid in thread 0. s_red: Q*32 doubles. */
template <int Q>
__device__ __forceinline__ void block_sum(double (&v)[Q], double *s_red)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    double w = warp_sum(v[q]);
    if (lane == 0) s_red[q*32 + wid] = w;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      double w = (lane < n_warps) ? s_red[q*32 + lane] : 0.0;
      v[q] = warp_sum(w);
    }
  }
  __syncthreads();
}

/* Grid-wide sum of Q values; the last CTA to arrive reduces the per-CTA
   partials in a fixed order and thread 0 of it calls fin(sums). */
template <int Q, class Fin>
__device__ __forceinline__ void grid_sum_finalize(double (&v)[Q], double *partials,
                                                  unsigned *ticket, Fin fin)
{
  __shared__ double s_red[Q*32];
  __shared__ int s_last;
  block_sum<Q>(v, s_red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < Q; q++)
      partials[q*gridDim.x + blockIdx.x] = v[q];
    __threadfence();
    unsigned t = atomicInc(ticket, gridDim.x - 1);   /* wraps back to 0 */
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double w[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      w[q] = 0.0;
      for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x)
        w[q] += __ldcg(partials + q*gridDim.x + i);
    }
    block_sum<Q>(w, s_red);
    if (threadIdx.x == 0) fin(w);
  }
}

/*--------------------------------------------------------------------------
 * Device-side solver logic
 *--------------------------------------------------------------------------*/

/* _convergence_test, src/alge/cs_sles_it.cpp:243-325 */
__device__ __forceinline__ int it_cvg_test(ItState *st, int n_iter, double residual)
{
  st->residual = residual;
  if (st->n_hist < HIST_MAX) st->hist[st->n_hist++] = residual;
  if (st->precision > 0 && residual < st->precision * st->r_norm)
    return CSB_SLES_CONVERGED;
  else if (n_iter >= st->n_max_iter)
    return CSB_SLES_MAX_ITERATION;
  else if (residual > st->initial_residual * 10000.0 && residual > 100.)
    return CSB_SLES_DIVERGED;
  if (isnan(residual) || isinf(residual))
    return CSB_SLES_DIVERGED;
  return CSB_SLES_ITERATING;
}

/* cs_sles_it.cpp:1529-1548 */
__device__ __forceinline__ void pcg_finalize_first(ItState *st, const double *s)
{
  double residual = sqrt(s[0]);
  st->ro_0 = s[1];
  st->ro_1 = s[2];
  st->initial_residual = residual;
  int cvg = it_cvg_test(st, 0, residual);
  if (cvg == CSB_SLES_ITERATING) {
    st->n_iter = 1;
    double d_ro_1 = (fabs(st->ro_1) > DBL_MIN) ? 1. / st->ro_1 : 0.;
    st->alpha = - st->ro_0 * d_ro_1;
    st->rk_gkm1 = st->ro_0;
  }
  st->cvg = cvg;
}

/* cs_sles_it.cpp:1595-1621 */
__device__ __forceinline__ void pcg_finalize_iter(ItState *st, const double *s)
{
  double residual = sqrt(s[0]);
  const double rk_gk = s[1], gk_sk = s[2];
  int cvg = CSB_SLES_ITERATING;
  if (st->n_iter > 1)
    cvg = it_cvg_test(st, st->n_iter, residual);
  if (cvg == CSB_SLES_ITERATING) {
    st->n_iter += 1;
    double beta = (fabs(st->rk_gkm1) > DBL_MIN) ? rk_gk / st->rk_gkm1 : 0.;
    st->rk_gkm1 = rk_gk;
    st->ro_1 = gk_sk - beta*beta*st->ro_1;
    st->ro_0 = rk_gk;
    double d_ro_1 = (fabs(st->ro_1) > DBL_MIN) ? 1. / st->ro_1 : 0.;
    st->alpha = - st->ro_0 * d_ro_1;
    st->beta = beta;
  }
  st->cvg = cvg;
}

/* Scalar part of one FCG iteration (cs_sles_it.cpp:926-980), s = (rk.rk, rk.vk = alpha_k,
   wk.vk = beta_k, vk.qk = gamma_k).  Leaves gk_rk1 in st->beta, ak_rk in st->alpha,
   rho_k in st->ro_0 and counts the iteration; the update kernel tests n_iter == 1 for the
   first-iteration form. */
__device__ __forceinline__ void fcg_finalize(ItState *st, const double *s)
{
  const double residual = sqrt(s[0]);
  const double alpha_k = s[1], beta_k = s[2], gamma_k = s[3];
  int cvg = CSB_SLES_ITERATING;
  if (st->n_iter > 0)
    cvg = it_cvg_test(st, st->n_iter, residual);
  else {
    st->initial_residual = residual;
    st->residual = residual;
    if (st->n_hist < HIST_MAX) st->hist[st->n_hist++] = residual;
  }
  if (cvg == CSB_SLES_ITERATING) {
    double gk_rk1 = 0., rho_k = beta_k;
    if (st->n_iter > 0) {
      const double rho_km1 = st->ro_0;
      gk_rk1 = (fabs(rho_km1) > DBL_MIN) ? gamma_k / rho_km1 : 0.;
      rho_k = beta_k - gamma_k * gk_rk1;
    }
    st->alpha = (fabs(rho_k) > DBL_MIN) ? alpha_k / rho_k : 0.;
    st->beta = gk_rk1;
    st->ro_0 = rho_k;
    st->n_iter += 1;
  }
  st->cvg = cvg;
}

/* IPCG step length (cs_sles_it.cpp:1122-1127, :1178-1183): s = (rk.rk, rk.dk, dk.zk) */
__device__ __forceinline__ void ipcg_finalize_alpha(ItState *st, const double *s)
{
  st->ro_0 = s[1];
  st->ro_1 = s[2];
  const double d_ro_1 = (fabs(st->ro_1) > DBL_MIN) ? 1. / st->ro_1 : 0.;
  st->alpha = - st->ro_0 * d_ro_1;
}

/* _convergence_test, src/alge/cs_multigrid.cpp:2927-3035 */
__device__ __forceinline__ void mg_finalize_residual(MgState *mg, const double *s)
{
  double residual = sqrt(s[0]);
  mg->residual = residual;
  const int cycle_id = mg->cycle_id;
  if (mg->n_cycles_res < CYCLE_MAX) mg->cycle_res[mg->n_cycles_res++] = residual;
  const double ir = (cycle_id == 1) ? residual : mg->initial_residual;
  int cvg = CSB_SLES_ITERATING;
  if (residual < mg->precision * mg->r_norm)
    cvg = CSB_SLES_CONVERGED;
  else if (cycle_id > mg->n_max_cycles)
    cvg = CSB_SLES_MAX_ITERATION;
  else if (residual > ir * 10000.0 && residual > 100.)
    cvg = CSB_SLES_DIVERGED;
  else if (isnan(residual) || isinf(residual))
    cvg = CSB_SLES_DIVERGED;
  mg->cvg = cvg;
}

/*--------------------------------------------------------------------------
 * SpMV, CSR-stream with fused epilogues
 *--------------------------------------------------------------------------*/

struct SpmvK {
  int           n_rows;
  const int    *row_index;
  const int    *col_id;
  const double *d_val;     /* nullptr: exclude diagonal */
  const double *x_val;
  const double *x;
  double       *y;
  const double *b;
  const double *ad_inv;
  double       *rk;
  double       *dk;
  ItState      *st;
  MgState      *mg;
  int           phase;
  double       *partials;
  unsigned     *ticket;
  int           smooth_mode;
  double        w0, w1;
  const double *qk;
};

template <int EPI>
__device__ __forceinline__ bool spmv_gated_off(const SpmvK &a)
{
  if (EPI == EPI_PCG_INIT || EPI == EPI_PCG_DOT3)
    return (a.st->cvg != CSB_SLES_ITERATING);
  if (EPI == EPI_RESIDUAL || EPI == EPI_RESIDUAL_NORM)
    return (a.mg != nullptr && a.mg->cvg != CSB_SLES_ITERATING);
  if (EPI == EPI_STORE)
    return (a.st != nullptr && a.st->cvg != CSB_SLES_ITERATING);
  if (EPI == EPI_SMOOTH)
    return (a.mg != nullptr && a.mg->cvg != CSB_SLES_ITERATING);
  if (EPI == EPI_FCG_DOT4)
    return (a.st->cvg != CSB_SLES_ITERATING);
  return false;
}

template <int EPI>
__device__ __forceinline__ void spmv_epilogue(const SpmvK &a, int r, bool active,
                                              double sii)
{
  /* sii = sum of off-diagonal products in CSR order; diagonal added last
     (src/alge/cs_matrix_spmv.cpp:2138-2142) */
  double v[3] = {0., 0., 0.};
  double v3 = 0.;     /* 4th sum of EPI_FCG_DOT4 */
  if (active) {
    const double xr = (a.d_val != nullptr || EPI == EPI_PCG_DOT3 || EPI == EPI_SMOOTH
                       || EPI == EPI_FCG_DOT4) ? a.x[r] : 0.;
    const double yr = (a.d_val != nullptr) ? sii + a.d_val[r]*xr : sii;
    if (EPI == EPI_STORE) {
      a.y[r] = yr;
    }
    else if (EPI == EPI_RESIDUAL || EPI == EPI_RESIDUAL_NORM) {
      const double w = a.b[r] - yr;
      a.y[r] = w;
      v[0] = w*w;
    }
    else if (EPI == EPI_PCG_INIT) {
      const double rr = yr - a.b[r];
      a.rk[r] = rr;
      a.dk[r] = rr * a.ad_inv[r];
    }
    else if (EPI == EPI_PCG_DOT3) {
      a.y[r] = yr;
      const double rr = a.rk[r];
      v[0] = rr*rr;
      v[1] = rr*xr;
      v[2] = xr*yr;
    }
    else if (EPI == EPI_SMOOTH) {
      /* yr = E.x (diagonal excluded) for the Jacobi variants, A.x for L1 */
      const double t = a.b[r] - yr;
      if (a.smooth_mode == SMOOTH_JACOBI)
        a.y[r] = t * a.ad_inv[r];
      else if (a.smooth_mode == SMOOTH_RELAXED)
        a.y[r] = a.w1*xr + a.w0*t*a.ad_inv[r];
      else
        a.y[r] = xr + t * a.ad_inv[r];
    }
    else if (EPI == EPI_FCG_DOT4) {
      a.y[r] = yr;
      const double rr = a.rk[r];
      v[0] = rr*rr;
      v[1] = rr*xr;
      v[2] = yr*xr;
      v3 = xr*a.qk[r];
    }
  }
  if (EPI == EPI_FCG_DOT4) {
    double w[4] = {v[0], v[1], v[2], v3};
    ItState *st = a.st;
    grid_sum_finalize<4>(w, a.partials, a.ticket,
                         [st](const double *s) { fcg_finalize(st, s); });
  }
  if (EPI == EPI_RESIDUAL_NORM) {
    double w[1] = {v[0]};
    MgState *mg = a.mg;
    grid_sum_finalize<1>(w, a.partials, a.ticket,
                         [mg](const double *s) { mg_finalize_residual(mg, s); });
  }
  else if (EPI == EPI_PCG_DOT3) {
    ItState *st = a.st;
    if (a.phase == 0)
      grid_sum_finalize<3>(v, a.partials, a.ticket,
                           [st](const double *s) { pcg_finalize_first(st, s); });
    else if (a.phase == 1)
      grid_sum_finalize<3>(v, a.partials, a.ticket,
                           [st](const double *s) { pcg_finalize_iter(st, s); });
    else     /* 2: IPCG, only the step length comes out of this product */
      grid_sum_finalize<3>(v, a.partials, a.ticket,
                           [st](const double *s) { ipcg_finalize_alpha(st, s); });
  }
}

template <int EPI>
__global__ void __launch_bounds__(TPB)
k_spmv_stream(const SpmvK a)
{
  extern __shared__ double s_prod[];
  if (spmv_gated_off<EPI>(a)) return;

  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TPB;
  const int r1 = min(r0 + TPB, a.n_rows);
  const int k0 = __ldg(a.row_index + r0);
  const int k1 = __ldg(a.row_index + r1);
  const int nk = k1 - k0;

  const int    *__restrict__ ci = a.col_id + k0;
  const double *__restrict__ xv = a.x_val + k0;
  const double *__restrict__ x = a.x;

  /* coalesced streaming of the CTA's entries, 4 independent loads in flight */
  int i = tid;
  for (; i + 3*TPB < nk; i += 4*TPB) {
    const int c0 = __ldg(ci + i),         c1 = __ldg(ci + i + TPB);
    const int c2 = __ldg(ci + i + 2*TPB), c3 = __ldg(ci + i + 3*TPB);
    const double m0 = __ldg(xv + i),         m1 = __ldg(xv + i + TPB);
    const double m2 = __ldg(xv + i + 2*TPB), m3 = __ldg(xv + i + 3*TPB);
    const double x0 = __ldg(x + c0), x1 = __ldg(x + c1);
    const double x2 = __ldg(x + c2), x3 = __ldg(x + c3);
    s_prod[i]         = m0*x0;
    s_prod[i + TPB]   = m1*x1;
    s_prod[i + 2*TPB] = m2*x2;
    s_prod[i + 3*TPB] = m3*x3;
  }
  for (; i < nk; i += TPB)
    s_prod[i] = __ldg(xv + i) * __ldg(x + __ldg(ci + i));
  __syncthreads();

  const int r = r0 + tid;
  const bool active = (r < r1);
  double sii = 0.0;
  if (active) {
    const int ks = __ldg(a.row_index + r) - k0;
    const int ke = __ldg(a.row_index + r + 1) - k0;
    for (int k = ks; k < ke; k++)
      sii += s_prod[k];
  }
  spmv_epilogue<EPI>(a, r, active, sii);
}

/* Fallback when a 256-row slab does not fit in shared memory: thread per row */
template <int EPI>
__global__ void __launch_bounds__(TPB)
k_spmv_scalar(const SpmvK a)
{
  if (spmv_gated_off<EPI>(a)) return;
  const int r = blockIdx.x * TPB + threadIdx.x;
  const bool active = (r < a.n_rows);
  double sii = 0.0;
  if (active) {
    const int ks = a.row_index[r], ke = a.row_index[r+1];
    for (int k = ks; k < ke; k++)
      sii += a.x_val[k] * a.x[a.col_id[k]];
  }
  spmv_epilogue<EPI>(a, r, active, sii);
}

template <int EPI>
bool spmv_dispatch(const SpmvK &k, const _csb_matrix_t *m)
{
  const int grid = (m->n_rows + TPB - 1) / TPB;
  if (grid < 1) return true;
  if (m->spmv_stream_ok) {
    smem_optin((const void *)k_spmv_stream<EPI>, m->spmv_smem);
    k_spmv_stream<EPI><<<grid, TPB, m->spmv_smem, ctx().stream>>>(k);
  }
  else
    k_spmv_scalar<EPI><<<grid, TPB, 0, ctx().stream>>>(k);
  return post_launch("spmv");
}

} // anonymous namespace

/* CSR matrix, diagonal excluded (_mat_vec_p_l_csr, cs_matrix_spmv.cpp:1936-1952): thread per
   row, entries in order, col_id == row skipped.  Not a hot path (Jacobi on CSR only). */
__global__ void __launch_bounds__(TPB)
k_spmv_csr_exdiag(int n_rows, const int *__restrict__ row_index, const int *__restrict__ col_id,
                  const double *__restrict__ val, const double *__restrict__ x,
                  double *__restrict__ y)
{
  const int ii = blockIdx.x*TPB + threadIdx.x;
  if (ii >= n_rows) return;
  double sii = 0.0;
  for (int jj = row_index[ii]; jj < row_index[ii+1]; jj++)
    if (col_id[jj] != ii) sii += (val[jj]*x[col_id[jj]]);
  y[ii] = sii;
}

bool launch_spmv_csr_exdiag(const _csb_matrix_t *m, const double *x, double *y)
{
  const int grid = (m->n_rows + TPB - 1) / TPB;
  if (grid < 1) return true;
  k_spmv_csr_exdiag<<<grid, TPB, 0, ctx().stream>>>(m->n_rows, m->row_index.p, m->col_id.p,
                                                    m->x_val.p, x, y);
  return post_launch("spmv_csr_exdiag");
}

/*--------------------------------------------------------------------------
 * 3x3 block-diagonal MSR (CS_MATRIX_BLOCK_D): _b_mat_vec_p_l_msr_3
 * (cs_matrix_spmv.cpp:2394-2455).  Thread per block row: y_i = D_i.x_i
 * (_dense_3_3_ax, :590-607) FIRST, then += e_ij * x_j in CSR order.
 * MODE -1: store only; 0: + (y.b) -> gamma; 1: + (y.y, y.b) -> alpha (BiCGStab).
 * Algorithmic bytes: 12 nnz + 4 n + 72 n (D) + 24 n (x) + 24 n (y).
 *--------------------------------------------------------------------------*/

template <int MODE>
__global__ void __launch_bounds__(TPB)
k_spmv_b3(int n_rows, const int *__restrict__ row_index, const int *__restrict__ col_id,
          const double *__restrict__ d_val, const double *__restrict__ x_val,
          const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ b,
          ItState *st, double *partials, unsigned *ticket)
{
  if (MODE >= 0 && st->cvg != CSB_SLES_ITERATING) return;
  const int ii = blockIdx.x*TPB + threadIdx.x;
  double v[2] = {0., 0.};
  if (ii < n_rows) {
    double y0, y1, y2;
    if (d_val != nullptr) {
      const double *a = d_val + (size_t)ii*9;
      const double x0 = x[ii*3], x1 = x[ii*3 + 1], x2 = x[ii*3 + 2];
      y0 = a[0]*x0 + a[1]*x1 + a[2]*x2;
      y1 = a[3]*x0 + a[4]*x1 + a[5]*x2;
      y2 = a[6]*x0 + a[7]*x1 + a[8]*x2;
    }
    else { y0 = 0.; y1 = 0.; y2 = 0.; }
    const int ks = row_index[ii], ke = row_index[ii+1];
    for (int jj = ks; jj < ke; jj++) {
      const double e = x_val[jj];
      const double *xc = x + (size_t)col_id[jj]*3;
      y0 += (e*xc[0]);
      y1 += (e*xc[1]);
      y2 += (e*xc[2]);
    }
    y[ii*3] = y0; y[ii*3 + 1] = y1; y[ii*3 + 2] = y2;
    if (MODE == 0)
      v[0] = y0*b[ii*3] + y1*b[ii*3 + 1] + y2*b[ii*3 + 2];
    else if (MODE == 1) {
      v[0] = y0*y0 + y1*y1 + y2*y2;
      v[1] = y0*b[ii*3] + y1*b[ii*3 + 1] + y2*b[ii*3 + 2];
    }
  }
  if (MODE == 0)
    grid_sum_finalize<2>(v, partials, ticket, [st](const double *s) {
      st->gamma = st->beta / s[0];          /* cs_sles_it.cpp:2830-2832 */
    });
  else if (MODE == 1)
    grid_sum_finalize<2>(v, partials, ticket, [st](const double *s) {
      const double ro_1 = s[0], ro_0 = s[1]; /* cs_sles_it.cpp:2847-2855 */
      if (fabs(ro_1) < 1.e-30) { st->cvg = CSB_SLES_BREAKDOWN; return; }
      double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
      st->alpha = ro_0 * d_ro_1;
    });
}

bool launch_spmv_bn(int mode, const _csb_matrix_t *m, ItState *st, const double *x, double *y,
                    const double *b, bool exclude_diag);

/* 3x3 blocks through the thread-per-component kernel (k_spmv_bn) instead of the
   thread-per-block-row one: chosen by measurement (profiles/r02/config4_times.md) */
static bool b3_by_component()
{
  static int v = -1;
  if (v < 0) { const char *e = getenv("CSB_SPMV_B3_BN"); v = (e != nullptr) ? (atoi(e) != 0) : 0; }
  return v != 0;
}

bool launch_spmv_b3(const _csb_matrix_t *m, const double *x, double *y, bool exclude_diag)
{
  if (b3_by_component()) return launch_spmv_bn(-1, m, nullptr, x, y, nullptr, exclude_diag);
  const int grid = (m->n_rows + TPB - 1) / TPB;
  if (grid < 1) return true;
  k_spmv_b3<-1><<<grid, TPB, 0, ctx().stream>>>(m->n_rows, m->row_index.p, m->col_id.p,
                                                exclude_diag ? nullptr : m->d_val.p,
                                                m->x_val.p, x, y, nullptr, nullptr, nullptr,
                                                nullptr);
  return post_launch("spmv_b3");
}

bool ensure_partials(int n_blocks);

bool launch_bicg_spmv_b3(int mode, const _csb_matrix_t *m, ItState *st, const double *x,
                         double *y, const double *b)
{
  if (b3_by_component()) return launch_spmv_bn(mode, m, st, x, y, b, false);
  const int grid = (m->n_rows + TPB - 1) / TPB;
  if (grid < 1) return true;
  if (!ensure_partials(grid)) return false;
  Context &c = ctx();
  if (mode == 0)
    k_spmv_b3<0><<<grid, TPB, 0, c.stream>>>(m->n_rows, m->row_index.p, m->col_id.p,
                                             m->d_val.p, m->x_val.p, x, y, b, st, c.partials,
                                             c.ticket);
  else
    k_spmv_b3<1><<<grid, TPB, 0, c.stream>>>(m->n_rows, m->row_index.p, m->col_id.p,
                                             m->d_val.p, m->x_val.p, x, y, b, st, c.partials,
                                             c.ticket);
  return post_launch("bicg_spmv_b3");
}

/*--------------------------------------------------------------------------
 * Block-diagonal MSR of any block size 2..9 (CS_MATRIX_BLOCK_D): _b_mat_vec_p_l_msr_generic
 * and _b_mat_vec_p_l_msr_6 (cs_matrix_spmv.cpp:2308-2380, :2478-2545).  Thread per (block
 * row, component): the db threads of a row read the same e_ij (broadcast) and db consecutive
 * x, so the accesses coalesce whatever the block size; a component's sum is the reference's,
 * term by term: (D_i.x_i)_k (_dense_b_ax :560-576 starts from 0., _dense_6_6_ax :622-670 from
 * the first product) FIRST, then += e_ij * x_j,k in CSR order.
 * MODE as k_spmv_b3.  Algorithmic bytes: 12 nnz + 4 n + 8 db^2 n (D) + 8 db n (x) + 8 db n (y).
 *--------------------------------------------------------------------------*/

template <int MODE>
__global__ void __launch_bounds__(TPB)
k_spmv_bn(int n_rows, int db, const int *__restrict__ row_index,
          const int *__restrict__ col_id, const double *__restrict__ d_val,
          const double *__restrict__ x_val, const double *__restrict__ x,
          double *__restrict__ y, const double *__restrict__ b, ItState *st,
          double *partials, unsigned *ticket)
{
  if (MODE >= 0 && st->cvg != CSB_SLES_ITERATING) return;
  const long long t = (long long)blockIdx.x*TPB + threadIdx.x;
  double v[2] = {0., 0.};
  if (t < (long long)n_rows*db) {
    const int ii = (int)(t / db), kk = (int)(t - (long long)ii*db);
    double yk = 0.;
    if (d_val != nullptr) {
      const double *a = d_val + (size_t)ii*db*db + (size_t)kk*db;
      const double *xi = x + (size_t)ii*db;
      if (db == 6 || db == 3) {                 /* _dense_6_6_ax / _dense_3_3_ax: no leading 0. */
        yk = a[0]*xi[0];
        for (int j = 1; j < db; j++) yk += a[j]*xi[j];
      }
      else
        for (int j = 0; j < db; j++) yk += a[j]*xi[j];
    }
    const int ks = row_index[ii], ke = row_index[ii+1];
    for (int jj = ks; jj < ke; jj++)
      yk += (x_val[jj]*x[(size_t)col_id[jj]*db + kk]);
    y[t] = yk;
    if (MODE == 0) v[0] = yk*b[t];
    else if (MODE == 1) { v[0] = yk*yk; v[1] = yk*b[t]; }
  }
  if (MODE == 0)
    grid_sum_finalize<2>(v, partials, ticket, [st](const double *s) {
      st->gamma = st->beta / s[0];          /* cs_sles_it.cpp:2830-2832 */
    });
  else if (MODE == 1)
    grid_sum_finalize<2>(v, partials, ticket, [st](const double *s) {
      const double ro_1 = s[0], ro_0 = s[1]; /* cs_sles_it.cpp:2847-2855 */
      if (fabs(ro_1) < 1.e-30) { st->cvg = CSB_SLES_BREAKDOWN; return; }
      double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
      st->alpha = ro_0 * d_ro_1;
    });
}

/* mode -1: y = A.x (or E.x); 0 / 1: the two BiCGStab products with their dot products */
bool launch_spmv_bn(int mode, const _csb_matrix_t *m, ItState *st, const double *x, double *y,
                    const double *b, bool exclude_diag)
{
  const long long nt = (long long)m->n_rows*m->db;
  const int grid = (int)((nt + TPB - 1) / TPB);
  if (grid < 1) return true;
  Context &c = ctx();
  if (mode >= 0 && !ensure_partials(grid)) return false;
  const double *d = exclude_diag ? nullptr : m->d_val.p;
  if (mode < 0)
    k_spmv_bn<-1><<<grid, TPB, 0, c.stream>>>(m->n_rows, m->db, m->row_index.p, m->col_id.p, d,
                                              m->x_val.p, x, y, nullptr, nullptr, nullptr,
                                              nullptr);
  else if (mode == 0)
    k_spmv_bn<0><<<grid, TPB, 0, c.stream>>>(m->n_rows, m->db, m->row_index.p, m->col_id.p, d,
                                             m->x_val.p, x, y, b, st, c.partials, c.ticket);
  else
    k_spmv_bn<1><<<grid, TPB, 0, c.stream>>>(m->n_rows, m->db, m->row_index.p, m->col_id.p, d,
                                             m->x_val.p, x, y, b, st, c.partials, c.ticket);
  return post_launch("spmv_bn");
}

/* point Jacobi of a block matrix: ad_inv[i] = 1 / ad[i*db + i%db]  (cs_sles_pc.cpp:383-388) */
__global__ void __launch_bounds__(TPB)
k_ad_inv_block(int n, int db, const double *__restrict__ d, double *__restrict__ ad_inv)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
    ad_inv[i] = 1.0 / d[(size_t)i*db + (i % db)];
}

bool launch_ad_inv_block(int n_rows, int db, const double *d_val, double *ad_inv)
{
  const int n = n_rows*db;
  if (n < 1) return true;
  int g = (n + TPB - 1) / TPB;
  if (g > 65535) g = 65535;
  k_ad_inv_block<<<g, TPB, 0, ctx().stream>>>(n, db, d_val, ad_inv);
  return post_launch("ad_inv_block");
}

void matrix_plan_spmv(_csb_matrix_t *m)
{
  /* largest number of off-diagonal entries in any slab of 256 rows */
  size_t max_nk = 0;
  const std::vector<csb_lnum_t> &ri = m->h_row_index;
  for (csb_lnum_t r0 = 0; r0 < m->n_rows; r0 += TPB) {
    csb_lnum_t r1 = r0 + TPB;
    if (r1 > m->n_rows) r1 = m->n_rows;
    size_t nk = (size_t)(ri[r1] - ri[r0]);
    if (nk > max_nk) max_nk = nk;
  }
  m->spmv_rows_per_block = TPB;
  m->spmv_smem = max_nk * sizeof(double);
  if (m->spmv_smem < 16) m->spmv_smem = 16;
  m->spmv_stream_ok = (m->spmv_smem <= 200*1024);
}

bool ensure_partials(int n_blocks);

/*--------------------------------------------------------------------------
 * Stand-alone y = A.x, a thread per row with everything in registers (k_spmv_rows).
 * Measured on the 256^3 cube (profiles/r02): on B200 the row-per-thread access pattern is
 * served by L1 (a warp's 32 rows are ~2.7 KB of contiguous col_id / x_val; every 128-byte
 * line is consumed by neighbouring lanes within a few instructions), and 2048 resident
 * threads per SM with 3 x 8 independent loads each keep more bytes in flight than the
 * shared-memory staged kernels did -- the reference's own `_mat_vect_p_l_msr`
 * (cs_matrix_spmv_cuda.cu:415-439) ran at 0.98 of the measured copy rate, the bulk-copy
 * pipeline at 0.90, k_spmv_stream at 0.75.  This kernel keeps that access pattern and adds:
 * batches of 8 entries whose index / value loads, then gathers, are all issued before the
 * first use (the compiler cannot hoist them out of a variable-trip loop by itself);
 * optionally streaming (evict-first) loads of the matrix so that x stays in L2; the CPU path's
 * arithmetic (no FMA, CSR order, diagonal last) so that y is bit-identical to the oracle.
 *--------------------------------------------------------------------------*/

namespace {

/* matrix entries are read once: STREAM selects ld.global.cs (evict-first) for them so that
   x keeps its place in L2; measured against plain read-only loads (CSB_SPMV_ROWS_CS) */
template <bool STREAM> __device__ __forceinline__ int ld_mat(const int *p)
{ return STREAM ? __ldcs(p) : __ldg(p); }
template <bool STREAM> __device__ __forceinline__ double ld_mat(const double *p)
{ return STREAM ? __ldcs(p) : __ldg(p); }

template <bool EXDIAG, bool STREAM, int B, int MINB>
__global__ void __launch_bounds__(TPB, MINB)
k_spmv_rows(int n_rows, const int *__restrict__ row_index, const int *__restrict__ col_id,
            const double *__restrict__ d_val, const double *__restrict__ x_val,
            const double *__restrict__ x, double *__restrict__ y)
{
  const int ii = blockIdx.x*TPB + threadIdx.x;
  if (ii >= n_rows) return;
  const int ks = __ldg(row_index + ii), ke = __ldg(row_index + ii + 1);
  double dx = 0.0;
  if (!EXDIAG) dx = ld_mat<STREAM>(d_val + ii) * __ldg(x + ii);
  double sii = 0.0;
  int k = ks;
  for (; k + B <= ke; k += B) {
    int c[B]; double v[B], g[B];
#pragma unroll
    for (int j = 0; j < B; j++) { c[j] = ld_mat<STREAM>(col_id + k + j); v[j] = ld_mat<STREAM>(x_val + k + j); }
#pragma unroll
    for (int j = 0; j < B; j++) g[j] = __ldg(x + c[j]);
#pragma unroll
    for (int j = 0; j < B; j++) sii += v[j] * g[j];
  }
  if (k < ke) {
    int c[B]; double v[B], g[B];
#pragma unroll
    for (int j = 0; j < B - 1; j++) {
      const bool p = (k + j < ke);
      c[j] = p ? ld_mat<STREAM>(col_id + k + j) : ii;
      v[j] = p ? ld_mat<STREAM>(x_val + k + j) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < B - 1; j++) g[j] = __ldg(x + c[j]);
#pragma unroll
    for (int j = 0; j < B - 1; j++) if (k + j < ke) sii += v[j] * g[j];
  }
  y[ii] = EXDIAG ? sii : sii + dx;
}

template <bool EXDIAG, bool STREAM, int B, int MINB>
void rows_launch(int grid, const _csb_matrix_t *m, const SpmvK &k)
{
  k_spmv_rows<EXDIAG, STREAM, B, MINB><<<grid, TPB, 0, ctx().stream>>>(
    m->n_rows, k.row_index, k.col_id, m->d_val.p, k.x_val, k.x, k.y);
}

int env_flag(const char *name, int dflt)
{
  const char *e = getenv(name);
  return (e != nullptr) ? atoi(e) : dflt;
}

/* 0 rows (default), 1 bulk-copy pipeline, 2 streaming kernel */
int spmv_kernel_choice()
{
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("CSB_SPMV_KERNEL");
    v = 0;
    if (e != nullptr && strcmp(e, "pipeline") == 0) v = 1;
    if (e != nullptr && strcmp(e, "stream") == 0) v = 2;
    /* earlier switch, still honoured */
    const char *o = getenv("CSB_SPMV_PIPELINE");
    if (e == nullptr && o != nullptr) v = (atoi(o) == 0) ? 2 : 1;
  }
  return v;
}

} // anonymous namespace

const char *spmv_kernel_name()
{
  const int v = spmv_kernel_choice();
  return v == 0 ? "k_spmv_rows" : (v == 1 ? "k_spmv_pipeline" : "k_spmv_stream");
}

bool launch_spmv(SpmvEpilogue epi, const SpmvArgs &a)
{
  const _csb_matrix_t *m = a.m;
  SpmvK k;
  k.n_rows = m->n_rows;
  k.row_index = m->row_index.p;
  k.col_id = m->col_id.p;
  k.d_val = a.exclude_diag ? nullptr : m->d_val.p;
  k.x_val = m->x_val.p;
  k.x = a.x; k.y = a.y; k.b = a.b; k.ad_inv = a.ad_inv;
  k.rk = a.rk; k.dk = a.dk; k.st = a.st; k.mg = a.mg; k.phase = a.phase;
  k.smooth_mode = a.smooth_mode; k.w0 = a.w0; k.w1 = a.w1; k.qk = a.qk;
  if (epi == EPI_RESIDUAL_NORM || epi == EPI_PCG_DOT3 || epi == EPI_FCG_DOT4)
    if (!ensure_partials((m->n_rows + TPB - 1) / TPB)) return false;
  k.partials = ctx().partials; k.ticket = ctx().ticket;   /* after a possible regrow */
  if (epi == EPI_STORE && spmv_kernel_choice() == 0) {
    const int grid = (m->n_rows + TPB - 1)/TPB;
    if (grid < 1) return true;
    /* measured variants (profiles/r02/spmv_rows_variants.md): batches of 8 entries at 48
       registers (5 CTAs per SM) or batches of 4 capped at 32 registers (8 CTAs per SM) */
    static const int cs = env_flag("CSB_SPMV_ROWS_CS", 0);
    static const int b4 = env_flag("CSB_SPMV_ROWS_B4", 0);
    const int v = (a.exclude_diag ? 4 : 0) + (cs ? 2 : 0) + (b4 ? 1 : 0);
    switch (v) {
    case 0: rows_launch<false, false, 8, 5>(grid, m, k); break;
    case 1: rows_launch<false, false, 4, 8>(grid, m, k); break;
    case 2: rows_launch<false, true, 8, 5>(grid, m, k); break;
    case 3: rows_launch<false, true, 4, 8>(grid, m, k); break;
    case 4: rows_launch<true, false, 8, 5>(grid, m, k); break;
    case 5: rows_launch<true, false, 4, 8>(grid, m, k); break;
    case 6: rows_launch<true, true, 8, 5>(grid, m, k); break;
    default: rows_launch<true, true, 4, 8>(grid, m, k); break;
    }
    return post_launch("spmv_rows");
  }
  if (epi == EPI_STORE && spmv_kernel_choice() == 1) {
    /* bulk-copy pipeline with producer / consumer warps (persistent.cu); the streaming
       kernel below remains for buffers the pipeline cannot take */
    bool launched = false;
    if (!launch_spmv_pipeline(m, a.x, a.y, a.exclude_diag, &launched)) return false;
    if (launched) return true;
  }
  switch (epi) {
  case EPI_STORE:         return spmv_dispatch<EPI_STORE>(k, m);
  case EPI_RESIDUAL:      return spmv_dispatch<EPI_RESIDUAL>(k, m);
  case EPI_RESIDUAL_NORM: return spmv_dispatch<EPI_RESIDUAL_NORM>(k, m);
  case EPI_PCG_INIT:      return spmv_dispatch<EPI_PCG_INIT>(k, m);
  case EPI_PCG_DOT3:      return spmv_dispatch<EPI_PCG_DOT3>(k, m);
  case EPI_SMOOTH:        return spmv_dispatch<EPI_SMOOTH>(k, m);
  case EPI_FCG_DOT4:      return spmv_dispatch<EPI_FCG_DOT4>(k, m);
  }
  return false;
}

/* partials buffer grows with the largest reduction grid seen */
bool ensure_partials(int n_blocks)
{
  Context &c = ctx();
  if (n_blocks <= c.partials_blocks) return true;
  if (c.partials) {
    cudaStreamSynchronize(c.stream);
    cudaFree(c.partials);
    c.partials = nullptr;
  }
  int nb = n_blocks + n_blocks/4 + 1024;
  if (!CSB_CUDA(cudaMalloc((void **)&c.partials, sizeof(double)*RED_MAX_Q*(size_t)nb)))
    return false;
  c.partials_blocks = nb;
  return true;
}

void reset_partials()
{
  ctx().partials_blocks = 0;
}

/*--------------------------------------------------------------------------
 * Elementwise / BLAS-1
 *--------------------------------------------------------------------------*/

namespace {

__global__ void __launch_bounds__(TPB)
k_ad_inv(int n, const double *__restrict__ d, double *__restrict__ ad_inv)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
    ad_inv[i] = 1.0 / d[i];
}

/* mode 0: x.y ; 1: x.x ; 2: (x.x, x.y, y.z) ; 3: (x.x, x.y) */
template <int MODE>
__global__ void __launch_bounds__(TPB)
k_dot(int n, const double *__restrict__ x, const double *__restrict__ y,
      const double *__restrict__ z, double *partials, unsigned *ticket, double *out)
{
  double v[3] = {0., 0., 0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    if (MODE == 0) v[0] += x[i]*y[i];
    else if (MODE == 1) v[0] += x[i]*x[i];
    else if (MODE == 2) { v[0] += x[i]*x[i]; v[1] += x[i]*y[i]; v[2] += y[i]*z[i]; }
    else { v[0] += x[i]*x[i]; v[1] += x[i]*y[i]; }
  }
  grid_sum_finalize<3>(v, partials, ticket,
                       [out](const double *s) { out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; });
}

__global__ void
k_it_state_init(ItState *st, const MgState *gate, double precision, double r_norm,
                int n_max_iter)
{
  st->precision = precision;
  st->r_norm = r_norm;
  st->residual = -1.;
  st->initial_residual = -1.;
  st->ro_0 = 0.; st->ro_1 = 0.; st->rk_gkm1 = 0.; st->alpha = 0.; st->beta = 0.;
  st->gamma = 1.; st->omega = 0.; st->betam1 = 1.;
  st->n_iter = 0;
  st->n_max_iter = n_max_iter;
  st->n_hist = 0;
  st->cvg = (gate != nullptr && gate->cvg != CSB_SLES_ITERATING) ? 2 : CSB_SLES_ITERATING;
}

/* cs_sles_it.cpp:1511-1515 */
__global__ void __launch_bounds__(TPB)
k_pcg_init_zero(int n, const ItState *st, const double *__restrict__ rhs,
                const double *__restrict__ ad_inv, double *__restrict__ vx,
                double *__restrict__ rk, double *__restrict__ dk)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    vx[i] = 0;
    const double r = -rhs[i];
    rk[i] = r;
    dk[i] = r * ad_inv[i];
  }
}

/* cs_sles_it.cpp:1550-1582 */
__global__ void __launch_bounds__(TPB)
k_pcg_update_first(int n, ItState *st, const double *__restrict__ ad_inv,
                   const double *__restrict__ dk, const double *__restrict__ zk,
                   double *__restrict__ vx, double *__restrict__ rk,
                   double *__restrict__ gk, double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double alpha = st->alpha;
  double v[1] = {0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    vx[i] += (alpha * dk[i]);
    const double r = rk[i] + (alpha * zk[i]);
    rk[i] = r;
    gk[i] = r * ad_inv[i];
    v[0] += r*r;
  }
  grid_sum_finalize<1>(v, partials, ticket, [st](const double *s) {
    st->cvg = it_cvg_test(st, st->n_iter, sqrt(s[0]));
  });
}

/* cs_sles_it.cpp:1623-1631 */
__global__ void __launch_bounds__(TPB)
k_pcg_update(int n, const ItState *st, const double *__restrict__ ad_inv,
             const double *__restrict__ sk, double *__restrict__ dk,
             double *__restrict__ gk, double *__restrict__ zk,
             double *__restrict__ rk, double *__restrict__ vx)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double alpha = st->alpha, beta = st->beta;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    const double d = gk[i] + (beta * dk[i]);
    dk[i] = d;
    vx[i] += alpha * d;
    const double z = sk[i] + (beta * zk[i]);
    zk[i] = z;
    const double r = rk[i] + alpha * z;
    rk[i] = r;
    gk[i] = r * ad_inv[i];
  }
}

} // anonymous namespace

__global__ void __launch_bounds__(TPB)
k_fill(int n, double v, double *__restrict__ x)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) x[i] = v;
}

bool launch_fill(int n, double v, double *x)
{
  if (n < 1) return true;
  k_fill<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, v, x);
  return post_launch("fill");
}

__global__ void __launch_bounds__(TPB)
k_fill_gated(int n, double v, double *__restrict__ x, const MgState *gate)
{
  if (gate != nullptr && gate->cvg != CSB_SLES_ITERATING) return;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) x[i] = v;
}

bool launch_fill_gated(int n, double v, double *x, const MgState *gate)
{
  if (n < 1) return true;
  k_fill_gated<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, v, x, gate);
  return post_launch("fill_gated");
}

bool launch_ad_inv(int n, const double *d, double *ad_inv)
{
  if (n < 1) return true;
  k_ad_inv<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, d, ad_inv);
  return post_launch("ad_inv");
}

bool launch_dot(int n, const double *x, const double *y, const double *z, int mode)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  switch (mode) {
  case 0: k_dot<0><<<g, TPB, 0, c.stream>>>(n, x, y, z, c.partials, c.ticket, c.red_out); break;
  case 1: k_dot<1><<<g, TPB, 0, c.stream>>>(n, x, y, z, c.partials, c.ticket, c.red_out); break;
  case 2: k_dot<2><<<g, TPB, 0, c.stream>>>(n, x, y, z, c.partials, c.ticket, c.red_out); break;
  default: k_dot<3><<<g, TPB, 0, c.stream>>>(n, x, y, z, c.partials, c.ticket, c.red_out); break;
  }
  return post_launch("dot");
}

bool launch_it_state_init(ItState *st, const MgState *gate, double precision,
                          double r_norm, int n_max_iter)
{
  k_it_state_init<<<1, 1, 0, ctx().stream>>>(st, gate, precision, r_norm, n_max_iter);
  return post_launch("it_state_init");
}

bool launch_pcg_init_zero(int n, ItState *st, const double *rhs, const double *ad_inv,
                          double *vx, double *rk, double *dk)
{
  if (n < 1) return true;
  k_pcg_init_zero<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, rhs, ad_inv, vx, rk, dk);
  return post_launch("pcg_init_zero");
}

bool launch_pcg_update_first(int n, ItState *st, const double *ad_inv, const double *dk,
                             const double *zk, double *vx, double *rk, double *gk)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_pcg_update_first<<<g, TPB, 0, c.stream>>>(n, st, ad_inv, dk, zk, vx, rk, gk,
                                              c.partials, c.ticket);
  return post_launch("pcg_update_first");
}

bool launch_pcg_update(int n, ItState *st, const double *ad_inv, const double *sk,
                       double *dk, double *gk, double *zk, double *rk, double *vx)
{
  if (n < 1) return true;
  k_pcg_update<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, ad_inv, sk, dk, gk, zk, rk, vx);
  return post_launch("pcg_update");
}

/*--------------------------------------------------------------------------
 * Whole PCG(+Jacobi) solve in ONE thread block, for the small levels of the
 * hierarchy and the coarsest solve (up to 500 iterations in the reference,
 * cs_multigrid.cpp:4762): avoids 2 launches per iteration on latency-bound
 * levels.  Same algorithm as the multi-kernel path.
 *--------------------------------------------------------------------------*/

namespace {

__device__ __forceinline__ double sb_row(const int *__restrict__ ri,
                                         const int *__restrict__ ci,
                                         const double *__restrict__ xv,
                                         const double *__restrict__ dv,
                                         const double *x, int r)
{
  double sii = 0.0;
  const int ks = ri[r], ke = ri[r+1];
  for (int k = ks; k < ke; k++)
    sii += xv[k] * x[ci[k]];
  return sii + dv[r]*x[r];
}

__global__ void __launch_bounds__(SB_TPB)
k_pcg_single_block(int n, const int *__restrict__ ri, const int *__restrict__ ci,
                   const double *__restrict__ dv, const double *__restrict__ xv,
                   ItState *st, const MgState *gate, double precision, double r_norm,
                   int n_max_iter, const double *__restrict__ ad_inv,
                   const double *__restrict__ rhs, double *vx, int vx_ini_nonzero,
                   double *work)
{
  __shared__ double s_red[3*32];
  __shared__ int s_cvg;
  __shared__ double s_alpha, s_beta;

  const int tid = threadIdx.x;
  double *rk = work, *dk = work + n, *gk = work + 2*n, *zk = work + 3*n, *sk = work + 4*n;

  if (tid == 0) {
    st->precision = precision; st->r_norm = r_norm;
    st->residual = -1.; st->initial_residual = -1.;
    st->ro_0 = 0.; st->ro_1 = 0.; st->rk_gkm1 = 0.; st->alpha = 0.; st->beta = 0.;
    st->n_iter = 0; st->n_max_iter = n_max_iter; st->n_hist = 0;
    st->cvg = (gate != nullptr && gate->cvg != CSB_SLES_ITERATING) ? 2 : CSB_SLES_ITERATING;
    s_cvg = st->cvg;
  }
  __syncthreads();
  if (s_cvg != CSB_SLES_ITERATING) return;

  /* residual and descent direction */
  if (vx_ini_nonzero) {
    for (int r = tid; r < n; r += SB_TPB) {
      const double rr = sb_row(ri, ci, xv, dv, vx, r) - rhs[r];
      rk[r] = rr;
      dk[r] = rr * ad_inv[r];
    }
  }
  else {
    for (int r = tid; r < n; r += SB_TPB) {
      vx[r] = 0;
      const double rr = -rhs[r];
      rk[r] = rr;
      dk[r] = rr * ad_inv[r];
    }
  }
  __syncthreads();

  double v[3] = {0., 0., 0.};
  for (int r = tid; r < n; r += SB_TPB) {
    const double z = sb_row(ri, ci, xv, dv, dk, r);
    zk[r] = z;
    const double rr = rk[r], d = dk[r];
    v[0] += rr*rr; v[1] += rr*d; v[2] += d*z;
  }
  block_sum<3>(v, s_red);
  if (tid == 0) {
    pcg_finalize_first(st, v);
    s_cvg = st->cvg; s_alpha = st->alpha;
  }
  __syncthreads();

  if (s_cvg == CSB_SLES_ITERATING) {
    const double alpha = s_alpha;
    double w[1] = {0.};
    for (int r = tid; r < n; r += SB_TPB) {
      vx[r] += (alpha * dk[r]);
      const double rr = rk[r] + (alpha * zk[r]);
      rk[r] = rr;
      gk[r] = rr * ad_inv[r];
      w[0] += rr*rr;
    }
    block_sum<1>(w, s_red);
    if (tid == 0) {
      st->cvg = it_cvg_test(st, st->n_iter, sqrt(w[0]));
      s_cvg = st->cvg;
    }
    __syncthreads();
  }

  while (s_cvg == CSB_SLES_ITERATING) {
    v[0] = 0.; v[1] = 0.; v[2] = 0.;
    for (int r = tid; r < n; r += SB_TPB) {
      const double s = sb_row(ri, ci, xv, dv, gk, r);
      sk[r] = s;
      const double rr = rk[r], g = gk[r];
      v[0] += rr*rr; v[1] += rr*g; v[2] += g*s;
    }
    block_sum<3>(v, s_red);
    if (tid == 0) {
      pcg_finalize_iter(st, v);
      s_cvg = st->cvg; s_alpha = st->alpha; s_beta = st->beta;
    }
    __syncthreads();
    if (s_cvg != CSB_SLES_ITERATING) break;
    const double alpha = s_alpha, beta = s_beta;
    for (int r = tid; r < n; r += SB_TPB) {
      const double d = gk[r] + (beta * dk[r]);
      dk[r] = d;
      vx[r] += alpha * d;
      const double z = sk[r] + (beta * zk[r]);
      zk[r] = z;
      const double rr = rk[r] + alpha * z;
      rk[r] = rr;
      gk[r] = rr * ad_inv[r];
    }
    __syncthreads();
  }
}

} // anonymous namespace

bool launch_pcg_single_block(const _csb_matrix_t *m, ItState *st, const MgState *gate,
                             double precision, double r_norm, int n_max_iter,
                             const double *ad_inv, const double *rhs, double *vx,
                             int vx_ini_nonzero, double *work)
{
  k_pcg_single_block<<<1, SB_TPB, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->d_val.p, m->x_val.p, st, gate,
    precision, r_norm, n_max_iter, ad_inv, rhs, vx, vx_ini_nonzero, work);
  return post_launch("pcg_single_block");
}

/*--------------------------------------------------------------------------
 * Jacobi (src/alge/cs_sles_it.cpp:2098-2270)
 *--------------------------------------------------------------------------*/

namespace {

__global__ void __launch_bounds__(TPB)
k_jacobi_first(int n, ItState *st, const double *__restrict__ rhs,
               const double *__restrict__ ad_inv, double *__restrict__ vx,
               double *__restrict__ rk, double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  double v[1] = {0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    const double r = rhs[i];
    const double x = r * ad_inv[i];
    vx[i] = x;
    rk[i] = x;
    v[0] += (r*r);
  }
  grid_sum_finalize<1>(v, partials, ticket, [st](const double *s) {
    st->n_iter = 1;
    double residual = (st->precision > 0.) ? sqrt(s[0]) : -1.;
    st->initial_residual = residual;
    st->cvg = it_cvg_test(st, 1, residual);
  });
}

__global__ void __launch_bounds__(TPB)
k_copy_gated(int n, const double *__restrict__ src, double *__restrict__ dst,
             const ItState *st)
{
  if (st != nullptr && st->cvg != CSB_SLES_ITERATING) return;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
    dst[i] = src[i];
}

/* one sweep: vx <- (rhs - (A-D).rk)/D ; res2 = sum (D (vx-rk))^2 ; rk <- vx */
__global__ void __launch_bounds__(TPB)
k_jacobi_sweep(SpmvK a)
{
  extern __shared__ double s_prod[];
  ItState *st = a.st;
  if (st->cvg != CSB_SLES_ITERATING) return;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TPB;
  const int r1 = min(r0 + TPB, a.n_rows);
  const int k0 = __ldg(a.row_index + r0);
  const int nk = __ldg(a.row_index + r1) - k0;
  const int *__restrict__ ci = a.col_id + k0;
  const double *__restrict__ xv = a.x_val + k0;
  const double *__restrict__ x = a.x;        /* rk (previous iterate) */
  for (int i = tid; i < nk; i += TPB)
    s_prod[i] = __ldg(xv + i) * __ldg(x + __ldg(ci + i));
  __syncthreads();
  const int r = r0 + tid;
  double v[1] = {0.};
  if (r < r1) {
    const int ks = __ldg(a.row_index + r) - k0;
    const int ke = __ldg(a.row_index + r + 1) - k0;
    double sii = 0.0;
    for (int k = ks; k < ke; k++) sii += s_prod[k];
    const double xn = (a.b[r] - sii) * a.ad_inv[r];
    const double d = a.d_val[r] * (xn - x[r]);
    a.y[r] = xn;              /* vx */
    v[0] = d*d;
  }
  grid_sum_finalize<1>(v, a.partials, a.ticket, [st](const double *s) {
    st->n_iter += 1;
    double residual = (st->precision > 0.) ? sqrt(s[0]) : -1.;
    if (st->n_iter == 1) st->initial_residual = residual;
    st->cvg = it_cvg_test(st, st->n_iter, residual);
  });
}

} // anonymous namespace

bool launch_jacobi_first(int n, ItState *st, const double *rhs, const double *ad_inv,
                         double *vx, double *rk)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_jacobi_first<<<g, TPB, 0, c.stream>>>(n, st, rhs, ad_inv, vx, rk, c.partials, c.ticket);
  return post_launch("jacobi_first");
}

namespace {
/* latest Jacobi iterate lives in rk when the iteration count is odd */
__global__ void __launch_bounds__(TPB)
k_jacobi_finish(int n, const ItState *st, const double *__restrict__ rk,
                double *__restrict__ vx)
{
  if ((st->n_iter & 1) == 0) return;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
    vx[i] = rk[i];
}
}

bool launch_jacobi_finish(int n, const ItState *st, const double *rk, double *vx)
{
  if (n < 1) return true;
  k_jacobi_finish<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, rk, vx);
  return post_launch("jacobi_finish");
}

bool launch_copy(int n, const double *src, double *dst, const ItState *st)
{
  if (n < 1) return true;
  k_copy_gated<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, src, dst, st);
  return post_launch("copy");
}

bool launch_jacobi_sweep(const _csb_matrix_t *m, ItState *st, const double *rhs,
                         const double *ad_inv, double *vx, double *rk)
{
  /* reads rk (previous iterate), writes vx; caller swaps/copies */
  SpmvK k;
  k.n_rows = m->n_rows; k.row_index = m->row_index.p; k.col_id = m->col_id.p;
  k.d_val = m->d_val.p; k.x_val = m->x_val.p; k.x = rk; k.y = vx; k.b = rhs;
  k.ad_inv = ad_inv; k.rk = nullptr; k.dk = nullptr; k.st = st; k.mg = nullptr;
  k.phase = 0; k.partials = ctx().partials; k.ticket = ctx().ticket;
  const int grid = (m->n_rows + TPB - 1) / TPB;
  if (grid < 1) return true;
  if (!ensure_partials(grid)) return false;
  k.partials = ctx().partials;
  if (!m->spmv_stream_ok) {
    set_error("jacobi sweep: row slab exceeds shared memory");
    return false;
  }
  smem_optin((const void *)k_jacobi_sweep, m->spmv_smem);
  k_jacobi_sweep<<<grid, TPB, m->spmv_smem, ctx().stream>>>(k);
  return post_launch("jacobi_sweep");
}

/*--------------------------------------------------------------------------
 * BiCGStab + Jacobi preconditioner (src/alge/cs_sles_it.cpp:2693-2880)
 *--------------------------------------------------------------------------*/

namespace {

constexpr double BICG_EPZERO = 1.e-30;   /* _epzero, cs_sles_it.cpp:169 */

/* dots (rk.rk, rk.res0) -> residual, beta, convergence test, omega
   (cs_sles_it.cpp:2780-2812) */
__device__ __forceinline__ void bicg_finalize_a(ItState *st, const double *s)
{
  double residual, beta;
  if (st->n_iter == 0) {
    beta = s[0];
    residual = sqrt(beta);
    st->initial_residual = residual;
  }
  else {
    residual = sqrt(s[0]);
    beta = s[1];
  }
  int cvg = it_cvg_test(st, st->n_iter, residual);
  if (cvg == CSB_SLES_ITERATING) {
    st->n_iter += 1;
    if (fabs(beta) < BICG_EPZERO || fabs(st->alpha) < BICG_EPZERO)
      cvg = CSB_SLES_BREAKDOWN;
    else {
      st->omega = beta*st->gamma / (st->alpha*st->betam1);
      st->betam1 = beta;
      st->beta = beta;
    }
  }
  st->cvg = cvg;
}

__global__ void __launch_bounds__(TPB)
k_bicg_init(int n, ItState *st, const double *__restrict__ rhs, double *__restrict__ vx,
            double *__restrict__ res0, double *__restrict__ rk, double *__restrict__ pk,
            double *__restrict__ uk, int have_ax)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    pk[i] = 0.0;
    uk[i] = 0.0;
    double r;
    if (have_ax)
      r = -res0[i] + rhs[i];     /* res0 holds A.vx on entry */
    else {
      vx[i] = 0;
      r = rhs[i];
    }
    res0[i] = r;
    rk[i] = r;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->alpha = 1.0; st->betam1 = 1.0; st->gamma = 1.0;
  }
}

__global__ void __launch_bounds__(TPB)
k_bicg_dots_a(int n, ItState *st, const double *__restrict__ rk,
              const double *__restrict__ res0, double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  double v[2] = {0., 0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    const double r = rk[i];
    v[0] += r*r;
    v[1] += r*res0[i];
  }
  grid_sum_finalize<2>(v, partials, ticket,
                       [st](const double *s) { bicg_finalize_a(st, s); });
}

__global__ void __launch_bounds__(TPB)
k_bicg_pk_zk(int n, const ItState *st, const double *__restrict__ rk,
             const double *__restrict__ uk, const double *__restrict__ ad_inv,
             double *__restrict__ pk, double *__restrict__ zk)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double omega = st->omega, alpha = st->alpha;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    const double p = rk[i] + omega*(pk[i] - alpha*uk[i]);
    pk[i] = p;
    zk[i] = p * ad_inv[i];
  }
}

/* y = A.x with one or two fused dots; MODE 0: (y.b) -> gamma ; 1: (y.y, y.b) -> alpha */
template <int MODE>
__global__ void __launch_bounds__(TPB)
k_bicg_spmv(SpmvK a)
{
  extern __shared__ double s_prod[];
  ItState *st = a.st;
  if (st->cvg != CSB_SLES_ITERATING) return;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TPB;
  const int r1 = min(r0 + TPB, a.n_rows);
  const int k0 = __ldg(a.row_index + r0);
  const int nk = __ldg(a.row_index + r1) - k0;
  const int *__restrict__ ci = a.col_id + k0;
  const double *__restrict__ xv = a.x_val + k0;
  const double *__restrict__ x = a.x;
  for (int i = tid; i < nk; i += TPB)
    s_prod[i] = __ldg(xv + i) * __ldg(x + __ldg(ci + i));
  __syncthreads();
  const int r = r0 + tid;
  double v[2] = {0., 0.};
  if (r < r1) {
    const int ks = __ldg(a.row_index + r) - k0;
    const int ke = __ldg(a.row_index + r + 1) - k0;
    double sii = 0.0;
    for (int k = ks; k < ke; k++) sii += s_prod[k];
    const double yr = sii + a.d_val[r]*x[r];
    a.y[r] = yr;
    if (MODE == 0) v[0] = yr * a.b[r];
    else { v[0] = yr*yr; v[1] = yr * a.b[r]; }
  }
  if (MODE == 0)
    grid_sum_finalize<2>(v, a.partials, a.ticket, [st](const double *s) {
      st->gamma = st->beta / s[0];          /* cs_sles_it.cpp:2830-2832 */
    });
  else
    grid_sum_finalize<2>(v, a.partials, a.ticket, [st](const double *s) {
      const double ro_1 = s[0], ro_0 = s[1]; /* cs_sles_it.cpp:2847-2855 */
      if (fabs(ro_1) < BICG_EPZERO) { st->cvg = CSB_SLES_BREAKDOWN; return; }
      double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
      st->alpha = ro_0 * d_ro_1;
    });
}

__global__ void __launch_bounds__(TPB)
k_bicg_upd1(int n, const ItState *st, const double *__restrict__ ad_inv,
            double *__restrict__ vx, double *__restrict__ rk, double *__restrict__ zk,
            const double *__restrict__ uk)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double gamma = st->gamma;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    vx[i] += (gamma * zk[i]);
    const double r = rk[i] - (gamma * uk[i]);
    rk[i] = r;
    zk[i] = r * ad_inv[i];
  }
}

/* final update of the iteration fused with next iteration's (rk.rk, rk.res0) */
__global__ void __launch_bounds__(TPB)
k_bicg_upd2(int n, ItState *st, double *__restrict__ vx, double *__restrict__ rk,
            const double *__restrict__ zk, const double *__restrict__ vk,
            const double *__restrict__ res0, double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double alpha = st->alpha;
  double v[2] = {0., 0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    vx[i] += (alpha * zk[i]);
    const double r = rk[i] - (alpha * vk[i]);
    rk[i] = r;
    v[0] += r*r;
    v[1] += r*res0[i];
  }
  grid_sum_finalize<2>(v, partials, ticket,
                       [st](const double *s) { bicg_finalize_a(st, s); });
}

bool bicg_spmv_launch(int mode, const _csb_matrix_t *m, ItState *st, const double *x,
                      double *y, const double *b)
{
  if (m->db == 3) return launch_bicg_spmv_b3(mode, m, st, x, y, b);
  if (m->db > 1) return launch_spmv_bn(mode, m, st, x, y, b, false);
  SpmvK k;
  k.n_rows = m->n_rows; k.row_index = m->row_index.p; k.col_id = m->col_id.p;
  k.d_val = m->d_val.p; k.x_val = m->x_val.p; k.x = x; k.y = y; k.b = b;
  k.ad_inv = nullptr; k.rk = nullptr; k.dk = nullptr; k.st = st; k.mg = nullptr;
  k.phase = 0;
  const int grid = (m->n_rows + TPB - 1) / TPB;
  if (grid < 1) return true;
  if (!ensure_partials(grid)) return false;
  k.partials = ctx().partials; k.ticket = ctx().ticket;
  if (!m->spmv_stream_ok) {
    set_error("bicgstab spmv: row slab exceeds shared memory");
    return false;
  }
  smem_optin(mode == 0 ? (const void *)k_bicg_spmv<0> : (const void *)k_bicg_spmv<1>,
             m->spmv_smem);
  if (mode == 0)
    k_bicg_spmv<0><<<grid, TPB, m->spmv_smem, ctx().stream>>>(k);
  else
    k_bicg_spmv<1><<<grid, TPB, m->spmv_smem, ctx().stream>>>(k);
  return post_launch("bicg_spmv");
}

} // anonymous namespace

bool launch_bicg_init(int n, ItState *st, const double *rhs, double *vx, double *res0,
                      double *rk, double *pk, double *uk, bool have_ax)
{
  if (n < 1) return true;
  k_bicg_init<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, rhs, vx, res0, rk, pk, uk,
                                                     have_ax ? 1 : 0);
  return post_launch("bicg_init");
}

bool launch_bicg_dots_a(int n, ItState *st, const double *rk, const double *res0)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_bicg_dots_a<<<g, TPB, 0, c.stream>>>(n, st, rk, res0, c.partials, c.ticket);
  return post_launch("bicg_dots_a");
}

bool launch_bicg_pk_zk(int n, ItState *st, const double *rk, const double *uk,
                       const double *ad_inv, double *pk, double *zk)
{
  if (n < 1) return true;
  k_bicg_pk_zk<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, rk, uk, ad_inv, pk, zk);
  return post_launch("bicg_pk_zk");
}

bool launch_bicg_spmv_b(const _csb_matrix_t *m, ItState *st, const double *zk, double *uk,
                        const double *res0)
{ return bicg_spmv_launch(0, m, st, zk, uk, res0); }

bool launch_bicg_upd1(int n, ItState *st, const double *ad_inv, double *vx, double *rk,
                      double *zk, const double *uk)
{
  if (n < 1) return true;
  k_bicg_upd1<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, ad_inv, vx, rk, zk, uk);
  return post_launch("bicg_upd1");
}

bool launch_bicg_spmv_c(const _csb_matrix_t *m, ItState *st, const double *zk, double *vk,
                        const double *rk)
{ return bicg_spmv_launch(1, m, st, zk, vk, rk); }

bool launch_bicg_upd2(int n, ItState *st, double *vx, double *rk, const double *zk,
                      const double *vk, const double *res0)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_bicg_upd2<<<g, TPB, 0, c.stream>>>(n, st, vx, rk, zk, vk, res0, c.partials, c.ticket);
  return post_launch("bicg_upd2");
}

/*--------------------------------------------------------------------------
 * Multigrid transfer operators and cycle bookkeeping
 *--------------------------------------------------------------------------*/

namespace {

/* Deterministic restriction: coarse row = sum of its fine rows in ascending
   fine-row order, i.e. exactly the order of the reference's sequential
   "simple sum" loop (src/alge/cs_grid.cpp:9025-9030); the reference's own
   device path uses atomics and is not reproducible. */
__global__ void __launch_bounds__(TPB)
k_restrict(int c_n_rows, const int *__restrict__ c_f_row_index,
           const int *__restrict__ c_f_row_ids, const double *__restrict__ f_var,
           double *__restrict__ c_var, const MgState *gate)
{
  if (gate != nullptr && gate->cvg != CSB_SLES_ITERATING) return;
  for (int ic = blockIdx.x*TPB + threadIdx.x; ic < c_n_rows; ic += gridDim.x*TPB) {
    double s = 0.;
    const int rs = c_f_row_index[ic], re = c_f_row_index[ic+1];
    for (int r = rs; r < re; r++)
      s += f_var[c_f_row_ids[r]];
    c_var[ic] = s;
  }
}

/* src/alge/cs_grid.cpp:9156-9166 */
__global__ void __launch_bounds__(TPB)
k_prolong_inc(int f_n_rows, const int *__restrict__ coarse_row,
              const double *__restrict__ c_var, double *__restrict__ f_var,
              const MgState *gate)
{
  if (gate != nullptr && gate->cvg != CSB_SLES_ITERATING) return;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < f_n_rows; i += gridDim.x*TPB) {
    const int ic = coarse_row[i];
    if (ic >= 0) f_var[i] += c_var[ic];
  }
}

__global__ void
k_mg_state_init(MgState *mg, double precision, double r_norm, int n_max_cycles)
{
  mg->precision = precision; mg->r_norm = r_norm;
  mg->residual = -1.; mg->initial_residual = -1.;
  mg->cvg = CSB_SLES_ITERATING;
  mg->cycle_id = 0; mg->n_max_cycles = n_max_cycles;
  mg->n_equiv_iter = 0; mg->n_cycles_res = 0; mg->diverged_level = -1;
  for (int l = 0; l < PROF_LEVELS; l++)
    for (int p = 0; p < PROF_PHASES; p++)
      mg->prof[l][p] = 0ull;
  for (int l = 0; l < PROF_LEVELS; l++)
    for (int p = 0; p < PROF_IT_PHASES; p++)
      mg->prof_it[l][p] = 0ull;
}

__global__ void
k_mg_cycle_begin(MgState *mg, int cycle_id)
{
  if (mg->cvg == CSB_SLES_ITERATING) mg->cycle_id = cycle_id;
}

/* *n_equiv_iter += n_iter * n_g_rows * denom_n_g_rows_0 (cs_multigrid.cpp:3341,
   :3414, :3514), int += double truncation included; a diverged smoother ends
   the cycle (:3285-3288). */
__global__ void
k_mg_accum_iter(MgState *mg, const ItState *st, int level, double n_rows_ratio,
                int take_initial_residual)
{
  if (st->cvg == 2) return;   /* skipped */
  /* gate already closed by the level-0 residual test: the reference still
     counts the level-0 descent smoother of that cycle (:3324-3333) */
  mg->n_equiv_iter = (int)((double)mg->n_equiv_iter + st->n_iter * n_rows_ratio);
  if (take_initial_residual && mg->cycle_id == 1)
    mg->initial_residual = st->initial_residual;
  if (st->cvg < CSB_SLES_BREAKDOWN && mg->cvg == CSB_SLES_ITERATING) {
    mg->cvg = CSB_SLES_DIVERGED;
    mg->diverged_level = level;
  }
}

} // anonymous namespace

bool launch_restrict(int c_n_rows, const int *c_f_row_index, const int *c_f_row_ids,
                     const double *f_var, double *c_var, const MgState *gate)
{
  if (c_n_rows < 1) return true;
  k_restrict<<<vec_grid(c_n_rows*4), TPB, 0, ctx().stream>>>(c_n_rows, c_f_row_index,
                                                            c_f_row_ids, f_var, c_var, gate);
  return post_launch("restrict");
}

bool launch_prolong_inc(int f_n_rows, const int *coarse_row, const double *c_var,
                        double *f_var, const MgState *gate)
{
  if (f_n_rows < 1) return true;
  k_prolong_inc<<<vec_grid(f_n_rows*4), TPB, 0, ctx().stream>>>(f_n_rows, coarse_row, c_var,
                                                               f_var, gate);
  return post_launch("prolong");
}

bool launch_mg_state_init(MgState *mg, double precision, double r_norm, int n_max_cycles)
{
  k_mg_state_init<<<1, 1, 0, ctx().stream>>>(mg, precision, r_norm, n_max_cycles);
  return post_launch("mg_state_init");
}

namespace {
__global__ void k_mg_set_initial_residual(MgState *mg, double v) { mg->initial_residual = v; }
}

/* a cs_multigrid_smoother reports no initial residual: HUGE_VAL (cs_multigrid.cpp:3270-3276) */
bool launch_mg_set_initial_residual(MgState *mg, double v)
{
  k_mg_set_initial_residual<<<1, 1, 0, ctx().stream>>>(mg, v);
  return post_launch("mg_set_initial_residual");
}

bool launch_mg_cycle_begin(MgState *mg, int cycle_id)
{
  k_mg_cycle_begin<<<1, 1, 0, ctx().stream>>>(mg, cycle_id);
  return post_launch("mg_cycle_begin");
}

bool launch_mg_accum_iter(MgState *mg, const ItState *st, int level, double n_rows_ratio,
                          int take_initial_residual)
{
  k_mg_accum_iter<<<1, 1, 0, ctx().stream>>>(mg, st, level, n_rows_ratio,
                                             take_initial_residual);
  return post_launch("mg_accum_iter");
}

/*--------------------------------------------------------------------------
 * Coarse level assembly (cs_grid.cpp _compute_coarse_quantities_msr_with_faces :6226-6990,
 * _compute_coarse_quantities_msr :5969-6090).
 *
 * What must be reproduced is a set of ORDERED sums: every coarse face (an upper entry of a
 * coarse row) adds up the fine faces between its two aggregates in the order "fine rows of the
 * aggregate ascending, entries of a fine row in CSR order", and the coarse diagonal adds up the
 * fine diagonal and every fine entry in that same order.  The reference walks this order with
 * one thread per coarse row and updates the face arrays as it goes.
 *
 * Here a WARP owns a coarse row and splits the work by what is parallel in it:
 *   stage:  the entries of the aggregate are enumerated in the canonical order -- fine rows
 *           one after the other, the lanes of the warp taking the entries of a row side by
 *           side -- and each lane computes ITS entry's target (the coarse row's entry it
 *           falls into, by binary search) and the values it contributes; target and values
 *           go to a shared-memory stage in canonical position;
 *   fold:   a lane owns a coarse ENTRY (lane k: entries k, k + 32, ...) and adds up, in
 *           canonical position order, the staged contributions aimed at it -- registers only,
 *           one store per coarse face, no read-modify-write of the face arrays, no atomics;
 *           one more lane runs the sequential chain of the diagonal over the same stage.
 * Aggregates larger than the stage are folded in several rounds (accumulators live in
 * registers across rounds).  The sums are the reference's, term by term (built with
 * -fmad=false), so the coarse matrices stay bit-identical to the CPU path.
 *--------------------------------------------------------------------------*/

namespace {

constexpr int CA_WARPS = 4;          /* coarse rows per CTA */
constexpr int CA_STAGE = 128;        /* staged contributions per round and warp */
constexpr int CA_OWN = 4;            /* coarse entries a lane can own: rows up to 128 entries */

/* position of id in the sorted row a[0..size), -1 if absent */
__device__ __forceinline__ int sorted_find(int size, int id, const int *__restrict__ a)
{
  int lo = 0, hi = size - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int v = a[mid];
    if (v < id) lo = mid + 1;
    else if (v > id) hi = mid - 1;
    else return mid;
  }
  return -1;
}

struct CaStage {
  int    target[CA_STAGE];           /* coarse entry (offset in its row); -1 none, -2 diagonal */
  double w[CA_STAGE];                /* value added to the diagonal chain / Galerkin value */
  double v[7][CA_STAGE];             /* face normal (3), xa0, xa0ij (3) */
};

/* Fold of one round: lane `lane` owns the coarse entries lane, lane + 32, ... of its row
   (those that are faces) and gathers the staged contributions aimed at them, in order. */
__device__ __forceinline__ void ca_fold7(const CaStage &S, int n_staged, int lane,
                                         const bool is_face[CA_OWN], double acc[CA_OWN][7])
{
#pragma unroll
  for (int o = 0; o < CA_OWN; o++) {
    if (!is_face[o]) continue;
    const int k = lane + 32*o;
    for (int p = 0; p < n_staged; p++) {
      if (S.target[p] == k) {
#pragma unroll
        for (int q = 0; q < 7; q++) acc[o][q] += S.v[q][p];
      }
    }
  }
}

/* Sums over the fine rows of every aggregate: cell centre and volume, the diagonal chain, and
   (faces) the coarse face normals, xa0 and xa0ij.  Rows with more than 32*CA_OWN entries are
   left to k_coarse_sums_wide. */
__global__ void __launch_bounds__(CA_WARPS*32)
k_coarse_sums(const CoarseAssemblyArgs a, int wide_limit, int *wide_rows, int *n_wide)
{
  __shared__ CaStage s_stage[CA_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int crow = blockIdx.x*CA_WARPS + wid;
  if (crow >= a.c_n_rows) return;                 /* whole warp */
  CaStage &S = s_stage[wid];

  const int c_beg = a.c_row_index[crow];
  const int c_len = a.c_row_index[crow+1] - c_beg;
  const int *__restrict__ c_cols = a.c_col_id + c_beg;
  if (c_len > wide_limit) {                       /* rare: handled by the wide kernel */
    if (lane == 0) wide_rows[atomicAdd(n_wide, 1)] = crow;
    return;
  }
  bool is_face[CA_OWN];
  double acc[CA_OWN][7];
#pragma unroll
  for (int o = 0; o < CA_OWN; o++) {
    const int k = lane + 32*o;
    is_face[o] = (k < c_len && c_cols[k] > crow);   /* faces are the upper entries */
#pragma unroll
    for (int q = 0; q < 7; q++) acc[o][q] = 0.;
  }
  double diag = 0.;                               /* lane 31: the coarse diagonal chain */

  const int r_beg = a.c_f_row_index[crow], r_end = a.c_f_row_index[crow+1];
  int n_staged = 0;
  for (int r = r_beg; r < r_end; r++) {
    const int frow = a.c_f_row_ids[r];
    const int e_beg = a.f_row_index[frow], e_end = a.f_row_index[frow+1];
    const double xi0 = a.f_cell_cen[frow*3], xi1 = a.f_cell_cen[frow*3+1],
                 xi2 = a.f_cell_cen[frow*3+2];
    /* the fine diagonal opens the row's part of the chain */
    int e = e_beg - 1;                            /* -1: the diagonal marker */
    while (e < e_end) {
      const int room = CA_STAGE - n_staged;
      const int n_now = min(room, e_end - e);
      const int mine = e + lane;
      for (int q = mine; q < e + n_now; q += 32) {
        const int p = n_staged + (q - e);
        if (q < e_beg) { S.target[p] = -2; S.w[p] = a.f_d_val[frow]; continue; }
        const int fcol = a.f_col_id[q];
        const int ccol = a.f_c_row[fcol];
        const double fx = a.f_x_val[q];
        S.w[p] = fx;
        int tgt = -1;
        if (ccol > crow) {
          tgt = sorted_find(c_len, ccol, c_cols);
          const int ff = a.f_cell_face[q];
          const int sg = a.f_cell_face_sgn[q];
          if (a.f_level == 0) {
            const double ss = sg * a.f_face_surf[ff];
            S.v[0][p] = ss * a.f_face_u_normal[ff*3];
            S.v[1][p] = ss * a.f_face_u_normal[ff*3+1];
            S.v[2][p] = ss * a.f_face_u_normal[ff*3+2];
            S.v[3][p] = fx;
            S.v[4][p] = fx * (a.f_cell_cen[fcol*3]   - xi0);
            S.v[5][p] = fx * (a.f_cell_cen[fcol*3+1] - xi1);
            S.v[6][p] = fx * (a.f_cell_cen[fcol*3+2] - xi2);
          }
          else {
            S.v[0][p] = sg * a.f_face_normal[ff*3];
            S.v[1][p] = sg * a.f_face_normal[ff*3+1];
            S.v[2][p] = sg * a.f_face_normal[ff*3+2];
            S.v[3][p] = a.f_xa0[ff];
            S.v[4][p] = sg * a.f_xa0ij[ff*3];
            S.v[5][p] = sg * a.f_xa0ij[ff*3+1];
            S.v[6][p] = sg * a.f_xa0ij[ff*3+2];
          }
        }
        S.target[p] = tgt;
      }
      n_staged += n_now;
      e += n_now;
      if (n_staged == CA_STAGE) {                 /* stage full: fold it */
        __syncwarp();
        ca_fold7(S, n_staged, lane, is_face, acc);
        if (lane == 31) for (int p = 0; p < n_staged; p++) diag += S.w[p];
        __syncwarp();
        n_staged = 0;
      }
    }
  }
  __syncwarp();
  ca_fold7(S, n_staged, lane, is_face, acc);
  if (lane == 31) {
    for (int p = 0; p < n_staged; p++) diag += S.w[p];
    a.c_d_val[crow] = diag;
  }
#pragma unroll
  for (int o = 0; o < CA_OWN; o++) {
    if (!is_face[o]) continue;
    const int face = a.c_cell_face[c_beg + lane + 32*o];
    a.c_face_normal[3*(size_t)face]   = acc[o][0];
    a.c_face_normal[3*(size_t)face+1] = acc[o][1];
    a.c_face_normal[3*(size_t)face+2] = acc[o][2];
    a.c_xa0[face] = acc[o][3];
    a.c_xa0ij[3*(size_t)face]   = acc[o][4];
    a.c_xa0ij[3*(size_t)face+1] = acc[o][5];
    a.c_xa0ij[3*(size_t)face+2] = acc[o][6];
  }
  /* cell centre and volume: four short chains over the fine rows, one lane each (the stage
     is reused for the exchange: every lane must be done reading it) */
  __syncwarp();
  if (lane < 4) {
    double sum = 0.;
    for (int r = r_beg; r < r_end; r++) {
      const int frow = a.c_f_row_ids[r];
      const double fv = a.f_cell_vol[frow];
      if (lane < 3) sum += a.f_cell_cen[frow*3 + lane]*fv;
      else sum += fv;
    }
    S.w[lane] = sum;
  }
  __syncwarp();
  if (lane < 3) a.c_cell_cen[crow*3 + lane] = S.w[lane] / S.w[3];
  else if (lane == 3) a.c_cell_vol[crow] = S.w[3];
}

/* Same sums for the coarse rows with more entries than a warp's lanes can own: one thread per
   row walks the canonical order and adds into the face arrays (zeroed by the launcher) -- a
   handful of rows at most on the hierarchies of this path. */
__global__ void __launch_bounds__(64)
k_coarse_sums_wide(const CoarseAssemblyArgs a, const int *__restrict__ wide_rows,
                   const int *__restrict__ n_wide)
{
  const int t = blockIdx.x*64 + threadIdx.x;
  if (t >= *n_wide) return;
  const int crow = wide_rows[t];
  const int c_beg = a.c_row_index[crow];
  const int c_len = a.c_row_index[crow+1] - c_beg;
  const int *__restrict__ c_cols = a.c_col_id + c_beg;
  for (int k = 0; k < c_len; k++)
    if (c_cols[k] > crow) {
      const int face = a.c_cell_face[c_beg + k];
      for (int q = 0; q < 3; q++) { a.c_face_normal[3*(size_t)face+q] = 0.; a.c_xa0ij[3*(size_t)face+q] = 0.; }
      a.c_xa0[face] = 0.;
    }
  double diag = 0., vol = 0., cen[3] = {0., 0., 0.};
  for (int r = a.c_f_row_index[crow]; r < a.c_f_row_index[crow+1]; r++) {
    const int frow = a.c_f_row_ids[r];
    const double fv = a.f_cell_vol[frow];
    const double xi[3] = {a.f_cell_cen[frow*3], a.f_cell_cen[frow*3+1], a.f_cell_cen[frow*3+2]};
    for (int q = 0; q < 3; q++) cen[q] += xi[q]*fv;
    vol += fv;
    diag += a.f_d_val[frow];
    for (int e = a.f_row_index[frow]; e < a.f_row_index[frow+1]; e++) {
      const int fcol = a.f_col_id[e];
      const int ccol = a.f_c_row[fcol];
      const double fx = a.f_x_val[e];
      diag += fx;
      if (ccol > crow) {
        const int face = a.c_cell_face[c_beg + sorted_find(c_len, ccol, c_cols)];
        const int ff = a.f_cell_face[e];
        const int sg = a.f_cell_face_sgn[e];
        double *fn = a.c_face_normal + 3*(size_t)face, *fxij = a.c_xa0ij + 3*(size_t)face;
        if (a.f_level == 0) {
          const double ss = sg * a.f_face_surf[ff];
          for (int q = 0; q < 3; q++) fn[q] += ss * a.f_face_u_normal[ff*3+q];
          a.c_xa0[face] += fx;
          for (int q = 0; q < 3; q++) fxij[q] += fx * (a.f_cell_cen[fcol*3+q] - xi[q]);
        }
        else {
          for (int q = 0; q < 3; q++) fn[q] += sg * a.f_face_normal[ff*3+q];
          a.c_xa0[face] += a.f_xa0[ff];
          for (int q = 0; q < 3; q++) fxij[q] += sg * a.f_xa0ij[ff*3+q];
        }
      }
    }
  }
  for (int q = 0; q < 3; q++) a.c_cell_cen[crow*3+q] = cen[q] / vol;
  a.c_cell_vol[crow] = vol;
  a.c_d_val[crow] = diag;
}

/* Extra-diagonal values from the face quantities (cs_grid.cpp:6848-6895): a thread per coarse
   ENTRY computes its value (they are independent of one another); the diagonal then takes
   them off in entry order, one lane per row walking the values its warp has just written. */
__global__ void __launch_bounds__(CA_WARPS*32)
k_coarse_values(const CoarseAssemblyArgs a)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int crow = blockIdx.x*CA_WARPS + wid;
  if (crow >= a.c_n_rows) return;
  const double xi0 = a.c_cell_cen[crow*3], xi1 = a.c_cell_cen[crow*3+1],
               xi2 = a.c_cell_cen[crow*3+2];
  const int beg = a.c_row_index[crow], end = a.c_row_index[crow+1];
  for (int k = beg + lane; k < end; k += 32) {
    const int ccol = a.c_col_id[k];
    const int face = a.c_cell_face[k];
    const double sg = (double)a.c_cell_face_sgn[k];
    const double n0 = a.c_face_normal[3*(size_t)face], n1 = a.c_face_normal[3*(size_t)face+1],
                 n2 = a.c_face_normal[3*(size_t)face+2];
    const double dist_n
      = sg * (  (a.c_cell_cen[ccol*3]   - xi0)*n0
              + (a.c_cell_cen[ccol*3+1] - xi1)*n1
              + (a.c_cell_cen[ccol*3+2] - xi2)*n2);
    const double xa0 = a.c_xa0[face];
    double val = xa0;
    if (fabs(dist_n) > 1e-12) {          /* cs_math_epzero */
      const double flux_n = a.c_xa0ij[3*(size_t)face]*n0 + a.c_xa0ij[3*(size_t)face+1]*n1
                          + a.c_xa0ij[3*(size_t)face+2]*n2;
      const double ratio = flux_n/dist_n;
      if (ratio >= val && ratio <= 0.) val = ratio;
    }
    a.c_x_val[k] = a.relax*val + (1. - a.relax)*xa0;
  }
  __syncwarp();
  if (lane == 0) {
    double d = a.c_d_val[crow];
    for (int k = beg; k < end; k++) d -= a.c_x_val[k];
    a.c_d_val[crow] = d;
  }
}

/* Galerkin coarse matrix, relaxation parameter <= 0 (_compute_coarse_quantities_msr, scalar
   case, cs_grid.cpp:6040-6090): the same stage / fold, one value per contribution; entries
   inside the aggregate join the diagonal chain. */
__global__ void __launch_bounds__(CA_WARPS*32)
k_coarse_galerkin(const CoarseAssemblyArgs a)
{
  __shared__ int    s_target[CA_WARPS][CA_STAGE];
  __shared__ double s_w[CA_WARPS][CA_STAGE];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int crow = blockIdx.x*CA_WARPS + wid;
  if (crow >= a.c_n_rows) return;
  int *tg = s_target[wid];
  double *w = s_w[wid];
  const int c_beg = a.c_row_index[crow];
  const int c_len = a.c_row_index[crow+1] - c_beg;
  const int *__restrict__ c_cols = a.c_col_id + c_beg;
  /* a lane's coarse entries: k = lane, lane + 32, ...; the sums live in c_x_val between
     rounds, so a row may have any number of entries */
  for (int k = lane; k < c_len; k += 32) a.c_x_val[c_beg + k] = 0.;
  double diag = 0.;
  int n_staged = 0;
  const int r_beg = a.c_f_row_index[crow], r_end = a.c_f_row_index[crow+1];
  for (int r = r_beg; r < r_end; r++) {
    const int frow = a.c_f_row_ids[r];
    const int e_beg = a.f_row_index[frow], e_end = a.f_row_index[frow+1];
    int e = e_beg - 1;                            /* -1: the fine diagonal opens the row */
    while (e < e_end) {
      const int n_now = min(CA_STAGE - n_staged, e_end - e);
      for (int q = e + lane; q < e + n_now; q += 32) {
        const int p = n_staged + (q - e);
        if (q < e_beg) { tg[p] = -2; w[p] = a.f_d_val[frow]; continue; }
        const int ccol = a.f_c_row[a.f_col_id[q]];
        w[p] = a.f_x_val[q];
        tg[p] = (ccol < 0) ? -1 : ((ccol == crow) ? -2 : sorted_find(c_len, ccol, c_cols));
      }
      n_staged += n_now;
      e += n_now;
      if (n_staged == CA_STAGE) {
        __syncwarp();
        for (int k = lane; k < c_len; k += 32) {
          double sum = a.c_x_val[c_beg + k];
          for (int p = 0; p < n_staged; p++) if (tg[p] == k) sum += w[p];
          a.c_x_val[c_beg + k] = sum;
        }
        if (lane == 31) for (int p = 0; p < n_staged; p++) if (tg[p] == -2) diag += w[p];
        __syncwarp();
        n_staged = 0;
      }
    }
  }
  __syncwarp();
  for (int k = lane; k < c_len; k += 32) {
    double sum = a.c_x_val[c_beg + k];
    for (int p = 0; p < n_staged; p++) if (tg[p] == k) sum += w[p];
    a.c_x_val[c_beg + k] = sum;
  }
  if (lane == 31) for (int p = 0; p < n_staged; p++) if (tg[p] == -2) diag += w[p];
  if (lane == 31) a.c_d_val[crow] = diag;
}

} // anonymous namespace

bool launch_coarse_galerkin(const CoarseAssemblyArgs &a)
{
  if (a.c_n_rows < 1) return true;
  k_coarse_galerkin<<<(a.c_n_rows + CA_WARPS - 1)/CA_WARPS, CA_WARPS*32, 0, ctx().stream>>>(a);
  return post_launch("coarse_galerkin");
}

/*----------------------------------------------------------------------------
 * Coarse level of a matrix mapped from native (face-based) arrays.
 *
 * The reference sums in face order: every coarse face adds up its fine faces by ascending
 * fine face id, every fine cell its faces by ascending face id, every coarse cell its
 * coarse faces by ascending coarse face id (cs_grid.cpp:5303-5372, :5479-5486, :5527-5534).
 * Nothing here is sorted or staged: a thread finds the next face of its set by scanning
 * the few entries it owns for the smallest face id above the last one it took, so the
 * order is the reference's whatever the entry order is.
 *----------------------------------------------------------------------------*/

namespace {

__global__ void k_native_cells(NativeCoarseArgs a)
{
  const int ic = blockIdx.x*blockDim.x + threadIdx.x;
  if (ic >= a.c_n_rows) return;
  double vol = 0., cx = 0., cy = 0., cz = 0.;
  for (int r = a.c_f_row_index[ic]; r < a.c_f_row_index[ic+1]; r++) {
    const int ii = a.c_f_row_ids[r];
    const double v = a.f_cell_vol[ii];
    vol += v;
    cx += v*a.f_cell_cen[3*(size_t)ii];
    cy += v*a.f_cell_cen[3*(size_t)ii+1];
    cz += v*a.f_cell_cen[3*(size_t)ii+2];
  }
  a.c_cell_vol[ic] = vol;
  a.c_cell_cen[3*(size_t)ic]   = cx / vol;
  a.c_cell_cen[3*(size_t)ic+1] = cy / vol;
  a.c_cell_cen[3*(size_t)ic+2] = cz / vol;
}

/* non-differential part of a fine row: diagonal plus its faces' values in face order */
__global__ void k_native_w1(NativeCoarseArgs a)
{
  const int ii = blockIdx.x*blockDim.x + threadIdx.x;
  if (ii >= a.f_n_rows) return;
  const int s = a.f_row_index[ii], e = a.f_row_index[ii+1];
  double w = a.f_d_val[ii];
  int last = -1;
  while (true) {
    int best = INT_MAX, bk = -1;
    for (int k = s; k < e; k++) {
      const int f = a.f_cell_face[k];
      if (f > last && f < best) { best = f; bk = k; }
    }
    if (bk < 0) break;
    w += a.f_x_val[bk];
    last = best;
  }
  a.w1[ii] = w;
}

__global__ void k_native_faces(NativeCoarseArgs a)
{
  const int cf = blockIdx.x*blockDim.x + threadIdx.x;
  if (cf >= a.c_n_faces) return;
  const int ic = a.c_face_cell[2*(size_t)cf], jc = a.c_face_cell[2*(size_t)cf+1];
  double xa0 = 0., nx = 0., ny = 0., nz = 0., ax = 0., ay = 0., az = 0.;
  int last = -1;
  while (true) {
    /* next fine face between the two aggregates */
    int best = INT_MAX, bk = -1;
    for (int r = a.c_f_row_index[ic]; r < a.c_f_row_index[ic+1]; r++) {
      const int ii = a.c_f_row_ids[r];
      for (int k = a.f_row_index[ii]; k < a.f_row_index[ii+1]; k++) {
        if (a.f_c_row[a.f_col_id[k]] != jc) continue;
        const int f = a.f_cell_face[k];
        if (f > last && f < best) { best = f; bk = k; }
      }
    }
    if (bk < 0) break;
    last = best;
    const size_t f = (size_t)best;
    const int c0 = a.f_face_cell[2*f], c1 = a.f_face_cell[2*f+1];
    const double sgn = (a.f_c_row[c0] == ic) ? 1. : -1.;
    if (a.f_level == 0) {
      const double v = a.f_x_val[bk];                       /* f_xa0 = xa */
      const double sx = sgn * (a.f_cell_cen[3*(size_t)c1]   - a.f_cell_cen[3*(size_t)c0]);
      const double sy = sgn * (a.f_cell_cen[3*(size_t)c1+1] - a.f_cell_cen[3*(size_t)c0+1]);
      const double sz = sgn * (a.f_cell_cen[3*(size_t)c1+2] - a.f_cell_cen[3*(size_t)c0+2]);
      xa0 += v;
      const double sgn_surf = sgn*a.f_face_surf[f];
      nx += a.f_face_u_normal[3*f]   * sgn_surf;
      ny += a.f_face_u_normal[3*f+1] * sgn_surf;
      nz += a.f_face_u_normal[3*f+2] * sgn_surf;
      ax += v * sx;
      ay += v * sy;
      az += v * sz;
    }
    else {
      xa0 += a.f_xa0[f];
      nx += sgn*a.f_face_normal[3*f];
      ny += sgn*a.f_face_normal[3*f+1];
      nz += sgn*a.f_face_normal[3*f+2];
      ax += sgn*a.f_xa0ij[3*f];
      ay += sgn*a.f_xa0ij[3*f+1];
      az += sgn*a.f_xa0ij[3*f+2];
    }
  }
  a.c_xa0[cf] = xa0;
  a.c_face_normal[3*(size_t)cf] = nx; a.c_face_normal[3*(size_t)cf+1] = ny;
  a.c_face_normal[3*(size_t)cf+2] = nz;
  a.c_xa0ij[3*(size_t)cf] = ax; a.c_xa0ij[3*(size_t)cf+1] = ay; a.c_xa0ij[3*(size_t)cf+2] = az;

  /* P1-interpolated value, clipped to [xa0, 0], relaxed towards the P0 value */
  const double *ci = a.c_cell_cen + 3*(size_t)ic, *cj = a.c_cell_cen + 3*(size_t)jc;
  const double dsigjg = (cj[0] - ci[0])*nx + (cj[1] - ci[1])*ny + (cj[2] - ci[2])*nz;
  const double dsxaij = ax*nx + ay*ny + az*nz;
  double x = xa0;
  if (fabs(dsigjg) > 1.e-12) {
    const double agij = dsxaij/dsigjg;
    x = (agij < xa0 || agij > 0.) ? xa0 : agij;
  }
  a.c_xa[cf] = a.relax*x + (1. - a.relax)*xa0;
}

/* coarse diagonal (fine rows ascending, then the coarse faces ascending) and the MSR values */
__global__ void k_native_diag(NativeCoarseArgs a)
{
  const int ic = blockIdx.x*blockDim.x + threadIdx.x;
  if (ic >= a.c_n_rows) return;
  double d = 0.;
  for (int r = a.c_f_row_index[ic]; r < a.c_f_row_index[ic+1]; r++) {
    const double w = a.w1[a.c_f_row_ids[r]];
    if (w > 0.) d += w;
  }
  const int s = a.c_row_index[ic], e = a.c_row_index[ic+1];
  int last = -1;
  while (true) {
    int best = INT_MAX, bk = -1;
    for (int k = s; k < e; k++) {
      const int f = a.c_cell_face[k];
      if (f > last && f < best) { best = f; bk = k; }
    }
    if (bk < 0) break;
    const double x = a.c_xa[best];
    d -= x;
    a.c_x_val[bk] = x;
    last = best;
  }
  a.c_d_val[ic] = d;
}

__global__ void k_native_to_msr(int nnz, const int *c2f, const double *xa, double *x_val)
{
  const int k = blockIdx.x*blockDim.x + threadIdx.x;
  if (k < nnz) x_val[k] = xa[c2f[k]];
}

} // anonymous namespace

bool launch_coarse_native(const NativeCoarseArgs &a)
{
  Context &c = ctx();
  if (a.c_n_rows < 1) return true;
  const int nt = 128;
  k_native_cells<<<(a.c_n_rows + nt - 1)/nt, nt, 0, c.stream>>>(a);
  if (!post_launch("native_cells")) return false;
  if (a.f_n_rows > 0) {
    k_native_w1<<<(a.f_n_rows + nt - 1)/nt, nt, 0, c.stream>>>(a);
    if (!post_launch("native_w1")) return false;
  }
  if (a.c_n_faces > 0) {
    k_native_faces<<<(a.c_n_faces + nt - 1)/nt, nt, 0, c.stream>>>(a);
    if (!post_launch("native_faces")) return false;
  }
  k_native_diag<<<(a.c_n_rows + nt - 1)/nt, nt, 0, c.stream>>>(a);
  return post_launch("native_diag");
}

namespace {
/* a thread per slot: sums of a few terms each (a row's diagonal gets one term per pair
   declared on that row), read through the plan in arrival order */
__global__ void k_assembler_fold(int n_rows, int n_slots, const int *__restrict__ index,
                                 const int *__restrict__ src, const double *__restrict__ val,
                                 double *d_val, double *x_val)
{
  const int s = blockIdx.x*blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int b = index[s], e = index[s+1];
  if (b == e) return;
  double *out = (s < n_rows) ? d_val + s : x_val + (s - n_rows);
  double acc = *out;
  for (int j = b; j < e; j++) acc += val[src[j]];
  *out = acc;
}
} // anonymous namespace

bool launch_assembler_fold(int n_rows, int nnz, const int *index, const int *src,
                           const double *val, double *d_val, double *x_val)
{
  const long n_slots = (long)n_rows + (long)nnz;
  if (n_slots < 1) return true;
  k_assembler_fold<<<(unsigned)((n_slots + 255)/256), 256, 0, ctx().stream>>>(
    n_rows, (int)n_slots, index, src, val, d_val, x_val);
  return post_launch("assembler_fold");
}

namespace {
/* The face operators below gather per cell in ascending face order (the order of the
   reference's scatter loops).  fperm[row_index[c] + t] = position, within row c, of the entry
   with the t-th smallest face id (255 = no more entries), built once per mesh association by
   k_face_order; without it (rows longer than 255 entries) the next face is found by a scan of
   the row. */
struct FaceOrder {
  const int *c2f;
  const unsigned char *fperm;
  int s, e, t, last;
  __device__ FaceOrder(const int *c2f_, const unsigned char *fperm_, int s_, int e_)
    : c2f(c2f_), fperm(fperm_), s(s_), e(e_), t(0), last(-1) {}
  /* next face of the row: its id and the row entry that holds it */
  __device__ bool next(int &face, int &bk)
  {
    if (fperm != nullptr) {
      if (s + t >= e) return false;
      const unsigned int p = fperm[s + t];
      if (p == 255u) return false;
      t++;
      bk = s + (int)p;
      face = c2f[bk];
      return true;
    }
    int best = INT_MAX;
    bk = -1;
    for (int k = s; k < e; k++) {
      const int f = c2f[k];
      if (f > last && f < best) { best = f; bk = k; }
    }
    if (bk < 0) return false;
    last = best;
    face = best;
    return true;
  }
};

__global__ void k_face_order(int n_rows, const int *__restrict__ row_index,
                             const int *__restrict__ c2f, unsigned char *fperm, int *too_long)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  if (e - s > 255) { *too_long = 1; return; }
  FaceOrder fo(c2f, nullptr, s, e);
  int t = 0, face, bk;
  while (fo.next(face, bk)) fperm[s + t++] = (unsigned char)(bk - s);
  for (; s + t < e; t++) fperm[s + t] = 255;        /* entries repeating a face id */
}

/* cs_diffusion_potential without flux reconstruction (cs_convection_diffusion.cpp:8869-8889,
   :8975-8982): the reference scatters face fluxes to the two cells in face order; a thread owns a
   cell and takes its interior faces by ascending face id (the next one is the smallest face id
   above the last one among the row's entries), then its boundary faces (stored ascending) --
   the same additions in the same order, no atomics. */
__global__ void k_diffusion_potential(int n_rows, const int *__restrict__ row_index,
                                      const int *__restrict__ col_id,
                                      const int *__restrict__ c2f,
                                      const unsigned char *__restrict__ fperm,
                                      const short *__restrict__ c2f_sgn,
                                      const int *__restrict__ cell_b_faces_idx,
                                      const int *__restrict__ cell_b_faces, int init, int inc,
                                      const double *__restrict__ pvar,
                                      const double *__restrict__ cofaf,
                                      const double *__restrict__ cofbf,
                                      const double *__restrict__ i_visc,
                                      const double *__restrict__ b_visc, double *diverg)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  const double pc = pvar[c];
  double d = (init >= 1) ? 0. : diverg[c];
  FaceOrder fo(c2f, fperm, s, e);
  while (true) {
    int best, bk;
    if (!fo.next(best, bk)) break;
    const double pj = pvar[col_id[bk]];
    if (c2f_sgn[bk] > 0) d += i_visc[best]*(pc - pj);      /* this cell is the face's first */
    else d += -(i_visc[best]*(pj - pc));
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      const int f = cell_b_faces[k];
      const double flux_d = inc*cofaf[f] + cofbf[f]*pc;
      d += b_visc[f]*flux_d;
    }
  diverg[c] = d;
}
} // anonymous namespace

/* face-order permutation of the matrix's rows, built on first use after a mesh association
   (nullptr: not available, the kernels scan) */
static const unsigned char *face_order(const _csb_matrix_t *m)
{
  if (m->fperm_state == 1) return m->fperm.p;
  if (m->fperm_state < 0 || m->n_rows < 1 || m->nnz < 1) return nullptr;
  m->fperm_state = -1;
  DevBuf<int> flag;
  int h_flag = 1;
  if (!flag.alloc(1) || !m->fperm.alloc((size_t)m->nnz)
      || cudaMemsetAsync(flag.p, 0, sizeof(int), ctx().stream) != cudaSuccess) return nullptr;
  k_face_order<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(m->n_rows, m->row_index.p,
                                                                m->c2f.p, m->fperm.p, flag.p);
  if (!post_launch("face_order")
      || cudaMemcpyAsync(&h_flag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream)
         != cudaSuccess
      || cudaStreamSynchronize(ctx().stream) != cudaSuccess) return nullptr;
  if (h_flag != 0) return nullptr;
  m->fperm_state = 1;
  return m->fperm.p;
}

bool launch_diffusion_potential(const _csb_matrix_t *m, int init, int inc, const double *pvar,
                                const double *cofaf, const double *cofbf, const double *i_visc,
                                const double *b_visc, double *diverg)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const unsigned char *fperm = face_order(m);
  k_diffusion_potential<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, fperm, m->c2f_sgn.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr, init, inc, pvar,
    cofaf, cofbf, i_visc, b_visc, diverg);
  return post_launch("diffusion_potential");
}

namespace {

/* (xb - xa) . xc, cs_math_3_distance_dot_product (cs_math.h:171-176) */
__device__ __forceinline__ double ddp3(const double *xa, const double *xb, const double *xc)
{
  return ((xb[0] - xa[0])*xc[0]+(xb[1] - xa[1])*xc[1]+(xb[2] - xa[2])*xc[2]);
}

/* cocg of the least-squares gradient, one cell per thread: sum of dc (x) dc / |dc|^2 over the
   cell's neighbours in adjacency (= MSR) order, then of n (x) n over its boundary faces,
   ACCUMULATED IN SINGLE PRECISION like the reference's cs_cocg_t (each term is a double
   expression added to the float), inverted by the float Gauss-Jordan elimination with row
   pivoting of cs_math_matrix_gauss_inverse (cs_math.h:542-590). */
__global__ void k_lsq_cocg(int n_rows, const int *__restrict__ row_index,
                           const int *__restrict__ col_id, const double *__restrict__ cell_cen,
                           const int *__restrict__ cell_b_faces_idx,
                           const int *__restrict__ cell_b_faces,
                           const double *__restrict__ b_un, const double *__restrict__ b_dist,
                           const double *__restrict__ diipb, float *cocg)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const double cx = cell_cen[3*(size_t)c], cy = cell_cen[3*(size_t)c+1], cz = cell_cen[3*(size_t)c+2];
  for (int k = row_index[c]; k < row_index[c+1]; k++) {
    const size_t j = (size_t)col_id[k];
    const double d0 = cell_cen[3*j] - cx, d1 = cell_cen[3*j+1] - cy, d2 = cell_cen[3*j+2] - cz;
    const double ddc = 1. / (d0*d0 + d1*d1 + d2*d2);
    g[0] = (float)((double)g[0] + d0 * d0 * ddc);
    g[1] = (float)((double)g[1] + d1 * d1 * ddc);
    g[2] = (float)((double)g[2] + d2 * d2 * ddc);
    g[3] = (float)((double)g[3] + d0 * d1 * ddc);
    g[4] = (float)((double)g[4] + d1 * d2 * ddc);
    g[5] = (float)((double)g[5] + d0 * d2 * ddc);
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      /* direction n + II'/b_dist (cs_gradient.cpp:2632-2636); n alone when II' = 0 */
      const size_t f = (size_t)cell_b_faces[k];
      double d[3] = {b_un[3*f], b_un[3*f+1], b_un[3*f+2]};
      if (diipb != nullptr) {
        const double unddij = 1. / b_dist[f];
        d[0] = d[0] + unddij*diipb[3*f]; d[1] = d[1] + unddij*diipb[3*f+1];
        d[2] = d[2] + unddij*diipb[3*f+2];
      }
      g[0] = (float)((double)g[0] + d[0] * d[0]);
      g[1] = (float)((double)g[1] + d[1] * d[1]);
      g[2] = (float)((double)g[2] + d[2] * d[2]);
      g[3] = (float)((double)g[3] + d[0] * d[1]);
      g[4] = (float)((double)g[4] + d[1] * d[2]);
      g[5] = (float)((double)g[5] + d[0] * d[2]);
    }
  /* rows are exchanged through their index, like the pointer swaps of the reference */
  float t[3][6] = {{g[0], g[3], g[5], 1.f, 0.f, 0.f},
                   {g[3], g[1], g[4], 0.f, 1.f, 0.f},
                   {g[5], g[4], g[2], 0.f, 0.f, 1.f}};
  int p[3] = {0, 1, 2};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    for (int j = i+1; j < 3; j++) {
      float pii = 0.f, pji = 0.f;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        if (p[i] == r) pii = t[r][i];
        if (p[j] == r) pji = t[r][i];
      }
      if (fabsf(pii) < fabsf(pji)) { const int tmp = p[j]; p[j] = p[i]; p[i] = tmp; }
    }
    float ri[6];
#pragma unroll
    for (int r = 0; r < 3; r++)
      if (p[i] == r) {
#pragma unroll
        for (int k = 0; k < 6; k++) ri[k] = t[r][k];
      }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (j != i) {
#pragma unroll
        for (int r = 0; r < 3; r++)
          if (p[j] == r) {
            const float sc = t[r][i] / ri[i];
#pragma unroll
            for (int k = 0; k < 6; k++) {
              const float prod = ri[k] * sc;
              t[r][k] = t[r][k] - prod;
            }
          }
      }
    }
  }
  float b[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int r = 0; r < 3; r++)
      if (p[i] == r) {
        const float sc = 1.f / t[r][i];
#pragma unroll
        for (int j = 0; j < 3; j++) b[i][j] = t[r][3+j] * sc;
      }
  float *o = cocg + 6*(size_t)c;
  o[0] = b[0][0]; o[1] = b[1][1]; o[2] = b[2][2];
  o[3] = b[0][1]; o[4] = b[1][2]; o[5] = b[0][2];
}

/* the gradient itself: right-hand side gathered per cell in the reference's scatter order
   (interior faces by ascending id -- both cells of a face add the SAME vector dc * pfac
   computed in the face's orientation -- then boundary faces), times the cell's cocg */
__global__ void k_lsq_gradient(int n_rows, const int *__restrict__ row_index,
                               const int *__restrict__ col_id, const int *__restrict__ c2f,
                               const unsigned char *__restrict__ fperm,
                               const short *__restrict__ c2f_sgn,
                               const double *__restrict__ cell_cen,
                               const int *__restrict__ cell_b_faces_idx,
                               const int *__restrict__ cell_b_faces,
                               const double *__restrict__ b_un,
                               const double *__restrict__ b_dist,
                               const float *__restrict__ cocg, int inc,
                               const double *__restrict__ pvar,
                               const double *__restrict__ coefa,
                               const double *__restrict__ coefb,
                               const double *__restrict__ diipb,
                               const double *__restrict__ val_f,
                               const double *__restrict__ i_face_cog,
                               const double *__restrict__ b_face_cog,
                               const double *__restrict__ f_ext, double *grad)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  const double pc = pvar[c];
  double r0 = 0., r1 = 0., r2 = 0.;
  FaceOrder fo(c2f, fperm, s, e);
  while (true) {
    int best, bk;
    if (!fo.next(best, bk)) break;
    const int o = col_id[bk];
    const bool first = (c2f_sgn[bk] > 0);
    const size_t ii = (size_t)(first ? c : o), jj = (size_t)(first ? o : c);
    const double d0 = cell_cen[3*jj] - cell_cen[3*ii], d1 = cell_cen[3*jj+1] - cell_cen[3*ii+1],
                 d2 = cell_cen[3*jj+2] - cell_cen[3*ii+2];
    double pfac;
    if (f_ext == nullptr)
      pfac = (pvar[jj] - pvar[ii]) / (d0*d0 + d1*d1 + d2*d2);
    else {                  /* _lsq_scalar_gradient_hyd_p (cs_gradient.cpp:3420-3445) */
      const double dvarij = pvar[jj] - pvar[ii];
      const double *cog = i_face_cog + 3*(size_t)best;
      pfac =   dvarij
             + ddp3(cog, cell_cen + 3*ii, f_ext + 3*ii)
             - ddp3(cog, cell_cen + 3*jj, f_ext + 3*jj);
      pfac /= (d0*d0 + d1*d1 + d2*d2);
    }
    r0 += d0 * pfac; r1 += d1 * pfac; r2 += d2 * pfac;
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      const int f = cell_b_faces[k];
      /* face value from the value at I' (_compute_bc_flux_grad,
         cs_boundary_conditions_set_coeffs.cpp:168-170); I' = I when II' = 0 */
      const double vf = (val_f != nullptr) ? val_f[f] : inc*coefa[f] + coefb[f]*pc;
      const double ddif = 1. / b_dist[f];
      double d0 = b_un[3*(size_t)f], d1 = b_un[3*(size_t)f+1], d2 = b_un[3*(size_t)f+2];
      if (diipb != nullptr) {
        d0 = d0 + ddif*diipb[3*(size_t)f]; d1 = d1 + ddif*diipb[3*(size_t)f+1];
        d2 = d2 + ddif*diipb[3*(size_t)f+2];
      }
      double pfac;
      if (f_ext == nullptr)
        pfac = (vf - pc) * ddif;
      else {                /* (:3505-3527; no porosity term) */
        const double if_dot_fext = ddp3(cell_cen + 3*(size_t)c, b_face_cog + 3*(size_t)f,
                                        f_ext + 3*(size_t)c);
        pfac = ddif * (vf - pc - 0. - if_dot_fext);
      }
      r0 += d0 * pfac; r1 += d1 * pfac; r2 += d2 * pfac;
    }
  const float *g = cocg + 6*(size_t)c;
  if (f_ext == nullptr) {
    grad[3*(size_t)c]   = g[0]*r0 + g[3]*r1 + g[5]*r2;
    grad[3*(size_t)c+1] = g[3]*r0 + g[1]*r1 + g[4]*r2;
    grad[3*(size_t)c+2] = g[5]*r0 + g[4]*r1 + g[2]*r2;
  }
  else {
    const double *fx = f_ext + 3*(size_t)c;
    grad[3*(size_t)c]   = g[0]*r0 + g[3]*r1 + g[5]*r2 + fx[0];
    grad[3*(size_t)c+1] = g[3]*r0 + g[1]*r1 + g[4]*r2 + fx[1];
    grad[3*(size_t)c+2] = g[5]*r0 + g[4]*r1 + g[2]*r2 + fx[2];
  }
}

/* cs_gradient_boundary_iprime_lsq_s (cs_gradient_boundary.cpp:197-640; no limiter, no
   weighting, standard neighbourhood): value of pvar at I' of the boundary faces, and the face
   values _compute_bc_flux_grad derives from them (cs_boundary_conditions_set_coeffs.cpp:
   168-192).  The local least-squares system of a face depends on its cell only, so one thread
   per cell builds it once (neighbours in adjacency order, then the cell's boundary faces in
   ascending order with the direction n + (1 - b) II'/b_dist; cocg in single precision, its
   cofactors in float arithmetic, the solve in double, :560-582) and serves the cell's faces;
   faces with |II'|^2 < 1e-6 b_dist^2 take the cell value (:285-291).  coefa enters the local
   system without inc, as in the reference.
   f_ext != nullptr: hyd_p_flag = 1 -- the interior differences gain (I - F).f_I - (J - F).f_J
   with F the centre of the shared face (c2f = cell_i_faces), the boundary terms
   (b - 1) (F_b - I).f_I, f_I is added to the local gradient, and val_f gains b (F_b - I').f_I.
   diipb == nullptr: II' = 0 (every face takes the cell value).  val_f may be nullptr. */
__global__ void k_b_iprime_lsq(int n_rows, const int *__restrict__ row_index,
                               const int *__restrict__ col_id, const int *__restrict__ c2f,
                               const double *__restrict__ cell_cen,
                               const int *__restrict__ cell_b_faces_idx,
                               const int *__restrict__ cell_b_faces,
                               const double *__restrict__ b_un,
                               const double *__restrict__ b_dist,
                               const double *__restrict__ diipb,
                               const double *__restrict__ i_face_cog,
                               const double *__restrict__ b_face_cog,
                               const double *__restrict__ f_ext, double clip_coeff, int inc,
                               const double *__restrict__ pvar,
                               const double *__restrict__ coefa,
                               const double *__restrict__ coefb, double *val_ip, double *val_f)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int bs = cell_b_faces_idx[c], be = cell_b_faces_idx[c+1];
  if (bs >= be) return;
  const double p_c = pvar[c];
  const double cen[3] = {cell_cen[3*(size_t)c], cell_cen[3*(size_t)c+1], cell_cen[3*(size_t)c+2]};
  double fx[3] = {0., 0., 0.};
  if (f_ext != nullptr) {
    fx[0] = f_ext[3*(size_t)c]; fx[1] = f_ext[3*(size_t)c+1]; fx[2] = f_ext[3*(size_t)c+2];
  }
  bool any = false;
  if (diipb != nullptr)
    for (int k = bs; k < be; k++) {
      const size_t f = (size_t)cell_b_faces[k];
      const double *ip = diipb + 3*f;
      const double bd = b_dist[f];
      if (!(ip[0]*ip[0] + ip[1]*ip[1] + ip[2]*ip[2] < (bd*bd) * (1e-3 * 1e-3))) any = true;
    }
  double gx = 0., gy = 0., gz = 0.;
  /* bounds of the simple limiter (:612-628; clip_coeff < 0: none): the cell's, its neighbours'
     and its boundary faces' unreconstructed values */
  double var_min = p_c, var_max = p_c;
  if (any) {
    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    double r0 = 0., r1 = 0., r2 = 0.;
    for (int k = row_index[c]; k < row_index[c+1]; k++) {
      const size_t j = (size_t)col_id[k];
      var_min = fmin(var_min, pvar[j]);
      var_max = fmax(var_max, pvar[j]);
      const double cj[3] = {cell_cen[3*j], cell_cen[3*j+1], cell_cen[3*j+2]};
      const double d0 = cj[0] - cen[0], d1 = cj[1] - cen[1], d2 = cj[2] - cen[2];
      const double inv_d2 = 1. / (d0*d0 + d1*d1 + d2*d2);
      g[0] = (float)((double)g[0] + d0 * d0 * inv_d2);
      g[1] = (float)((double)g[1] + d1 * d1 * inv_d2);
      g[2] = (float)((double)g[2] + d2 * d2 * inv_d2);
      g[3] = (float)((double)g[3] + d0 * d1 * inv_d2);
      g[4] = (float)((double)g[4] + d1 * d2 * inv_d2);
      g[5] = (float)((double)g[5] + d0 * d2 * inv_d2);
      double t = pvar[j] - p_c;
      if (f_ext != nullptr) {
        const double *cog = i_face_cog + 3*(size_t)c2f[k];
        const double dot_i = ddp3(cog, cen, fx);
        const double dot_j = ddp3(cog, cj, f_ext + 3*j);
        t += dot_i - dot_j;
      }
      t *= inv_d2;
      r0 += d0 * t; r1 += d1 * t; r2 += d2 * t;
    }
    for (int k = bs; k < be; k++) {
      const size_t f = (size_t)cell_b_faces[k];
      const double a = coefa[f], b = coefb[f];
      const double inv_d = 1. / b_dist[f];
      const double s_ip = (1. - b) * inv_d;
      const double d0 = b_un[3*f] + s_ip * diipb[3*f], d1 = b_un[3*f+1] + s_ip * diipb[3*f+1],
                   d2 = b_un[3*f+2] + s_ip * diipb[3*f+2];
      g[0] = (float)((double)g[0] + d0 * d0);
      g[1] = (float)((double)g[1] + d1 * d1);
      g[2] = (float)((double)g[2] + d2 * d2);
      g[3] = (float)((double)g[3] + d0 * d1);
      g[4] = (float)((double)g[4] + d1 * d2);
      g[5] = (float)((double)g[5] + d0 * d2);
      double t = (a + (b-1.)*p_c);
      if (f_ext != nullptr) {
        const double dot = ddp3(cen, b_face_cog + 3*f, fx);
        t += (b-1.) * dot;
      }
      t *= inv_d;
      r0 += d0 * t; r1 += d1 * t; r2 += d2 * t;
      const double p_face = a + b*p_c;
      var_min = fmin(var_min, p_face);
      var_max = fmax(var_max, p_face);
    }
    if (clip_coeff >= 0 && (clip_coeff < 1. || clip_coeff > 1.)) {
      const double d_min = var_min - p_c, d_max = var_max - p_c;
      var_min = p_c + d_min*clip_coeff;
      var_max = p_c + d_max*clip_coeff;
    }
    const double a00 = (double)(g[1]*g[2] - g[4]*g[4]);
    const double a01 = (double)(g[4]*g[5] - g[3]*g[2]);
    const double a02 = (double)(g[3]*g[4] - g[1]*g[5]);
    const double a11 = (double)(g[0]*g[2] - g[5]*g[5]);
    const double a12 = (double)(g[3]*g[5] - g[0]*g[4]);
    const double a22 = (double)(g[0]*g[1] - g[3]*g[3]);
    const double inv_det = 1. / (g[0]*a00 + g[3]*a01 + g[5]*a02);
    gx = (a00 * r0 + a01 * r1 + a02 * r2) * inv_det;
    gy = (a01 * r0 + a11 * r1 + a12 * r2) * inv_det;
    gz = (a02 * r0 + a12 * r1 + a22 * r2) * inv_det;
    if (f_ext != nullptr) { gx += fx[0]; gy += fx[1]; gz += fx[2]; }
  }
  for (int k = bs; k < be; k++) {
    const size_t f = (size_t)cell_b_faces[k];
    double ip[3] = {0., 0., 0.};
    double vip = p_c;
    if (diipb != nullptr) {
      ip[0] = diipb[3*f]; ip[1] = diipb[3*f+1]; ip[2] = diipb[3*f+2];
      const double bd = b_dist[f];
      if (!(ip[0]*ip[0] + ip[1]*ip[1] + ip[2]*ip[2] < (bd*bd) * (1e-3 * 1e-3))) {
        const double shift = gx*ip[0] + gy*ip[1] + gz*ip[2];
        vip = p_c + shift;
        if (clip_coeff >= 0) {
          if (vip < var_min) vip = var_min;
          if (vip > var_max) vip = var_max;
        }
      }
    }
    val_ip[f] = vip;
    if (val_f != nullptr) {
      double vf = inc * coefa[f] + coefb[f] * vip;
      if (f_ext != nullptr) {
        const double x_ip[3] = {cen[0] + ip[0], cen[1] + ip[1], cen[2] + ip[2]};
        vf += coefb[f] * ddp3(x_ip, b_face_cog + 3*f, fx);
      }
      val_f[f] = vf;
    }
  }
}

/* _reconstruct_scalar_gradient (cs_gradient.cpp:4167-4540, the branch without hydrostatic
   pressure :4428-4506; no cell weighting, no disabled cell): the second half of the
   Green-Gauss gradient with least-squares face values (imrgra = 4, CS_GRADIENT_GREEN_LSQ).
   One thread per cell gathers its faces in the reference's scatter order (interior faces by
   ascending id, then boundary faces): (pfac + rfac) S n with the face weight and the OF vector,
   divided by the cell volume, then the warped-face correction when corr_grad_lin is given. */
__global__ void k_green_reconstruct(int n_rows, const int *__restrict__ row_index,
                                    const int *__restrict__ col_id,
                                    const int *__restrict__ c2f,
                                    const unsigned char *__restrict__ fperm,
                                    const short *__restrict__ c2f_sgn,
                                    const int *__restrict__ cell_b_faces_idx,
                                    const int *__restrict__ cell_b_faces,
                                    const double *__restrict__ weight,
                                    const double *__restrict__ dofij,
                                    const double *__restrict__ i_face_surf,
                                    const double *__restrict__ i_face_u_normal,
                                    const double *__restrict__ b_face_surf,
                                    const double *__restrict__ b_un,
                                    const double *__restrict__ cell_vol,
                                    const double *__restrict__ corr_grad_lin, int inc,
                                    const double *__restrict__ pvar,
                                    const double *__restrict__ coefa,
                                    const double *__restrict__ coefb,
                                    const double *__restrict__ val_f,
                                    const double *__restrict__ cell_cen,
                                    const double *__restrict__ i_face_cog,
                                    const double *__restrict__ f_ext,
                                    const double *__restrict__ r_grad, double *grad)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  const double pc = pvar[c];
  double g0 = 0., g1 = 0., g2 = 0.;
  FaceOrder fo(c2f, fperm, s, e);
  while (true) {
    int best, bk;
    if (!fo.next(best, bk)) break;
    const size_t f = (size_t)best;
    const int o = col_id[bk];
    const bool first = (c2f_sgn[bk] > 0);
    const size_t c1 = (size_t)(first ? c : o), c2 = (size_t)(first ? o : c);
    const double w_face = weight[f];
    const double dp = pvar[c2] - pvar[c1];
    const double *of = dofij + 3*f, *ra = r_grad + 3*c1, *rb = r_grad + 3*c2;
    const double *n = i_face_u_normal + 3*f;
    double t_first, t_second, t_rec;
    if (f_ext == nullptr) {
      t_rec = 0.5 * (  of[0]*(ra[0]+rb[0])
                    + of[1]*(ra[1]+rb[1])
                    + of[2]*(ra[2]+rb[2]));
      t_first = (1.0-w_face) * dp;
      t_second = -w_face * dp;
    }
    else {                /* hydrostatic branch (cs_gradient.cpp:4318-4398; no porosity) */
      const double *f1 = f_ext + 3*c1, *f2 = f_ext + 3*c2, *cog = i_face_cog + 3*f;
      double f_mid[3], r_first[3], r_second[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        f_mid[k] = 0.5 * (f1[k] + f2[k]);
        r_first[k] = cog[k] - cell_cen[3*c1+k];
        r_second[k] = cog[k] - cell_cen[3*c2+k];
      }
      t_first =  w_face
                 * (  r_first[0]*f1[0]
                    + r_first[1]*f1[1]
                    + r_first[2]*f1[2]
                    + 0.)
            +  (1.0 - w_face)
                 * (  r_second[0]*f2[0]
                    + r_second[1]*f2[1]
                    + r_second[2]*f2[2]
                    + 0.);
      t_second = t_first;
      t_first += (1.0-w_face) * dp;
      t_second -=      w_face  * dp;
      t_rec =   - w_face         * (r_first[0]*f_mid[0] + r_first[1]*f_mid[1] + r_first[2]*f_mid[2])
           -  (1.0 - w_face)    * (r_second[0]*f_mid[0] + r_second[1]*f_mid[1] + r_second[2]*f_mid[2])
           + (  of[0] * (ra[0]+rb[0])
              + of[1] * (ra[1]+rb[1])
              + of[2] * (ra[2]+rb[2])) * 0.5;
    }
    if (first) {
      const double pf = (t_first + t_rec) * i_face_surf[f];
      g0 += pf * n[0]; g1 += pf * n[1]; g2 += pf * n[2];
    }
    else {
      const double pf = (t_second + t_rec) * i_face_surf[f];
      g0 += -(pf * n[0]); g1 += -(pf * n[1]); g2 += -(pf * n[2]);
    }
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      const size_t f = (size_t)cell_b_faces[k];
      const double vf = (val_f != nullptr) ? val_f[f] : inc*coefa[f] + coefb[f]*pc;
      const double t_b = (vf - pc) * b_face_surf[f];
      g0 += t_b * b_un[3*f]; g1 += t_b * b_un[3*f+1]; g2 += t_b * b_un[3*f+2];
    }
  const double inv_vol = 1. / cell_vol[c];
  g0 *= inv_vol; g1 *= inv_vol; g2 *= inv_vol;
  if (corr_grad_lin != nullptr) {
    const double *m = corr_grad_lin + 9*(size_t)c;
    double r[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double a = 0.;
      a += m[3*i] * g0; a += m[3*i+1] * g1; a += m[3*i+2] * g2;
      r[i] = a;
    }
    g0 = r[0]; g1 = r[1]; g2 = r[2];
  }
  grad[3*(size_t)c] = g0; grad[3*(size_t)c+1] = g1; grad[3*(size_t)c+2] = g2;
}

/* cs_divergence (src/alge/cs_divergence.cpp:1224-1290): diverg[c] (+)= sum of the cell's
   interior face fluxes in ascending face order (+ the flux when the cell is the face's first,
   - otherwise), then of its boundary face fluxes: the reference's scatter in face order, as a
   gather */
__global__ void k_divergence(int n_rows, const int *__restrict__ row_index,
                             const int *__restrict__ c2f,
                             const unsigned char *__restrict__ fperm, const short *__restrict__ c2f_sgn,
                             const int *__restrict__ cell_b_faces_idx,
                             const int *__restrict__ cell_b_faces, int init,
                             const double *__restrict__ i_massflux,
                             const double *__restrict__ b_massflux, double *diverg)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  double d = (init >= 1) ? 0. : diverg[c];
  FaceOrder fo(c2f, fperm, s, e);
  while (true) {
    int best, bk;
    if (!fo.next(best, bk)) break;
    if (c2f_sgn[bk] > 0) d += i_massflux[best];
    else d += -i_massflux[best];
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++)
      d += b_massflux[cell_b_faces[k]];
  diverg[c] = d;
}

/* cs_face_diffusion_potential (cs_convection_diffusion.cpp:8075-8310): the face fluxes
   themselves.  Every interior face has exactly one MSR entry with a positive sign, in the row
   of its first cell: the thread of that cell computes the face's flux (grad nullptr:
   eqp->ircflu = 0, :8236-8246; else the reconstructed form :8254-8296, always the II' - JJ'
   formula) and adds it to i_massflux[f]; it also serves the cell's boundary faces (:8298-8303).
   No two threads touch the same face. */
__global__ void k_face_diffusion_potential(int n_rows, const int *__restrict__ row_index,
                                           const int *__restrict__ col_id,
                                           const int *__restrict__ c2f,
                                           const short *__restrict__ c2f_sgn,
                                           const int *__restrict__ cell_b_faces_idx,
                                           const int *__restrict__ cell_b_faces,
                                           const double *__restrict__ diipf,
                                           const double *__restrict__ djjpf,
                                           const double *__restrict__ i_dist,
                                           const double *__restrict__ i_face_surf, int init,
                                           int inc, const double *__restrict__ pvar,
                                           const double *__restrict__ grad,
                                           const double *__restrict__ cofaf,
                                           const double *__restrict__ cofbf,
                                           const double *__restrict__ i_visc,
                                           const double *__restrict__ b_visc,
                                           const double *__restrict__ visel,
                                           const double *__restrict__ val_ip,
                                           double *i_massflux, double *b_massflux)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const double pc = pvar[c];
  for (int k = row_index[c]; k < row_index[c+1]; k++) {
    if (c2f_sgn[k] <= 0) continue;
    const size_t f = (size_t)c2f[k], c1 = (size_t)col_id[k];
    double flux = i_visc[f]*(pc - pvar[c1]);
    if (grad != nullptr) {
      const double *g0 = grad + 3*(size_t)c, *g1 = grad + 3*c1;
      const double *ip = diipf + 3*f, *jp = djjpf + 3*f;
      const double dpxf = 0.5*(visel[c]*g0[0] + visel[c1]*g1[0]);
      const double dpyf = 0.5*(visel[c]*g0[1] + visel[c1]*g1[1]);
      const double dpzf = 0.5*(visel[c]*g0[2] + visel[c1]*g1[2]);
      const double dijx = ip[0] - jp[0], dijy = ip[1] - jp[1], dijz = ip[2] - jp[2];
      flux = flux + (dpxf*dijx + dpyf*dijy + dpzf*dijz) * i_face_surf[f]/i_dist[f];
    }
    i_massflux[f] = ((init >= 1) ? 0. : i_massflux[f]) + flux;
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      const int f = cell_b_faces[k];
      const double flux_d = inc*cofaf[f] + cofbf[f]*(val_ip != nullptr ? val_ip[f] : pc);
      b_massflux[f] = ((init >= 1) ? 0. : b_massflux[f]) + b_visc[f]*flux_d;
    }
}

/* Iterative Green-Gauss gradient (imrgra = 0, CS_GRADIENT_GREEN_ITER; cs_gradient.cpp:
   _compute_cell_cocg_it :1776-1860, _initialize_scalar_gradient :1329-1590,
   _iterative_scalar_gradient :1891-2300, branches without hydrostatic pressure, no weighting).
   One thread per cell gathers its faces in the reference's scatter order.
   k_cocg_it: identity - 0.5 S OF (x) n / V summed over the cell's interior faces, inverted by
   Cramer's rule (cs_math.h:1587-1610). */
__global__ void k_cocg_it(int n_rows, const int *__restrict__ row_index,
                          const int *__restrict__ c2f, const unsigned char *__restrict__ fperm,
                          const short *__restrict__ c2f_sgn, const double *__restrict__ dofij,
                          const double *__restrict__ i_face_surf,
                          const double *__restrict__ i_face_u_normal,
                          const double *__restrict__ cell_vol, double *cocg_it)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  double a[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
  const double inv_vol = (cell_vol[c] > 0) ? 1. / cell_vol[c] : 0.;
  FaceOrder fo(c2f, fperm, s, e);
  while (true) {
    int best, bk;
    if (!fo.next(best, bk)) break;
    const size_t f = (size_t)best;
    const bool first = (c2f_sgn[bk] > 0);
    const double half_s = -0.5 * i_face_surf[f];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double t_i = half_s * dofij[3*f+i];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const double t_ij = t_i*i_face_u_normal[3*f+j];
        if (first) a[i][j] += t_ij * inv_vol;
        else a[i][j] += -t_ij * inv_vol;
      }
    }
  }
  const double a00 = a[1][1]*a[2][2] - a[2][1]*a[1][2];
  const double a01 = a[2][1]*a[0][2] - a[0][1]*a[2][2];
  const double a02 = a[0][1]*a[1][2] - a[1][1]*a[0][2];
  const double a10 = a[2][0]*a[1][2] - a[1][0]*a[2][2];
  const double a11 = a[0][0]*a[2][2] - a[2][0]*a[0][2];
  const double a12 = a[1][0]*a[0][2] - a[0][0]*a[1][2];
  const double a20 = a[1][0]*a[2][1] - a[2][0]*a[1][1];
  const double a21 = a[2][0]*a[0][1] - a[0][0]*a[2][1];
  const double a22 = a[0][0]*a[1][1] - a[1][0]*a[0][1];
  const double inv_det = 1. / (a[0][0]*a00 + a[1][0]*a01 + a[2][0]*a02);
  double *o = cocg_it + 9*(size_t)c;
  o[0] = a00 * inv_det; o[1] = a01 * inv_det; o[2] = a02 * inv_det;
  o[3] = a10 * inv_det; o[4] = a11 * inv_det; o[5] = a12 * inv_det;
  o[6] = a20 * inv_det; o[7] = a21 * inv_det; o[8] = a22 * inv_det;
}

/* sweep 0 (r_grad == nullptr): the non-reconstructed Green-Gauss gradient, out = grad;
   later sweeps: out = rhs / V with rhs = -grad V + the reconstructed face values
   (0.5 OF.(grad_I + grad_J) at interior faces, b II'.grad_I at boundary faces) */
__global__ void k_green_iter_sweep(int n_rows, const int *__restrict__ row_index,
                                   const int *__restrict__ col_id,
                                   const int *__restrict__ c2f,
                                   const unsigned char *__restrict__ fperm,
                                   const short *__restrict__ c2f_sgn,
                                   const int *__restrict__ cell_b_faces_idx,
                                   const int *__restrict__ cell_b_faces,
                                   const double *__restrict__ weight,
                                   const double *__restrict__ dofij,
                                   const double *__restrict__ i_face_surf,
                                   const double *__restrict__ i_face_u_normal,
                                   const double *__restrict__ b_face_surf,
                                   const double *__restrict__ b_un,
                                   const double *__restrict__ cell_vol,
                                   const double *__restrict__ diipb, double inc,
                                   const double *__restrict__ pvar,
                                   const double *__restrict__ coefa,
                                   const double *__restrict__ coefb,
                                   const double *__restrict__ r_grad, double *out)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  const double pc = pvar[c];
  double g0 = 0.0, g1 = 0.0, g2 = 0.0;
  if (r_grad != nullptr) {
    const double *g = r_grad + 3*(size_t)c;
    const double vol = cell_vol[c];
    g0 = -g[0] * vol; g1 = -g[1] * vol; g2 = -g[2] * vol;
  }
  FaceOrder fo(c2f, fperm, s, e);
  while (true) {
    int best, bk;
    if (!fo.next(best, bk)) break;
    const size_t f = (size_t)best;
    const int o = col_id[bk];
    const bool first = (c2f_sgn[bk] > 0);
    const size_t c1 = (size_t)(first ? c : o), c2 = (size_t)(first ? o : c);
    const double w_face = weight[f];
    const double *n = i_face_u_normal + 3*f;
    double pf;
    if (r_grad == nullptr) {
      if (first) pf = (1.0-w_face) * (pvar[c2]-pvar[c1]) * i_face_surf[f];
      else pf = -w_face * (pvar[c2]-pvar[c1]) * i_face_surf[f];
    }
    else {
      const double *of = dofij + 3*f, *ra = r_grad + 3*c1, *rb = r_grad + 3*c2;
      pf = 0.5 * (  of[0] * (ra[0]+rb[0])
                  + of[1] * (ra[1]+rb[1])
                  + of[2] * (ra[2]+rb[2]));
      if (first) pf += (1.0-w_face) * (pvar[c2] - pvar[c1]);
      else pf -= w_face * (pvar[c2] - pvar[c1]);
      pf *= i_face_surf[f];
    }
    if (first) { g0 += pf * n[0]; g1 += pf * n[1]; g2 += pf * n[2]; }
    else { g0 += -pf * n[0]; g1 += -pf * n[1]; g2 += -pf * n[2]; }
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      const size_t f = (size_t)cell_b_faces[k];
      double t_b;
      if (r_grad == nullptr)
        t_b = inc*coefa[f] + (coefb[f]-1.0)*pc;
      else {
        const double *g = r_grad + 3*(size_t)c;
        double ip0 = 0., ip1 = 0., ip2 = 0.;
        if (diipb != nullptr) { ip0 = diipb[3*f]; ip1 = diipb[3*f+1]; ip2 = diipb[3*f+2]; }
        t_b =   coefa[f] * inc
               + coefb[f] * (ip0 * g[0] + ip1 * g[1] + ip2 * g[2]);
        t_b += (coefb[f] -1.0) * pc;
      }
      t_b *= b_face_surf[f];
      g0 += t_b * b_un[3*f]; g1 += t_b * b_un[3*f+1]; g2 += t_b * b_un[3*f+2];
    }
  const double inv_vol = 1. / cell_vol[c];
  g0 *= inv_vol; g1 *= inv_vol; g2 *= inv_vol;
  out[3*(size_t)c] = g0; out[3*(size_t)c+1] = g1; out[3*(size_t)c+2] = g2;
}

/* grad += cocg_it^T-style product of the scaled right-hand side (:2252-2260) */
__global__ void k_green_iter_update(int n_rows, const double *__restrict__ cocg_it,
                                    const double *__restrict__ rhs, double *grad)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const double *a = cocg_it + 9*(size_t)c, *r = rhs + 3*(size_t)c;
  double *g = grad + 3*(size_t)c;
  g[0] += r[0] * a[0] + r[1] * a[3] + r[2] * a[6];
  g[1] += r[0] * a[1] + r[1] * a[4] + r[2] * a[7];
  g[2] += r[0] * a[2] + r[1] * a[5] + r[2] * a[8];
}

/* _scalar_gradient_clipping (cs_gradient.cpp:808-1215), standard neighbourhood.
   k_grad_clip_factor -- mode 0 (CS_GRADIENT_LIMIT_CELL): factor of the cell from its own
   gradient; 1 (_FACE): the same with the mean gradient of the two cells, written to `factor`
   for k_grad_clip_face_min (the smallest factor of the cell and its neighbours,
   _gradient_update_face_clip_factor :691-758); 2 (_CELL_RC): values reconstructed at I' / J'
   of the cell's interior faces kept within climgp times the range of the neighbouring cell
   values and boundary face values.  Maxima and minima only: no summation order to keep. */
__global__ void k_grad_clip_factor(int mode, int n_rows, const int *__restrict__ row_index,
                                   const int *__restrict__ col_id, const int *__restrict__ c2f,
                                   const short *__restrict__ c2f_sgn,
                                   const int *__restrict__ cell_b_faces_idx,
                                   const int *__restrict__ cell_b_faces,
                                   const double *__restrict__ cell_cen,
                                   const double *__restrict__ diipf,
                                   const double *__restrict__ djjpf, double climgp, int inc,
                                   const double *__restrict__ pvar,
                                   const double *__restrict__ coefa,
                                   const double *__restrict__ coefb,
                                   const double *__restrict__ val_f,
                                   const double *__restrict__ grad, double *out)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  const double pc = pvar[c];
  const double g0 = grad[3*(size_t)c], g1 = grad[3*(size_t)c+1], g2 = grad[3*(size_t)c+2];
  if (mode == 0 || mode == 1) {
    double g_var = 0, p_var = 0;
    for (int k = s; k < e; k++) {
      const size_t j = (size_t)col_id[k];
      const double d0 = cell_cen[3*(size_t)c] - cell_cen[3*j],
                   d1 = cell_cen[3*(size_t)c+1] - cell_cen[3*j+1],
                   d2 = cell_cen[3*(size_t)c+2] - cell_cen[3*j+2];
      double gv;
      if (mode == 0)
        gv = fabs(d0*g0 + d1*g1 + d2*g2);
      else {
        const double m0 = g0 + grad[3*j], m1 = g1 + grad[3*j+1], m2 = g2 + grad[3*j+2];
        gv = 0.5 * fabs(d0*m0 + d1*m1 + d2*m2);
      }
      const double pv = fabs(pc - pvar[j]);
      g_var = fmax(g_var, gv);
      p_var = fmax(p_var, pv);
    }
    double lim = 1.;
    if (g_var > climgp * p_var) lim = climgp * p_var/g_var;
    out[c] = lim;
    return;
  }
  double v_min = pc, v_max = pc;
  for (int k = s; k < e; k++) {
    const double pj = pvar[col_id[k]];
    v_min = fmin(v_min, pj);
    v_max = fmax(v_max, pj);
  }
  if (cell_b_faces_idx != nullptr)
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      const int f = cell_b_faces[k];
      const double vf = (val_f != nullptr) ? val_f[f] : inc*coefa[f] + coefb[f]*pc;
      v_min = fmin(v_min, vf);
      v_max = fmax(v_max, vf);
    }
  v_min = (v_min - pc) * climgp;
  v_max = (v_max - pc) * climgp;
  double lim_c = 1.;
  for (int k = s; k < e; k++) {
    const double *off = ((c2f_sgn[k] > 0) ? diipf : djjpf) + 3*(size_t)c2f[k];
    double lim_f = 1.;
    const double dv = off[0]*g0 + off[1]*g1 + off[2]*g2;
    if (dv < v_min)
      lim_f = fabs(v_min) / fabs(dv);
    else if (dv > v_max)
      lim_f = fabs(v_max) / fabs(dv);
    lim_c = fmin(lim_c, lim_f);
  }
  out[c] = lim_c;
}

__global__ void k_grad_clip_face_min(int n_rows, const int *__restrict__ row_index,
                                     const int *__restrict__ col_id,
                                     const double *__restrict__ factor, double *clip_factor)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  double lim = factor[c];
  for (int k = row_index[c]; k < row_index[c+1]; k++)
    lim = fmin(lim, factor[col_id[k]]);
  clip_factor[c] = fmin(1., lim);
}

__global__ void k_grad_clip_apply(int n_rows, const double *__restrict__ clip_factor,
                                  double *grad)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const double lim = clip_factor[c];
  if (lim < 1) {
    grad[3*(size_t)c] *= lim; grad[3*(size_t)c+1] *= lim; grad[3*(size_t)c+2] *= lim;
  }
}

/* cs_diffusion_potential with flux reconstruction (:8921-8961): the face flux is computed in
   the face's orientation by both of its cells, identically */
__global__ void k_diffusion_potential_rc(int n_rows, const int *__restrict__ row_index,
                                         const int *__restrict__ col_id,
                                         const int *__restrict__ c2f,
                                         const unsigned char *__restrict__ fperm,
                                         const short *__restrict__ c2f_sgn,
                                         const int *__restrict__ cell_b_faces_idx,
                                         const int *__restrict__ cell_b_faces,
                                         const double *__restrict__ diipf,
                                         const double *__restrict__ djjpf,
                                         const double *__restrict__ i_dist,
                                         const double *__restrict__ i_face_surf, int irecmf,
                                         int init, int inc, const double *__restrict__ pvar,
                                         const double *__restrict__ grad,
                                         const double *__restrict__ cofaf,
                                         const double *__restrict__ cofbf,
                                         const double *__restrict__ i_visc,
                                         const double *__restrict__ b_visc,
                                         const double *__restrict__ visel,
                                         const double *__restrict__ val_ip, double *diverg)
{
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c >= n_rows) return;
  const int s = row_index[c], e = row_index[c+1];
  double d = (init >= 1) ? 0. : diverg[c];
  FaceOrder fo(c2f, fperm, s, e);
  while (true) {
    int best, bk;
    if (!fo.next(best, bk)) break;
    const size_t f = (size_t)best;
    const int o = col_id[bk];
    const bool first = (c2f_sgn[bk] > 0);
    const size_t c0 = (size_t)(first ? c : o), c1 = (size_t)(first ? o : c);
    const double *g0 = grad + 3*c0, *g1 = grad + 3*c1;
    const double *ip = diipf + 3*f, *jp = djjpf + 3*f;
    double flux = i_visc[f]*(pvar[c0] - pvar[c1]);
    if (irecmf == 0) {
      const double dijx = ip[0] - jp[0], dijy = ip[1] - jp[1], dijz = ip[2] - jp[2];
      const double dpxf = 0.5*(visel[c0]*g0[0] + visel[c1]*g1[0]);
      const double dpyf = 0.5*(visel[c0]*g0[1] + visel[c1]*g1[1]);
      const double dpzf = 0.5*(visel[c0]*g0[2] + visel[c1]*g1[2]);
      flux += (dpxf*dijx + dpyf*dijy + dpzf*dijz) * i_face_surf[f]/i_dist[f];
    }
    else
      flux += i_visc[f] * (  (g0[0]*ip[0] + g0[1]*ip[1] + g0[2]*ip[2])
                           - (g1[0]*jp[0] + g1[1]*jp[1] + g1[2]*jp[2]));
    if (first) d += flux;
    else d += -flux;
  }
  if (cell_b_faces_idx != nullptr) {
    const double pc = pvar[c];
    for (int k = cell_b_faces_idx[c]; k < cell_b_faces_idx[c+1]; k++) {
      const int f = cell_b_faces[k];
      const double flux_d = inc*cofaf[f] + cofbf[f]*(val_ip != nullptr ? val_ip[f] : pc);
      d += b_visc[f]*flux_d;
    }
  }
  diverg[c] = d;
}

} // anonymous namespace

bool launch_lsq_cocg(const _csb_matrix_t *m, float *cocg)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  k_lsq_cocg<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->cell_cen.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr,
    m->b_face_u_normal.p, m->b_dist.p, (bnd && m->has_diipb) ? m->diipb.p : nullptr, cocg);
  return post_launch("lsq_cocg");
}

bool launch_b_iprime_lsq(const _csb_matrix_t *m, int inc, const double *pvar,
                         const double *coefa, const double *coefb, const double *f_ext,
                         double clip_coeff, double *val_ip, double *val_f)
{
  if (m->n_rows < 1 || !m->has_b_faces || m->n_b_faces < 1) return true;
  k_b_iprime_lsq<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, m->cell_cen.p, m->cell_b_faces_idx.p,
    m->cell_b_faces.p, m->b_face_u_normal.p, m->b_dist.p, m->has_diipb ? m->diipb.p : nullptr,
    m->i_face_cog.p, m->b_face_cog.p, f_ext, clip_coeff, inc, pvar, coefa, coefb, val_ip, val_f);
  return post_launch("b_iprime_lsq");
}

bool launch_lsq_gradient(const _csb_matrix_t *m, const float *cocg, int inc, const double *pvar,
                         const double *coefa, const double *coefb, const double *val_f,
                         const double *f_ext, double *grad)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const unsigned char *fperm = face_order(m);
  k_lsq_gradient<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, fperm, m->c2f_sgn.p, m->cell_cen.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr,
    m->b_face_u_normal.p, m->b_dist.p, cocg, inc, pvar, coefa, coefb,
    (bnd && m->has_diipb) ? m->diipb.p : nullptr, val_f, m->i_face_cog.p, m->b_face_cog.p,
    f_ext, grad);
  return post_launch("lsq_gradient");
}

bool launch_cocg_it(const _csb_matrix_t *m, double *cocg_it)
{
  if (m->n_rows < 1) return true;
  const unsigned char *fperm = face_order(m);
  k_cocg_it<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->c2f.p, fperm, m->c2f_sgn.p, m->dofij.p, m->face_surf.p,
    m->face_u_normal.p, m->cell_vol.p, cocg_it);
  return post_launch("cocg_it");
}

bool launch_green_iter_sweep(const _csb_matrix_t *m, int inc, const double *pvar,
                             const double *coefa, const double *coefb, const double *r_grad,
                             double *out)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const unsigned char *fperm = face_order(m);
  k_green_iter_sweep<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, fperm, m->c2f_sgn.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr, m->g_weight.p,
    m->dofij.p, m->face_surf.p, m->face_u_normal.p, m->b_face_surf.p, m->b_face_u_normal.p,
    m->cell_vol.p, (bnd && m->has_diipb) ? m->diipb.p : nullptr, (double)inc, pvar, coefa, coefb,
    r_grad, out);
  return post_launch("green_iter_sweep");
}

bool launch_green_iter_update(const _csb_matrix_t *m, const double *cocg_it, const double *rhs,
                              double *grad)
{
  if (m->n_rows < 1) return true;
  k_green_iter_update<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(m->n_rows, cocg_it, rhs,
                                                                       grad);
  return post_launch("green_iter_update");
}

/* gradient limiter applied to grad in place; work: 2 n_rows doubles */
bool launch_gradient_clip(const _csb_matrix_t *m, int clip_mode, double climgp, int inc,
                          const double *pvar, const double *coefa, const double *coefb,
                          const double *val_f, double *grad, double *work)
{
  if (m->n_rows < 1 || clip_mode < 0 || climgp < 0) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const int g = (m->n_rows + 127)/128;
  double *clip_factor = work, *factor = work + m->n_rows;
  k_grad_clip_factor<<<g, 128, 0, ctx().stream>>>(
    clip_mode, m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, m->c2f_sgn.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr, m->cell_cen.p,
    m->diipf.p, m->djjpf.p, climgp, inc, pvar, coefa, coefb, val_f, grad,
    (clip_mode == 1) ? factor : clip_factor);
  if (!post_launch("grad_clip_factor")) return false;
  if (clip_mode == 1) {
    k_grad_clip_face_min<<<g, 128, 0, ctx().stream>>>(m->n_rows, m->row_index.p, m->col_id.p,
                                                      factor, clip_factor);
    if (!post_launch("grad_clip_face_min")) return false;
  }
  k_grad_clip_apply<<<g, 128, 0, ctx().stream>>>(m->n_rows, clip_factor, grad);
  return post_launch("grad_clip_apply");
}

bool launch_divergence(const _csb_matrix_t *m, int init, const double *i_massflux,
                       const double *b_massflux, double *diverg)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const unsigned char *fperm = face_order(m);
  k_divergence<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->c2f.p, fperm, m->c2f_sgn.p, bnd ? m->cell_b_faces_idx.p : nullptr,
    bnd ? m->cell_b_faces.p : nullptr, init, i_massflux, b_massflux, diverg);
  return post_launch("divergence");
}

bool launch_face_diffusion_potential(const _csb_matrix_t *m, int init, int inc,
                                     const double *pvar, const double *grad,
                                     const double *cofaf, const double *cofbf,
                                     const double *i_visc, const double *b_visc,
                                     const double *visel, const double *val_ip,
                                     double *i_massflux, double *b_massflux)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  k_face_diffusion_potential<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, m->c2f_sgn.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr, m->diipf.p,
    m->djjpf.p, m->i_dist.p, m->face_surf.p, init, inc, pvar, grad, cofaf, cofbf, i_visc,
    b_visc, visel, val_ip, i_massflux, b_massflux);
  return post_launch("face_diffusion_potential");
}

bool launch_green_reconstruct(const _csb_matrix_t *m, int inc, const double *pvar,
                              const double *coefa, const double *coefb, const double *val_f,
                              const double *f_ext, const double *r_grad, double *grad)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const unsigned char *fperm = face_order(m);
  k_green_reconstruct<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, fperm, m->c2f_sgn.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr, m->g_weight.p,
    m->dofij.p, m->face_surf.p, m->face_u_normal.p, m->b_face_surf.p, m->b_face_u_normal.p,
    m->cell_vol.p, m->has_corr_grad_lin ? m->corr_grad_lin.p : nullptr, inc, pvar, coefa, coefb,
    val_f, m->cell_cen.p, m->i_face_cog.p, f_ext, r_grad, grad);
  return post_launch("green_reconstruct");
}

bool launch_diffusion_potential_rc(const _csb_matrix_t *m, int irecmf, int init, int inc,
                                   const double *pvar, const double *grad, const double *cofaf,
                                   const double *cofbf, const double *i_visc,
                                   const double *b_visc, const double *visel,
                                   const double *val_ip, double *diverg)
{
  if (m->n_rows < 1) return true;
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const unsigned char *fperm = face_order(m);
  k_diffusion_potential_rc<<<(m->n_rows + 127)/128, 128, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, fperm, m->c2f_sgn.p,
    bnd ? m->cell_b_faces_idx.p : nullptr, bnd ? m->cell_b_faces.p : nullptr, m->diipf.p,
    m->djjpf.p, m->i_dist.p, m->face_surf.p, irecmf, init, inc, pvar, grad, cofaf, cofbf,
    i_visc, b_visc, visel, val_ip, diverg);
  return post_launch("diffusion_potential_rc");
}

bool launch_native_to_msr(int nnz, const int *c2f, const double *xa, double *x_val)
{
  if (nnz < 1) return true;
  k_native_to_msr<<<(nnz + 255)/256, 256, 0, ctx().stream>>>(nnz, c2f, xa, x_val);
  return post_launch("native_to_msr");
}

/* stages: 1 = sums over fine rows (cell centres, volumes, faces, diagonal),
   2 = extra-diagonal values; ghost cell centres / volumes of a distributed
   level are synchronised between the two (cs_grid.cpp:6561-6581) */
bool launch_coarse_assembly(const CoarseAssemblyArgs &a, int stages)
{
  Context &c = ctx();
  if (a.c_n_rows < 1) return true;
  const int grid = (a.c_n_rows + CA_WARPS - 1)/CA_WARPS;
  if (stages & 1) {
    /* list of the rows too wide for a warp (normally empty): n_wide, then the row ids */
    int *wide = static_cast<int *>(dev_alloc(sizeof(int)*((size_t)a.c_n_rows + 1)));
    if (wide == nullptr) return false;
    bool ok = CSB_CUDA(cudaMemsetAsync(wide, 0, sizeof(int), c.stream));
    if (ok) {
      /* CSB_CA_WIDE_LIMIT: test hook, sends rows with more entries than this to the wide kernel */
      static const int lim = env_flag("CSB_CA_WIDE_LIMIT", 32*CA_OWN);
      k_coarse_sums<<<grid, CA_WARPS*32, 0, c.stream>>>(a, lim < 32*CA_OWN ? lim : 32*CA_OWN,
                                                         wide + 1, wide);
      ok = post_launch("coarse_sums");
    }
    if (ok) {
      /* at most the rows with more than 128 entries; the grid covers them all, threads
         beyond *n_wide return at once */
      k_coarse_sums_wide<<<(a.c_n_rows + 63)/64, 64, 0, c.stream>>>(a, wide + 1, wide);
      ok = post_launch("coarse_sums_wide");
    }
    dev_free(wide);
    if (!ok) return false;
  }
  if (stages & 2) {
    k_coarse_values<<<grid, CA_WARPS*32, 0, c.stream>>>(a);
    return post_launch("coarse_values");
  }
  return true;
}

/*--------------------------------------------------------------------------
 * Multigrid smoothers without convergence test (cs_multigrid_smoother.cpp:143-373,
 * :804-905): the sweeps themselves are k_spmv_stream<EPI_SMOOTH> (product and update
 * fused, one pass over the matrix per sweep); these are the pieces around them.
 *--------------------------------------------------------------------------*/

namespace {

/* _setup_ad_l1_inv, MSR (src/alge/cs_sles_it_priv.cpp:236-251) */
__global__ void __launch_bounds__(TPB)
k_l1_inv(int n, const int *__restrict__ row_index, const double *__restrict__ d_val,
         const double *__restrict__ x_val, double *__restrict__ l1_inv)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    double s1 = d_val[i];
    const int ks = row_index[i], ke = row_index[i+1];
    for (int k = ks; k < ke; k++)
      s1 += fabs(x_val[k]);
    l1_inv[i] = 1.0 / s1;
  }
}

/* first sweep from vx = 0 (cs_multigrid_smoother.cpp:196-199, :322-327, :858-861) */
__global__ void __launch_bounds__(TPB)
k_smoother_first(int n, int relaxed, double w0, const double *__restrict__ rhs,
                 const double *__restrict__ ad_inv, double *__restrict__ vx,
                 const MgState *gate)
{
  if (gate != nullptr && gate->cvg != CSB_SLES_ITERATING) return;
  if (relaxed) {
    for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
      vx[i] = w0 * rhs[i]*ad_inv[i];
  }
  else {
    for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
      vx[i] = rhs[i]*ad_inv[i];
  }
}

__global__ void __launch_bounds__(TPB)
k_copy_mg_gated(int n, const double *__restrict__ src, double *__restrict__ dst,
                const MgState *gate)
{
  if (gate != nullptr && gate->cvg != CSB_SLES_ITERATING) return;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
    dst[i] = src[i];
}

/* *n_equiv_iter += n_iter * n_g_rows * denom_n_g_rows_0 (cs_multigrid.cpp:3700, :3790, :3905):
   integer product, then double, then the int += double truncation */
__global__ void
k_mg_accum_n(MgState *mg, int n_iter, int n_rows, double denom)
{
  if (mg->cvg != CSB_SLES_ITERATING) return;
  const double x = (double)((unsigned long long)n_iter * (unsigned long long)n_rows) * denom;
  mg->n_equiv_iter = (int)((double)mg->n_equiv_iter + x);
}

__global__ void
k_gate_from_it(MgState *gate, const ItState *st)
{
  gate->cvg = (st->cvg == CSB_SLES_ITERATING) ? CSB_SLES_ITERATING : CSB_SLES_CONVERGED;
}

/* FCG start (cs_sles_it.cpp:880-905): rk holds A.vx on entry when have_residual */
__global__ void __launch_bounds__(TPB)
k_fcg_init(int n, const double *__restrict__ rhs, double *__restrict__ vx,
           double *__restrict__ rk, double *__restrict__ qk, int have_residual)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    if (have_residual)
      rk[i] = rhs[i] - rk[i];
    else {
      vx[i] = 0.;
      rk[i] = rhs[i];
    }
    qk[i] = 0.;
  }
}

/* FCG vector updates (cs_sles_it.cpp:940-978); st->n_iter was already counted by
   fcg_finalize, so n_iter == 1 is the first iteration (dk, qk not yet defined) */
__global__ void __launch_bounds__(TPB)
k_fcg_update(int n, const ItState *st, const double *__restrict__ vk,
             const double *__restrict__ wk, double *__restrict__ dk,
             double *__restrict__ qk, double *__restrict__ rk, double *__restrict__ vx)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double ak_rk = st->alpha, gk_rk1 = st->beta;
  if (st->n_iter == 1) {
    for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
      const double v = vk[i], w = wk[i];
      dk[i] = v;
      vx[i] = vx[i] + ak_rk * v;
      qk[i] = w;
      rk[i] = rk[i] - ak_rk * w;
    }
  }
  else {
    for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
      const double d = vk[i] - gk_rk1 * dk[i];
      dk[i] = d;
      vx[i] = vx[i] + ak_rk * d;
      const double q = wk[i] - gk_rk1 * qk[i];
      qk[i] = q;
      rk[i] = rk[i] - ak_rk * q;
    }
  }
}

/* Single-reduction PCG with a general preconditioner (_conjugate_gradient_sr,
   cs_sles_it.cpp:1229-1415): the vector phases without the fused Jacobi scaling of
   k_pcg_update_first / k_pcg_update; the preconditioner runs between them. */

/* :1278-1290; rk holds A.vx on entry when have_ax */
__global__ void __launch_bounds__(TPB)
k_pcg_pc_init(int n, const double *__restrict__ rhs, double *__restrict__ vx,
              double *__restrict__ rk, int have_ax)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    if (have_ax)
      rk[i] -= rhs[i];
    else {
      vx[i] = 0.;
      rk[i] = -rhs[i];
    }
  }
}

/* :1325-1341 */
__global__ void __launch_bounds__(TPB)
k_pcg_pc_update_first(int n, ItState *st, const double *__restrict__ dk,
                      const double *__restrict__ zk, double *__restrict__ vx,
                      double *__restrict__ rk, double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double alpha = st->alpha;
  double v[1] = {0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    vx[i] += (alpha * dk[i]);
    const double r = rk[i] + (alpha * zk[i]);
    rk[i] = r;
    v[0] += r*r;
  }
  grid_sum_finalize<1>(v, partials, ticket, [st](const double *s) {
    st->cvg = it_cvg_test(st, st->n_iter, sqrt(s[0]));
  });
}

/* :1388-1394 */
__global__ void __launch_bounds__(TPB)
k_pcg_pc_update(int n, const ItState *st, const double *__restrict__ gk,
                const double *__restrict__ sk, double *__restrict__ dk,
                double *__restrict__ zk, double *__restrict__ rk, double *__restrict__ vx)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double alpha = st->alpha, beta = st->beta;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    const double d = gk[i] + (beta * dk[i]);
    dk[i] = d;
    vx[i] += alpha * d;
    const double z = sk[i] + (beta * zk[i]);
    zk[i] = z;
    rk[i] += alpha * z;
  }
}

/* Inexact-preconditioning CG with a general preconditioner (_conjugate_gradient_ip,
   cs_sles_it.cpp:1016-1203): the reductions that follow the preconditioner.
   first = 1 (:1090-1103): dk = gk, (rk.rk, rk.gk) -> initial residual and test, rk_gk_m1;
   first = 0 (:1140-1163): (rk.rk, rk.gk, gk.rkm1) -> test, beta = (rk.gk - rkm1.gk)/rk_gk_m1 */
__global__ void __launch_bounds__(TPB)
k_ipcg_dots(int n, ItState *st, int first, const double *__restrict__ rk,
            const double *__restrict__ gk, const double *__restrict__ rkm1,
            double *__restrict__ dk, double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  double v[3] = {0., 0., 0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    const double r = rk[i], g = gk[i];
    v[0] += r*r;
    v[1] += r*g;
    if (first) dk[i] = g;
    else v[2] += g*rkm1[i];
  }
  grid_sum_finalize<3>(v, partials, ticket, [st, first](const double *s) {
    const double residual = sqrt(s[0]);
    if (first) {
      st->rk_gkm1 = s[1];
      st->initial_residual = residual;
      const int cvg = it_cvg_test(st, 0, residual);
      if (cvg == CSB_SLES_ITERATING) st->n_iter = 1;
      st->cvg = cvg;
    }
    else {
      int cvg = CSB_SLES_ITERATING;
      if (st->n_iter > 1) cvg = it_cvg_test(st, st->n_iter, residual);
      if (cvg == CSB_SLES_ITERATING) {
        st->n_iter += 1;
        const double rk_gk = s[1], rkm1_gk = s[2];
        st->beta = (fabs(st->rk_gkm1) > DBL_MIN) ? (rk_gk - rkm1_gk) / st->rk_gkm1 : 0.;
        st->rk_gkm1 = rk_gk;
      }
      st->cvg = cvg;
    }
  });
}

/* dk = gk + beta dk (:1166-1168) */
__global__ void __launch_bounds__(TPB)
k_ipcg_dir(int n, const ItState *st, const double *__restrict__ gk, double *__restrict__ dk)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double beta = st->beta;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB)
    dk[i] = gk[i] + (beta * dk[i]);
}

/* vx += alpha dk; rkm1 = rk; rk += alpha zk (:1129-1137, :1185-1193); first = 1 adds the
   residual norm and convergence test that close the first iteration (:1139-1146) */
__global__ void __launch_bounds__(TPB)
k_ipcg_update(int n, ItState *st, int first, const double *__restrict__ dk,
              const double *__restrict__ zk, double *__restrict__ vx,
              double *__restrict__ rk, double *__restrict__ rkm1, double *partials,
              unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  const double alpha = st->alpha;
  double v[1] = {0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    vx[i] += (alpha * dk[i]);
    const double ro = rk[i];
    rkm1[i] = ro;
    const double r = ro + (alpha * zk[i]);
    rk[i] = r;
    v[0] += r*r;
  }
  if (first)
    grid_sum_finalize<1>(v, partials, ticket, [st](const double *s) {
      st->cvg = it_cvg_test(st, st->n_iter, sqrt(s[0]));
    });
}

} // anonymous namespace

bool launch_ipcg_dots(int n, ItState *st, bool first, const double *rk, const double *gk,
                      const double *rkm1, double *dk)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_ipcg_dots<<<g, TPB, 0, c.stream>>>(n, st, first ? 1 : 0, rk, gk, rkm1, dk, c.partials,
                                       c.ticket);
  return post_launch("ipcg_dots");
}

bool launch_ipcg_dir(int n, ItState *st, const double *gk, double *dk)
{
  if (n < 1) return true;
  k_ipcg_dir<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, gk, dk);
  return post_launch("ipcg_dir");
}

bool launch_ipcg_update(int n, ItState *st, bool first, const double *dk, const double *zk,
                        double *vx, double *rk, double *rkm1)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_ipcg_update<<<g, TPB, 0, c.stream>>>(n, st, first ? 1 : 0, dk, zk, vx, rk, rkm1,
                                         c.partials, c.ticket);
  return post_launch("ipcg_update");
}

bool launch_pcg_pc_init(int n, const double *rhs, double *vx, double *rk, bool have_ax)
{
  if (n < 1) return true;
  k_pcg_pc_init<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, rhs, vx, rk, have_ax ? 1 : 0);
  return post_launch("pcg_pc_init");
}

bool launch_pcg_pc_update_first(int n, ItState *st, const double *dk, const double *zk,
                                double *vx, double *rk)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_pcg_pc_update_first<<<g, TPB, 0, c.stream>>>(n, st, dk, zk, vx, rk, c.partials, c.ticket);
  return post_launch("pcg_pc_update_first");
}

bool launch_pcg_pc_update(int n, ItState *st, const double *gk, const double *sk, double *dk,
                          double *zk, double *rk, double *vx)
{
  if (n < 1) return true;
  k_pcg_pc_update<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, gk, sk, dk, zk, rk, vx);
  return post_launch("pcg_pc_update");
}

/*--------------------------------------------------------------------------
 * 3x3 block Jacobi (_block_3_jacobi, src/alge/cs_sles_it.cpp:2284-2440)
 *--------------------------------------------------------------------------*/

namespace {

/* cs_math_matrix_gauss_inverse<3> (src/base/cs_math.h:540-586): Gauss-Jordan on [A | I]
   with row pivoting; the reference swaps row pointers, here a row permutation -- the same
   rows see the same operations in the same order, so the inverse is bit-identical. */
__global__ void __launch_bounds__(TPB)
k_gauss_inverse_33(int n_blocks, const double *__restrict__ ad, double *__restrict__ ad_inv)
{
  for (int b = blockIdx.x*TPB + threadIdx.x; b < n_blocks; b += gridDim.x*TPB) {
    double t[3][6];
    int p[3] = {0, 1, 2};
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) {
        t[i][j] = ad[(size_t)b*9 + i*3 + j];
        t[i][j+3] = 0.;
      }
      t[i][i+3] = 1.;
    }
    for (int i = 0; i < 3; i++) {
      for (int j = i+1; j < 3; j++) {
        if (fabs(t[p[i]][i]) < fabs(t[p[j]][i])) {
          const int tmp = p[j]; p[j] = p[i]; p[i] = tmp;
        }
      }
      for (int j = 0; j < 3; j++) {
        if (j != i) {
          const double s = t[p[j]][i] / t[p[i]][i];
          for (int k = 0; k < 6; k++)
            t[p[j]][k] -= t[p[i]][k] * s;
        }
      }
    }
    for (int i = 0; i < 3; i++) {
      const double s = 1. / t[p[i]][i];
      for (int j = 0; j < 3; j++)
        ad_inv[(size_t)b*9 + i*3 + j] = t[p[i]][3+j] * s;
    }
  }
}

/* One sweep after vxx = E.vx: vx = ad_inv.(rhs - vxx) (_mat_c_m_b_33,
   cs_sles_it_priv.h:272-289) and residual^2 = sum (D.(x_new - x_old))^2
   (cs_sles_it.cpp:2396-2410); first = 1: the simplified sweep from vx = 0 (:2338-2359). */
__global__ void __launch_bounds__(TPB)
k_b3_jacobi_update(int n_blocks, ItState *st, const double *__restrict__ ad,
                   const double *__restrict__ ad_inv, const double *__restrict__ rhs,
                   const double *__restrict__ vxx, double *__restrict__ vx, int first,
                   double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  double v[1] = {0.};
  for (int ii = blockIdx.x*TPB + threadIdx.x; ii < n_blocks; ii += gridDim.x*TPB) {
    double xo[3], aux[3], xn[3];
    for (int k = 0; k < 3; k++) {
      xo[k] = first ? 0. : vx[(size_t)ii*3 + k];
      const double b = first ? 0. : vxx[(size_t)ii*3 + k];
      aux[k] = (rhs[(size_t)ii*3 + k] - b);
    }
    const double *mat = ad_inv + (size_t)ii*9;
    for (int r = 0; r < 3; r++)
      xn[r] = mat[3*r + 0]*aux[0] + mat[3*r + 1]*aux[1] + mat[3*r + 2]*aux[2];
    for (int k = 0; k < 3; k++) vx[(size_t)ii*3 + k] = xn[k];
    for (int jj = 0; jj < 3; jj++) {
      double r = 0.0;
      for (int kk = 0; kk < 3; kk++)
        r += ad[(size_t)ii*9 + jj*3 + kk] * (xn[kk] - xo[kk]);
      v[0] += (r*r);
    }
  }
  grid_sum_finalize<1>(v, partials, ticket, [st](const double *s) {
    st->n_iter += 1;
    const double residual = sqrt(s[0]);
    if (st->n_iter == 1) st->initial_residual = residual;
    st->cvg = it_cvg_test(st, st->n_iter, residual);
  });
}

/* cs_math_matrix_gauss_inverse<6> (cs_math.h:540-586), the block inverse of 6x6 diagonal
   blocks (_setup_ad_block_nn_inv, cs_sles_it_priv.cpp:205-214) */
__global__ void __launch_bounds__(64)
k_gauss_inverse_66(int n_blocks, const double *__restrict__ ad, double *__restrict__ ad_inv)
{
  for (int b = blockIdx.x*64 + threadIdx.x; b < n_blocks; b += gridDim.x*64) {
    double t[6][12];
    int p[6] = {0, 1, 2, 3, 4, 5};
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < 6; j++) {
        t[i][j] = ad[(size_t)b*36 + i*6 + j];
        t[i][j+6] = 0.;
      }
      t[i][i+6] = 1.;
    }
    for (int i = 0; i < 6; i++) {
      for (int j = i+1; j < 6; j++) {
        if (fabs(t[p[i]][i]) < fabs(t[p[j]][i])) {
          const int tmp = p[j]; p[j] = p[i]; p[i] = tmp;
        }
      }
      for (int j = 0; j < 6; j++) {
        if (j != i) {
          const double s = t[p[j]][i] / t[p[i]][i];
          for (int k = 0; k < 12; k++)
            t[p[j]][k] -= t[p[i]][k] * s;
        }
      }
    }
    for (int i = 0; i < 6; i++) {
      const double s = 1. / t[p[i]][i];
      for (int j = 0; j < 6; j++)
        ad_inv[(size_t)b*36 + i*6 + j] = t[p[i]][6+j] * s;
    }
  }
}

/* _fact_lu (cs_sles_it_priv.cpp:118-160): what _setup_ad_block_nn_inv stores in ad_inv for
   block sizes other than 3 and 6 -- the LU factors, which _block_jacobi then applies with
   _mat_c_m_b as if they were the inverse.  Reproduced as it is: the results are the
   reference's for every block size. */
__global__ void __launch_bounds__(TPB)
k_fact_lu(int n_blocks, int db, const double *__restrict__ ad, double *__restrict__ ad_inv)
{
  for (int b = blockIdx.x*TPB + threadIdx.x; b < n_blocks; b += gridDim.x*TPB) {
    double *ai = ad_inv + (size_t)db*db*b;
    const double *a = ad + (size_t)db*db*b;
    ai[0] = a[0];
    for (int ii = 1; ii < db; ii++) {
      ai[ii] = a[ii];
      ai[ii*db] = a[ii*db]/a[0];
    }
    for (int ii = 1; ii < db - 1; ii++) {
      ai[ii + ii*db] = a[ii + ii*db];
      for (int kk = 0; kk < ii; kk++)
        ai[ii + ii*db] -= ai[ii*db + kk]*ai[kk*db + ii];
      for (int jj = ii + 1; jj < db; jj++) {
        ai[ii*db + jj] = a[ii*db + jj];
        ai[jj*db + ii] = a[jj*db + ii] / ai[ii*db + ii];
        for (int kk = 0; kk < ii; kk++) {
          ai[ii*db + jj] -= ai[ii*db + kk]*ai[kk*db + jj];
          ai[jj*db + ii] -= ai[jj*db + kk]*ai[kk*db + ii]/ai[ii*db + ii];
        }
      }
    }
    ai[db*db - 1] = a[db*db - 1];
    for (int kk = 0; kk < db - 1; kk++)
      ai[db*db - 1] -= ai[(db-1)*db + kk]*ai[kk*db + db - 1];
  }
}

/* _block_jacobi sweep (cs_sles_it.cpp:2460-2620) after vxx = E.vx: vx = mat.(rhs - vxx)
   (_mat_c_m_b, cs_sles_it_priv.h:302-322: x = 0; x += aux_j * mat_ij) and
   residual^2 = sum (D.(x_new - x_old))^2; first = 1: the simplified sweep from vx = 0. */
__global__ void __launch_bounds__(TPB)
k_bn_jacobi_update(int n_blocks, int db, ItState *st, const double *__restrict__ ad,
                   const double *__restrict__ ad_inv, const double *__restrict__ rhs,
                   const double *__restrict__ vxx, double *__restrict__ vx, int first,
                   double *partials, unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  double v[1] = {0.};
  for (int ii = blockIdx.x*TPB + threadIdx.x; ii < n_blocks; ii += gridDim.x*TPB) {
    double xo[9], aux[9], xn[9];
    for (int k = 0; k < db; k++) {
      xo[k] = first ? 0. : vx[(size_t)ii*db + k];
      const double b = first ? 0. : vxx[(size_t)ii*db + k];
      aux[k] = (rhs[(size_t)ii*db + k] - b);
    }
    const double *mat = ad_inv + (size_t)ii*db*db;
    for (int r = 0; r < db; r++) {
      double x = 0;
      for (int j = 0; j < db; j++) x += aux[j]*mat[r*db + j];
      xn[r] = x;
    }
    for (int k = 0; k < db; k++) vx[(size_t)ii*db + k] = xn[k];
    for (int jj = 0; jj < db; jj++) {
      double r = 0.0;
      for (int kk = 0; kk < db; kk++)
        r += ad[(size_t)ii*db*db + jj*db + kk] * (xn[kk] - xo[kk]);
      v[0] += (r*r);
    }
  }
  grid_sum_finalize<1>(v, partials, ticket, [st](const double *s) {
    st->n_iter += 1;
    const double residual = sqrt(s[0]);
    if (st->n_iter == 1) st->initial_residual = residual;
    st->cvg = it_cvg_test(st, st->n_iter, residual);
  });
}

} // anonymous namespace

bool launch_gauss_inverse_33(int n_blocks, const double *ad, double *ad_inv)
{
  if (n_blocks < 1) return true;
  k_gauss_inverse_33<<<vec_grid(n_blocks*4), TPB, 0, ctx().stream>>>(n_blocks, ad, ad_inv);
  return post_launch("gauss_inverse_33");
}

bool launch_b3_jacobi_update(int n_blocks, ItState *st, const double *ad, const double *ad_inv,
                             const double *rhs, const double *vxx, double *vx, bool first)
{
  Context &c = ctx();
  const int g = vec_grid(n_blocks*4);
  if (!ensure_partials(g)) return false;
  k_b3_jacobi_update<<<g, TPB, 0, c.stream>>>(n_blocks, st, ad, ad_inv, rhs, vxx, vx,
                                              first ? 1 : 0, c.partials, c.ticket);
  return post_launch("b3_jacobi_update");
}

namespace {
/* _block_relaxed_jacobi (cs_multigrid_smoother.cpp:626-790), one thread per (block row,
   component): vx_n = (1 - w) vx_np + w (ad_inv_row . (rhs - vxx)), first = 1: the sweep from
   zero, vx_n = w (rhs . ad_inv_row) */
__global__ void __launch_bounds__(TPB)
k_bn_relaxed_update(long long n, int db, double wk, const double *__restrict__ ad_inv,
                    const double *__restrict__ rhs, const double *__restrict__ vxx,
                    const double *__restrict__ vx_np, double *__restrict__ vx_n, int first)
{
  const long long t = (long long)blockIdx.x*TPB + threadIdx.x;
  if (t >= n) return;
  const long long ii = t / db;
  const int jj = (int)(t - ii*db);
  const double *inv = ad_inv + (size_t)db*db*ii + (size_t)db*jj;
  const double *b = rhs + (size_t)db*ii;
  double v = 0;
  if (first) {
    for (int kk = 0; kk < db; kk++) v += b[kk] * inv[kk];
    vx_n[t] = wk * v;
  }
  else {
    const double *e = vxx + (size_t)db*ii;
    for (int kk = 0; kk < db; kk++) v += inv[kk] * (b[kk] - e[kk]);
    vx_n[t] = (1.0 - wk)*vx_np[t] + wk*v;
  }
}
}

bool launch_bn_relaxed_update(int n_blocks, int db, double wk, const double *ad_inv,
                              const double *rhs, const double *vxx, const double *vx_np,
                              double *vx_n, bool first)
{
  const long long n = (long long)n_blocks*db;
  if (n < 1) return true;
  k_bn_relaxed_update<<<(unsigned)((n + TPB - 1)/TPB), TPB, 0, ctx().stream>>>(
    n, db, wk, ad_inv, rhs, vxx, vx_np, vx_n, first ? 1 : 0);
  return post_launch("bn_relaxed_update");
}

/* block "inverse" of cs_sles_it_setup for block Jacobi with db != 3: Gauss-Jordan inverse for
   6x6 blocks, LU factors otherwise (_setup_ad_block_nn_inv, cs_sles_it_priv.cpp:178-222) */
bool launch_block_nn_inverse(int n_blocks, int db, const double *ad, double *ad_inv)
{
  if (n_blocks < 1) return true;
  if (db == 3) return launch_gauss_inverse_33(n_blocks, ad, ad_inv);
  if (db == 6) {
    int g = (n_blocks + 63)/64; if (g > 65535) g = 65535;
    k_gauss_inverse_66<<<g, 64, 0, ctx().stream>>>(n_blocks, ad, ad_inv);
  }
  else
    k_fact_lu<<<vec_grid(n_blocks*4), TPB, 0, ctx().stream>>>(n_blocks, db, ad, ad_inv);
  return post_launch("block_nn_inverse");
}

bool launch_bn_jacobi_update(int n_blocks, int db, ItState *st, const double *ad,
                             const double *ad_inv, const double *rhs, const double *vxx,
                             double *vx, bool first)
{
  Context &c = ctx();
  const int g = vec_grid(n_blocks*4);
  if (!ensure_partials(g)) return false;
  k_bn_jacobi_update<<<g, TPB, 0, c.stream>>>(n_blocks, db, st, ad, ad_inv, rhs, vxx, vx,
                                              first ? 1 : 0, c.partials, c.ticket);
  return post_launch("bn_jacobi_update");
}

bool launch_l1_inv(const _csb_matrix_t *m, double *l1_inv)
{
  if (m->n_rows < 1) return true;
  k_l1_inv<<<vec_grid(m->n_rows*4), TPB, 0, ctx().stream>>>(m->n_rows, m->row_index.p,
                                                           m->d_val.p, m->x_val.p, l1_inv);
  return post_launch("l1_inv");
}

bool launch_smoother_first(int n, int smooth_mode, double w0, const double *rhs,
                           const double *ad_inv, double *vx, const MgState *gate)
{
  if (n < 1) return true;
  k_smoother_first<<<vec_grid(n), TPB, 0, ctx().stream>>>(
    n, smooth_mode == SMOOTH_RELAXED ? 1 : 0, w0, rhs, ad_inv, vx, gate);
  return post_launch("smoother_first");
}

bool launch_copy_gated(int n, const double *src, double *dst, const MgState *gate)
{
  if (n < 1) return true;
  k_copy_mg_gated<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, src, dst, gate);
  return post_launch("copy_gated");
}

/*--------------------------------------------------------------------------
 * K-cycle vector steps (_multigrid_k_cycle, src/alge/cs_multigrid.cpp:4300-4470)
 *--------------------------------------------------------------------------*/

namespace {

/* mode 0: y = b - a*x (:4320)   1: y = a*y (:4336)   2: y = a*y + c*x (:4379)
   3: y += x + b (:4460, vx += z1 + z2)   4: y -= a*x*c   5: y = x*b (GCR, sles.cu) */
template <int MODE>
__global__ void __launch_bounds__(TPB)
k_kc_vec(int n, double a, double c, const double *__restrict__ x,
         const double *__restrict__ b, double *__restrict__ y)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    if (MODE == 0) y[i] = b[i] - a * x[i];
    else if (MODE == 1) y[i] = a * y[i];
    else if (MODE == 2) y[i] = a * y[i] + c * x[i];
    else if (MODE == 3) y[i] += x[i] + b[i];
    else if (MODE == 4) y[i] -= a * x[i] * c;            /* GCR solution update (:3471) */
    else y[i] = x[i] * b[i];                             /* Jacobi preconditioner: rk * ad_inv */
  }
}

/* the same vector operations with their coefficients read from device memory (the K-cycle's
   inner Krylov coefficients stay on the device, see k_kc_scalars) */
template <int MODE>
__global__ void __launch_bounds__(TPB)
k_kc_vec_d(int n, const double *__restrict__ pa, const double *__restrict__ pc,
           const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ y)
{
  const double a = *pa, c = (pc != nullptr) ? *pc : 0.;
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    if (MODE == 0) y[i] = b[i] - a * x[i];
    else if (MODE == 1) y[i] = a * y[i];
    else y[i] = a * y[i] + c * x[i];
  }
}

/* Coefficients of the two-step Krylov acceleration of a K-cycle level visit
   (_multigrid_k_cycle, cs_multigrid.cpp:4300-4400) from the reductions left in red[0..2]:
   sc = {rho1, ar1, |rt|^2, |r|^2, gamma, beta, ar1_ar2, ar2} */
__global__ void k_kc_scalars(int mode, const double *__restrict__ red, double *sc)
{
  if (mode == 0) {                    /* after (v.v, v.vx, vx.rhs) */
    const double rho1 = red[1], alpha1 = red[2];
    sc[0] = rho1;
    sc[1] = alpha1 / rho1;
  }
  else if (mode == 1) sc[2] = red[0];
  else if (mode == 2) sc[3] = red[0];
  else if (mode == 3) { sc[4] = red[1]; sc[5] = red[2]; }      /* gamma, beta */
  else {                              /* after vx2.rt */
    const double alpha2 = red[0];
    const double rho1 = sc[0], ar1 = sc[1], gamma = sc[4], beta = sc[5];
    const double rho2 = beta - (gamma * gamma) / rho1;
    const double ar2 = alpha2 / rho2;
    sc[6] = ar1 - (gamma / rho1) * ar2;
    sc[7] = ar2;
  }
}

/* cs_grid_prolong_row_var, overwrite form (cs_grid.cpp:9180-9190) */
__global__ void __launch_bounds__(TPB)
k_prolong_set(int f_n_rows, const int *__restrict__ coarse_row,
              const double *__restrict__ c_var, double *__restrict__ f_var)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < f_n_rows; i += gridDim.x*TPB) {
    const int ic = coarse_row[i];
    f_var[i] = (ic >= 0) ? c_var[ic] : 0.;
  }
}

} // anonymous namespace

bool launch_kc_vec(int mode, int n, double a, double c, const double *x, const double *b,
                   double *y)
{
  if (n < 1) return true;
  const int g = vec_grid(n);
  cudaStream_t st = ctx().stream;
  switch (mode) {
  case 0: k_kc_vec<0><<<g, TPB, 0, st>>>(n, a, c, x, b, y); break;
  case 1: k_kc_vec<1><<<g, TPB, 0, st>>>(n, a, c, x, b, y); break;
  case 2: k_kc_vec<2><<<g, TPB, 0, st>>>(n, a, c, x, b, y); break;
  case 3: k_kc_vec<3><<<g, TPB, 0, st>>>(n, a, c, x, b, y); break;
  case 4: k_kc_vec<4><<<g, TPB, 0, st>>>(n, a, c, x, b, y); break;
  default: k_kc_vec<5><<<g, TPB, 0, st>>>(n, a, c, x, b, y); break;
  }
  return post_launch("kc_vec");
}

bool launch_kc_vec_d(int mode, int n, const double *pa, const double *pc, const double *x,
                     const double *b, double *y)
{
  if (n < 1) return true;
  const int g = vec_grid(n);
  cudaStream_t st = ctx().stream;
  switch (mode) {
  case 0: k_kc_vec_d<0><<<g, TPB, 0, st>>>(n, pa, pc, x, b, y); break;
  case 1: k_kc_vec_d<1><<<g, TPB, 0, st>>>(n, pa, pc, x, b, y); break;
  default: k_kc_vec_d<2><<<g, TPB, 0, st>>>(n, pa, pc, x, b, y); break;
  }
  return post_launch("kc_vec_d");
}

bool launch_kc_scalars(int mode, double *sc)
{
  k_kc_scalars<<<1, 1, 0, ctx().stream>>>(mode, ctx().red_out, sc);
  return post_launch("kc_scalars");
}

bool launch_prolong_set(int f_n_rows, const int *coarse_row, const double *c_var, double *f_var)
{
  if (f_n_rows < 1) return true;
  k_prolong_set<<<vec_grid(f_n_rows*4), TPB, 0, ctx().stream>>>(f_n_rows, coarse_row, c_var, f_var);
  return post_launch("prolong_set");
}

bool launch_mg_accum_n(MgState *mg, int n_iter, int n_rows, double denom_n_g_rows_0)
{
  k_mg_accum_n<<<1, 1, 0, ctx().stream>>>(mg, n_iter, n_rows, denom_n_g_rows_0);
  return post_launch("mg_accum_n");
}

bool launch_gate_from_it(MgState *gate, const ItState *st)
{
  k_gate_from_it<<<1, 1, 0, ctx().stream>>>(gate, st);
  return post_launch("gate_from_it");
}

bool launch_fcg_init(int n, ItState *st, const double *rhs, double *vx, double *rk,
                     double *qk, bool have_residual)
{
  (void)st;
  if (n < 1) return true;
  k_fcg_init<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, rhs, vx, rk, qk,
                                                    have_residual ? 1 : 0);
  return post_launch("fcg_init");
}

bool launch_fcg_update(int n, ItState *st, const double *vk, const double *wk, double *dk,
                       double *qk, double *rk, double *vx)
{
  if (n < 1) return true;
  k_fcg_update<<<vec_grid(n), TPB, 0, ctx().stream>>>(n, st, vk, wk, dk, qk, rk, vx);
  return post_launch("fcg_update");
}

/*--------------------------------------------------------------------------
 * Matrix coefficients straight into the MSR arrays in HBM
 * (cs_matrix_compute_coeffs, src/alge/cs_matrix_building.cpp:1948-2283, iconvp = 0):
 * the solve's inputs become 8 bytes per face and per cell instead of the 12 bytes per
 * entry of an uploaded matrix.
 *--------------------------------------------------------------------------*/

namespace {

/* _sym_coeffs_msr (:136-206): thread per cell; interior faces in the order of the MSR row,
   then the cell's boundary faces in ascending face id -- the order of the reference's
   sequential boundary loop -- then relaxation and the no-Dirichlet shift (:2243-2262). */
__global__ void __launch_bounds__(TPB)
k_sym_coeffs_msr(int n_rows, const int *__restrict__ row_index,
                 const int *__restrict__ cell_i_faces, const int *__restrict__ c2b_idx,
                 const int *__restrict__ c2b, int idiffp, int ndircp, double thetap,
                 double relaxp, const double *__restrict__ cofbfp,
                 const double *__restrict__ rovsdt, const double *__restrict__ i_visc,
                 const double *__restrict__ b_visc, double *__restrict__ da,
                 double *__restrict__ ea)
{
  const double rf = 1. / relaxp;
  const double epsi_p1 = 1. + 1.e-7;
  for (int c_id = blockIdx.x*TPB + threadIdx.x; c_id < n_rows; c_id += gridDim.x*TPB) {
    double _da = rovsdt[c_id];
    const int ks = row_index[c_id], ke = row_index[c_id+1];
    if (idiffp) {
      for (int k = ks; k < ke; k++) {
        const double _xa = -thetap*i_visc[cell_i_faces[k]];
        ea[k] = _xa;
        _da -= _xa;
      }
      if (c2b_idx != nullptr)
        for (int b = c2b_idx[c_id]; b < c2b_idx[c_id+1]; b++) {
          const int f_id = c2b[b];
          _da += thetap*b_visc[f_id]*cofbfp[f_id];
        }
    }
    else
      for (int k = ks; k < ke; k++) ea[k] = 0.;
    if (relaxp < 1) _da *= rf;
    if (ndircp <= 0) _da *= epsi_p1;
    da[c_id] = _da;
  }
}

/* _coeffs_msr<is_thermal> (:407-520), convection-diffusion scalar: upwind convective part
   from the face mass flux (sign of the face relative to the cell from the mesh association),
   boundary terms in ascending face id; xcpp != nullptr multiplies the convective part by Cp. */
__global__ void __launch_bounds__(TPB)
k_coeffs_msr(int n_rows, const int *__restrict__ row_index, const int *__restrict__ col_id,
             const int *__restrict__ cell_i_faces, const short *__restrict__ c2f_sgn,
             const int *__restrict__ c2b_idx, const int *__restrict__ c2b, int idiffp,
             int ndircp, double thetap, double relaxp, const double *__restrict__ coefbp,
             const double *__restrict__ cofbfp, const double *__restrict__ rovsdt,
             const double *__restrict__ i_massflux, const double *__restrict__ b_massflux,
             const double *__restrict__ i_visc, const double *__restrict__ b_visc,
             const double *__restrict__ xcpp, double *__restrict__ da, double *__restrict__ ea)
{
  const double rf = 1. / relaxp;
  const double epsi_p1 = 1. + 1.e-7;
  const bool is_thermal = (xcpp != nullptr);
  for (int c_id = blockIdx.x*TPB + threadIdx.x; c_id < n_rows; c_id += gridDim.x*TPB) {
    double _da = rovsdt[c_id];
    const int ks = row_index[c_id], ke = row_index[c_id+1];
    for (int k = ks; k < ke; k++) {
      const int f_id = cell_i_faces[k];
      const double _i_massflux = i_massflux[f_id];
      if (c2f_sgn[k] > 0) {
        const double flui = 0.5*(_i_massflux - fabs(_i_massflux));
        const double cpi = is_thermal ? xcpp[c_id] : 1.0;
        const double xai = thetap*(cpi*flui - idiffp*i_visc[f_id]);
        ea[k] = xai;
        _da -= xai + (1.-thetap) * cpi * _i_massflux;
      }
      else {
        const double fluj = -0.5*(_i_massflux + fabs(_i_massflux));
        const double cpj = is_thermal ? xcpp[col_id[k]] : 1.0;
        const double xaj = thetap*(cpj*fluj - idiffp*i_visc[f_id]);
        ea[k] = xaj;
        _da -= xaj - (1.-thetap) * cpj * _i_massflux;
      }
    }
    if (c2b_idx != nullptr)
      for (int b = c2b_idx[c_id]; b < c2b_idx[c_id+1]; b++) {
        const int f_id = c2b[b];
        const double _b_massflux = b_massflux[f_id];
        const double flui = 0.5*(_b_massflux - fabs(_b_massflux));
        const double cpi = is_thermal ? xcpp[c_id] : 1.0;
        const double bfac =          cpi *(  flui * thetap * (coefbp[f_id] - 1.)
                                           - (1. - thetap) * _b_massflux)
                            + idiffp * thetap * b_visc[f_id] * cofbfp[f_id];
        _da += bfac;
      }
    if (relaxp < 1) _da *= rf;
    if (ndircp <= 0) _da *= epsi_p1;
    da[c_id] = _da;
  }
}

/* Vector variable, 3x3 diagonal blocks and scalar extra-diagonal (CS_MATRIX_BLOCK_D):
   _sym_coeffs_msr<3> (:250-353) when iconvp = 0, _coeffs_msr<3> (:560-692) otherwise; then
   the relaxation of all 9 block entries and the diagonal shift (:2551-2568). */
__global__ void __launch_bounds__(TPB)
k_coeffs_msr_v3(int n_rows, const int *__restrict__ row_index,
                const int *__restrict__ cell_i_faces, const short *__restrict__ c2f_sgn,
                const int *__restrict__ c2b_idx, const int *__restrict__ c2b, int iconvp,
                int idiffp, int ndircp, double thetap, double relaxp,
                const double *__restrict__ coefbp, const double *__restrict__ cofbfp,
                const double *__restrict__ fimp, const double *__restrict__ i_massflux,
                const double *__restrict__ b_massflux, const double *__restrict__ i_visc,
                const double *__restrict__ b_visc, double *__restrict__ da,
                double *__restrict__ ea)
{
  const double rf = 1. / relaxp;
  const double epsi_p1 = 1. + 1.e-7;
  for (int c_id = blockIdx.x*TPB + threadIdx.x; c_id < n_rows; c_id += gridDim.x*TPB) {
    double _da[9];
    for (int q = 0; q < 9; q++) _da[q] = (fimp != nullptr) ? fimp[(size_t)c_id*9 + q] : 0.;
    const int ks = row_index[c_id], ke = row_index[c_id+1];
    for (int k = ks; k < ke; k++) {
      const int f_id = cell_i_faces[k];
      if (!iconvp) {
        if (idiffp) {
          const double _xa = -thetap*i_visc[f_id];
          ea[k] = _xa;
          _da[0] -= _xa; _da[4] -= _xa; _da[8] -= _xa;
        }
        else
          ea[k] = 0.;
      }
      else {
        const double _i_massflux = i_massflux[f_id];
        double fac;
        if (c2f_sgn[k] > 0) {
          const double flui = 0.5*(_i_massflux - fabs(_i_massflux));
          const double xai = thetap*(flui - idiffp*i_visc[f_id]);
          ea[k] = xai;
          fac = xai + (1.-thetap) * _i_massflux;
        }
        else {
          const double fluj = -0.5*(_i_massflux + fabs(_i_massflux));
          const double xaj = thetap*(fluj - idiffp*i_visc[f_id]);
          ea[k] = xaj;
          fac = xaj - (1.-thetap) * _i_massflux;
        }
        _da[0] -= fac; _da[4] -= fac; _da[8] -= fac;
      }
    }
    if (c2b_idx != nullptr && (iconvp || idiffp))
      for (int b = c2b_idx[c_id]; b < c2b_idx[c_id+1]; b++) {
        const int f_id = c2b[b];
        if (!iconvp) {
          for (int q = 0; q < 9; q++)
            _da[q] += thetap * b_visc[f_id] * cofbfp[(size_t)f_id*9 + q];
        }
        else {
          const double _b_massflux = b_massflux[f_id];
          const double flui = 0.5*(_b_massflux - fabs(_b_massflux));
          for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
              const double d_ij = ((i == j) ? 1. : 0.);
              const double diag = d_ij * (thetap * flui + (1. - thetap) * _b_massflux);
              _da[3*i+j] += - diag + thetap * (  flui * coefbp[(size_t)f_id*9 + 3*i + j]
                                               + idiffp * b_visc[f_id] * cofbfp[(size_t)f_id*9 + 3*i + j]);
            }
        }
      }
    if (relaxp < 1)
      for (int q = 0; q < 9; q++) _da[q] *= rf;
    if (ndircp <= 0) {
      _da[0] = epsi_p1 * _da[0]; _da[4] = epsi_p1 * _da[4]; _da[8] = epsi_p1 * _da[8];
    }
    for (int q = 0; q < 9; q++) da[(size_t)c_id*9 + q] = _da[q];
  }
}

} // anonymous namespace

bool launch_coeffs_msr_v3(const _csb_matrix_t *m, int iconvp, int idiffp, int ndircp,
                          double thetap, double relaxp, const double *coefbp,
                          const double *cofbfp, const double *fimp, const double *i_massflux,
                          const double *b_massflux, const double *i_visc, const double *b_visc,
                          double *d_val, double *x_val)
{
  if (m->n_rows < 1) return true;
  k_coeffs_msr_v3<<<vec_grid(m->n_rows*4), TPB, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->c2f.p, m->c2f_sgn.p,
    m->has_b_faces ? m->cell_b_faces_idx.p : nullptr, m->cell_b_faces.p, iconvp, idiffp,
    ndircp, thetap, relaxp, coefbp, cofbfp, fimp, i_massflux, b_massflux, i_visc, b_visc,
    d_val, x_val);
  return post_launch("coeffs_msr_v3");
}

bool launch_coeffs_msr(const _csb_matrix_t *m, int idiffp, int ndircp, double thetap,
                       double relaxp, const double *coefbp, const double *cofbfp,
                       const double *rovsdt, const double *i_massflux, const double *b_massflux,
                       const double *i_visc, const double *b_visc, const double *xcpp,
                       double *d_val, double *x_val)
{
  if (m->n_rows < 1) return true;
  k_coeffs_msr<<<vec_grid(m->n_rows*4), TPB, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->col_id.p, m->c2f.p, m->c2f_sgn.p,
    m->has_b_faces ? m->cell_b_faces_idx.p : nullptr, m->cell_b_faces.p, idiffp, ndircp,
    thetap, relaxp, coefbp, cofbfp, rovsdt, i_massflux, b_massflux, i_visc, b_visc, xcpp,
    d_val, x_val);
  return post_launch("coeffs_msr");
}

bool launch_sym_coeffs_msr(const _csb_matrix_t *m, int idiffp, int ndircp, double thetap,
                           double relaxp, const double *cofbfp, const double *rovsdt,
                           const double *i_visc, const double *b_visc, double *d_val,
                           double *x_val)
{
  if (m->n_rows < 1) return true;
  k_sym_coeffs_msr<<<vec_grid(m->n_rows*4), TPB, 0, ctx().stream>>>(
    m->n_rows, m->row_index.p, m->c2f.p,
    m->has_b_faces ? m->cell_b_faces_idx.p : nullptr, m->cell_b_faces.p, idiffp, ndircp,
    thetap, relaxp, cofbfp, rovsdt, i_visc, b_visc, d_val, x_val);
  return post_launch("sym_coeffs_msr");
}

/*--------------------------------------------------------------------------
 * Halo packing (cs_halo_sync_pack, src/base/cs_halo.cpp:1739-1834;
 * _gather_full_cu, src/base/cs_halo_cuda.cu:74-160)
 *--------------------------------------------------------------------------*/

namespace {
__global__ void __launch_bounds__(TPB)
k_halo_pack(int n_send, const int *__restrict__ send_list, const double *__restrict__ var,
            double *__restrict__ buf)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n_send; i += gridDim.x*TPB)
    buf[i] = var[send_list[i]];
}
}

/*--------------------------------------------------------------------------
 * Index helpers of the coarse-level merge (sles.cu, mg_merge_coarse_levels)
 *--------------------------------------------------------------------------*/

namespace {
/* mode 0: out[i] = shift + i                  (merged ids of the local rows)
   mode 1: out[i] = map[in[i]]                 (column ids -> merged ids)
   mode 2: out[i] = in[i+1] - in[i]            (row lengths)
   mode 3: out[i] = in[i] >= 0 ? in[i] + shift : in[i]   (coarse rows -> merged coarse rows) */
__global__ void __launch_bounds__(TPB)
k_merge_index(int mode, int n, int shift, const int *__restrict__ in,
              const int *__restrict__ map, int *__restrict__ out)
{
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) {
    if (mode == 0) out[i] = shift + i;
    else if (mode == 1) out[i] = map[in[i]];
    else if (mode == 2) out[i] = in[i+1] - in[i];
    else out[i] = (in[i] >= 0) ? in[i] + shift : in[i];
  }
}
}

bool launch_merge_index(int mode, int n, int shift, const int *in, const int *map, int *out)
{
  if (n < 1) return true;
  k_merge_index<<<vec_grid(n), TPB, 0, ctx().stream>>>(mode, n, shift, in, map, out);
  return post_launch("merge_index");
}

/*--------------------------------------------------------------------------
 * Gauss-Seidel family: the rows of ONE dependency level (sles.cu, GsPlan).  A row's update is
 * the reference's, term by term: vx0 = rhs; vx0 -= (a_ij * vx_j) over the row's entries in
 * order (all of them, the lower part up to the first column > i, or the upper part from the
 * last entry down to the first column < i); vx_i = vx0 * ad_inv_i
 * (cs_multigrid_smoother.cpp:1095-1107, :1243-1256, :1436-1449, :1566-1579).  Rows of a level
 * do not depend on each other, every row they read from an earlier level is final: the
 * sequential sweep, reproduced bit for bit.  res (solver form, cs_sles_it.cpp:4003-4024,
 * :4262-4284): r_i = d_i (vx_i_new - vx_i_old), summed in squares by k_gs_residual.
 *--------------------------------------------------------------------------*/

namespace {

struct GsK {
  const int *rows, *row_index, *col_id;
  const double *x_val, *ad_inv, *rhs, *d_val;
  double *vx, *res;
  int mode;
  const MgState *gate;
  const ItState *st;
};

__device__ __forceinline__ void gs_row(const GsK &a, int ii)
{
  const int ks = a.row_index[ii], ke = a.row_index[ii+1];
  double vx0 = a.rhs[ii];
  if (a.mode == 0)
    for (int jj = ks; jj < ke; jj++) vx0 -= (a.x_val[jj]*a.vx[a.col_id[jj]]);
  else if (a.mode == 1)
    for (int jj = ks; jj < ke; jj++) {
      if (a.col_id[jj] > ii) break;
      vx0 -= (a.x_val[jj]*a.vx[a.col_id[jj]]);
    }
  else
    for (int jj = ke - 1; jj > ks - 1; jj--) {
      if (a.col_id[jj] < ii) break;
      vx0 -= (a.x_val[jj]*a.vx[a.col_id[jj]]);
    }
  vx0 *= a.ad_inv[ii];
  if (a.res != nullptr) a.res[ii] = a.d_val[ii] * (vx0 - a.vx[ii]);
  a.vx[ii] = vx0;
}

__global__ void __launch_bounds__(TPB)
k_gs_rows(GsK a, int n)
{
  if (a.gate != nullptr && a.gate->cvg != CSB_SLES_ITERATING) return;
  if (a.st != nullptr && a.st->cvg != CSB_SLES_ITERATING) return;
  const int t = blockIdx.x*TPB + threadIdx.x;
  if (t < n) gs_row(a, a.rows[t]);
}

/* a run of consecutive small levels in ONE CTA (levels l0 .. l1-1 in the given direction,
   a __syncthreads between levels): coarse matrices and the first / last wavefronts of fine
   ones, where a launch per level would be all latency */
__global__ void __launch_bounds__(1024)
k_gs_chain(GsK a, const int *__restrict__ level_start, int l0, int l1, int backward)
{
  if (a.gate != nullptr && a.gate->cvg != CSB_SLES_ITERATING) return;
  if (a.st != nullptr && a.st->cvg != CSB_SLES_ITERATING) return;
  for (int k = 0; k < l1 - l0; k++) {
    const int l = backward ? l1 - 1 - k : l0 + k;
    const int s = level_start[l], e = level_start[l+1];
    for (int t = s + threadIdx.x; t < e; t += 1024) gs_row(a, a.rows[t]);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(TPB)
k_gs_residual(int n, ItState *st, const double *__restrict__ res, double *partials,
              unsigned *ticket)
{
  if (st->cvg != CSB_SLES_ITERATING) return;
  double v[1] = {0.};
  for (int i = blockIdx.x*TPB + threadIdx.x; i < n; i += gridDim.x*TPB) v[0] += (res[i]*res[i]);
  grid_sum_finalize<1>(v, partials, ticket, [st](const double *s) {
    st->n_iter += 1;
    double residual = (st->precision > 0.) ? sqrt(s[0]) : -1.;   /* of the previous iterate */
    if (st->n_iter == 1) st->initial_residual = residual;
    st->cvg = it_cvg_test(st, st->n_iter, residual);
  });
}

GsK gs_args(const int *rows, const _csb_matrix_t *m, const double *ad_inv, const double *rhs,
            double *vx, int mode, const MgState *gate, const ItState *st, double *res)
{
  GsK a;
  a.rows = rows; a.row_index = m->row_index.p; a.col_id = m->col_id.p; a.x_val = m->x_val.p;
  a.ad_inv = ad_inv; a.rhs = rhs; a.d_val = m->d_val.p; a.vx = vx; a.res = res; a.mode = mode;
  a.gate = gate; a.st = st;
  return a;
}
}

bool launch_gs_rows(int n, const int *rows, const _csb_matrix_t *m, const double *ad_inv,
                    const double *rhs, double *vx, int mode, const MgState *gate,
                    const ItState *st, double *res)
{
  if (n < 1) return true;
  k_gs_rows<<<(n + TPB - 1)/TPB, TPB, 0, ctx().stream>>>(
    gs_args(rows, m, ad_inv, rhs, vx, mode, gate, st, res), n);
  return post_launch("gs_rows");
}

bool launch_gs_chain(const int *rows, const int *level_start, int l0, int l1, bool backward,
                     const _csb_matrix_t *m, const double *ad_inv, const double *rhs,
                     double *vx, int mode, const MgState *gate, const ItState *st, double *res)
{
  if (l1 <= l0) return true;
  k_gs_chain<<<1, 1024, 0, ctx().stream>>>(
    gs_args(rows, m, ad_inv, rhs, vx, mode, gate, st, res), level_start, l0, l1,
    backward ? 1 : 0);
  return post_launch("gs_chain");
}

bool launch_gs_residual(int n, ItState *st, const double *res)
{
  Context &c = ctx();
  const int g = vec_grid(n);
  if (!ensure_partials(g)) return false;
  k_gs_residual<<<g, TPB, 0, c.stream>>>(n, st, res, c.partials, c.ticket);
  return post_launch("gs_residual");
}

bool launch_halo_pack(int n_send, const int *send_list, const double *var, double *buf)
{
  if (n_send < 1) return true;
  k_halo_pack<<<vec_grid(n_send*4), TPB, 0, ctx().stream>>>(n_send, send_list, var, buf);
  return post_launch("halo_pack");
}

} // namespace csb
