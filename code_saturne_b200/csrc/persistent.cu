/*============================================================================
 * persistent.cu -- persistent cooperative kernels: a whole cs_sles_it PCG
 * solve, or a whole cs_multigrid V-cycle solve (all cycles, all levels), in
 * ONE launch.
 *
 * Why: the hot path is a long chain of short HBM-bound phases (SpMV + 3 dots,
 * 5-vector update, residual, restrict, prolong ...) over a hierarchy whose
 * levels shrink by 3x each; with one kernel per phase the chain costs ~290
 * launches per cycle, each with its own tail (block reduction + atomic ticket)
 * and, on the coarse levels, pure launch latency.  Here the grid is sized to
 * the machine (k CTAs per SM, all co-resident: cooperative launch), every CTA
 * walks the same device-side control flow, phases are separated by a grid
 * barrier (~2 us) and reductions are "every CTA sums all per-CTA partials in
 * the same fixed order": all CTAs obtain bit-identical Krylov scalars and take
 * identical branches, with no host round trip and no broadcast.
 *
 * Arithmetic and evaluation order per entry are those of the reference:
 *   SpMV        _mat_vec_p_l_msr            src/alge/cs_matrix_spmv.cpp:2103-2165
 *   PCG         _conjugate_gradient_sr_p0   src/alge/cs_sles_it.cpp:1441-1645
 *   V-cycle     _multigrid_v_cycle          src/alge/cs_multigrid.cpp:3127-3542
 *   cycle test  _convergence_test           src/alge/cs_multigrid.cpp:2927-3035
 *   restrict / prolong                      src/alge/cs_grid.cpp:8987-9210
 *============================================================================*/

#include "csb_internal.h"
#include "persistent.h"

#include <cfloat>
#include <cmath>

namespace csb {

namespace {

constexpr int PT = 256;   /* threads per CTA */
#ifndef CSB_MIN_CTAS
#define CSB_MIN_CTAS 3
#endif

/*--------------------------------------------------------------------------
 * Grid-wide barrier (all CTAs are co-resident: cooperative launch).
 * bar[0] = arrival counter, bar[1] = generation.
 *--------------------------------------------------------------------------*/

__device__ __forceinline__ void grid_barrier(unsigned *bar)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned *gen = bar + 1;
    const unsigned my_gen = *gen;
    __threadfence();
    if (atomicAdd(bar, 1u) == gridDim.x - 1) {
      bar[0] = 0;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    }
    else {
      while (*gen == my_gen) { __nanosleep(20); }
    }
    __threadfence();      /* acquire; also invalidates this SM's L1 */
  }
  __syncthreads();
}

template <bool SINGLE>
__device__ __forceinline__ void phase_barrier(unsigned *bar)
{
  if (SINGLE) __syncthreads();
  else grid_barrier(bar);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

/* CTA-wide sum of Q values, result broadcast to every thread. */
template <int Q>
__device__ __forceinline__ void cta_sum_bcast(double (&v)[Q], double *s_red /* Q*8 + Q */)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const double w = warp_sum(v[q]);
    if (lane == 0) s_red[q*8 + wid] = w;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      double w = (lane < PT/32) ? s_red[q*8 + lane] : 0.0;
      w = warp_sum(w);
      if (lane == 0) s_red[Q*8 + q] = w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < Q; q++) v[q] = s_red[Q*8 + q];
  __syncthreads();
}

/* Deterministic sum over the grid and, with a communicator, over all ranks:
   the result is identical in every thread of every CTA of every rank.
   SINGLE: only this CTA takes part on this rank.
   Cross-rank stage: block 0 stores the rank's totals into slot [parity][rank]
   of every peer's header and publishes the epoch in the matching flag; every
   CTA polls the (local) flags of all ranks and adds the totals in rank order.
   Two slots suffice: a rank cannot start all-reduce e+2 before every rank has
   contributed to e+1, i.e. has finished reading e. */
template <bool SINGLE, int Q>
__device__ __forceinline__ void all_sum(double (&v)[Q], double *partials, unsigned *bar,
                                        double *s_red, const CommDev *C)
{
  unsigned e = 0;
  if (C != nullptr)      /* read before the grid barrier: block 0 bumps it after */
    e = *(volatile unsigned *)&C->local->x_epoch + 1;
  cta_sum_bcast<Q>(v, s_red);
  if (!SINGLE) {
    const unsigned G = gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < Q; q++) partials[q*G + blockIdx.x] = v[q];
    }
    grid_barrier(bar);
#pragma unroll
    for (int q = 0; q < Q; q++) {
      double w = 0.0;
      for (unsigned i = threadIdx.x; i < G; i += PT)
        w += __ldcg(partials + q*G + i);
      v[q] = w;
    }
    cta_sum_bcast<Q>(v, s_red);
  }
  if (C == nullptr) return;

  const int R = C->n_ranks, me = C->rank, par = e & 1;
  CommHeader *H = (CommHeader *)C->peer[me];
  if ((SINGLE || blockIdx.x == 0) && threadIdx.x == 0) {
    *(volatile unsigned *)&C->local->x_epoch = e;
    for (int r = 0; r < R; r++) {
      CommHeader *P = (CommHeader *)C->peer[r];
#pragma unroll
      for (int q = 0; q < Q; q++) P->xsum[par][me][q] = v[q];
    }
    __threadfence_system();
    for (int r = 0; r < R; r++)
      *(volatile unsigned *)&((CommHeader *)C->peer[r])->xflag[par][me] = e;
  }
  double *s_x = s_red + 40;        /* [Q][COMM_MAX_RANKS] */
  if (threadIdx.x < R) {
    spin_until_ge(&H->xflag[par][threadIdx.x], e, C->timeout_ns, 6);
    __threadfence();
#pragma unroll
    for (int q = 0; q < Q; q++)
      s_x[q*COMM_MAX_RANKS + threadIdx.x] = __ldcg(&H->xsum[par][threadIdx.x][q]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < Q; q++) {
    double w = 0.0;
    for (int r = 0; r < R; r++) w += s_x[q*COMM_MAX_RANKS + r];
    v[q] = w;
  }
  __syncthreads();
}

/*--------------------------------------------------------------------------
 * Ghost exchange of one level, fused into the solver kernel
 * (cs_halo_sync_var: pack + send/recv + wait, src/base/cs_halo.cpp:1739-2220).
 * Every CTA gathers its share of the send list and stores it straight into the
 * neighbours' ghost buffers; the last CTA to finish publishes the exchange
 * count in each neighbour's flag word; then every CTA waits for the flags of
 * its own neighbours.  Returns the buffer holding this exchange's ghost values
 * (ghost column c is at [c - n_rows]); nullptr when there is nothing to wait for.
 * x must be complete (call after a grid barrier).
 *--------------------------------------------------------------------------*/

template <bool SINGLE>
__device__ __forceinline__ const double *halo_exchange(const MgLevelDesc &L, const double *x)
{
  const DistLevel *D = L.dist;
  if (D == nullptr) return nullptr;
  const int n_dom = D->n_dom;
  if (n_dom == 0) return nullptr;
  const unsigned k = *(volatile unsigned *)D->count + 1;
  const int par = k & 1;
  const int first = SINGLE ? threadIdx.x : blockIdx.x*PT + threadIdx.x;
  const int step = SINGLE ? PT : gridDim.x*PT;
  const int *__restrict__ send_list = D->send_list;
  for (int i = 0; i < n_dom; i++) {
    const int s0 = D->send_index[i], s1 = D->send_index[i+1];
    double *dst = D->push_dst[i][par];
    for (int e = s0 + first; e < s1; e += step)
      dst[e - s0] = x[send_list[e]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    bool last = true;
    if (!SINGLE) last = (atomicInc(D->ticket, gridDim.x - 1) == gridDim.x - 1);
    if (last) {
      *(volatile unsigned *)D->count = k;
      __threadfence_system();
      for (int i = 0; i < n_dom; i++)
        *(volatile unsigned *)D->flag_remote[i] = k;
    }
  }
  if (threadIdx.x < n_dom) {
    spin_until_ge(D->flag_local + threadIdx.x, k, D->timeout_ns, 7);
    __threadfence();
  }
  __syncthreads();
  return D->gbuf[par];
}

/*--------------------------------------------------------------------------
 * SpMV over the slabs owned by this CTA.
 * A slab = L.slab_rows (<= 256) consecutive rows; its off-diagonal entries
 * are contiguous in memory: streamed with coalesced loads, products staged in
 * shared memory, then one thread per row sums its products in CSR order and
 * calls epi(row, off_diagonal_sum).
 *--------------------------------------------------------------------------*/

template <bool SINGLE, class Epi>
__device__ __forceinline__ void spmv_slabs(const MgLevelDesc &L, const double *x,
                                           const double *ghost, double *s_prod, Epi epi)
{
  /* column c < n_rows: x[c]; ghost column: ghost[c - n_rows] */
  const double *xg = (ghost != nullptr) ? ghost - L.n_rows : x;
  const int tid = threadIdx.x;
  const int first = SINGLE ? 0 : blockIdx.x;
  const int step = SINGLE ? 1 : gridDim.x;
  const int n_slabs = L.n_slabs, slab_rows = L.slab_rows, n_rows = L.n_rows;
  const int    *__restrict__ row_index = L.row_index;
  const int    *__restrict__ col_id = L.col_id;
  const double *__restrict__ x_val = L.x_val;
  for (int slab = first; slab < n_slabs; slab += step) {
    const int r0 = slab * slab_rows;
    const int r1 = min(r0 + slab_rows, n_rows);
    const int k0 = __ldg(row_index + r0);
    const int nk = __ldg(row_index + r1) - k0;
    const int    *__restrict__ ci = col_id + k0;
    const double *__restrict__ xv = x_val + k0;
    int i = tid;
    for (; i + 3*PT < nk; i += 4*PT) {
      const int c0 = __ldg(ci + i),        c1 = __ldg(ci + i + PT);
      const int c2 = __ldg(ci + i + 2*PT), c3 = __ldg(ci + i + 3*PT);
      const double m0 = __ldg(xv + i),        m1 = __ldg(xv + i + PT);
      const double m2 = __ldg(xv + i + 2*PT), m3 = __ldg(xv + i + 3*PT);
      const double x0 = (c0 < n_rows) ? x[c0] : xg[c0], x1 = (c1 < n_rows) ? x[c1] : xg[c1];
      const double x2 = (c2 < n_rows) ? x[c2] : xg[c2], x3 = (c3 < n_rows) ? x[c3] : xg[c3];
      s_prod[i]        = m0*x0;
      s_prod[i + PT]   = m1*x1;
      s_prod[i + 2*PT] = m2*x2;
      s_prod[i + 3*PT] = m3*x3;
    }
    for (; i < nk; i += PT) {
      const int c = __ldg(ci + i);
      s_prod[i] = __ldg(xv + i) * ((c < n_rows) ? x[c] : xg[c]);
    }
    __syncthreads();
    const int r = r0 + tid;
    if (r < r1) {
      const int ks = __ldg(row_index + r) - k0;
      const int ke = __ldg(row_index + r + 1) - k0;
      double sii = 0.0;
      for (int k = ks; k < ke; k++) sii += s_prod[k];
      epi(r, sii);
    }
    __syncthreads();
  }
}

/* elementwise loop over the rows owned by this CTA */
template <bool SINGLE, class F>
__device__ __forceinline__ void row_loop(int n, F f)
{
  const int first = SINGLE ? threadIdx.x : blockIdx.x*PT + threadIdx.x;
  const int step = SINGLE ? PT : gridDim.x*PT;
  for (int i = first; i < n; i += step) f(i);
}

/* _convergence_test, src/alge/cs_sles_it.cpp:276-324 (pure part) */
__device__ __forceinline__ int it_cvg_eval(double precision, double r_norm, int n_max_iter,
                                           double initial_residual, int n_iter,
                                           double residual)
{
  if (precision > 0 && residual < precision * r_norm) return CSB_SLES_CONVERGED;
  else if (n_iter >= n_max_iter) return CSB_SLES_MAX_ITERATION;
  else if (residual > initial_residual * 10000.0 && residual > 100.) return CSB_SLES_DIVERGED;
  if (isnan(residual) || isinf(residual)) return CSB_SLES_DIVERGED;
  return CSB_SLES_ITERATING;
}

__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

/* lead thread only: add the time since the previous mark to prof[level][phase] */
__device__ __forceinline__ void prof_mark(MgState *ms, unsigned long long &t_last, int level,
                                          int phase)
{
  const unsigned long long t = global_ns();
  if (level < PROF_LEVELS) ms->prof[level][phase] += t - t_last;
  t_last = t;
}

struct PcgResult { int n_iter; int cvg; double residual; double initial_residual; };

/*--------------------------------------------------------------------------
 * PCG + Jacobi, single reduction (cs_sles_it.cpp:1441-1645).
 * Every thread of every participating CTA returns the same PcgResult.
 * smoother: skip the SpMV + dots that only feed the convergence test after
 * the last iteration (vx is not affected).
 * On exit vx is complete only after the caller's next barrier.
 *--------------------------------------------------------------------------*/

template <bool SINGLE>
__device__ __noinline__ PcgResult pcg_run(const MgLevelDesc &L, const double *rhs, double *vx,
                             bool vx_ini_nonzero, int n_max_iter, double precision,
                             double r_norm, bool smoother, const PersistentCommon &P,
                             ItState *st, double *s_prod, double *s_red)
{
  const int n = L.n_rows;
  const size_t ws = P.work_stride;
  double *rk = P.work, *dk = P.work + ws, *gk = P.work + 2*ws, *zk = P.work + 3*ws,
         *sk = P.work + 4*ws;
  const double *__restrict__ ad_inv = L.ad_inv;
  const double *__restrict__ d_val = L.d_val;
  const bool rec = (st != nullptr) && (SINGLE || blockIdx.x == 0) && threadIdx.x == 0;
  double *part0 = P.partials, *part1 = P.partials + 4*(size_t)gridDim.x;

  PcgResult R;
  R.n_iter = 0; R.cvg = CSB_SLES_ITERATING; R.residual = -1.; R.initial_residual = -1.;
  if (rec) st->n_hist = 0;

  /* residual and descent direction (:1502-1516) */
  const CommDev *comm = P.comm;
  if (vx_ini_nonzero) {
    const double *gh = halo_exchange<SINGLE>(L, vx);
    spmv_slabs<SINGLE>(L, vx, gh, s_prod, [&](int r, double sii) {
      const double rr = (sii + d_val[r]*vx[r]) - rhs[r];
      rk[r] = rr;
      dk[r] = rr * ad_inv[r];
    });
  }
  else {
    row_loop<SINGLE>(n, [&](int i) {
      vx[i] = 0;
      const double rr = -rhs[i];
      rk[i] = rr;
      dk[i] = rr * ad_inv[i];
    });
  }
  phase_barrier<SINGLE>(P.barrier);

  /* zk = A.dk and the first 3 dots (:1521-1535) */
  double v[3] = {0., 0., 0.};
  const double *gh_d = halo_exchange<SINGLE>(L, dk);
  spmv_slabs<SINGLE>(L, dk, gh_d, s_prod, [&](int r, double sii) {
    const double d = dk[r];
    const double z = sii + d_val[r]*d;
    zk[r] = z;
    const double rr = rk[r];
    v[0] += rr*rr; v[1] += rr*d; v[2] += d*z;
  });
  all_sum<SINGLE, 3>(v, part0, P.barrier, s_red, comm);

  double ro_0 = v[1], ro_1 = v[2], rk_gkm1 = 0., alpha = 0., beta = 0.;
  R.residual = sqrt(v[0]);
  R.initial_residual = R.residual;
  R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, 0, R.residual);
  if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;

  if (R.cvg == CSB_SLES_ITERATING) {
    R.n_iter = 1;
    const double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    rk_gkm1 = ro_0;
    double w[1] = {0.};
    row_loop<SINGLE>(n, [&](int i) {      /* :1550-1558 */
      vx[i] += (alpha * dk[i]);
      const double rr = rk[i] + (alpha * zk[i]);
      rk[i] = rr;
      gk[i] = rr * ad_inv[i];
      w[0] += rr*rr;
    });
    all_sum<SINGLE, 1>(w, part1, P.barrier, s_red, comm);
    if (SINGLE) __syncthreads();
    R.residual = sqrt(w[0]);
    R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, 1, R.residual);
    if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;
  }

  while (R.cvg == CSB_SLES_ITERATING) {
    if (smoother && R.n_iter >= n_max_iter) break;
    v[0] = 0.; v[1] = 0.; v[2] = 0.;
    const double *gh_g = halo_exchange<SINGLE>(L, gk);
    spmv_slabs<SINGLE>(L, gk, gh_g, s_prod, [&](int r, double sii) {   /* :1591-1600 */
      const double g = gk[r];
      const double s = sii + d_val[r]*g;
      sk[r] = s;
      const double rr = rk[r];
      v[0] += rr*rr; v[1] += rr*g; v[2] += g*s;
    });
    all_sum<SINGLE, 3>(v, part0, P.barrier, s_red, comm);
    const double residual = sqrt(v[0]);
    const double rk_gk = v[1], gk_sk = v[2];
    R.residual = residual;
    if (R.n_iter > 1) {
      R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, R.n_iter,
                          residual);
      if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = residual;
    }
    if (R.cvg != CSB_SLES_ITERATING) break;
    R.n_iter += 1;
    beta = (fabs(rk_gkm1) > DBL_MIN) ? rk_gk / rk_gkm1 : 0.;    /* :1614-1621 */
    rk_gkm1 = rk_gk;
    ro_1 = gk_sk - beta*beta*ro_1;
    ro_0 = rk_gk;
    const double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    if (smoother && R.n_iter >= n_max_iter) {
      /* last sweep of a smoother: only vx is used afterwards */
      row_loop<SINGLE>(n, [&](int i) {
        const double d = gk[i] + (beta * dk[i]);
        vx[i] += alpha * d;
      });
    }
    else {
      row_loop<SINGLE>(n, [&](int i) {        /* :1623-1631 */
        const double d = gk[i] + (beta * dk[i]);
        dk[i] = d;
        vx[i] += alpha * d;
        const double z = sk[i] + (beta * zk[i]);
        zk[i] = z;
        const double rr = rk[i] + alpha * z;
        rk[i] = rr;
        gk[i] = rr * ad_inv[i];
      });
    }
    phase_barrier<SINGLE>(P.barrier);
  }

  if (rec) {
    st->precision = precision; st->r_norm = r_norm;
    st->residual = R.residual; st->initial_residual = R.initial_residual;
    st->n_iter = R.n_iter; st->n_max_iter = n_max_iter; st->cvg = R.cvg;
    st->alpha = alpha; st->beta = beta; st->ro_0 = ro_0; st->ro_1 = ro_1;
    st->rk_gkm1 = rk_gkm1;
  }
  return R;
}

/*--------------------------------------------------------------------------
 * Kernel 1: one PCG solve (cs_sles_it_solve with CS_SLES_PCG + Jacobi)
 *--------------------------------------------------------------------------*/

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_pcg_persistent(const MgLevelDesc L, const double *rhs, double *vx, int vx_ini_nonzero,
                 int n_max_iter, double precision, double r_norm,
                 const PersistentCommon P, ItState *st, int single)
{
  extern __shared__ double s_dyn[];
  double *s_red = s_dyn;
  double *s_prod = s_dyn + PERSISTENT_S_RED;
  if (single) {
    if (blockIdx.x == 0)
      pcg_run<true>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, false,
                    P, st, s_prod, s_red);
  }
  else
    pcg_run<false>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, false,
                   P, st, s_prod, s_red);
}

/*--------------------------------------------------------------------------
 * Kernel 2: whole multigrid solve
 *--------------------------------------------------------------------------*/

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_mg_persistent(const MgSolveParams S, const PersistentCommon P)
{
  extern __shared__ double s_dyn[];
  double *s_red = s_dyn;
  double *s_prod = s_dyn + PERSISTENT_S_RED;

  const MgLevelDesc *__restrict__ LV = S.levels;
  MgState *ms = S.ms;
  const bool lead = (blockIdx.x == 0 && threadIdx.x == 0);
  const int coarsest = S.n_levels - 1;
  const double denom_n_rows_0 = 1.0 / LV[0].n_g_rows;     /* global rows (:3341) */

  int cvg = CSB_SLES_ITERATING;
  int n_equiv_iter = 0;
  double initial_residual = -1., residual = -1.;
  int n_cycles = 0;
  unsigned long long t_last = lead ? global_ns() : 0ull;

  for (n_cycles = 0; cvg == CSB_SLES_ITERATING; n_cycles++) {
    const int cycle_id = n_cycles + 1;
    bool end_cycle = false;

    /* ---------------- descent (cs_multigrid.cpp:3231-3365) ---------------- */
    for (int level = 0; level < coarsest; level++) {
      const MgLevelDesc &L = LV[level];
      const MgLevelDesc &C = LV[level+1];
      const double *rhs_lv = (level == 0) ? S.rhs : L.rhs;
      double *vx_lv = (level == 0) ? S.vx : L.vx;
      const bool ini = (level == 0) ? ((n_cycles > 0) || S.vx_ini_nonzero) : false;
      const bool single = (L.n_rows <= S.single_rows);

      PcgResult R;
      if (single) {
        if (blockIdx.x == 0)
          R = pcg_run<true>(L, rhs_lv, vx_lv, ini, S.n_max_iter[0],
                            S.precision*S.precision_mult[0], S.r_norm, true, P,
                            L.st_descent, s_prod, s_red);
        grid_barrier(P.barrier);
        /* other CTAs need the iteration count for n_equiv_iter */
        R.n_iter = *((volatile int *)&L.st_descent->n_iter);
        R.cvg = *((volatile int *)&L.st_descent->cvg);
        R.initial_residual = *((volatile double *)&L.st_descent->initial_residual);
      }
      else {
        R = pcg_run<false>(L, rhs_lv, vx_lv, ini, S.n_max_iter[0],
                           S.precision*S.precision_mult[0], S.r_norm, true, P,
                           L.st_descent, s_prod, s_red);
        grid_barrier(P.barrier);
      }
      if (lead) prof_mark(ms, t_last, level, MG_PROF_DESCENT);
      if (level == 0 && cycle_id == 1) initial_residual = R.initial_residual;
      if (R.cvg < CSB_SLES_BREAKDOWN) { end_cycle = true; cvg = CSB_SLES_DIVERGED; break; }

      /* wr = rhs - A.vx (:3296-3301), fine-level norm for the cycle test */
      double w[1] = {0.};
      {
        const double *__restrict__ d_val = L.d_val;
        double *wr = S.wr;
        const double *gh = halo_exchange<false>(L, vx_lv);
        spmv_slabs<false>(L, vx_lv, gh, s_prod, [&](int r, double sii) {
          const double y = sii + d_val[r]*vx_lv[r];
          const double t = rhs_lv[r] - y;
          wr[r] = t;
          w[0] += t*t;
        });
      }
      if (level == 0) {
        all_sum<false, 1>(w, P.partials + 8*(size_t)gridDim.x, P.barrier, s_red, P.comm);
        residual = sqrt(w[0]);
        const double ir = (cycle_id == 1) ? residual : initial_residual;   /* :2967 */
        if (lead && ms->n_cycles_res < CYCLE_MAX) ms->cycle_res[ms->n_cycles_res++] = residual;
        if (residual < S.precision * S.r_norm) cvg = CSB_SLES_CONVERGED;
        else if (cycle_id > S.n_max_cycles) cvg = CSB_SLES_MAX_ITERATION;
        else if (residual > ir * 10000.0 && residual > 100.) cvg = CSB_SLES_DIVERGED;
        else if (isnan(residual) || isinf(residual)) cvg = CSB_SLES_DIVERGED;
      }
      else
        grid_barrier(P.barrier);
      if (lead) prof_mark(ms, t_last, level, MG_PROF_RESIDUAL);

      n_equiv_iter = (int)((double)n_equiv_iter
                           + R.n_iter * (L.n_g_rows * denom_n_rows_0));
      if (cvg != CSB_SLES_ITERATING) { end_cycle = true; break; }

      /* restriction: coarse row = sum of its fine rows, ascending fine order
         (== the sequential loop of cs_grid.cpp:9025-9030) */
      {
        const int *__restrict__ cfi = L.c_f_row_index;
        const int *__restrict__ cfr = L.c_f_row_ids;
        const double *wr = S.wr;
        double *c_rhs = C.rhs;
        row_loop<false>(C.n_rows, [&](int ic) {
          double s = 0.;
          const int rs = cfi[ic], re = cfi[ic+1];
          for (int q = rs; q < re; q++) s += wr[cfr[q]];
          c_rhs[ic] = s;
        });
      }
      grid_barrier(P.barrier);
      if (lead) prof_mark(ms, t_last, level, MG_PROF_RESTRICT);
    }

    /* ---------------- coarsest level (:3367-3429) ---------------- */
    if (!end_cycle) {
      const MgLevelDesc &L = LV[coarsest];
      const double *rhs_lv = (coarsest == 0) ? S.rhs : L.rhs;
      double *vx_lv = (coarsest == 0) ? S.vx : L.vx;
      const bool single = (L.n_rows <= S.single_rows);
      PcgResult R;
      if (single) {
        if (blockIdx.x == 0)
          R = pcg_run<true>(L, rhs_lv, vx_lv, false, S.n_max_iter[2],
                            S.precision*S.precision_mult[2], S.r_norm, true, P,
                            L.st_descent, s_prod, s_red);
        grid_barrier(P.barrier);
        R.n_iter = *((volatile int *)&L.st_descent->n_iter);
        R.cvg = *((volatile int *)&L.st_descent->cvg);
      }
      else {
        R = pcg_run<false>(L, rhs_lv, vx_lv, false, S.n_max_iter[2],
                           S.precision*S.precision_mult[2], S.r_norm, true, P,
                           L.st_descent, s_prod, s_red);
        grid_barrier(P.barrier);
      }
      if (lead) prof_mark(ms, t_last, coarsest, MG_PROF_COARSEST);
      n_equiv_iter = (int)((double)n_equiv_iter
                           + R.n_iter * (L.n_g_rows * denom_n_rows_0));
      if (R.cvg < CSB_SLES_BREAKDOWN) { end_cycle = true; cvg = CSB_SLES_DIVERGED; }
      if (coarsest == 0) {   /* degenerate one-level hierarchy: plain PCG solve */
        residual = *((volatile double *)&L.st_descent->residual);
        cvg = (R.cvg == CSB_SLES_ITERATING) ? CSB_SLES_MAX_ITERATION : R.cvg;
      }
    }

    /* ---------------- ascent (:3431-3524) ---------------- */
    if (!end_cycle) {
      for (int level = coarsest - 1; level > -1; level--) {
        const MgLevelDesc &L = LV[level];
        const MgLevelDesc &C = LV[level+1];
        double *vx_lv = (level == 0) ? S.vx : L.vx;
        {
          const int *__restrict__ coarse_row = L.coarse_row;
          const double *c_vx = C.vx;
          row_loop<false>(L.n_rows, [&](int i) {       /* cs_grid.cpp:9156-9166 */
            const int ic = coarse_row[i];
            if (ic >= 0) vx_lv[i] += c_vx[ic];
          });
        }
        grid_barrier(P.barrier);
        if (lead) prof_mark(ms, t_last, level, MG_PROF_PROLONG);
        if (level > 0) {
          const bool single = (L.n_rows <= S.single_rows);
          PcgResult R;
          if (single) {
            if (blockIdx.x == 0)
              R = pcg_run<true>(L, L.rhs, vx_lv, true, S.n_max_iter[1],
                                S.precision*S.precision_mult[1], S.r_norm, true, P,
                                L.st_ascent, s_prod, s_red);
            grid_barrier(P.barrier);
            R.n_iter = *((volatile int *)&L.st_ascent->n_iter);
            R.cvg = *((volatile int *)&L.st_ascent->cvg);
          }
          else {
            R = pcg_run<false>(L, L.rhs, vx_lv, true, S.n_max_iter[1],
                               S.precision*S.precision_mult[1], S.r_norm, true, P,
                               L.st_ascent, s_prod, s_red);
            grid_barrier(P.barrier);
          }
          if (lead) prof_mark(ms, t_last, level, MG_PROF_ASCENT);
          n_equiv_iter = (int)((double)n_equiv_iter
                               + R.n_iter * (L.n_g_rows * denom_n_rows_0));
          if (R.cvg < CSB_SLES_BREAKDOWN) { cvg = CSB_SLES_DIVERGED; break; }
        }
      }
    }
  }

  if (lead) {
    ms->precision = S.precision; ms->r_norm = S.r_norm;
    ms->residual = residual; ms->initial_residual = initial_residual;
    ms->cvg = cvg; ms->cycle_id = n_cycles; ms->n_max_cycles = S.n_max_cycles;
    ms->n_equiv_iter = n_equiv_iter;
  }
}


} // anonymous namespace

void plan_slabs(const std::vector<csb_lnum_t> &ri, int n_rows, int &slab_rows, int &n_slabs,
                size_t &max_entries)
{
  slab_rows = PT;
  while (true) {
    max_entries = 0;
    for (int r0 = 0; r0 < n_rows; r0 += slab_rows) {
      const int r1 = (r0 + slab_rows < n_rows) ? r0 + slab_rows : n_rows;
      const size_t nk = (size_t)(ri[r1] - ri[r0]);
      if (nk > max_entries) max_entries = nk;
    }
    if (max_entries * sizeof(double) <= 40*1024 || slab_rows <= 32) break;
    slab_rows /= 2;
  }
  n_slabs = (n_rows + slab_rows - 1) / slab_rows;
}

size_t persistent_smem_bytes(size_t max_slab_entries)
{
  return (PERSISTENT_S_RED + (max_slab_entries > 2 ? max_slab_entries : 2)) * sizeof(double);
}

/* number of co-resident CTAs for the given dynamic shared memory size */
int persistent_grid(int which, size_t smem)
{
  Context &c = ctx();
  int occ = 0;
  if (which == 0) {
    if (smem > 48*1024)
      cudaFuncSetAttribute(k_pcg_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pcg_persistent, PT, smem);
  }
  else {
    if (smem > 48*1024)
      cudaFuncSetAttribute(k_mg_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mg_persistent, PT, smem);
  }
  if (occ < 1) {
    set_error("persistent kernel does not fit on an SM with %zu bytes of shared memory", smem);
    return 0;
  }
  static int cap = -1;
  if (cap < 0) {
    const char *s = getenv("CSB_CTAS_PER_SM");
    cap = (s != nullptr) ? atoi(s) : 4;
  }
  if (occ > cap && cap > 0) occ = cap;
  int g = occ * c.sm_count / (c.device_share > 1 ? c.device_share : 1);
  return (g > 0) ? g : 1;
}

bool launch_pcg_persistent(const MgLevelDesc &L, const double *rhs, double *vx,
                           bool vx_ini_nonzero, int n_max_iter, double precision,
                           double r_norm, const PersistentCommon &P, ItState *st,
                           int grid, size_t smem, bool single, bool cooperative)
{
  int ini = vx_ini_nonzero ? 1 : 0;
  int sgl = single ? 1 : 0;
  void *args[] = {(void *)&L, (void *)&rhs, (void *)&vx, (void *)&ini, (void *)&n_max_iter,
                  (void *)&precision, (void *)&r_norm, (void *)&P, (void *)&st, (void *)&sgl};
  /* ranks sharing one device (in-process test groups) cannot both launch
     cooperatively; their grids are sized to be co-resident together */
  cudaError_t e = cooperative
    ? cudaLaunchCooperativeKernel((void *)k_pcg_persistent, dim3(grid), dim3(PT), args, smem,
                                  ctx().stream)
    : cudaLaunchKernel((void *)k_pcg_persistent, dim3(grid), dim3(PT), args, smem,
                       ctx().stream);
  ctx().launches += 1;
  if (e != cudaSuccess) {
    set_error("cooperative launch of k_pcg_persistent failed: %s", cudaGetErrorString(e));
    return false;
  }
  return true;
}

bool launch_mg_persistent(const MgSolveParams &S, const PersistentCommon &P, int grid,
                          size_t smem, bool cooperative)
{
  void *args[] = {(void *)&S, (void *)&P};
  cudaError_t e = cooperative
    ? cudaLaunchCooperativeKernel((void *)k_mg_persistent, dim3(grid), dim3(PT), args, smem,
                                  ctx().stream)
    : cudaLaunchKernel((void *)k_mg_persistent, dim3(grid), dim3(PT), args, smem,
                       ctx().stream);
  ctx().launches += 1;
  if (e != cudaSuccess) {
    set_error("cooperative launch of k_mg_persistent failed: %s", cudaGetErrorString(e));
    return false;
  }
  return true;
}

} // namespace csb
