/*============================================================================
 * sles.cu -- host drivers of the device-resident solvers:
 *   cs_sles_it  (PCG+Jacobi single-reduction, Jacobi, BiCGStab)
 *   cs_multigrid V-cycle over the aggregation hierarchy
 *
 * The host only enqueues kernels.  Krylov scalars, convergence tests and the
 * multigrid cycle gate live in device memory (ItState / MgState), so there is
 * no host<->device round trip per iteration (the reference's device path does
 * a stream sync + D2H of 1-5 doubles per dot product,
 * src/alge/cs_sles_it_cuda.cu:1049-1095).  Once a solver converges the
 * remaining enqueued kernels of that solve return immediately.
 *============================================================================*/

#include "csb_internal.h"
#include "grid_host.h"
#include "persistent.h"

#include <chrono>
#include <cmath>
#include <cstddef>
#include <algorithm>
#include <omp.h>

using namespace csb;

namespace {

double wall()
{
  return std::chrono::duration<double>(
    std::chrono::steady_clock::now().time_since_epoch()).count();
}

int env_int(const char *name, int dflt)
{
  const char *s = getenv(name);
  return (s != nullptr) ? atoi(s) : dflt;
}

/* CSB_SETUP_TRACE=1: wall-clock marks of the setup stages on stderr (tuning aid) */
struct SetupTrace {
  bool on; double t_last;
  SetupTrace() : on(env_int("CSB_SETUP_TRACE", 0) != 0), t_last(wall()) {}
  void mark(const char *what, int level = -1) {
    if (!on) return;
    const double t = wall();
    fprintf(stderr, "[csb200 setup] %-28s level %2d  %8.2f ms\n", what, level, (t - t_last)*1e3);
    t_last = t;
  }
};

/* stage a host vector on the device when needed */
struct StagedVec {
  double         *d = nullptr;
  double         *h = nullptr;     /* caller's buffer when it is not used in place */
  size_t          n = 0;
  /* A device buffer is used in place when the bulk copies of the persistent kernels
     can read it: 16-byte aligned, and an even length (they read whole 16-byte units).
     Otherwise the vector is staged in `keep`, a buffer that lives with the solver object
     (no cudaMalloc / cudaFree per solve). */
  bool init(double *p, size_t count, bool copy_in, DevBuf<double> &keep) {
    n = count;
    if (is_device_ptr(p) && ((uintptr_t)p & 15) == 0 && (count & 1) == 0) { d = p; return true; }
    h = p;
    if (keep.n < count && !keep.alloc(count)) return false;
    d = keep.p;
    if (copy_in && count > 0)
      return CSB_CUDA(cudaMemcpyAsync(d, p, sizeof(double)*count, cudaMemcpyDefault,
                                      ctx().stream));
    return true;
  }
  bool copy_back(size_t count) {
    if (h == nullptr) return true;
    return CSB_CUDA(cudaMemcpyAsync(h, d, sizeof(double)*count, cudaMemcpyDefault,
                                    ctx().stream));
  }
};

/* distance between the work vectors of the persistent kernels: even, so that every
   vector starts on a 16-byte boundary */
inline size_t work_stride_of(size_t n_cols_ext) { return (n_cols_ext + 1) & ~(size_t)1; }

const size_t IT_HEAD = offsetof(ItState, hist);
const size_t MG_HEAD = offsetof(MgState, cycle_res);

/* col_id on the host (row_index always is); defined with the multigrid below */
bool matrix_host_mirror_structure(csb_matrix_t *m);

/* Ghost-exchange descriptor of one matrix/level for the persistent kernels:
   ghost buffers and flag words in the arena, the neighbours' matching
   addresses obtained through the mailbox.  Collective over all ranks. */
struct DistLevelHost {
  DistLevel           h;
  DevBuf<unsigned>    words;      /* count, ticket */
  DevBuf<unsigned>    bmask;      /* rows with ghost columns */
  DevBuf<double>      bsum;       /* their parked sums */
  DevBuf<DistLevel>   d;          /* device copy */
};

/* rows of m that have ghost columns (one bit each); false if a row does not keep its ghost
   columns behind its local ones (the exchange / product overlap sums the local head of such
   rows first) */
bool ghost_row_mask(const csb_matrix_t *m, std::vector<unsigned> &mask)
{
  mask.assign(((size_t)m->n_rows + 31) / 32 + 1, 0u);
  if (!m->host_structure_ready() || m->h_row_index.size() < (size_t)m->n_rows + 1) {
    set_error("ghost_row_mask: matrix structure is not mirrored on the host");
    return false;
  }
  const int *ri = m->h_row_index.data(), *ci = m->host_col_id();
  bool sorted = true;
  #pragma omp parallel for schedule(static, 4096) reduction(&&: sorted)
  for (int i = 0; i < m->n_rows; i++) {
    bool seen_ghost = false;
    for (int k = ri[i]; k < ri[i+1]; k++) {
      if (ci[k] >= m->n_rows) seen_ghost = true;
      else if (seen_ghost) sorted = false;
    }
    if (seen_ghost) {
      #pragma omp atomic
      mask[(size_t)i >> 5] |= 1u << (i & 31);
    }
  }
  if (!sorted)
    set_error("matrix rows must list their ghost columns (ids >= n_rows) after the local "
              "ones, as cs_matrix_t does (sorted column ids)");
  return sorted;
}

bool build_dist_level(csb_pcomm_t *pc, const csb_halo_t *halo, const csb_matrix_t *m,
                      DistLevelHost &out)
{
  Context &c = ctx();
  DistLevel &D = out.h;
  memset(&D, 0, sizeof(D));
  const int nd = (halo != nullptr) ? halo->n_c_domains : 0;
  if (nd > COMM_MAX_DOM) {
    set_error("halo with %d neighbour ranks: at most %d supported", nd, COMM_MAX_DOM);
    return false;
  }
  if (halo != nullptr && halo->pcomm != pc) {
    set_error("matrix halo is not attached to the solver's communicator (csb_halo_set_pcomm)");
    return false;
  }
  D.n_dom = nd;
  D.n_send = (halo != nullptr) ? halo->n_send : 0;
  D.n_ghost = (halo != nullptr) ? halo->n_ghost : 0;
  D.send_list = (halo != nullptr) ? halo->send_list.p : nullptr;
  for (int i = 0; i <= nd && halo != nullptr; i++) D.send_index[i] = halo->send_index[i];
  const size_t gb = sizeof(double) * (size_t)(D.n_ghost > 0 ? D.n_ghost : 1);
  D.gbuf[0] = (double *)pcomm_arena_alloc(pc, gb);
  D.gbuf[1] = (double *)pcomm_arena_alloc(pc, gb);
  D.flag_local = (unsigned *)pcomm_arena_alloc(pc, sizeof(unsigned)*COMM_MAX_DOM);
  if (D.gbuf[0] == nullptr || D.gbuf[1] == nullptr || D.flag_local == nullptr) return false;
  if (!out.words.alloc(2)) return false;
  if (!CSB_CUDA(cudaMemsetAsync(out.words.p, 0, 2*sizeof(unsigned), c.stream))) return false;
  D.count = out.words.p;
  D.ticket = out.words.p + 1;
  D.timeout_ns = pc->h.timeout_ns;
  {
    std::vector<unsigned> mask;
    if (!ghost_row_mask(m, mask)) return false;
    if (!out.bmask.upload(mask.data(), mask.size())
        || !out.bsum.alloc((size_t)std::max(m->n_rows, 1))
        || !CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;    /* mask dies here */
    D.bmask = out.bmask.p;
    D.bsum = out.bsum.p;
    /* Two forms of the SpMV phase on a partitioned level (persistent.cu, spmv_phase_g):
       the exchange is waited for first and ghost columns are read in line (default), or the
       rows with ghost columns are parked while the peer stores are in flight and finished
       in a fix-up pass after the wait (CSB_HALO_OVERLAP=1).  Measured on B200 pairs over
       NVLink (profiles/r02/scale_v2_overlap_ab.md): the neighbours' stores land within a few
       microseconds, about what the fix-up pass costs, so the overlapped form does not pay
       on this fabric; both are kept and tested. */
    D.overlap = (env_int("CSB_HALO_OVERLAP", 0) != 0) ? 1 : 0;
  }

  std::vector<long long> snd((size_t)nd*8 + 8, 0), rcv((size_t)nd*8 + 8, 0);
  for (int i = 0; i < nd; i++) {
    snd[(size_t)i*8 + 0] = ((char *)D.gbuf[0] - pc->arena) + 8ll*halo->index[i];
    snd[(size_t)i*8 + 1] = ((char *)D.gbuf[1] - pc->arena) + 8ll*halo->index[i];
    snd[(size_t)i*8 + 2] = ((char *)D.flag_local - pc->arena) + 4ll*i;
  }
  if (!pcomm_mailbox(pc, nd, nd > 0 ? halo->c_domain_rank.data() : nullptr, snd.data(),
                     rcv.data())) return false;
  for (int i = 0; i < nd; i++) {
    char *base = pc->h.peer[halo->c_domain_rank[i]];
    D.push_dst[i][0] = (double *)(base + rcv[(size_t)i*8 + 0]);
    D.push_dst[i][1] = (double *)(base + rcv[(size_t)i*8 + 1]);
    D.flag_remote[i] = (unsigned *)(base + rcv[(size_t)i*8 + 2]);
  }
  if (!out.d.upload(&D, 1)) return false;
  return CSB_CUDA(cudaStreamSynchronize(c.stream));
}

} // anonymous namespace

/*----------------------------------------------------------------------------
 * cs_sles_it
 *----------------------------------------------------------------------------*/

struct _csb_sles_it_t {
  csb_sles_it_type_t  type = CSB_SLES_PCG;
  int                 poly_degree = 0;
  int                 n_max_iter = 10000;
  bool                update_stats = false;
  void               *nccl_comm = nullptr;
  csb_pcomm_t        *pcomm = nullptr;
  DistLevelHost       dist;
  size_t              arena_mark = 0;
  bool                arena_used = false;
  /* preconditioner handed over by csb_sles_it_transfer_pc (owned), FCG only:
     cs_sles_it_t::_pc, src/alge/cs_sles_it_priv.h */
  csb_multigrid_t    *pc_mg = nullptr;
  bool                is_smoother = false;   /* created by csb_multigrid_smoother_create */
  bool                fcg_mg_persistent = false;  /* FCG + V-cycle in one launch (setup) */
  DevBuf<double>      pc_tmp;                /* polynomial preconditioner work vector */
  /* true when the solve goes through a general preconditioner (kernel-per-phase FCG /
     single-reduction PCG): polynomial (poly_degree >= 1), transferred multigrid, or none
     (FCG with poly_degree < 0) -- cs_sles_it_create, cs_sles_it.cpp:4676-4690 */
  bool general_pc() const {
    if (is_smoother) return false;
    if (type == CSB_SLES_FCG || type == CSB_SLES_IPCG) return (poly_degree != 0);
    if (type == CSB_SLES_PCG) return (poly_degree > 0 || (poly_degree < 0 && pc_mg != nullptr));
    return false;
  }
  /* setup data (cs_sles_it_setup_t, src/alge/cs_sles_it_priv.h) */
  const csb_matrix_t *a = nullptr;
  bool                setup_done = false;
  DevBuf<double>      ad_inv;
  DevBuf<double>      work;
  DevBuf<ItState>     st;
  ItState            *h_st = nullptr;     /* pinned mirror */
  /* GCR (restart_interval, cs_sles_it.cpp:4699) and the fallback solver
     (fallback_cvg / fallback_n_max_iter / fallback, cs_sles_it.cpp:4736-4749) */
  int                 restart_interval = 20;
  int                 fallback_cvg = CSB_SLES_DIVERGED;
  int                 fallback_n_max_iter = 0;
  _csb_sles_it_t     *fallback = nullptr;
  bool                fallback_last = false;
  const _csb_sles_it_t *shared = nullptr;  /* cs_sles_it_set_shareable: setup data of another context */
  /* persistent-kernel launch data */
  MgLevelDesc         desc;
  DevBuf<double>      partials;
  DevBuf<double>      stage_rhs, stage_vx;   /* host rhs / vx staged here, kept across solves */
  DevBuf<unsigned>    barrier;
  int                 p_grid = 0;
  size_t              p_smem = 0;
  /* stats (cs_sles_it_priv.h:177-190) */
  unsigned n_setups = 0, n_solves = 0;
  unsigned n_iterations_last = 0, n_iterations_min = 0, n_iterations_max = 0;
  unsigned long long n_iterations_tot = 0;
  double t_setup = 0, t_solve = 0;
  ~_csb_sles_it_t();
};

/* Dependency levels of the Gauss-Seidel sweeps of one matrix: rows sorted by level, a level's
   rows in ascending order.  level(i) = 1 + max level(j) over the neighbours j < i of row i in
   the SYMMETRISED structure (j in row i or i in row j), so that a row runs after every
   lower-numbered row it reads (the values the sequential sweep has already updated) and
   before every higher-numbered row that reads it or that it reads (the values the sequential
   sweep has not touched yet).  Forward sweeps take the levels in ascending order, backward
   sweeps in descending order: exactly the reference's one-thread result
   (cs_multigrid_smoother.cpp:1053-1630, cs_sles_it.cpp:3940-4360). */
struct GsPlan {
  csb::DevBuf<int> rows, level_start_d;
  std::vector<int> level_start;
};

namespace csb { void gs_plan_free(GsPlan *p) { delete p; } }

namespace {

/* restriction through the bulk-copy ring of the persistent kernel (CSB_RESTRICT_RING=0: the
   register-batched gather form) */
bool restrict_ring_enabled()
{
  static int v = env_int("CSB_RESTRICT_RING", 1);
  return v != 0;
}

int sb_rows_threshold()
{
  static int v = env_int("CSB_SB_ROWS", 512);   /* measured: 796-row level faster on the grid, 263-row level in one CTA */
  return v;
}

bool is_smoother_type(int t)
{
  return (t == CSB_SLES_JACOBI || t == CSB_SLES_L1_JACOBI || t == CSB_SLES_R_JACOBI
          || t == CSB_SLES_RJ2 || t == CSB_SLES_RJ3 || t == CSB_SLES_P_GAUSS_SEIDEL
          || t == CSB_SLES_P_SYM_GAUSS_SEIDEL || t == CSB_SLES_TS_F_GAUSS_SEIDEL
          || t == CSB_SLES_TS_B_GAUSS_SEIDEL);
}

bool is_gs_type(int t)
{
  return (t == CSB_SLES_P_GAUSS_SEIDEL || t == CSB_SLES_P_SYM_GAUSS_SEIDEL
          || t == CSB_SLES_TS_F_GAUSS_SEIDEL || t == CSB_SLES_TS_B_GAUSS_SEIDEL);
}

/* Sweep schedule of the Jacobi-family smoothers: relaxation weights
   (cs_multigrid_smoother.cpp:259-281) and the sweep count rounded up to a multiple of the
   schedule length (cs_multigrid_smoother_setup, :1888-1898). */
struct SmootherPlan {
  int    mode = SMOOTH_JACOBI;
  int    wk_m = 1;
  double w[3] = {1., 1., 1.};
  int    n_iter = 0;
};

SmootherPlan smoother_plan(int type, int n_max_iter)
{
  SmootherPlan p;
  p.n_iter = n_max_iter;
  if (type == CSB_SLES_L1_JACOBI) p.mode = SMOOTH_L1;
  else if (type != CSB_SLES_JACOBI) {
    p.mode = SMOOTH_RELAXED;
    p.w[0] = 2./3.;
    if (type == CSB_SLES_RJ2) { p.w[0] = 1.7319; p.w[1] = 0.5695; p.wk_m = 2; }
    if (type == CSB_SLES_RJ3) { p.w[0] = 2.2473; p.w[1] = 0.8573; p.w[2] = 0.5296; p.wk_m = 3; }
    if (p.n_iter % p.wk_m) p.n_iter += p.wk_m - (p.n_iter % p.wk_m);
  }
  return p;
}

const GsPlan *gs_plan_get(const csb_matrix_t *m)
{
  if (m->gs_plan != nullptr) return m->gs_plan;
  if (m->db != 1) { set_error("Gauss-Seidel: scalar (db = 1) matrices only"); return nullptr; }
  if (!matrix_host_mirror_structure(const_cast<csb_matrix_t *>(m))) return nullptr;
  const int n = m->n_rows;
  const int *ri = m->h_row_index.data();
  const int *ci = m->host_col_id();
  std::vector<int> lvl((size_t)std::max(n, 1), 0);
  int n_lv = 0;
  for (int i = 0; i < n; i++) {
    int l = lvl[i];                                   /* pushed by rows j < i that hold i */
    for (int k = ri[i]; k < ri[i+1]; k++) {
      const int j = ci[k];
      if (j < i && lvl[j] + 1 > l) l = lvl[j] + 1;
    }
    lvl[i] = l;
    for (int k = ri[i]; k < ri[i+1]; k++) {
      const int j = ci[k];
      if (j > i && j < n && lvl[j] < l + 1) lvl[j] = l + 1;
    }
    if (l + 1 > n_lv) n_lv = l + 1;
  }
  GsPlan *p = new GsPlan;
  p->level_start.assign((size_t)n_lv + 1, 0);
  for (int i = 0; i < n; i++) p->level_start[(size_t)lvl[i] + 1]++;
  for (int l = 0; l < n_lv; l++) p->level_start[(size_t)l + 1] += p->level_start[l];
  std::vector<int> pos(p->level_start.begin(), p->level_start.end() - 1), rows((size_t)std::max(n, 1));
  for (int i = 0; i < n; i++) rows[(size_t)pos[lvl[i]]++] = i;
  if (!p->rows.upload(rows.data(), (size_t)n)
      || !p->level_start_d.upload(p->level_start.data(), p->level_start.size())
      || !CSB_CUDA(cudaStreamSynchronize(ctx().stream))) { delete p; return nullptr; }
  m->gs_plan = p;
  return p;
}

/* One sweep over all rows in dependency-level order.  Levels of more than 1024 rows take a
   launch each; runs of smaller ones share one single-CTA launch. */
bool gs_sweep(const csb_matrix_t *m, const GsPlan *p, bool backward, int mode,
              const double *ad_inv, const double *rhs, double *vx, const MgState *gate,
              const ItState *st, double *res)
{
  const int nl = (int)p->level_start.size() - 1;
  const int chain_max = 1024;
  int k = 0;
  while (k < nl) {
    const int l = backward ? nl - 1 - k : k;
    const int cnt = p->level_start[l+1] - p->level_start[l];
    if (cnt > chain_max) {
      if (!launch_gs_rows(cnt, p->rows.p + p->level_start[l], m, ad_inv, rhs, vx, mode, gate,
                          st, res)) return false;
      k++;
      continue;
    }
    int k1 = k + 1;
    while (k1 < nl) {
      const int l1 = backward ? nl - 1 - k1 : k1;
      if (p->level_start[l1+1] - p->level_start[l1] > chain_max) break;
      k1++;
    }
    const int la = backward ? nl - k1 : k, lb = backward ? nl - k : k1;   /* [la, lb) */
    if (!launch_gs_chain(p->rows.p, p->level_start_d.p, la, lb, backward, m, ad_inv, rhs, vx,
                         mode, gate, st, res)) return false;
    k = k1;
  }
  return true;
}

/* Gauss-Seidel smoothers (fixed sweeps, no convergence test): _p_gauss_seidel_msr,
   _p_sym_gauss_seidel_msr, _ts_f_gauss_seidel_msr, _ts_b_gauss_seidel_msr
   (cs_multigrid_smoother.cpp:1053, :1179, :1389, :1519). */
bool gs_smoother_enqueue(const csb_matrix_t *m, int type, int n_max_iter, const double *ad_inv,
                         const double *rhs, double *vx, bool vx_ini_nonzero,
                         const MgState *gate)
{
  const GsPlan *p = gs_plan_get(m);
  if (p == nullptr) return false;
  const int n = m->n_rows, n_ext = m->n_cols_ext;
  if (type == CSB_SLES_TS_F_GAUSS_SEIDEL || type == CSB_SLES_TS_B_GAUSS_SEIDEL) {
    /* one truncated sweep; ghost values (all values without vx_ini) are zero */
    if (!vx_ini_nonzero) { if (!launch_fill_gated(n_ext, 0.0, vx, gate)) return false; }
    else if (n_ext > n && !launch_fill_gated(n_ext - n, 0.0, vx + n, gate)) return false;
    const bool fwd = (type == CSB_SLES_TS_F_GAUSS_SEIDEL);
    return gs_sweep(m, p, !fwd, fwd ? 1 : 2, ad_inv, rhs, vx, gate, nullptr, nullptr);
  }
  for (int it = 0; it < n_max_iter; it++) {
    if (it == 0 && !vx_ini_nonzero) { if (!launch_fill_gated(n_ext, 0.0, vx, gate)) return false; }
    else if (m->halo != nullptr && csb_halo_sync_var(m->halo, vx) != 0) return false;
    if (!gs_sweep(m, p, false, 0, ad_inv, rhs, vx, gate, nullptr, nullptr)) return false;
    if (type == CSB_SLES_P_SYM_GAUSS_SEIDEL) {
      if (m->halo != nullptr && csb_halo_sync_var(m->halo, vx) != 0) return false;
      if (!gs_sweep(m, p, true, 0, ad_inv, rhs, vx, gate, nullptr, nullptr)) return false;
    }
  }
  return true;
}

/* Gauss-Seidel solvers with the convergence test of cs_sles_it (_p_gauss_seidel_msr
   cs_sles_it.cpp:3940-4110, _p_sym_gauss_seidel_msr :4124-4360): the residual of an
   iteration is sqrt(sum (d_i (vx_i_new - vx_i_old))^2) over the forward sweep (GS) or the
   backward sweep (symmetric GS).  res: n_rows doubles. */
bool gs_enqueue(const csb_matrix_t *m, int type, ItState *st, const double *ad_inv,
                double precision, double r_norm, int n_max_iter, const double *rhs,
                double *vx, bool vx_ini_nonzero, double *res, ItState *h_st)
{
  const GsPlan *p = gs_plan_get(m);
  if (p == nullptr) return false;
  Context &c = ctx();
  const int n = m->n_rows, n_ext = m->n_cols_ext;
  const bool sym = (type == CSB_SLES_P_SYM_GAUSS_SEIDEL);
  if (!launch_it_state_init(st, nullptr, precision, r_norm, n_max_iter)) return false;
  const int poll = 4;
  for (int it = 1; it <= n_max_iter; it++) {
    if (it == 1 && !vx_ini_nonzero) { if (!launch_fill(n_ext, 0.0, vx)) return false; }
    else if (m->halo != nullptr && csb_halo_sync_var(m->halo, vx) != 0) return false;
    if (!gs_sweep(m, p, false, 0, ad_inv, rhs, vx, nullptr, st, sym ? nullptr : res))
      return false;
    if (sym) {
      if (m->halo != nullptr && csb_halo_sync_var(m->halo, vx) != 0) return false;
      if (!gs_sweep(m, p, true, 0, ad_inv, rhs, vx, nullptr, st, res)) return false;
    }
    if (!launch_gs_residual(n, st, res)) return false;
    if (h_st != nullptr && (it % poll) == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, c.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
      if (h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return true;
}

/* cs_multigrid_smoother on a block-diagonal matrix (db > 1): _block_3_jacobi / _block_jacobi
   (cs_multigrid_smoother.cpp:390-500, :507-620: the solver forms' sweep without the residual)
   and _block_relaxed_jacobi (:626-790) for R_JACOBI / RJ2 / RJ3.  inv: the block_nn_inverse
   of the setup.  work: two vectors of ws doubles (vx_tmp, vxx). */
bool block_smoother_enqueue(const csb_matrix_t *m, int type, int n_max_iter, const double *inv,
                            const double *rhs, double *vx, bool vx_ini_nonzero, double *work,
                            size_t ws, ItState *st)
{
  const int db = m->db, nb = m->n_rows;
  double *vx_tmp = work, *vxx = work + ws;
  auto e_product = [&](const double *x) {
    return (db == 3) ? launch_spmv_b3(m, x, vxx, true)
                     : launch_spmv_bn(-1, m, nullptr, x, vxx, nullptr, true);
  };
  if (type == CSB_SLES_JACOBI) {
    for (int it = 0; it < n_max_iter; it++) {
      /* the sweep kernels of the solver form, their convergence test made void (a fresh
         state per sweep: iteration 1 of 2 never stops and never diverges) */
      if (!launch_it_state_init(st, nullptr, 0., 1., 2)) return false;
      const bool first = (it == 0 && !vx_ini_nonzero);
      if (!first && !e_product(vx)) return false;
      const bool ok = (db == 3)
        ? launch_b3_jacobi_update(nb, st, m->d_val.p, inv, rhs, vxx, vx, first)
        : launch_bn_jacobi_update(nb, db, st, m->d_val.p, inv, rhs, vxx, vx, first);
      if (!ok) return false;
    }
    return true;
  }
  const SmootherPlan p = smoother_plan(type, n_max_iter);
  double *vx_n = vx, *vx_np = vx_tmp;
  int iter_ini = 0;
  if (!vx_ini_nonzero) {
    if (!launch_bn_relaxed_update(nb, db, p.w[0], inv, rhs, nullptr, nullptr, vx_n, true))
      return false;
    iter_ini = 1;
  }
  for (int it = iter_ini; it < p.n_iter; it++) {
    const double wk = p.w[it % p.wk_m];
    double *t = vx_n; vx_n = vx_np; vx_np = t;
    if (!e_product(vx_np)) return false;
    if (!launch_bn_relaxed_update(nb, db, wk, inv, rhs, vxx, vx_np, vx_n, false)) return false;
  }
  if (vx_n != vx && !launch_copy(nb*db, vx_n, vx, nullptr)) return false;
  return true;
}

/* Fixed number of Jacobi-family sweeps without convergence test (_jacobi, _relaxed_jacobi,
   _l1_jacobi: cs_multigrid_smoother.cpp:143-226, :244-373, :804-905).  Each sweep is ONE
   kernel: the (diagonal-excluded, or full for L1) product in CSR order with the update in
   its epilogue, ping-ponging between vx and tmp like the reference's vx_n / vx_np.
   inv: 1/d, or the L1 inverse for CSB_SLES_L1_JACOBI.  tmp: n_cols_ext doubles. */
bool smoother_enqueue(const csb_matrix_t *m, int type, int n_max_iter, const double *inv,
                      const double *rhs, double *vx, bool vx_ini_nonzero, double *tmp,
                      const MgState *gate)
{
  if (is_gs_type(type))
    return gs_smoother_enqueue(m, type, n_max_iter, inv, rhs, vx, vx_ini_nonzero, gate);
  const SmootherPlan p = smoother_plan(type, n_max_iter);
  const int n = m->n_rows;
  double *vx_n = vx, *vx_np = tmp;
  int iter_ini = 0;
  if (!vx_ini_nonzero) {
    if (!launch_smoother_first(n, p.mode, p.w[0], rhs, inv, vx_n, gate)) return false;
    iter_ini = 1;
  }
  for (int it = iter_ini; it < p.n_iter; it++) {
    const double wk = p.w[it % p.wk_m];
    double *t = vx_n; vx_n = vx_np; vx_np = t;
    SpmvArgs sa;
    sa.m = m; sa.exclude_diag = (p.mode != SMOOTH_L1);
    sa.x = vx_np; sa.y = vx_n; sa.b = rhs; sa.ad_inv = inv;
    sa.mg = const_cast<MgState *>(gate);
    sa.smooth_mode = p.mode; sa.w0 = wk; sa.w1 = 1.0 - wk;
    if (!launch_spmv(EPI_SMOOTH, sa)) return false;
  }
  if (vx_n != vx)
    if (!launch_copy_gated(n, vx_n, vx, gate)) return false;
  return true;
}

/* Enqueue a PCG+Jacobi solve (cs_sles_it.cpp:1441-1645).
   smoother = true: fixed-sweep use inside the V-cycle; the SpMV + dots whose
   only purpose is the convergence test after the last iteration are not
   issued (they do not change vx).
   Returns false on launch error.  work: 5*n_cols_ext doubles. */
bool pcg_enqueue(const csb_matrix_t *m, ItState *st, const MgState *gate,
                 const double *ad_inv, double precision, double r_norm, int n_max_iter,
                 const double *rhs, double *vx, bool vx_ini_nonzero, bool smoother,
                 double *work, ItState *h_st)
{
  const int n = m->n_rows;
  const size_t ws = (size_t)m->n_cols_ext;
  Context &c = ctx();

  if (n <= sb_rows_threshold() && m->halo == nullptr)
    return launch_pcg_single_block(m, st, gate, precision, r_norm, n_max_iter, ad_inv,
                                   rhs, vx, vx_ini_nonzero ? 1 : 0, work);

  double *rk = work, *dk = work + ws, *gk = work + 2*ws, *zk = work + 3*ws,
         *sk = work + 4*ws;

  if (!launch_it_state_init(st, gate, precision, r_norm, n_max_iter)) return false;

  SpmvArgs a;
  a.m = m; a.st = st; a.ad_inv = ad_inv; a.rk = rk; a.dk = dk; a.b = rhs;
  if (vx_ini_nonzero) {
    a.x = vx; a.y = nullptr;
    if (!launch_spmv(EPI_PCG_INIT, a)) return false;
  }
  else if (!launch_pcg_init_zero(n, st, rhs, ad_inv, vx, rk, dk)) return false;

  a.x = dk; a.y = zk; a.phase = 0;
  if (!launch_spmv(EPI_PCG_DOT3, a)) return false;
  if (!launch_pcg_update_first(n, st, ad_inv, dk, zk, vx, rk, gk)) return false;

  a.x = gk; a.y = sk; a.phase = 1;
  const int poll = 16;
  for (int it = 1; ; it++) {
    /* iteration "it" is complete on the device here */
    if (smoother && it >= n_max_iter) break;
    if (!launch_spmv(EPI_PCG_DOT3, a)) return false;
    if (it >= n_max_iter) break;     /* that was the final convergence test */
    if (!launch_pcg_update(n, st, ad_inv, sk, dk, gk, zk, rk, vx)) return false;
    if (!smoother && h_st != nullptr && (it % poll) == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, c.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
      if (h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return true;
}

/* Jacobi (cs_sles_it.cpp:2098-2270); iterate t lives in vx when t is even,
   in rk when t is odd (ping-pong instead of the reference's rk <- vx copy). */
bool jacobi_enqueue(const csb_matrix_t *m, ItState *st, const double *ad_inv,
                    double precision, double r_norm, int n_max_iter, const double *rhs,
                    double *vx, bool vx_ini_nonzero, double *rk, ItState *h_st);

bool bicgstab_enqueue(const csb_matrix_t *m, ItState *st, const double *ad_inv,
                      double precision, double r_norm, int n_max_iter, const double *rhs,
                      double *vx, bool vx_ini_nonzero, double *work, ItState *h_st);

} // anonymous namespace

namespace csb {
bool launch_jacobi_finish(int n, const ItState *st, const double *rk, double *vx);
}

namespace {

bool jacobi_enqueue(const csb_matrix_t *m, ItState *st, const double *ad_inv,
                    double precision, double r_norm, int n_max_iter, const double *rhs,
                    double *vx, bool vx_ini_nonzero, double *rk, ItState *h_st)
{
  const int n = m->n_rows;
  Context &c = ctx();
  if (!launch_it_state_init(st, nullptr, precision, r_norm, n_max_iter)) return false;
  int t = 0;
  if (!vx_ini_nonzero) {
    if (!launch_jacobi_first(n, st, rhs, ad_inv, vx, rk)) return false;
    t = 1;
  }
  const int poll = 16;
  while (t < n_max_iter) {
    t += 1;
    double *src = ((t - 1) % 2 == 0) ? vx : rk;
    double *dst = (t % 2 == 0) ? vx : rk;
    if (t == 2 && !vx_ini_nonzero) src = rk;   /* iterate 1 is in both */
    if (!launch_jacobi_sweep(m, st, rhs, ad_inv, dst, src)) return false;
    if (h_st != nullptr && (t % poll) == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, c.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
      if (h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return launch_jacobi_finish(n, st, rk, vx);
}

bool bicgstab_enqueue(const csb_matrix_t *m, ItState *st, const double *ad_inv,
                      double precision, double r_norm, int n_max_iter, const double *rhs,
                      double *vx, bool vx_ini_nonzero, double *work, ItState *h_st)
{
  const int n = m->n_rows*m->db;                      /* cs_sles_it_priv.cpp: n_rows*db */
  const size_t ws = work_stride_of((size_t)m->n_cols_ext*m->db);
  Context &c = ctx();
  double *res0 = work, *rk = work + ws, *pk = work + 2*ws, *zk = work + 3*ws,
         *uk = work + 4*ws, *vk = work + 5*ws;
  if (!launch_it_state_init(st, nullptr, precision, r_norm, n_max_iter)) return false;
  if (vx_ini_nonzero) {
    if (m->db == 3) { if (!launch_spmv_b3(m, vx, res0, false)) return false; }
    else if (m->db > 1) { if (!launch_spmv_bn(-1, m, nullptr, vx, res0, nullptr, false)) return false; }
    else {
      SpmvArgs a; a.m = m; a.x = vx; a.y = res0;
      if (!launch_spmv(EPI_STORE, a)) return false;
    }
  }
  if (!launch_bicg_init(n, st, rhs, vx, res0, rk, pk, uk, vx_ini_nonzero)) return false;
  if (!launch_bicg_dots_a(n, st, rk, res0)) return false;
  const int poll = 8;
  for (int it = 1; it <= n_max_iter; it++) {
    if (!launch_bicg_pk_zk(n, st, rk, uk, ad_inv, pk, zk)) return false;
    if (!launch_bicg_spmv_b(m, st, zk, uk, res0)) return false;
    if (!launch_bicg_upd1(n, st, ad_inv, vx, rk, zk, uk)) return false;
    if (!launch_bicg_spmv_c(m, st, zk, vk, rk)) return false;
    if (!launch_bicg_upd2(n, st, vx, rk, zk, vk, res0)) return false;
    if (h_st != nullptr && (it % poll) == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, c.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
      if (h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return true;
}

bool dots_to_host(int n, const double *x, const double *y, const double *z, int mode,
                  double r[3]);

/* y = A.x for any diagonal block size */
bool spmv_any(const csb_matrix_t *m, const double *x, double *y)
{
  if (m->db == 3) return launch_spmv_b3(m, x, y, false);
  if (m->db > 1) return launch_spmv_bn(-1, m, nullptr, x, y, nullptr, false);
  SpmvArgs a; a.m = m; a.x = x; a.y = y;
  return launch_spmv(EPI_STORE, a);
}

/* _convergence_test (cs_sles_it.cpp:243-325) on the host, for the solvers whose scalars are
   read back every iteration */
int it_cvg_test_host(ItState *st, int n_iter, double residual)
{
  st->residual = residual;
  st->n_iter = n_iter;
  if (st->n_hist < HIST_MAX) st->hist[st->n_hist++] = residual;
  if (st->precision > 0 && residual < st->precision * st->r_norm) return CSB_SLES_CONVERGED;
  else if (n_iter >= st->n_max_iter) return CSB_SLES_MAX_ITERATION;
  else if (residual > st->initial_residual * 10000.0 && residual > 100.) return CSB_SLES_DIVERGED;
  if (std::isnan(residual) || std::isinf(residual)) return CSB_SLES_DIVERGED;
  return CSB_SLES_ITERATING;
}

/* Preconditioned GCR with restart (_gcr, cs_sles_it.cpp:3272-3480): the solver the reference
   falls back to on a device when BiCGStab breaks down (:5383-5385), and a solver type of its
   own.  Host-driven: the projection coefficients of the modified Gram-Schmidt loop, the step
   lengths and the residual norm are read back (every one of them decides what the next
   launch does, as in the reference), the small triangular system Gamma is inverted on the
   host exactly as there; every vector operation is the reference's expression.
   ad_inv: Jacobi preconditioner (NULL: none).  work: (1 + 2*restart) vectors of stride ws. */
bool gcr_enqueue(const csb_matrix_t *m, ItState *st, ItState *h, const double *ad_inv,
                 int restart, double precision, double r_norm, int n_max_iter,
                 const double *rhs, double *vx, bool vx_ini_nonzero, double *work)
{
  Context &cx = ctx();
  const int n = m->n_rows*m->db;
  const size_t ws = work_stride_of((size_t)m->n_cols_ext*m->db);
  double *rk = work, *zk = work + ws, *ck = work + ws*(1 + (size_t)restart);
  std::vector<double> alpha((size_t)restart), gkj((size_t)(restart + 1)*restart/2),
                      gkj_inv((size_t)(restart + 1)*restart/2);
  memset(h, 0, sizeof(ItState));
  h->precision = precision; h->r_norm = r_norm; h->n_max_iter = n_max_iter;
  h->residual = -1.; h->initial_residual = -1.;
  int cvg = CSB_SLES_ITERATING;
  unsigned n_restart = 0;
  bool have_vx = vx_ini_nonzero;
  double r[3];

  while (cvg == CSB_SLES_ITERATING) {
    int n_c_iter = 0;
    if (have_vx) {
      if (!spmv_any(m, vx, rk)) return false;                       /* rk = A.x0 - rhs */
      if (!launch_kc_vec(2, n, 1.0, -1.0, rhs, nullptr, rk)) return false;
    }
    else {
      if (!launch_fill(m->n_cols_ext*m->db, 0.0, vx)) return false;
      if (!launch_fill(n, 0.0, rk)) return false;
      if (!launch_kc_vec(2, n, 1.0, -1.0, rhs, nullptr, rk)) return false;
      have_vx = true;
    }
    if (!dots_to_host(n, rk, nullptr, nullptr, 1, r)) return false;
    double residual = sqrt(r[0]);
    if (n_restart == 0) h->initial_residual = residual;

    while (cvg == CSB_SLES_ITERATING && n_c_iter < restart) {
      double *zk_n = zk + (size_t)n_c_iter*ws, *ck_n = ck + (size_t)n_c_iter*ws;
      if (ad_inv != nullptr) { if (!launch_kc_vec(5, n, 0., 0., rk, ad_inv, zk_n)) return false; }
      else if (!launch_copy(n, rk, zk_n, nullptr)) return false;
      if (!spmv_any(m, zk_n, ck_n)) return false;
      for (int jj = 0; jj < n_c_iter; jj++) {
        const double *ck_j = ck + (size_t)jj*ws;
        const int ii_jn = (n_c_iter + 1)*n_c_iter/2 + jj;
        if (!dots_to_host(n, ck_j, ck_n, nullptr, 0, r)) return false;
        gkj[ii_jn] = r[0];
        if (!launch_kc_vec(2, n, 1.0, -gkj[ii_jn], ck_j, nullptr, ck_n)) return false;
      }
      const int iter_shift = (n_c_iter + 1)*n_c_iter/2 + n_c_iter;
      if (!dots_to_host(n, ck_n, nullptr, nullptr, 1, r)) return false;
      gkj[iter_shift] = sqrt(r[0]);
      if (fabs(gkj[iter_shift]) > 0) {
        const double scale = 1. / gkj[iter_shift];
        if (!launch_kc_vec(1, n, scale, 0., nullptr, nullptr, ck_n)) return false;
        if (!dots_to_host(n, ck_n, rk, nullptr, 0, r)) return false;
        alpha[n_c_iter] = r[0];
      }
      else
        alpha[n_c_iter] = 0.;
      if (!launch_kc_vec(2, n, 1.0, -alpha[n_c_iter], ck_n, nullptr, rk)) return false;
      if (!dots_to_host(n, rk, nullptr, nullptr, 1, r)) return false;
      residual = sqrt(r[0]);
      n_c_iter += 1;
      cvg = it_cvg_test_host(h, (int)(n_restart*(unsigned)restart) + n_c_iter, residual);
      if (cvg != CSB_SLES_ITERATING) break;
    }

    /* inversion of Gamma (:3424-3449) */
    if (n_c_iter == 1 && !(fabs(alpha[0]) > 0))
      gkj_inv[0] = 1.0;
    else {
      std::fill(gkj_inv.begin(), gkj_inv.end(), 0.0);
      for (int kk = 0; kk < n_c_iter; kk++) {
        for (int ii = 0; ii < kk; ii++)
          for (int jj = 0; jj < kk; jj++)
            gkj_inv[(kk + 1)*kk/2 + ii]
              += ((ii <= jj) ? gkj_inv[(jj + 1)*jj/2 + ii] : 0.0) * gkj[(kk + 1)*kk/2 + jj];
        for (int jj = 0; jj < kk; jj++)
          gkj_inv[(kk + 1)*kk/2 + jj] /= - gkj[(kk + 1)*kk/2 + kk];
        gkj_inv[(kk + 1)*kk/2 + kk] = 1.0 / gkj[(kk + 1)*kk/2 + kk];
      }
    }
    /* the solution (:3453-3468) */
    for (int kk = 0; kk < n_c_iter; kk++)
      for (int jj = 0; jj <= kk; jj++)
        if (!launch_kc_vec(4, n, alpha[kk], gkj_inv[(kk + 1)*kk/2 + jj], zk + (size_t)jj*ws,
                           nullptr, vx)) return false;
    n_restart += 1;
  }
  h->cvg = cvg;
  /* the common tail of csb_sles_it_solve reads the device state */
  return CSB_CUDA(cudaMemcpyAsync(st, h, sizeof(ItState), cudaMemcpyHostToDevice, cx.stream))
      && CSB_CUDA(cudaStreamSynchronize(cx.stream));
}

/* _multigrid_v_cycle_pc as a sequence of launches (defined with the multigrid below) */
bool mg_cycle_pc_enqueue(csb_multigrid_t *mg, double precision, double r_norm,
                         const double *rhs_d, double *vx_d, bool vx_ini_nonzero);
int  mg_n_levels(const csb_multigrid_t *mg);
bool mg_is_k_cycle(const csb_multigrid_t *mg);
MgState *mg_gate(csb_multigrid_t *mg);
/* FCG + one V-cycle per iteration as one persistent launch (k_fcg_mg_persistent): can this
   hierarchy run there (V-cycle, Jacobi-family cs_multigrid_smoother smoothers, PCG coarse
   solver); shared memory / fill the kernel's parameters (defined with the multigrid below) */
bool mg_fcg_persistent_ok(const csb_multigrid_t *mg);
size_t mg_persistent_smem(const csb_multigrid_t *mg);
void mg_fcg_params(csb_multigrid_t *mg, FcgMgParams &S, PersistentCommon &P);
bool mg_fetch_state(csb_multigrid_t *mg);     /* device-side timers -> host mirror (async) */

/* x_out = M^-1 x_in for the general preconditioners of cs_sles_it:
     multigrid   _multigrid_pc_apply (cs_multigrid.cpp:1500-1600): one V-cycle, identity for
                 a one-level hierarchy (:1518-1541); tolerance = the outer solver's
     polynomial  _sles_pc_poly_apply_poly (cs_sles_pc.cpp:549-610): x = r/d then `degree`
                 times x = (r - E.x)/d, i.e. degree+1 Jacobi sweeps from zero -- the fused
                 sweep kernel of the smoothers
     none        cs_sles_pc_none: copy */
bool pc_apply_enqueue(csb_sles_it_t *c, const csb_matrix_t *a, double precision,
                      double r_norm, const double *x_in, double *x_out)
{
  csb_multigrid_t *mg = c->pc_mg;
  if (mg != nullptr && mg_n_levels(mg) > 1)
    return mg_cycle_pc_enqueue(mg, precision, r_norm, x_in, x_out, false);
  if (mg == nullptr && c->poly_degree > 0)
    return smoother_enqueue(a, CSB_SLES_JACOBI, (c->poly_degree > 2 ? 2 : c->poly_degree) + 1,
                            c->ad_inv.p, x_in, x_out, false, c->pc_tmp.p, nullptr);
  return launch_copy(a->n_rows, x_in, x_out, c->st.p);
}

bool pc_is_mg(const csb_sles_it_t *c) { return (c->pc_mg != nullptr && mg_n_levels(c->pc_mg) > 1); }

/* 3x3 block Jacobi (_block_3_jacobi, cs_sles_it.cpp:2284-2440) on a CS_MATRIX_BLOCK_D
   matrix: vxx = E.vx (k_spmv_b3, diagonal blocks excluded), then one kernel for
   vx = D^-1.(rhs - vxx) and the residual ||D.(x_new - x_old)||, tested every sweep on the
   device.  vxx: n_cols_ext*3 doubles. */
bool block3_jacobi_enqueue(const csb_matrix_t *m, ItState *st, const double *ad_inv,
                           double precision, double r_norm, int n_max_iter, const double *rhs,
                           double *vx, bool vx_ini_nonzero, double *vxx, ItState *h_st)
{
  Context &c = ctx();
  if (!launch_it_state_init(st, nullptr, precision, r_norm, n_max_iter)) return false;
  int t = 0;
  const int db = m->db;
  auto update = [&](bool first) {
    return (db == 3)
      ? launch_b3_jacobi_update(m->n_rows, st, m->d_val.p, ad_inv, rhs, vxx, vx, first)
      : launch_bn_jacobi_update(m->n_rows, db, st, m->d_val.p, ad_inv, rhs, vxx, vx, first);
  };
  if (!vx_ini_nonzero) {
    if (!update(true)) return false;
    t = 1;
  }
  const int poll = 16;
  while (t < n_max_iter) {
    t += 1;
    if (db == 3) { if (!launch_spmv_b3(m, vx, vxx, true)) return false; }
    else if (!launch_spmv_bn(-1, m, nullptr, vx, vxx, nullptr, true)) return false;
    if (!update(false)) return false;
    if (h_st != nullptr && (t % poll) == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, c.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
      if (h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return true;
}

/* FCG with a general preconditioner (_flexible_conjugate_gradient, cs_sles_it.cpp:832-985).
   Per iteration: the preconditioner's launches, one SpMV with the 4 dot products
   and the scalar recurrences in its epilogue, one fused 6-vector update; Krylov scalars and
   the convergence state stay on the device, the host only polls the state every few
   iterations.  Launches issued after convergence are gated off on the device. */
bool fcg_pc_enqueue(csb_sles_it_t *c, const csb_matrix_t *a, double precision, double r_norm,
                    const double *rhs, double *vx, bool vx_ini_nonzero)
{
  const int n = a->n_rows;
  const size_t ws = work_stride_of((size_t)a->n_cols_ext);
  Context &cx = ctx();
  double *rk = c->work.p, *vk = rk + ws, *wk = rk + 2*ws, *dk = rk + 3*ws, *qk = rk + 4*ws;
  ItState *st = c->st.p;
  const bool mgpc = pc_is_mg(c);

  if (!launch_it_state_init(st, nullptr, precision, r_norm, c->n_max_iter)) return false;
  if (vx_ini_nonzero) {
    SpmvArgs sa; sa.m = a; sa.x = vx; sa.y = rk;
    if (!launch_spmv(EPI_STORE, sa)) return false;
  }
  if (!launch_fcg_init(n, st, rhs, vx, rk, qk, vx_ini_nonzero)) return false;
  if (mgpc && !launch_mg_state_init(mg_gate(c->pc_mg), precision, r_norm, 1)) return false;

  /* the K-cycle reads reductions back on the host: never run it past convergence */
  const int poll = (mgpc && mg_is_k_cycle(c->pc_mg)) ? 1 : (mgpc ? 2 : 16);
  for (int it = 0; it <= c->n_max_iter; it++) {
    if (!pc_apply_enqueue(c, a, precision, r_norm, rk, vk)) return false;
    SpmvArgs sa;
    sa.m = a; sa.x = vk; sa.y = wk; sa.rk = rk; sa.qk = qk; sa.st = st;
    if (!launch_spmv(EPI_FCG_DOT4, sa)) return false;
    if (mgpc && !launch_gate_from_it(mg_gate(c->pc_mg), st)) return false;
    if (!launch_fcg_update(n, st, vk, wk, dk, qk, rk, vx)) return false;
    if ((it + 1) % poll == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(c->h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, cx.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(cx.stream))) return false;
      if (c->h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return true;
}

/* Inexact-preconditioning CG with a general preconditioner (_conjugate_gradient_ip,
   cs_sles_it.cpp:1016-1203) -- the solver the reference pairs with the K-cycle.  Per
   iteration: the preconditioner's launches, a 3-dot reduction (test, beta), the direction
   update, one SpMV with the 2 dots of the step length in its epilogue, one fused update. */
bool ipcg_pc_enqueue(csb_sles_it_t *c, const csb_matrix_t *a, double precision, double r_norm,
                     const double *rhs, double *vx, bool vx_ini_nonzero)
{
  const int n = a->n_rows;
  const size_t ws = work_stride_of((size_t)a->n_cols_ext);
  Context &cx = ctx();
  double *rk = c->work.p, *rkm1 = rk + ws, *dk = rk + 2*ws, *gk = rk + 3*ws, *zk = rk + 4*ws;
  ItState *st = c->st.p;
  const bool mgpc = pc_is_mg(c);

  if (!launch_it_state_init(st, nullptr, precision, r_norm, c->n_max_iter)) return false;
  if (vx_ini_nonzero) {
    SpmvArgs s0; s0.m = a; s0.x = vx; s0.y = rk;
    if (!launch_spmv(EPI_STORE, s0)) return false;
  }
  if (!launch_pcg_pc_init(n, rhs, vx, rk, vx_ini_nonzero)) return false;
  if (mgpc && !launch_mg_state_init(mg_gate(c->pc_mg), precision, r_norm, 1)) return false;

  SpmvArgs sa;
  sa.m = a; sa.st = st; sa.rk = rk; sa.x = dk; sa.y = zk; sa.phase = 2;

  if (!pc_apply_enqueue(c, a, precision, r_norm, rk, gk)) return false;
  if (!launch_ipcg_dots(n, st, true, rk, gk, rkm1, dk)) return false;
  if (!launch_spmv(EPI_PCG_DOT3, sa)) return false;
  if (!launch_ipcg_update(n, st, true, dk, zk, vx, rk, rkm1)) return false;
  if (mgpc && !launch_gate_from_it(mg_gate(c->pc_mg), st)) return false;

  const int poll = (mgpc && mg_is_k_cycle(c->pc_mg)) ? 1 : (mgpc ? 2 : 16);
  if (poll == 1) {
    if (!CSB_CUDA(cudaMemcpyAsync(c->h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, cx.stream)))
      return false;
    if (!CSB_CUDA(cudaStreamSynchronize(cx.stream))) return false;
    if (c->h_st->cvg != CSB_SLES_ITERATING) return true;
  }
  for (int it = 1; ; it++) {
    if (!pc_apply_enqueue(c, a, precision, r_norm, rk, gk)) return false;
    if (!launch_ipcg_dots(n, st, false, rk, gk, rkm1, dk)) return false;
    if (it >= c->n_max_iter) break;     /* that was the final convergence test */
    if (!launch_ipcg_dir(n, st, gk, dk)) return false;
    if (!launch_spmv(EPI_PCG_DOT3, sa)) return false;
    if (!launch_ipcg_update(n, st, false, dk, zk, vx, rk, rkm1)) return false;
    if (mgpc && !launch_gate_from_it(mg_gate(c->pc_mg), st)) return false;
    if ((it % poll) == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(c->h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, cx.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(cx.stream))) return false;
      if (c->h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return true;
}

/* Single-reduction PCG with a general preconditioner (_conjugate_gradient_sr,
   cs_sles_it.cpp:1229-1415): pcg_enqueue with the fused Jacobi scaling replaced by the
   preconditioner's launches; the SpMV epilogue (3 dots + scalar recurrences) is shared. */
bool pcg_pc_enqueue(csb_sles_it_t *c, const csb_matrix_t *a, double precision, double r_norm,
                    const double *rhs, double *vx, bool vx_ini_nonzero)
{
  const int n = a->n_rows;
  const size_t ws = work_stride_of((size_t)a->n_cols_ext);
  Context &cx = ctx();
  double *rk = c->work.p, *dk = rk + ws, *gk = rk + 2*ws, *zk = rk + 3*ws, *sk = rk + 4*ws;
  ItState *st = c->st.p;
  const bool mgpc = pc_is_mg(c);

  if (!launch_it_state_init(st, nullptr, precision, r_norm, c->n_max_iter)) return false;
  if (vx_ini_nonzero) {
    SpmvArgs sa; sa.m = a; sa.x = vx; sa.y = rk;
    if (!launch_spmv(EPI_STORE, sa)) return false;
  }
  if (!launch_pcg_pc_init(n, rhs, vx, rk, vx_ini_nonzero)) return false;
  if (mgpc && !launch_mg_state_init(mg_gate(c->pc_mg), precision, r_norm, 1)) return false;

  if (!pc_apply_enqueue(c, a, precision, r_norm, rk, dk)) return false;
  SpmvArgs sa;
  sa.m = a; sa.st = st; sa.rk = rk;
  sa.x = dk; sa.y = zk; sa.phase = 0;
  if (!launch_spmv(EPI_PCG_DOT3, sa)) return false;
  if (!launch_pcg_pc_update_first(n, st, dk, zk, vx, rk)) return false;
  if (mgpc && !launch_gate_from_it(mg_gate(c->pc_mg), st)) return false;

  sa.x = gk; sa.y = sk; sa.phase = 1;
  const int poll = (mgpc && mg_is_k_cycle(c->pc_mg)) ? 1 : (mgpc ? 2 : 16);
  for (int it = 1; ; it++) {
    if (!pc_apply_enqueue(c, a, precision, r_norm, rk, gk)) return false;
    if (!launch_spmv(EPI_PCG_DOT3, sa)) return false;
    if (it >= c->n_max_iter) break;     /* that was the final convergence test */
    if (mgpc && !launch_gate_from_it(mg_gate(c->pc_mg), st)) return false;
    if (!launch_pcg_pc_update(n, st, gk, sk, dk, zk, rk, vx)) return false;
    if ((it % poll) == 0) {
      if (!CSB_CUDA(cudaMemcpyAsync(c->h_st, st, IT_HEAD, cudaMemcpyDeviceToHost, cx.stream)))
        return false;
      if (!CSB_CUDA(cudaStreamSynchronize(cx.stream))) return false;
      if (c->h_st->cvg != CSB_SLES_ITERATING) break;
    }
  }
  return true;
}

} // anonymous namespace

extern "C" {

csb_sles_it_t *
csb_sles_it_create(csb_sles_it_type_t solver_type, int poly_degree, int n_max_iter,
                   int update_stats)
{
  if (!ctx().ready) { set_error("csb_sles_it_create: csb_init not called"); return nullptr; }
  if (solver_type != CSB_SLES_PCG && solver_type != CSB_SLES_FCG
      && solver_type != CSB_SLES_IPCG
      && solver_type != CSB_SLES_JACOBI && solver_type != CSB_SLES_BICGSTAB
      && solver_type != CSB_SLES_GCR
      && solver_type != CSB_SLES_P_GAUSS_SEIDEL && solver_type != CSB_SLES_P_SYM_GAUSS_SEIDEL) {
    set_error("csb_sles_it_create: solver type %d is not available on this path",
              (int)solver_type);
    return nullptr;
  }
  /* poly_degree 0: Jacobi; < 0 with PCG: no preconditioner (_conjugate_gradient_sr_npc,
     cs_sles_it.cpp:1673-1860) */
  /* FCG with poly_degree < 0: the preconditioner is handed over afterwards
     (csb_sles_it_transfer_pc: one multigrid V-cycle), identity if none is */
  if (solver_type != CSB_SLES_JACOBI && !is_gs_type((int)solver_type) && poly_degree != 0
      && !(solver_type == CSB_SLES_GCR && poly_degree < 0)       /* GCR without preconditioner */
      && !(solver_type == CSB_SLES_PCG || solver_type == CSB_SLES_FCG
           || solver_type == CSB_SLES_IPCG)) {
    set_error("csb_sles_it_create: poly_degree %d: polynomial (> 0), none or transferred "
              "(< 0) preconditioners are available with PCG, FCG and IPCG only on this path",
              poly_degree);
    return nullptr;
  }
  csb_sles_it_t *c = new csb_sles_it_t;
  c->type = solver_type;
  c->poly_degree = poly_degree;
  c->n_max_iter = n_max_iter;
  c->update_stats = (update_stats != 0);
  /* fallback mechanism (cs_sles_it_create, cs_sles_it.cpp:4736-4749) */
  c->fallback_cvg = (solver_type == CSB_SLES_BICGSTAB) ? CSB_SLES_MAX_ITERATION
                                                       : CSB_SLES_DIVERGED;
  c->fallback_n_max_iter = n_max_iter;
  return c;
}

/* cs_sles_it_set_fallback_threshold (cs_sles_it.cpp:5791-5797): the solve switches to the
   fallback solver (GCR on a device, :5383-5385) when its convergence state is < threshold */
void csb_sles_it_set_fallback_threshold(csb_sles_it_t *c, csb_sles_convergence_state_t threshold,
                                        int n_iter_max)
{
  if (c == nullptr) return;
  c->fallback_cvg = (int)threshold;
  c->fallback_n_max_iter = n_iter_max;
  /* a fallback context made by an earlier solve would keep its own limit */
  delete c->fallback;
  c->fallback = nullptr;
  c->fallback_last = false;
}

/* cs_sles_it_set_restart_interval (cs_sles_it.cpp:5810-5817) */
void csb_sles_it_set_restart_interval(csb_sles_it_t *c, int interval)
{
  if (c == nullptr || interval < 1) return;
  c->restart_interval = interval;
  c->setup_done = false;
}

/* 1 when the last csb_sles_it_solve of this context went through its fallback solver */
int csb_sles_it_used_fallback(const csb_sles_it_t *c)
{
  return (c != nullptr && c->fallback != nullptr && c->fallback->n_solves > 0
          && c->fallback_last) ? 1 : 0;
}

int csb_sles_it_set_reduce_comm(csb_sles_it_t *c, void *nccl_comm)
{
  if (c == nullptr) return -1;
  c->nccl_comm = nccl_comm;
  return 0;
}

int csb_sles_it_set_pcomm(csb_sles_it_t *c, csb_pcomm_t *comm)
{
  if (c == nullptr) return -1;
  c->pcomm = comm;
  return 0;
}

void csb_sles_it_setup(void *context, const char *name, const csb_matrix_t *a, int verbosity)
{
  (void)name; (void)verbosity;
  csb_sles_it_t *c = static_cast<csb_sles_it_t *>(context);
  if (c == nullptr || a == nullptr) { set_error("csb_sles_it_setup: null argument"); return; }
  double t0 = wall();
  c->a = a;
  c->setup_done = false;
  const size_t n_wa = (c->type == CSB_SLES_GCR) ? 1 + 2*(size_t)c->restart_interval
    : (c->type == CSB_SLES_BICGSTAB) ? 6
    : ((c->type == CSB_SLES_PCG || c->type == CSB_SLES_FCG || c->type == CSB_SLES_IPCG) ? 5 : 1);
  if (is_gs_type((int)c->type) && a->halo != nullptr) {
    set_error("csb_sles_it_setup: the Gauss-Seidel solvers run on one rank on this path (the "
              "smoother forms take a halo)");
    return;
  }
  if (a->db != 1 && c->type != CSB_SLES_BICGSTAB && c->type != CSB_SLES_JACOBI
      && c->type != CSB_SLES_GCR) {
    set_error("csb_sles_it_setup: block matrices (db %d) are solved with CSB_SLES_BICGSTAB, "
              "CSB_SLES_GCR or CSB_SLES_JACOBI (block Jacobi) on this path", a->db);
    return;
  }
  if (a->db > 1 && c->type == CSB_SLES_JACOBI) {
    /* block_nn_inverse (cs_sles_it_setup :4999, cs_sles_it_priv.cpp:174-222): Gauss-Jordan
       inverses of 3x3 and 6x6 blocks, LU factors for the other sizes */
    if (a->halo != nullptr) { set_error("csb_sles_it_setup: block Jacobi runs on one rank"); return; }
    if (!c->ad_inv.alloc((size_t)a->n_rows*a->db*a->db)) return;
    if (!launch_block_nn_inverse(a->n_rows, a->db, a->d_val.p, c->ad_inv.p)) return;
    if (!c->work.alloc(work_stride_of((size_t)a->n_cols_ext*a->db))) return;
    if (!c->st.alloc(1)) return;
    if (c->h_st == nullptr)
      if (!CSB_CUDA(cudaMallocHost((void **)&c->h_st, sizeof(ItState)))) return;
    c->setup_done = true;
    c->n_setups += 1;
    c->t_setup += wall() - t0;
    return;
  }
  if (!c->ad_inv.alloc((size_t)a->n_rows*a->db)) return;
  if (a->db == 1 && c->type == CSB_SLES_PCG && c->poly_degree < 0) {
    /* no preconditioner: gk = rk * 1 is rk, the single-reduction recurrences are those of
       _conjugate_gradient_sr_npc term for term */
    if (!launch_fill(a->n_rows, 1.0, c->ad_inv.p)) return;
  }
  else if (a->db == 1) { if (!launch_ad_inv(a->n_rows, a->d_val.p, c->ad_inv.p)) return; }
  else if (!launch_ad_inv_block(a->n_rows, a->db, a->d_val.p, c->ad_inv.p)) return;
  if (!c->work.alloc(n_wa * work_stride_of((size_t)a->n_cols_ext*a->db))) return;
  if (!c->st.alloc(1)) return;
  if (c->h_st == nullptr)
    if (!CSB_CUDA(cudaMallocHost((void **)&c->h_st, sizeof(ItState)))) return;
  if (c->general_pc()) {
    /* general preconditioner: kernel per phase (fcg_pc_enqueue, pcg_pc_enqueue) */
    const bool dist_fcg_mg = (a->halo != nullptr && c->type == CSB_SLES_FCG
                              && c->pc_mg != nullptr && c->pcomm != nullptr);
    if (a->halo != nullptr && !dist_fcg_mg) {
      set_error("csb_sles_it_setup: on a partitioned matrix the general preconditioners are "
                "available as FCG + multigrid (csb_sles_it_transfer_pc) with a peer "
                "communicator (csb_sles_it_set_pcomm)");
      return;
    }
    if (dist_fcg_mg) csb_multigrid_set_pcomm(c->pc_mg, c->pcomm);
    if (c->pc_mg == nullptr && c->poly_degree > 0
        && !c->pc_tmp.alloc(work_stride_of((size_t)a->n_cols_ext))) return;
    SetupTrace tr;
    if (c->pc_mg != nullptr) {
      tr.mark("sles_it: ad_inv, work");
      csb_multigrid_setup(c->pc_mg, name, a, verbosity);
      tr.mark("sles_it: multigrid setup");
      if (!csb_multigrid_is_setup(c->pc_mg)) return;
      c->fcg_mg_persistent = false;
      if (c->type == CSB_SLES_FCG && mg_fcg_persistent_ok(c->pc_mg)
          && env_int("CSB_FCG_MG_MODE_MULTI", 0) == 0) {
        c->p_smem = mg_persistent_smem(c->pc_mg);
        c->p_grid = persistent_grid(4, c->p_smem);
        if (c->p_grid < 1) return;
        if (!c->partials.alloc(12*(size_t)c->p_grid) || !c->barrier.alloc(2)) return;
        if (!CSB_CUDA(cudaMemsetAsync(c->barrier.p, 0, 2*sizeof(unsigned), ctx().stream))) return;
        c->fcg_mg_persistent = true;
      }
      if (a->halo != nullptr && !c->fcg_mg_persistent) {
        set_error("csb_sles_it_setup: FCG + multigrid on a partitioned matrix needs the "
                  "persistent form (V-cycle, Jacobi-family smoothers with precision_mult < 0, "
                  "PCG coarse solver)");
        return;
      }
    }
    c->setup_done = true;
    c->n_setups += 1;
    c->t_setup += wall() - t0;
    return;
  }
  if (a->halo != nullptr && ((c->type != CSB_SLES_PCG && c->type != CSB_SLES_FCG && c->type != CSB_SLES_IPCG)
                             || c->pcomm == nullptr)) {
    set_error("csb_sles_it_setup: a matrix with a halo needs CSB_SLES_PCG / _FCG and a peer "
              "communicator (csb_sles_it_set_pcomm) on this path");
    return;
  }
  if (c->type == CSB_SLES_PCG || c->type == CSB_SLES_FCG || c->type == CSB_SLES_IPCG) {
    size_t max_entries = 0;
    MgLevelDesc &d = c->desc;
    memset(&d, 0, sizeof(d));
    d.n_g_rows = (double)a->n_rows;
    if (a->halo != nullptr) {
      if (c->arena_used) pcomm_arena_release(c->pcomm, c->arena_mark);
      c->arena_mark = pcomm_arena_mark(c->pcomm);
      c->arena_used = true;
      if (!matrix_host_mirror_structure(const_cast<csb_matrix_t *>(a))
          || !build_dist_level(c->pcomm, a->halo, a, c->dist)) return;
      d.dist = c->dist.d.p;
      if (!pcomm_allreduce(c->pcomm, &d.n_g_rows, 1)) return;
    }
    plan_slabs(a->h_row_index, a->n_rows, d.slab_rows, d.n_slabs, max_entries);
    d.gather_depth = plan_gather_depth(a->h_row_index, a->n_rows, &d.row_len_max);
    d.n_rows = a->n_rows;
    d.row_index = a->row_index.p; d.col_id = a->col_id.p;
    d.d_val = a->d_val.p; d.x_val = a->x_val.p; d.ad_inv = c->ad_inv.p;
    c->p_smem = persistent_smem_bytes(max_entries);
    c->p_grid = persistent_grid(c->type == CSB_SLES_FCG ? 2 : (c->type == CSB_SLES_IPCG ? 3 : 0), c->p_smem);
    if (c->p_grid < 1) return;
    if (!c->partials.alloc(12*(size_t)c->p_grid) || !c->barrier.alloc(2)) return;
    if (!CSB_CUDA(cudaMemsetAsync(c->barrier.p, 0, 2*sizeof(unsigned), ctx().stream))) return;
  }
  c->setup_done = true;
  c->n_setups += 1;
  c->t_setup += wall() - t0;
}

csb_sles_convergence_state_t
csb_sles_it_solve(void *context, const char *name, const csb_matrix_t *a, int verbosity,
                  double precision, double r_norm, int *n_iter, double *residual,
                  const csb_real_t *rhs, csb_real_t *vx_ini, csb_real_t *vx,
                  size_t aux_size, void *aux_vectors)
{
  csb_sles_it_t *c = static_cast<csb_sles_it_t *>(context);
  Context &cx = ctx();
  cx.err.clear();
  if (n_iter) *n_iter = 0;
  if (c == nullptr || a == nullptr || rhs == nullptr || vx == nullptr) {
    set_error("csb_sles_it_solve: null argument");
    return CSB_SLES_BREAKDOWN;
  }
  if (vx_ini != nullptr && vx_ini != vx) {
    set_error("csb_sles_it_solve: vx_ini must be vx or NULL (cs_sles_it.cpp:5182-5184)");
    return CSB_SLES_BREAKDOWN;
  }
  if (!c->setup_done || c->a != a)
    csb_sles_it_setup(c, name, a, verbosity);
  if (!c->setup_done) return CSB_SLES_BREAKDOWN;

  double t0 = wall();
  const bool ini = (vx_ini == vx);
  StagedVec s_rhs, s_vx;
  if (!s_rhs.init(const_cast<double *>(rhs), (size_t)a->n_rows*a->db, true, c->stage_rhs))
    return CSB_SLES_BREAKDOWN;
  if (!s_vx.init(vx, (size_t)a->n_cols_ext*a->db, ini, c->stage_vx)) return CSB_SLES_BREAKDOWN;

  bool ok = false;
  switch (c->type) {
  case CSB_SLES_PCG:
    if (c->general_pc())
      ok = pcg_pc_enqueue(c, a, precision, r_norm, s_rhs.d, s_vx.d, ini);
    else if (a->halo != nullptr || env_int("CSB_PCG_MODE_MULTI", 0) == 0) {
      /* whole solve in one persistent cooperative kernel (per rank) */
      PersistentCommon P;
      P.work = c->work.p; P.work_stride = work_stride_of((size_t)a->n_cols_ext);
      P.partials = c->partials.p; P.barrier = c->barrier.p;
      P.comm = (a->halo != nullptr) ? c->pcomm->d : nullptr;
      const bool single = (a->n_rows <= sb_rows_threshold());
      if (a->halo != nullptr) pcomm_host_barrier(c->pcomm);
      ok = launch_pcg_persistent(c->desc, s_rhs.d, s_vx.d, ini, c->n_max_iter, precision,
                                 r_norm, P, c->st.p, c->p_grid, c->p_smem, single,
                                 cx.device_share <= 1);
      if (a->halo != nullptr) pcomm_host_barrier(c->pcomm);   /* all ranks launched */
    }
    else
      ok = pcg_enqueue(a, c->st.p, nullptr, c->ad_inv.p, precision, r_norm, c->n_max_iter,
                       s_rhs.d, s_vx.d, ini, false, c->work.p, c->h_st);
    break;
  case CSB_SLES_FCG:
  case CSB_SLES_IPCG:
    if (c->general_pc() && c->type == CSB_SLES_FCG && c->fcg_mg_persistent) {
      /* FCG + cs_multigrid_pc: the whole solve in one persistent cooperative kernel */
      FcgMgParams S;
      PersistentCommon P;
      mg_fcg_params(c->pc_mg, S, P);
      S.st = c->st.p;
      S.rhs = s_rhs.d; S.vx = s_vx.d; S.vx_ini_nonzero = ini ? 1 : 0;
      S.n_max_iter = c->n_max_iter; S.precision = precision; S.r_norm = r_norm;
      S.single_rows = sb_rows_threshold();
      S.small_rows = env_int("CSB_PCG_SMALL_ROWS", 448);
      S.fwork = c->work.p; S.fstride = work_stride_of((size_t)a->n_cols_ext);
      P.partials = c->partials.p; P.barrier = c->barrier.p;
      if (!launch_mg_state_init(S.ms, precision, r_norm, 1)) return CSB_SLES_BREAKDOWN;
      if (a->halo != nullptr) pcomm_host_barrier(c->pcomm);   /* nobody inside cudaFree */
      ok = launch_fcg_mg_persistent(S, P, c->p_grid, c->p_smem, cx.device_share <= 1);
      if (a->halo != nullptr) pcomm_host_barrier(c->pcomm);   /* all ranks launched */
      ok = ok && mg_fetch_state(c->pc_mg);
      break;
    }
    if (c->general_pc()) {
      ok = (c->type == CSB_SLES_IPCG)
        ? ipcg_pc_enqueue(c, a, precision, r_norm, s_rhs.d, s_vx.d, ini)
        : fcg_pc_enqueue(c, a, precision, r_norm, s_rhs.d, s_vx.d, ini);
      break;
    }
    {
      /* whole solve in one persistent cooperative kernel (per rank) */
      PersistentCommon P;
      P.work = c->work.p; P.work_stride = work_stride_of((size_t)a->n_cols_ext);
      P.partials = c->partials.p; P.barrier = c->barrier.p;
      P.comm = (a->halo != nullptr) ? c->pcomm->d : nullptr;
      const bool single = (a->n_rows <= sb_rows_threshold());
      if (a->halo != nullptr) pcomm_host_barrier(c->pcomm);
      ok = launch_fcg_persistent(c->desc, s_rhs.d, s_vx.d, ini, c->n_max_iter, precision,
                                 r_norm, P, c->st.p, c->p_grid, c->p_smem, single,
                                 cx.device_share <= 1, c->type == CSB_SLES_IPCG);
      if (a->halo != nullptr) pcomm_host_barrier(c->pcomm);
    }
    break;
  case CSB_SLES_JACOBI:
    if (a->db > 1) {
      /* _block_3_jacobi for 3x3 blocks, _block_jacobi otherwise (cs_sles_it.cpp:5058-5064) */
      ok = block3_jacobi_enqueue(a, c->st.p, c->ad_inv.p, precision, r_norm, c->n_max_iter,
                                 s_rhs.d, s_vx.d, ini, c->work.p, c->h_st);
      break;
    }
    ok = jacobi_enqueue(a, c->st.p, c->ad_inv.p, precision, r_norm, c->n_max_iter,
                        s_rhs.d, s_vx.d, ini, c->work.p, c->h_st);
    break;
  case CSB_SLES_BICGSTAB:
    ok = bicgstab_enqueue(a, c->st.p, c->ad_inv.p, precision, r_norm, c->n_max_iter,
                          s_rhs.d, s_vx.d, ini, c->work.p, c->h_st);
    break;
  case CSB_SLES_GCR:
    ok = gcr_enqueue(a, c->st.p, c->h_st, (c->poly_degree < 0) ? nullptr : c->ad_inv.p,
                     c->restart_interval, precision, r_norm, c->n_max_iter, s_rhs.d, s_vx.d,
                     ini, c->work.p);
    break;
  case CSB_SLES_P_GAUSS_SEIDEL:
  case CSB_SLES_P_SYM_GAUSS_SEIDEL:
    ok = gs_enqueue(a, (int)c->type, c->st.p, c->ad_inv.p, precision, r_norm, c->n_max_iter,
                    s_rhs.d, s_vx.d, ini, c->work.p, c->h_st);
    break;
  default:
    set_error("csb_sles_it_solve: solver type %d", (int)c->type);
    break;
  }
  if (!ok) return CSB_SLES_BREAKDOWN;

  if (!CSB_CUDA(cudaMemcpyAsync(c->h_st, c->st.p, sizeof(ItState), cudaMemcpyDeviceToHost,
                                cx.stream))) return CSB_SLES_BREAKDOWN;
  if (!s_vx.copy_back((size_t)a->n_rows*a->db)) return CSB_SLES_BREAKDOWN;
  if (!CSB_CUDA(cudaStreamSynchronize(cx.stream))) return CSB_SLES_BREAKDOWN;

  int cvg = c->h_st->cvg;
  if (cvg == CSB_SLES_ITERATING) cvg = CSB_SLES_MAX_ITERATION;
  /* reported iteration count = value at the last convergence test */
  int n_it = c->h_st->n_iter;
  if (n_iter) *n_iter = n_it;
  if (residual) *residual = c->h_st->residual;

  c->n_solves += 1;
  if (c->n_iterations_tot == 0 || c->n_iterations_min > (unsigned)n_it)
    c->n_iterations_min = n_it;
  if (c->n_iterations_max < (unsigned)n_it) c->n_iterations_max = n_it;
  c->n_iterations_last = n_it;
  c->n_iterations_tot += n_it;
  c->t_solve += wall() - t0;

  /* _fallback (cs_sles_it.cpp:4464-4528, call site :5383-5418): GCR on a device, from zero
     unless the first solver merely broke down, its iterations added to the count */
  c->fallback_last = false;
  if (cvg < c->fallback_cvg && c->type != CSB_SLES_GCR && !c->is_smoother) {
    if (c->fallback == nullptr) {
      c->fallback = csb_sles_it_create(CSB_SLES_GCR, 0, c->fallback_n_max_iter,
                                       c->update_stats ? 1 : 0);
      if (c->fallback == nullptr) return (csb_sles_convergence_state_t)cvg;
    }
    int fb_iter = 0;
    csb_real_t *fb_ini = (cvg < CSB_SLES_BREAKDOWN) ? nullptr : vx;
    const csb_sles_convergence_state_t fcvg
      = csb_sles_it_solve(c->fallback, name, a, verbosity, precision, r_norm, &fb_iter,
                          residual, rhs, fb_ini, vx, aux_size, aux_vectors);
    csb_sles_it_free(c->fallback);
    if (n_iter) *n_iter = fb_iter + n_it;
    c->fallback_last = true;
    return fcvg;
  }
  return (csb_sles_convergence_state_t)cvg;
}

void csb_sles_it_free(void *context)
{
  csb_sles_it_t *c = static_cast<csb_sles_it_t *>(context);
  if (c == nullptr) return;
  SetupTrace tr;
  c->ad_inv.release();
  c->work.release();
  if (c->pc_mg != nullptr) csb_multigrid_free(c->pc_mg);
  tr.mark("csb_sles_it_free");
  if (c->arena_used && c->pcomm != nullptr) {
    pcomm_arena_release(c->pcomm, c->arena_mark);
    c->arena_used = false;
  }
  c->a = nullptr;
  c->setup_done = false;
}

void csb_sles_it_destroy(void **context)
{
  if (context == nullptr || *context == nullptr) return;
  csb_sles_it_t *c = static_cast<csb_sles_it_t *>(*context);
  delete c;
  *context = nullptr;
}

double csb_sles_it_get_last_initial_residual(const csb_sles_it_t *c)
{
  return (c != nullptr && c->h_st != nullptr) ? c->h_st->initial_residual : -1.;
}

int csb_sles_it_get_residual_history(const csb_sles_it_t *c, double *hist, int max_n)
{
  if (c == nullptr || c->h_st == nullptr) return 0;
  int n = c->h_st->n_hist;
  if (n > max_n) n = max_n;
  for (int i = 0; i < n; i++) hist[i] = c->h_st->hist[i];
  return n;
}

void csb_sles_it_log(const void *context, int log_type)
{
  (void)log_type;
  const csb_sles_it_t *c = static_cast<const csb_sles_it_t *>(context);
  if (c == nullptr) return;
  const char *tn = (c->type == CSB_SLES_PCG) ? "Conjugate Gradient"
                 : (c->type == CSB_SLES_FCG) ? "Flexible Conjugate Gradient"
                 : (c->type == CSB_SLES_IPCG) ? "Inexact Conjugate Gradient"
                 : (c->type == CSB_SLES_JACOBI) ? "Jacobi"
                 : (c->type == CSB_SLES_L1_JACOBI) ? "L1 Jacobi"
                 : (c->type == CSB_SLES_R_JACOBI) ? "Relaxed Jacobi"
                 : (c->type == CSB_SLES_RJ2) ? "Scheduled relaxed Jacobi (M=2)"
                 : (c->type == CSB_SLES_RJ3) ? "Scheduled relaxed Jacobi (M=3)"
                 : (c->type == CSB_SLES_P_GAUSS_SEIDEL) ? "Local Gauss-Seidel"
                 : (c->type == CSB_SLES_P_SYM_GAUSS_SEIDEL) ? "Local symmetric Gauss-Seidel"
                 : (c->type == CSB_SLES_TS_F_GAUSS_SEIDEL) ? "Truncated forward Gauss-Seidel"
                 : (c->type == CSB_SLES_TS_B_GAUSS_SEIDEL) ? "Truncated backward Gauss-Seidel"
                 : "BiCGstab";
  printf("  Solver type:                       %s (B200 device)\n"
         "  Number of setups:                  %12u\n"
         "  Number of calls:                   %12u\n"
         "  Minimum number of iterations:      %12u\n"
         "  Maximum number of iterations:      %12u\n"
         "  Total number of iterations:        %12llu\n"
         "  Total setup time:                  %12.3f\n"
         "  Total solution time:               %12.3f\n",
         tn, c->n_setups, c->n_solves, c->n_iterations_min, c->n_iterations_max,
         c->n_iterations_tot, c->t_setup, c->t_solve);
}

void *csb_sles_it_copy(const void *context)
{
  const csb_sles_it_t *c = static_cast<const csb_sles_it_t *>(context);
  if (c == nullptr) return nullptr;
  csb_sles_it_t *d = c->is_smoother
    ? csb_multigrid_smoother_create(c->type, c->poly_degree, c->n_max_iter)
    : csb_sles_it_create(c->type, c->poly_degree, c->n_max_iter, c->update_stats ? 1 : 0);
  if (d) {
    d->nccl_comm = c->nccl_comm; d->pcomm = c->pcomm;
    /* the preconditioner is cloned with its settings (cs_sles_pc_clone, cs_sles_it.cpp:4920) */
    if (c->pc_mg != nullptr)
      d->pc_mg = static_cast<csb_multigrid_t *>(csb_multigrid_copy(c->pc_mg));
  }
  return d;
}

} /* extern "C" */

/*----------------------------------------------------------------------------
 * cs_multigrid
 *----------------------------------------------------------------------------*/

namespace {

struct MgLevel {
  const csb_matrix_t *m = nullptr;
  csb_matrix_t       *owned = nullptr;
  /* towards the next (coarser) level */
  DevBuf<int>         coarse_row;                 /* this level's rows -> next level */
  DevBuf<int>         c_f_row_index, c_f_row_ids; /* next level rows -> this level's rows */
  std::vector<int>    h_coarse_row;
  int                 r_slab_rows = 0, r_n_slabs = 0;   /* slab plan of (c_f_row_index, c_f_row_ids) */
  size_t              r_max_entries = 0;
  /* solve data */
  DevBuf<double>      ad_inv, rhs, vx;
  DevBuf<double>      l1_inv;                     /* L1 Jacobi smoother only */
  DevBuf<double>      kwork;                      /* K-cycle: the 10 vectors of this level
                                                     (mgd->rhs_vx[level*10 + k], cs_multigrid.cpp:1950) */
  size_t              kstride = 0;
  /* geometry of this level used to coarsen it (level > 0; level 0 uses the
     matrix's mesh association) */
  DevBuf<int>         cell_face;
  DevBuf<short>       cell_face_sgn;
  DevBuf<double>      cell_cen, cell_vol, face_normal, xa0, xa0ij;
  int                 n_faces = 0;
  /* hierarchy of a matrix mapped from native arrays: the level's faces (2 cells each,
     numbered like _coarsen_face_cell numbers them) and their values; host copies feed the
     face-list aggregation of the next step */
  DevBuf<int>         face_cell;
  DevBuf<double>      xa;
  HostVec<int>        h_face_cell;
  HostVec<double>     h_xa;
  /* distributed hierarchy */
  csb_halo_t         *halo = nullptr;       /* this level's halo (owned when level > 0) */
  bool                halo_owned = false;
  double              n_g_rows = 0.;        /* rows over all ranks */
  DistLevelHost       dist;
  /* coarse-level merge: this level and the ones below live on rank 0 only */
  bool                merged = false;       /* rank 0: rows of all ranks, no halo */
  bool                stub = false;         /* other ranks: their part of the merge level */
  double             *arena_rhs = nullptr;  /* rank 0, merge level: peer-writable rhs */
  double             *arena_vx = nullptr;   /* other ranks, merge level: peer-writable vx */
  ~MgLevel() {
    if (owned) csb_matrix_destroy(&owned);
    if (halo_owned && halo) csb_halo_destroy(&halo);
  }
};

} // anonymous namespace

struct _csb_multigrid_t {
  csb_multigrid_type_t type = CSB_MULTIGRID_V_CYCLE;
  /* settings, defaults of cs_multigrid_create (cs_multigrid.cpp:4689-4764) */
  int     aggregation_limit = 3;
  int     coarsening_type = 0;
  int     n_levels_max = 25;
  unsigned long long n_g_rows_min = 30;
  double  p0p1_relax = 0.95;
  double  k_cycle_threshold = 0.;      /* cs_multigrid.cpp:4707, :4724 */
  csb::DevBuf<double> kc_scal;         /* K-cycle: 8 device scalars per level (k_kc_scalars) */
  int     n_blocks = 0;      /* 0: as many as OpenMP threads (the reference's behaviour) */
  csb_sles_it_type_t solver_type[3] = {CSB_SLES_PCG, CSB_SLES_PCG, CSB_SLES_PCG};
  int     n_max_cycles = 100;
  int     n_max_iter[3] = {2, 10, 500};
  int     poly_degree[3] = {0, 0, 0};
  double  precision_mult[3] = {1., 1., 1.};
  void   *nccl_comm = nullptr;
  csb_pcomm_t *pcomm = nullptr;
  size_t  arena_mark = 0;
  bool    arena_used = false;
  size_t  work_stride = 0;             /* distance between the PCG work vectors in aux */
  /* cs_multigrid_set_merge_options (cs_multigrid.cpp:4689-4694: stride 1 = no merging) */
  int     merge_stride = 1;
  int     merge_mean_threshold = 300;
  unsigned long long merge_glob_threshold = 500;
  int     merge_level = -1;              /* set by setup */
  MergeDev          h_merge;
  DevBuf<MergeDev>  d_merge;
  DevBuf<unsigned>  merge_count;
  /* setup data */
  std::vector<MgLevel *> levels;
  DevBuf<ItState>  it_states;   /* 2 per level */
  DevBuf<MgState>  mg_state;
  MgState         *h_mg = nullptr;   /* pinned */
  DevBuf<double>   wr, aux;
  /* persistent-kernel launch data */
  DevBuf<MgLevelDesc> descs;
  DevBuf<double>   partials;
  DevBuf<double>   stage_rhs, stage_vx;      /* host rhs / vx staged here, kept across solves */
  DevBuf<unsigned> barrier;
  int              p_grid = 0;
  size_t           p_smem = 0;
  bool             setup_done = false;
  const csb_matrix_t *a = nullptr;
  /* results / stats */
  int      n_cycles_last = 0;
  unsigned n_setups = 0, n_solves = 0;
  unsigned long long n_cycles_tot = 0;
  double   t_setup_tot = 0, t_solve_tot = 0;
  double   t_setup[4] = {0, 0, 0, 0};
  ~_csb_multigrid_t() {
    for (MgLevel *l : levels) delete l;
    if (h_mg) cudaFreeHost(h_mg);
  }
};

_csb_sles_it_t::~_csb_sles_it_t()
{
  if (h_st) cudaFreeHost(h_st);
  delete pc_mg;
  delete fallback;
}

namespace {

int mg_n_levels(const csb_multigrid_t *mg) { return (int)mg->levels.size(); }
bool mg_is_k_cycle(const csb_multigrid_t *mg) { return mg->type == CSB_MULTIGRID_K_CYCLE; }
MgState *mg_gate(csb_multigrid_t *mg) { return mg->mg_state.p; }

bool mg_fcg_persistent_ok(const csb_multigrid_t *mg)
{
  if (mg->type != CSB_MULTIGRID_V_CYCLE || mg->levels.size() < 2) return false;
  for (int i = 0; i < 2; i++)
    if (!is_smoother_type((int)mg->solver_type[i]) || is_gs_type((int)mg->solver_type[i])
        || mg->precision_mult[i] >= 0 || mg->n_max_iter[i] < 1) return false;
  return (mg->solver_type[2] == CSB_SLES_PCG && mg->poly_degree[2] == 0);
}

size_t mg_persistent_smem(const csb_multigrid_t *mg) { return mg->p_smem; }

bool mg_fetch_state(csb_multigrid_t *mg)
{
  return CSB_CUDA(cudaMemcpyAsync(mg->h_mg, mg->mg_state.p, sizeof(MgState),
                                  cudaMemcpyDeviceToHost, ctx().stream));
}

void mg_fcg_params(csb_multigrid_t *mg, FcgMgParams &S, PersistentCommon &P)
{
  memset(&S, 0, sizeof(S));
  S.n_levels = (int)mg->levels.size();
  S.levels = mg->descs.p;
  S.ms = mg->mg_state.p;
  for (int i = 0; i < 2; i++) {
    const SmootherPlan p = smoother_plan((int)mg->solver_type[i], mg->n_max_iter[i]);
    S.sm[i].mode = p.mode; S.sm[i].wk_m = p.wk_m; S.sm[i].n_iter = p.n_iter;
    for (int k = 0; k < 3; k++) S.sm[i].w[k] = p.w[k];
  }
  S.coarse_n_max_iter = mg->n_max_iter[2];
  S.coarse_precision_mult = mg->precision_mult[2];
  S.wr = mg->wr.p;
  S.tmp = mg->aux.p;          /* free while no level-local PCG runs */
  /* the coarsest-level PCG works in the hierarchy's own vectors */
  P.work = mg->aux.p;
  P.work_stride = mg->work_stride;
  S.merge = (mg->merge_level >= 0) ? mg->d_merge.p : nullptr;
  S.merge_level = mg->merge_level;
  P.partials = nullptr; P.barrier = nullptr;
  P.comm = (mg->levels[0]->m->halo != nullptr && mg->pcomm != nullptr) ? mg->pcomm->d : nullptr;
}

/* One level's descent (which = 0) or ascent (1) smoother of the single-cycle form, as
   _multigrid_setup_sles builds it (cs_multigrid.cpp:2150-2175): a cs_multigrid_smoother
   when precision_mult < 0 and the type is one, cs_sles_it PCG + Jacobi otherwise. */
bool mg_level_smooth(csb_multigrid_t *mg, int which, int level, const double *rhs_lv,
                     double *vx_lv, bool ini, double precision, double r_norm,
                     bool count = true)   /* false: not part of n_equiv_iter */
{
  MgLevel *f = mg->levels[level];
  MgState *ms = mg->mg_state.p;
  const int type = (int)mg->solver_type[which];
  const double ratio = (double)f->m->n_rows / (double)mg->levels[0]->m->n_rows;
  if (type == CSB_SLES_PCG) {
    ItState *st = mg->it_states.p + 2*level + which;
    return pcg_enqueue(f->m, st, ms, f->ad_inv.p, precision*mg->precision_mult[which], r_norm,
                       mg->n_max_iter[which], rhs_lv, vx_lv, ini, true, mg->aux.p, nullptr)
           && launch_mg_accum_iter(ms, st, level, ratio, 0);
  }
  const double *inv = (type == CSB_SLES_L1_JACOBI) ? f->l1_inv.p : f->ad_inv.p;
  if (!smoother_enqueue(f->m, type, mg->n_max_iter[which], inv, rhs_lv, vx_lv, ini,
                        mg->aux.p, ms)) return false;
  const int n_it = (type == CSB_SLES_TS_F_GAUSS_SEIDEL || type == CSB_SLES_TS_B_GAUSS_SEIDEL)
    ? 1 : smoother_plan(type, mg->n_max_iter[which]).n_iter;
  return !count || launch_mg_accum_n(ms, n_it,
                              f->m->n_rows, 1.0 / (double)mg->levels[0]->m->n_rows);
}

/* (x.x, x.y, y.z) -- or x.y / x.x alone -- on the device, result read back (GCR's
   Gram-Schmidt coefficients, declared above) */
bool dots_to_host(int n, const double *x, const double *y, const double *z, int mode,
                  double r[3])
{
  Context &c = ctx();
  if (!launch_dot(n, x, y, z, mode)) return false;
  if (!CSB_CUDA(cudaMemcpyAsync(c.red_out_h, c.red_out, 3*sizeof(double),
                                cudaMemcpyDeviceToHost, c.stream))) return false;
  if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
  r[0] = c.red_out_h[0]; r[1] = c.red_out_h[1]; r[2] = c.red_out_h[2];
  return true;
}

/* _multigrid_k_cycle (cs_multigrid.cpp:4051-4480), single-cycle form: pre-smoothing from
   zero, residual, restriction; next level solved by the coarse solver, or by one or two
   recursive K-cycles combined by two steps of a Krylov acceleration (the second one skipped
   when ||rt|| < trsh ||r||); overwrite prolongation, residual, post-smoothing from zero, sum.
   Kernel per phase; the inner Krylov coefficients are computed and consumed on the device
   (k_kc_scalars, k_kc_vec_d); one read per level visit -- the two norms of the threshold
   test -- comes back to the host, which takes the reference's branch (none with a zero
   threshold). */
bool k_cycle_enqueue(csb_multigrid_t *mg, int level, double precision, double r_norm,
                     const double *rhs, double *vx)
{
  const int coarsest = (int)mg->levels.size() - 1;
  MgLevel *f = mg->levels[level], *cl = mg->levels[level+1];
  const csb_matrix_t *fm = f->m, *cm = cl->m;
  const int f_n = fm->n_rows, c_n = cm->n_rows;
  MgState *ms = mg->mg_state.p;
  const double trsh = mg->k_cycle_threshold;
  /* slots of mgd->rhs_vx (cs_multigrid.cpp:4192, :4262-4263, :4311-4312, :4345-4346, :4396-4398) */
  double *rt_lv = f->kwork.p, *rb_lv = rt_lv + f->kstride, *z1_lv = rt_lv + 2*f->kstride,
         *z2_lv = rt_lv + 3*f->kstride;
  double *vx_lv1 = cl->kwork.p + 4*cl->kstride, *rhs_lv1 = cl->kwork.p + 5*cl->kstride,
         *v_lv1 = cl->kwork.p + 6*cl->kstride, *rt_lv1 = cl->kwork.p + 7*cl->kstride,
         *vx2_lv1 = cl->kwork.p + 8*cl->kstride, *w_lv1 = cl->kwork.p + 9*cl->kstride;

  /* only the level-0 visit counts in n_equiv_iter: deeper visits add to a local of the
     recursion in the reference (:4281-4282) */
  if (!mg_level_smooth(mg, 0, level, rhs, vx, false, precision, r_norm, level == 0))
    return false;
  SpmvArgs sa;
  sa.m = fm; sa.x = vx; sa.y = rt_lv; sa.b = rhs;
  if (!launch_spmv(EPI_RESIDUAL, sa)) return false;
  if (!launch_restrict(c_n, f->c_f_row_index.p, f->c_f_row_ids.p, rt_lv, rhs_lv1, nullptr))
    return false;

  if (level == coarsest - 1) {
    ItState *st = mg->it_states.p + 2*coarsest;
    if (!pcg_enqueue(cm, st, ms, cl->ad_inv.p, precision*mg->precision_mult[2], r_norm,
                     mg->n_max_iter[2], rhs_lv1, vx_lv1, false, true, mg->aux.p, nullptr))
      return false;
    if (level == 0
        && !launch_mg_accum_iter(ms, st, coarsest,
                                 (double)c_n / (double)mg->levels[0]->m->n_rows, 0)) return false;
  }
  else {
    if (!k_cycle_enqueue(mg, level + 1, precision, r_norm, rhs_lv1, vx_lv1)) return false;
    SpmvArgs sc;
    sc.m = cm; sc.x = vx_lv1; sc.y = v_lv1;
    if (!launch_spmv(EPI_STORE, sc)) return false;
    /* the coefficients of the inner Krylov steps stay on the device (ks, 8 doubles of this
       level visit): only the two norms of the threshold test come back, in one read */
    double *ks = mg->kc_scal.p + 8*(size_t)level;
    if (!launch_dot(c_n, v_lv1, vx_lv1, rhs_lv1, 2) || !launch_kc_scalars(0, ks)) /* _dot_xy_yz */
      return false;
    if (!launch_kc_vec_d(0, c_n, ks + 1, nullptr, v_lv1, rhs_lv1, rt_lv1)) return false;
    double rt_lv1_norm = 1.0, r_lv1_norm = 0.0;
    if (trsh > 0) {                                                         /* _dot_xx_yy */
      Context &c = ctx();
      if (!launch_dot(c_n, rt_lv1, nullptr, nullptr, 1) || !launch_kc_scalars(1, ks)
          || !launch_dot(c_n, rhs_lv1, nullptr, nullptr, 1) || !launch_kc_scalars(2, ks)
          || !CSB_CUDA(cudaMemcpyAsync(c.red_out_h, ks + 2, 2*sizeof(double),
                                       cudaMemcpyDeviceToHost, c.stream))
          || !CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
      rt_lv1_norm = c.red_out_h[0];
      r_lv1_norm = c.red_out_h[1];
    }
    if (rt_lv1_norm < trsh * trsh * r_lv1_norm) {
      if (!launch_kc_vec_d(1, c_n, ks + 1, nullptr, nullptr, nullptr, vx_lv1)) return false;
    }
    else {
      if (!k_cycle_enqueue(mg, level + 1, precision, r_norm, rt_lv1, vx2_lv1)) return false;
      sc.x = vx2_lv1; sc.y = w_lv1;
      if (!launch_spmv(EPI_STORE, sc)) return false;
      if (!launch_dot(c_n, v_lv1, vx2_lv1, w_lv1, 2) || !launch_kc_scalars(3, ks)  /* _dot_xu_xv_xw */
          || !launch_dot(c_n, vx2_lv1, rt_lv1, nullptr, 0) || !launch_kc_scalars(4, ks))
        return false;
      if (!launch_kc_vec_d(2, c_n, ks + 6, ks + 7, vx2_lv1, nullptr, vx_lv1)) return false;
    }
  }

  if (!launch_fill(f_n, 0.0, z2_lv)) return false;
  if (!launch_prolong_set(f_n, f->coarse_row.p, vx_lv1, z1_lv)) return false;
  sa.x = z1_lv; sa.y = rb_lv; sa.b = rt_lv;
  if (!launch_spmv(EPI_RESIDUAL, sa)) return false;
  if (!mg_level_smooth(mg, 1, level, rb_lv, z2_lv, true, precision, r_norm, level == 0))
    return false;
  return launch_kc_vec(3, f_n, 0., 0., z1_lv, z2_lv, vx);
}

/* _multigrid_v_cycle_pc (cs_multigrid.cpp:3569-4000): one V-cycle without convergence
   tests -- descent smoother from zero (from vx at level 0 when vx_ini_nonzero), residual,
   restriction; coarsest-level PCG; prolongation and the ascent smoother on every level,
   level 0 included.  A sequence of launches, no host synchronisation; every launch is
   gated on mg_state.cvg (closed by the outer solver once it has converged). */
bool mg_cycle_pc_enqueue(csb_multigrid_t *mg, double precision, double r_norm,
                         const double *rhs_d, double *vx_d, bool vx_ini_nonzero)
{
  const int n_levels = (int)mg->levels.size();
  const int coarsest = n_levels - 1;
  MgState *ms = mg->mg_state.p;
  const double denom = 1.0 / (double)mg->levels[0]->m->n_rows;
  bool ok = true;

  if (mg->type == CSB_MULTIGRID_K_CYCLE && n_levels > 1)
    return k_cycle_enqueue(mg, 0, precision, r_norm, rhs_d, vx_d);   /* starts from zero */

  for (int level = 0; ok && level < coarsest; level++) {
    MgLevel *f = mg->levels[level], *cl = mg->levels[level+1];
    const double *rhs_lv = (level == 0) ? rhs_d : f->rhs.p;
    double *vx_lv = (level == 0) ? vx_d : f->vx.p;
    ok = mg_level_smooth(mg, 0, level, rhs_lv, vx_lv, (level == 0) ? vx_ini_nonzero : false,
                         precision, r_norm);
    SpmvArgs sa;
    sa.m = f->m; sa.x = vx_lv; sa.y = mg->wr.p; sa.b = rhs_lv; sa.mg = ms;
    ok = ok && launch_spmv(EPI_RESIDUAL, sa);
    ok = ok && launch_restrict(cl->m->n_rows, f->c_f_row_index.p, f->c_f_row_ids.p,
                               mg->wr.p, cl->rhs.p, ms);
  }

  if (ok) {
    MgLevel *cl = mg->levels[coarsest];
    const double *rhs_lv = (coarsest == 0) ? rhs_d : cl->rhs.p;
    double *vx_lv = (coarsest == 0) ? vx_d : cl->vx.p;
    ItState *st = mg->it_states.p + 2*coarsest;
    ok = pcg_enqueue(cl->m, st, ms, cl->ad_inv.p, precision*mg->precision_mult[2], r_norm,
                     mg->n_max_iter[2], rhs_lv, vx_lv, false, true, mg->aux.p, nullptr);
    ok = ok && launch_mg_accum_iter(ms, st, coarsest, (double)cl->m->n_rows * denom, 0);
  }

  for (int level = coarsest - 1; ok && level > -1; level--) {
    MgLevel *f = mg->levels[level], *cl = mg->levels[level+1];
    const double *rhs_lv = (level == 0) ? rhs_d : f->rhs.p;
    double *vx_lv = (level == 0) ? vx_d : f->vx.p;
    ok = launch_prolong_inc(f->m->n_rows, f->coarse_row.p, cl->vx.p, vx_lv, ms);
    ok = ok && mg_level_smooth(mg, 1, level, rhs_lv, vx_lv, true, precision, r_norm);
  }
  return ok;
}

void mg_clear_hierarchy(csb_multigrid_t *mg)
{
  for (MgLevel *l : mg->levels) delete l;
  mg->levels.clear();
  if (mg->arena_used && mg->pcomm != nullptr) {
    pcomm_arena_release(mg->pcomm, mg->arena_mark);
    mg->arena_used = false;
  }
  mg->wr.release();
  mg->aux.release();
  mg->it_states.release();
  mg->merge_level = -1;
  mg->setup_done = false;
  mg->a = nullptr;
}

} // anonymous namespace
namespace {
bool matrix_host_mirror_structure(csb_matrix_t *m)
{
  if (m->host_structure_ready()) return true;
  m->h_col_id.resize(std::max<size_t>((size_t)m->nnz, 1));
  if (m->nnz > 0 &&
      !CSB_CUDA(cudaMemcpyAsync(m->h_col_id.data(), m->col_id.p, sizeof(int)*(size_t)m->nnz,
                                cudaMemcpyDeviceToHost, ctx().stream))) return false;
  return CSB_CUDA(cudaStreamSynchronize(ctx().stream));
}
}
namespace {

/* make sure the host mirrors of a matrix are available for aggregation */
bool matrix_host_mirror(csb_matrix_t *m)
{
  Context &c = ctx();
  if (!m->host_structure_ready()) {
    m->h_col_id.resize(std::max<size_t>((size_t)m->nnz, 1));
    if (m->nnz > 0 &&
        !CSB_CUDA(cudaMemcpyAsync(m->h_col_id.data(), m->col_id.p, sizeof(int)*(size_t)m->nnz,
                                  cudaMemcpyDeviceToHost, c.stream))) return false;
  }
  if (!m->h_vals_valid && (m->b_d_val == nullptr || m->b_x_val == nullptr)) {
    m->h_d_val.resize(std::max<size_t>((size_t)m->n_rows, 1));
    m->h_x_val.resize(std::max<size_t>((size_t)m->nnz, 1));
    if (m->n_rows > 0 &&
        !CSB_CUDA(cudaMemcpyAsync(m->h_d_val.data(), m->d_val.p, sizeof(double)*(size_t)m->n_rows,
                                  cudaMemcpyDeviceToHost, c.stream))) return false;
    if (m->nnz > 0 &&
        !CSB_CUDA(cudaMemcpyAsync(m->h_x_val.data(), m->x_val.p, sizeof(double)*(size_t)m->nnz,
                                  cudaMemcpyDeviceToHost, c.stream))) return false;
    m->h_vals_valid = true;
  }
  return CSB_CUDA(cudaStreamSynchronize(c.stream));
}

/* Coarse halo of a distributed level: _coarsen_halo (src/alge/cs_grid.cpp:924-1329).
   f_c_row has f_n_cols_ext entries; the local part is the aggregation map, the
   ghost part is filled here with coarse ghost column ids (c_n_rows + position in
   the coarse ghost sections).  Collective over the neighbour ranks. */
csb_halo_t *coarsen_halo(csb_pcomm_t *pc, const csb_halo_t *fh, int f_n_rows, int c_n_rows,
                         std::vector<int> &f_c_row)
{
  Context &c = ctx();
  const int nd = fh->n_c_domains;
  const int n_f_send = fh->n_send, n_f_ghost = fh->n_ghost;
  for (int ii = f_n_rows; ii < f_n_rows + n_f_ghost; ii++) f_c_row[ii] = -1;      /* :960-962 */

  /* sub_num shifted by 1 so that penalised rows (-1) map to slot 0 (:971-977) */
  std::vector<int> sub_num((size_t)c_n_rows + 1, -1), coarse_send((size_t)std::max(n_f_send, 1), -1);
  std::vector<int> c_send_list, c_send_index((size_t)nd + 1, 0);
  sub_num[0] = -2;
  const int *send_list = fh->h_send_list.data();
  for (int d = 0; d < nd; d++) {                                                   /* :988-1061 */
    int sub_count = 0;
    for (int ii = fh->send_index[d]; ii < fh->send_index[d+1]; ii++) {
      const int jj = f_c_row[send_list[ii]] + 1;
      if (sub_num[jj] == -1) {
        sub_num[jj] = sub_count++;
        c_send_list.push_back(jj - 1);                                             /* :1188-1197 */
      }
      coarse_send[ii] = sub_num[jj];
    }
    for (int ii = fh->send_index[d]; ii < fh->send_index[d+1]; ii++)
      sub_num[f_c_row[send_list[ii]] + 1] = -1;
    sub_num[0] = -2;
    c_send_index[d+1] = (int)c_send_list.size();
  }

  /* _exchange_halo_coarsening (:802-910): ids travel in send-list order */
  {
    DevBuf<int> d_send, d_ghost;
    HaloDev v;
    if (!halo_dev_view(fh, v)) return nullptr;
    v.send_list = nullptr;
    if (!d_send.upload(coarse_send.data(), (size_t)std::max(n_f_send, 1))
        || !d_ghost.alloc((size_t)std::max(n_f_ghost, 1))) return nullptr;
    if (!pcomm_window_exchange(pc, v, 4, 1, d_send.p, d_ghost.p)) return nullptr;
    if (n_f_ghost > 0 &&
        !CSB_CUDA(cudaMemcpyAsync(f_c_row.data() + f_n_rows, d_ghost.p, sizeof(int)*(size_t)n_f_ghost,
                                  cudaMemcpyDeviceToHost, c.stream))) return nullptr;
    if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return nullptr;
  }

  /* coarse ghost numbering, section by section (:1076-1140) */
  std::vector<int> c_index((size_t)nd + 1, 0);
  int n_elts = 0;
  for (int d = 0; d < nd; d++) {
    int jj = -1;
    for (int ii = f_n_rows + fh->index[d]; ii < f_n_rows + fh->index[d+1]; ii++) {
      if (f_c_row[ii] > -1) {
        if (f_c_row[ii] > jj) jj += 1;
        f_c_row[ii] += c_n_rows + n_elts;
      }
    }
    n_elts += jj + 1;
    c_index[d+1] = n_elts;
  }

  /* drop neighbours with nothing left to exchange (:1218-1304) */
  std::vector<int> ranks, idx(1, 0), sidx(1, 0), sl;
  for (int d = 0; d < nd; d++) {
    const int n_r = c_index[d+1] - c_index[d], n_s = c_send_index[d+1] - c_send_index[d];
    if (n_r + n_s == 0) continue;
    ranks.push_back(fh->c_domain_rank[d]);
    idx.push_back(c_index[d+1]);
    sl.insert(sl.end(), c_send_list.begin() + c_send_index[d], c_send_list.begin() + c_send_index[d+1]);
    sidx.push_back((int)sl.size());
  }
  int dummy = 0;
  csb_halo_t *ch = csb_halo_create(nullptr, fh->local_rank, (int)ranks.size(),
                                   ranks.empty() ? &dummy : ranks.data(), c_n_rows, sidx.data(),
                                   sl.empty() ? &dummy : sl.data(), idx.data());
  if (ch == nullptr) return nullptr;
  if (csb_halo_set_pcomm(ch, pc) != 0) { csb_halo_destroy(&ch); return nullptr; }
  return ch;
}


/* Coarse levels merged onto rank 0: cs_grid_merge_ranks (src/alge/cs_grid_rank_merge.cpp:1359)
   as the reference triggers it (cs_grid.cpp:8355-8372), with merge_stride >= n_ranks.
   Called when the partitioned hierarchy is complete.  For every level l >= Lm the local
   matrices of all ranks are concatenated in rank order on rank 0 (rows of rank r first get
   the ids shift_l[r] + i, cs_grid_rank_merge.cpp:1420-1448; ghost columns become the merged
   ids of the rows they mirror, obtained by one exchange of the shifted ids over the level's
   halo, :1450-1460).  Every row keeps its entries in their order, so the row sums -- and
   with them every smoother sweep -- are those of the partitioned level, bit for bit; the
   aggregates are the ones built before the merge (the reference aggregates the merged grid
   again, across the former rank boundaries).  On the other ranks the hierarchy ends with a
   stub of the merge level (their rows of its rhs / vx).  Collective. */
bool mg_merge_coarse_levels(csb_multigrid_t *mg, csb_pcomm_t *pc, int Lm, bool galerkin)
{
  (void)galerkin;
  Context &c = ctx();
  const int R = pc->h.n_ranks, me = pc->h.rank;
  const int nl = (int)mg->levels.size();
  std::vector<MgLevel *> merged;

  std::vector<long long> n_next(R, 0);      /* rows of level l + 1 per rank */
  for (int l = Lm; l < nl; l++) {
    MgLevel *lv = mg->levels[l];
    csb_matrix_t *m = const_cast<csb_matrix_t *>(lv->m);
    std::vector<long long> n_of(R), nnz_of(R), ext_of(R);
    if (!pcomm_allgather_count(pc, m->n_rows, n_of.data())
        || !pcomm_allgather_count(pc, m->nnz, nnz_of.data())) return false;
    long long n_tot = 0, nnz_tot = 0;
    std::vector<int> shift(R, 0);
    for (int r = 0; r < R; r++) { shift[r] = (int)n_tot; n_tot += n_of[r]; nnz_tot += nnz_of[r]; }
    std::vector<int> shift_next(R, 0);
    if (l + 1 < nl) {
      if (!pcomm_allgather_count(pc, mg->levels[l+1]->m->n_rows, n_next.data())) return false;
      long long t = 0;
      for (int r = 0; r < R; r++) { shift_next[r] = (int)t; t += n_next[r]; }
    }

    /* merged ids of the local rows and, through the halo, of the ghost columns */
    DevBuf<int> new_id, col_m, row_len, cr_m;
    const size_t nce = (size_t)std::max(m->n_cols_ext, 1);
    if (!new_id.alloc(nce) || !col_m.alloc((size_t)std::max(m->nnz, 1))
        || !row_len.alloc((size_t)std::max(m->n_rows, 1))
        || !cr_m.alloc((size_t)std::max(m->n_rows, 1))) return false;
    if (!launch_merge_index(0, m->n_rows, shift[me], nullptr, nullptr, new_id.p)) return false;
    {
      HaloDev hv;
      memset(&hv, 0, sizeof(hv));
      if (lv->halo != nullptr && !halo_dev_view(lv->halo, hv)) return false;
      if (!pcomm_window_exchange(pc, hv, 4, 1, new_id.p, new_id.p + m->n_rows)) return false;
    }
    if (!launch_merge_index(1, m->nnz, 0, m->col_id.p, new_id.p, col_m.p)
        || !launch_merge_index(2, m->n_rows, 0, m->row_index.p, nullptr, row_len.p))
      return false;
    if (l + 1 < nl
        && !launch_merge_index(3, m->n_rows, shift_next[me], lv->coarse_row.p, nullptr, cr_m.p))
      return false;

    MgLevel *ml = nullptr;
    csb_matrix_t *cm = nullptr;
    DevBuf<int> g_len, g_cr;
    if (me == 0) {
      ml = new MgLevel;
      cm = new csb_matrix_t;
      ml->owned = cm; ml->m = cm; ml->merged = true;
      ml->n_g_rows = (double)n_tot;
      cm->n_rows = (csb_lnum_t)n_tot; cm->n_cols_ext = (csb_lnum_t)n_tot;
      cm->nnz = (csb_lnum_t)nnz_tot; cm->symmetric = true;
      merged.push_back(ml);
      if (!cm->col_id.alloc((size_t)std::max<long long>(nnz_tot, 1))
          || !cm->x_val.alloc((size_t)std::max<long long>(nnz_tot, 1))
          || !cm->d_val.alloc((size_t)std::max<long long>(n_tot, 1))
          || !cm->row_index.alloc((size_t)n_tot + 1)
          || !g_len.alloc((size_t)std::max<long long>(n_tot, 1))
          || !g_cr.alloc((size_t)std::max<long long>(n_tot, 1))) return false;
    }
    bool ok = pcomm_gather_to_root(pc, 4, n_of.data(), row_len.p, me == 0 ? g_len.p : nullptr)
           && pcomm_gather_to_root(pc, 4, nnz_of.data(), col_m.p, me == 0 ? cm->col_id.p : nullptr)
           && pcomm_gather_to_root(pc, 8, nnz_of.data(), m->x_val.p, me == 0 ? cm->x_val.p : nullptr)
           && pcomm_gather_to_root(pc, 8, n_of.data(), m->d_val.p, me == 0 ? cm->d_val.p : nullptr);
    if (ok && l + 1 < nl)
      ok = pcomm_gather_to_root(pc, 4, n_of.data(), cr_m.p, me == 0 ? g_cr.p : nullptr);
    if (!ok || !CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;

    if (me == 0) {
      std::vector<int> h_len((size_t)n_tot);
      if (n_tot > 0 && !CSB_CUDA(cudaMemcpy(h_len.data(), g_len.p, sizeof(int)*(size_t)n_tot,
                                            cudaMemcpyDeviceToHost))) return false;
      cm->h_row_index.assign((size_t)n_tot + 1, 0);
      for (long long i = 0; i < n_tot; i++) cm->h_row_index[(size_t)i+1] = cm->h_row_index[(size_t)i] + h_len[(size_t)i];
      if (cm->h_row_index[(size_t)n_tot] != nnz_tot) {
        set_error("coarse level merge: gathered row lengths do not add up");
        return false;
      }
      if (!CSB_CUDA(cudaMemcpy(cm->row_index.p, cm->h_row_index.data(), sizeof(int)*((size_t)n_tot+1),
                               cudaMemcpyHostToDevice))) return false;
      matrix_plan_spmv(cm);
      if (l + 1 < nl) {
        /* merged fine -> coarse map and its inverse; the fine rows of a coarse row stay in
           ascending order (they all belong to one rank) */
        long long nc_tot = 0;
        for (int r = 0; r < R; r++) nc_tot += n_next[r];
        ml->h_coarse_row.resize((size_t)std::max<long long>(n_tot, 1));
        if (n_tot > 0 && !CSB_CUDA(cudaMemcpy(ml->h_coarse_row.data(), g_cr.p, sizeof(int)*(size_t)n_tot,
                                              cudaMemcpyDeviceToHost))) return false;
        std::vector<int> cfi((size_t)nc_tot + 1, 0), cfr((size_t)std::max<long long>(n_tot, 1));
        for (long long i = 0; i < n_tot; i++) {
          const int ic = ml->h_coarse_row[(size_t)i];
          if (ic >= 0) cfi[(size_t)ic + 1]++;
        }
        for (long long i = 0; i < nc_tot; i++) cfi[(size_t)i+1] += cfi[(size_t)i];
        {
          std::vector<int> fill(cfi.begin(), cfi.end() - 1);
          for (long long i = 0; i < n_tot; i++) {
            const int ic = ml->h_coarse_row[(size_t)i];
            if (ic >= 0) cfr[(size_t)fill[(size_t)ic]++] = (int)i;
          }
        }
        if (!ml->coarse_row.upload(ml->h_coarse_row.data(), (size_t)std::max<long long>(n_tot, 1))
            || !ml->c_f_row_index.upload(cfi.data(), (size_t)nc_tot + 1)
            || !ml->c_f_row_ids.upload(cfr.data(), (size_t)std::max<long long>(n_tot, 1))
            || !CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
      }
    }
    if (l == Lm) {
      /* what the kernels need at the boundary */
      MergeDev &M = mg->h_merge;
      memset(&M, 0, sizeof(M));
      M.is_root = (me == 0) ? 1 : 0; M.n_ranks = R; M.rank = me;
      M.n_local = m->n_rows;
      for (int r = 0; r < R; r++) { M.shift[r] = shift[r]; M.n_of[r] = (int)n_of[r]; }
      M.timeout_ns = pc->h.timeout_ns;
    }
  }

  /* swap the levels: rank 0 continues with the merged ones, the others stop at a stub */
  if (me == 0) {
    for (int l = Lm; l < nl; l++) { delete mg->levels[l]; mg->levels[l] = merged[(size_t)(l - Lm)]; }
  }
  else {
    for (int l = Lm + 1; l < nl; l++) delete mg->levels[l];
    mg->levels.resize((size_t)Lm + 1);
    mg->levels[Lm]->stub = true;
  }
  mg->merge_level = Lm;
  return true;
}

/* peer buffers and flags of the merge boundary (after the levels' solve-time data) */
bool mg_merge_build_dev(csb_multigrid_t *mg, csb_pcomm_t *pc)
{
  Context &c = ctx();
  MergeDev &M = mg->h_merge;
  const int R = M.n_ranks, me = M.rank;
  MgLevel *lv = mg->levels[(size_t)mg->merge_level];
  if (!mg->merge_count.alloc(1)
      || !CSB_CUDA(cudaMemsetAsync(mg->merge_count.p, 0, sizeof(unsigned), c.stream))) return false;
  M.count = mg->merge_count.p;
  M.flag_in = (unsigned *)pcomm_arena_alloc(pc, sizeof(unsigned)*COMM_MAX_RANKS);
  if (M.flag_in == nullptr) return false;
  std::vector<int> nbr;
  std::vector<long long> snd, rcv;
  if (me == 0) {
    const size_t n_tot = (size_t)std::max(lv->m->n_rows, 1);
    lv->arena_rhs = (double *)pcomm_arena_alloc(pc, sizeof(double)*(n_tot + 2));
    if (lv->arena_rhs == nullptr) return false;
    for (int r = 1; r < R; r++) {
      nbr.push_back(r);
      snd.push_back(((char *)lv->arena_rhs - pc->arena) + 8ll*M.shift[r]);
      snd.push_back(((char *)M.flag_in - pc->arena) + 4ll*r);
      for (int k = 2; k < 8; k++) snd.push_back(0);
    }
  }
  else {
    lv->arena_vx = (double *)pcomm_arena_alloc(pc, sizeof(double)*((size_t)std::max(M.n_local, 1) + 2));
    if (lv->arena_vx == nullptr) return false;
    nbr.push_back(0);
    snd.push_back((char *)lv->arena_vx - pc->arena);
    snd.push_back((char *)M.flag_in - pc->arena);
    for (int k = 2; k < 8; k++) snd.push_back(0);
  }
  rcv.assign(snd.size() + 8, 0);
  if (!pcomm_mailbox(pc, (int)nbr.size(), nbr.data(), snd.data(), rcv.data())) return false;
  if (me == 0) {
    for (int r = 1; r < R; r++) {
      char *base = pc->h.peer[r];
      M.peer_vx[r] = (double *)(base + rcv[(size_t)(r-1)*8 + 0]);
      M.peer_flag[r] = (unsigned *)(base + rcv[(size_t)(r-1)*8 + 1]);
    }
  }
  else {
    char *base = pc->h.peer[0];
    M.root_rhs_dst = (double *)(base + rcv[0]);
    M.root_flag = (unsigned *)(base + rcv[1]);
  }
  return true;
}
} // anonymous namespace

extern "C" {

csb_multigrid_t *csb_multigrid_create(csb_multigrid_type_t mg_type)
{
  if (!ctx().ready) { set_error("csb_multigrid_create: csb_init not called"); return nullptr; }
  if (mg_type != CSB_MULTIGRID_V_CYCLE && mg_type != CSB_MULTIGRID_K_CYCLE) {
    set_error("csb_multigrid_create: V-cycle and K-cycle are available on this path");
    return nullptr;
  }
  csb_multigrid_t *mg = new csb_multigrid_t;
  mg->type = mg_type;
  if (mg_type == CSB_MULTIGRID_K_CYCLE) {
    /* cs_multigrid_create :4718-4725: pairwise aggregation up to 4 rows, Galerkin coarse
       matrices, at most 10 levels of at least 256 rows, second inner iteration skipped when
       the first one reduced the residual by 0.25; :4777-4788 with the device substitution
       of :5083-5087: symmetric Gauss-Seidel (1 sweep) -> Jacobi (3 sweeps), PCG (50) */
    mg->coarsening_type = 3;
    mg->aggregation_limit = 4;
    mg->n_levels_max = 10;
    mg->n_g_rows_min = 256;
    mg->p0p1_relax = 0.;
    mg->k_cycle_threshold = 0.25;
    csb_multigrid_set_solver_options(mg, CSB_SLES_JACOBI, CSB_SLES_JACOBI, CSB_SLES_PCG,
                                     100, 3, 3, 50, -1, -1, 0, -1., -1., 1.);
  }
  return mg;
}

void
csb_multigrid_set_solver_options(csb_multigrid_t *mg,
                                 csb_sles_it_type_t descent_smoother_type,
                                 csb_sles_it_type_t ascent_smoother_type,
                                 csb_sles_it_type_t coarse_solver_type,
                                 int n_max_cycles, int n_max_iter_descent,
                                 int n_max_iter_ascent, int n_max_iter_coarse,
                                 int poly_degree_descent, int poly_degree_ascent,
                                 int poly_degree_coarse, double precision_mult_descent,
                                 double precision_mult_ascent, double precision_mult_coarse)
{
  if (mg == nullptr) return;
  mg->solver_type[0] = descent_smoother_type;
  mg->solver_type[1] = ascent_smoother_type;
  mg->solver_type[2] = coarse_solver_type;
  mg->n_max_cycles = n_max_cycles;
  mg->n_max_iter[0] = n_max_iter_descent;
  mg->n_max_iter[1] = n_max_iter_ascent;
  mg->n_max_iter[2] = n_max_iter_coarse;
  mg->poly_degree[0] = poly_degree_descent;
  mg->poly_degree[1] = poly_degree_ascent;
  mg->poly_degree[2] = poly_degree_coarse;
  mg->precision_mult[0] = precision_mult_descent;
  mg->precision_mult[1] = precision_mult_ascent;
  mg->precision_mult[2] = precision_mult_coarse;
}

void
csb_multigrid_set_coarsening_options(csb_multigrid_t *mg, int aggregation_limit,
                                     int coarsening_type, int n_max_levels,
                                     unsigned long long min_g_rows, double p0p1_relax,
                                     int postprocess_block_size)
{
  (void)postprocess_block_size;
  if (mg == nullptr) return;
  mg->aggregation_limit = aggregation_limit;
  mg->coarsening_type = coarsening_type;
  mg->n_levels_max = n_max_levels;
  mg->n_g_rows_min = min_g_rows;
  mg->p0p1_relax = p0p1_relax;
}

void csb_multigrid_set_aggregation_blocks(csb_multigrid_t *mg, int n_blocks)
{
  if (mg != nullptr) mg->n_blocks = (n_blocks > 0) ? n_blocks : 0;
}

void csb_multigrid_set_merge_options(csb_multigrid_t *mg, int rank_stride,
                                     int rows_mean_threshold,
                                     unsigned long long rows_glob_threshold)
{
  if (mg == nullptr) return;
  mg->merge_stride = rank_stride;
  mg->merge_mean_threshold = rows_mean_threshold;
  mg->merge_glob_threshold = rows_glob_threshold;
}

int csb_multigrid_get_merge_level(const csb_multigrid_t *mg)
{
  return (mg != nullptr) ? mg->merge_level : -1;
}

int csb_multigrid_set_reduce_comm(csb_multigrid_t *mg, void *nccl_comm)
{
  if (mg == nullptr) return -1;
  mg->nccl_comm = nccl_comm;
  return 0;
}

int csb_multigrid_set_pcomm(csb_multigrid_t *mg, csb_pcomm_t *comm)
{
  if (mg == nullptr) return -1;
  mg->pcomm = comm;
  return 0;
}

void csb_multigrid_free(void *context)
{
  csb_multigrid_t *mg = static_cast<csb_multigrid_t *>(context);
  if (mg != nullptr) mg_clear_hierarchy(mg);
}

void csb_multigrid_destroy(void **context)
{
  if (context == nullptr || *context == nullptr) return;
  delete static_cast<csb_multigrid_t *>(*context);
  *context = nullptr;
}

/* Hierarchy construction: cs_multigrid_setup -> _setup_hierarchy ->
   cs_grid_coarsen (cs_multigrid.cpp:5275, :2407; cs_grid.cpp:7980) on the
   reference's live path for a symmetric MSR matrix with mesh association
   ("msr_gather", cs_grid.cpp:8019-8022). */
void csb_multigrid_setup(void *context, const char *name, const csb_matrix_t *a,
                         int verbosity)
{
  (void)name;
  csb_multigrid_t *mg = static_cast<csb_multigrid_t *>(context);
  Context &c = ctx();
  if (mg == nullptr || a == nullptr) { set_error("csb_multigrid_setup: null argument"); return; }
  /* whatever an earlier (possibly failed) setup left behind goes first; any early return
     below leaves the object empty again (levels, arena mark), never half built */
  mg_clear_hierarchy(mg);
  struct SetupGuard {
    csb_multigrid_t *mg;
    ~SetupGuard() { if (!mg->setup_done) mg_clear_hierarchy(mg); }
  } setup_guard{mg};

  /* p0p1_relax <= 0: Galerkin coarse matrices (_compute_coarse_quantities_msr,
     cs_grid.cpp:8222-8223), which need no mesh quantities */
  const bool galerkin = (mg->p0p1_relax <= 0.);
  if (!a->symmetric || (!a->has_mesh && !galerkin)) {
    set_error("csb_multigrid_setup: this path needs a symmetric MSR matrix, with mesh "
              "association (cs_matrix_set_mesh_association) unless p0p1_relax <= 0, as the "
              "reference's pressure solve provides");
    return;
  }
  /* cs_grid_coarsening_t (cs_grid.h): 0 default (= SPD_DX for a symmetric matrix), 1 SPD_DX,
     3 SPD_PW (pairwise aggregation, with Galerkin coarse matrices on one rank here) */
  const bool pairwise = (mg->coarsening_type == 3);
  if (mg->coarsening_type != 0 && mg->coarsening_type != 1 && !pairwise) {
    set_error("csb_multigrid_setup: only CS_GRID_COARSENING_DEFAULT / SPD_DX / SPD_PW are "
              "available");
    return;
  }
  /* the reference confines its greedy aggregation to one block of rows per OpenMP thread
     (cs_grid.cpp:4374-4393): unless csb_multigrid_set_aggregation_blocks fixed a count,
     follow the thread count of this process (OMP_NUM_THREADS), as it does; pairwise
     aggregation is the reference's single-thread pairing */
  const bool pairwise_req = (mg->coarsening_type == 3);
  const int n_blocks = (mg->n_blocks > 0) ? mg->n_blocks
                     : (pairwise_req ? 1 : std::max(1, omp_get_max_threads()));
  if (pairwise && (!galerkin || a->halo != nullptr || n_blocks > 1)) {
    set_error("csb_multigrid_setup: CS_GRID_COARSENING_SPD_PW is available with Galerkin "
              "coarse matrices (p0p1_relax <= 0) on one rank, one aggregation block (the "
              "reference's single-thread pairing)");
    return;
  }
  /* smoothers: cs_sles_it PCG + Jacobi (cs_multigrid_create defaults), or -- in the
     kernel-per-phase forms (single cycle, or solver mode), precision_mult < 0 -- the
     Jacobi-family cs_multigrid_smoother types the reference's device settings select
     (cs_multigrid_set_solver_options, cs_multigrid.cpp:5070-5090); coarse solver: PCG */
  bool any_l1 = false;
  for (int i = 0; i < 3; i++) {
    const bool pcg = (mg->solver_type[i] == CSB_SLES_PCG && mg->poly_degree[i] == 0);
    const bool smoother = (i < 2 && is_smoother_type((int)mg->solver_type[i])
                           && mg->precision_mult[i] < 0);
    if (!pcg && !smoother) {
      set_error("csb_multigrid_setup: smoothers must be PCG + Jacobi, or (precision_mult < 0) "
                "Jacobi / L1 / relaxed / RJ2 / RJ3 smoothers; the coarse solver PCG + Jacobi");
      return;
    }
    if (mg->solver_type[i] == CSB_SLES_L1_JACOBI) any_l1 = true;
  }
  const bool kcycle = (mg->type == CSB_MULTIGRID_K_CYCLE);
  if (kcycle && (mg->n_max_cycles != 1 || a->halo != nullptr
                 || !is_smoother_type((int)mg->solver_type[0])
                 || !is_smoother_type((int)mg->solver_type[1]))) {
    set_error("csb_multigrid_setup: the K-cycle is available as a single cycle (n_max_cycles "
              "1: the preconditioner form, csb_multigrid_pc_create) with Jacobi-family "
              "smoothers on one rank");
    return;
  }
  const bool dist = (a->halo != nullptr);
  /* matrix mapped from native arrays (cs_grid_create_from_shared: use_faces, use_xa,
     cs_grid.cpp:7485-7491): coarse values by the face-order sums of
     _compute_coarse_quantities_native; with one aggregation block (the reference running
     one thread) the aggregation is the face-list one, _automatic_aggregation_fc
     (cs_grid.cpp:8106-8122); with Galerkin values only level 0 still has faces (:8055) */
  const bool native = (a->from_native && !galerkin);
  if (a->from_native && (dist || pairwise || a->n_rows != a->n_cols_ext)) {
    set_error("csb_multigrid_setup: a matrix created from native arrays is handled on one "
              "rank, with the DX aggregation");
    return;
  }
  csb_pcomm_t *pc = dist ? mg->pcomm : nullptr;
  if (dist && (pc == nullptr || a->halo->pcomm != pc)) {
    set_error("csb_multigrid_setup: a matrix with a halo needs csb_multigrid_set_pcomm and "
              "a halo attached to the same communicator (csb_halo_set_pcomm)");
    return;
  }
  if (dist) {
    mg->arena_mark = pcomm_arena_mark(pc);
    mg->arena_used = true;
  }

  const double t_start = wall();
  double t_aggr = 0, t_asm = 0, t_xfer = 0;
  SetupTrace tr;

  MgLevel *l0 = new MgLevel;
  l0->m = a;
  l0->halo = a->halo;
  mg->levels.push_back(l0);
  if (a->db != 1) {
    set_error("csb_multigrid_setup: block matrices (db %d) are not handled by the multigrid "
              "path yet", a->db);
    return;
  }
  mg->a = a;

  /* global row counts drive the stop criteria (cs_grid.cpp:1364-1374) */
  double g_rows = (double)a->n_rows;
  if (dist && !pcomm_allreduce(pc, &g_rows, 1)) return;
  l0->n_g_rows = g_rows;
  unsigned long long n_g_rows = (unsigned long long)g_rows, n_g_rows_prev;

  while (true) {
    n_g_rows_prev = n_g_rows;
    if ((int)mg->levels.size() >= mg->n_levels_max) break;

    MgLevel *f = mg->levels.back();
    csb_matrix_t *fm = const_cast<csb_matrix_t *>(f->m);
    const int f_level = (int)mg->levels.size() - 1;

    /* faces of this level when the hierarchy follows a native matrix */
    const bool f_faces = a->from_native && (native || f_level == 0);
    const bool use_fc = f_faces && n_blocks == 1;
    const int f_n_faces = (f_level == 0) ? fm->n_faces : f->n_faces;
    const int *f_h_face_cell = (f_level == 0) ? fm->h_face_cell.data() : f->h_face_cell.data();
    const double *f_h_xa = (f_level == 0) ? fm->h_xa.data() : f->h_xa.data();

    double t0 = wall();
    if (use_fc ? !matrix_host_mirror_structure(fm) : !matrix_host_mirror(fm)) return;
    t_xfer += wall() - t0;
    tr.mark("host mirror", f_level);

    /* aggregation on the host */
    t0 = wall();
    f->h_coarse_row.assign((size_t)std::max(fm->n_cols_ext, 1), -1);
    HostMsr pw;      /* pairwise: coarse values come out of the host recursion */
    const int c_n_rows = pairwise
      ? coarsen_pairwise_galerkin(f_level, fm->n_rows, fm->h_row_index.data(),
                                  fm->host_col_id(), fm->host_d_val(), fm->host_x_val(),
                                  mg->aggregation_limit, f->h_coarse_row.data(), pw)
      : use_fc
      ? aggregate_fc(f_level, fm->n_rows, fm->n_cols_ext, f_n_faces, f_h_face_cell,
                     fm->h_d_val.data(), f_h_xa, mg->aggregation_limit,
                     f->h_coarse_row.data())
      : aggregate_dx_msr(f_level, fm->n_rows, fm->h_row_index.data(), fm->host_col_id(),
                         fm->host_d_val(), fm->host_x_val(), mg->aggregation_limit,
                         n_blocks, f->h_coarse_row.data());
    t_aggr += wall() - t0;
    tr.mark("aggregation", f_level);

    /* stop criteria of _setup_hierarchy (cs_multigrid.cpp:2518-2529); they
       only depend on the (global) row count, so the rejected level is never
       assembled */
    g_rows = (double)c_n_rows;
    if (dist && !pcomm_allreduce(pc, &g_rows, 1)) return;
    n_g_rows = (unsigned long long)g_rows;
    bool add = true;
    if (n_g_rows == 0) add = false;
    else if (mg->levels.size() > 1) {
      if (n_g_rows < mg->n_g_rows_min || (double)n_g_rows > (0.8 * (double)n_g_rows_prev))
        add = false;
    }
    if (!add) { f->h_coarse_row.clear(); break; }

    /* coarse halo and coarse ids of the ghost rows */
    csb_halo_t *ch = nullptr;
    if (dist) {
      t0 = wall();
      ch = coarsen_halo(pc, f->halo, fm->n_rows, c_n_rows, f->h_coarse_row);
      if (ch == nullptr) return;
      t_xfer += wall() - t0;
    }
    const int c_n_ghost = (ch != nullptr) ? ch->n_ghost : 0;
    const size_t c_n_cols_ext = (size_t)c_n_rows + (size_t)c_n_ghost;

    t0 = wall();
    CoarseStructure cs;
    build_coarse_structure(fm->n_rows, fm->h_row_index.data(), fm->host_col_id(),
                           f->h_coarse_row.data(), c_n_rows, cs);
    /* the restriction as a 0/1 CSR matrix (coarse rows x fine rows): same slab pipeline as the
       matrix products in the persistent kernel */
    if (restrict_ring_enabled())
      plan_slabs(cs.c_f_row_index, c_n_rows, f->r_slab_rows, f->r_n_slabs, f->r_max_entries);
    HostVec<int> c_h_face_cell;
    if (native)
      number_coarse_faces_native(f_n_faces, f_h_face_cell, f->h_coarse_row.data(), c_n_rows,
                                 cs, c_h_face_cell);
    t_aggr += wall() - t0;
    tr.mark("coarse structure (host)", f_level);
    const int c_nnz = cs.c_row_index[c_n_rows];

    t0 = wall();
    MgLevel *cl = new MgLevel;
    csb_matrix_t *cm = new csb_matrix_t;
    cl->owned = cm; cl->m = cm;
    cl->halo = ch; cl->halo_owned = (ch != nullptr);
    cl->n_g_rows = g_rows;
    cm->halo = ch;
    cm->n_rows = c_n_rows; cm->n_cols_ext = (csb_lnum_t)c_n_cols_ext; cm->nnz = c_nnz;
    cm->symmetric = true;
    bool ok = f->coarse_row.upload(f->h_coarse_row.data(), (size_t)fm->n_cols_ext)
           && f->c_f_row_index.upload(cs.c_f_row_index.data(), (size_t)c_n_rows + 1)
           && f->c_f_row_ids.upload(cs.c_f_row_ids.data(), (size_t)fm->n_rows)
           && cm->row_index.upload(cs.c_row_index.data(), (size_t)c_n_rows + 1)
           && cm->col_id.upload(cs.c_col_id.data(), (size_t)c_nnz)
           && cm->d_val.alloc((size_t)c_n_rows) && cm->x_val.alloc((size_t)c_nnz);
    if (ok && !galerkin)
      ok = cl->cell_face.upload(cs.c_cell_face.data(), (size_t)c_nnz)
           && cl->cell_face_sgn.upload(cs.c_cell_face_sgn.data(), (size_t)c_nnz)
           && cl->cell_cen.alloc(3*c_n_cols_ext) && cl->cell_vol.alloc(c_n_cols_ext)
           && cl->face_normal.alloc(3*(size_t)cs.c_n_faces)
           && cl->xa0.alloc((size_t)cs.c_n_faces) && cl->xa0ij.alloc(3*(size_t)cs.c_n_faces);
    if (ok && native)
      ok = cl->face_cell.upload(c_h_face_cell.data(), 2*(size_t)cs.c_n_faces)
           && cl->xa.alloc((size_t)cs.c_n_faces);
    ok = ok && CSB_CUDA(cudaStreamSynchronize(c.stream));  /* cs.* vectors die below */
    if (!ok) { delete cl; return; }
    cl->n_faces = cs.c_n_faces;
    if (native) cl->h_face_cell.swap(c_h_face_cell);
    cm->h_row_index.swap(cs.c_row_index);
    cm->h_col_id.swap(cs.c_col_id);
    t_xfer += wall() - t0;
    tr.mark("alloc + upload", f_level);

    /* floating-point assembly on the device */
    t0 = wall();
    CoarseAssemblyArgs ca;
    ca.f_level = f_level;
    ca.f_n_rows = fm->n_rows; ca.c_n_rows = c_n_rows; ca.c_n_faces = cl->n_faces;
    ca.f_row_index = fm->row_index.p; ca.f_col_id = fm->col_id.p;
    ca.f_d_val = fm->d_val.p; ca.f_x_val = fm->x_val.p;
    ca.f_c_row = f->coarse_row.p;
    ca.c_f_row_index = f->c_f_row_index.p; ca.c_f_row_ids = f->c_f_row_ids.p;
    ca.c_row_index = cm->row_index.p; ca.c_col_id = cm->col_id.p;
    ca.c_d_val = cm->d_val.p; ca.c_x_val = cm->x_val.p;
    if (pairwise) {
      /* same structure (sorted unique coarse neighbours) whichever way it is built; the
         values are those of the recursion (sums of the intermediate matrix's entries) */
      if ((int)pw.row_index.size() != c_n_rows + 1 || pw.row_index[c_n_rows] != c_nnz
          || pw.col_id.size() < (size_t)c_nnz
          || !std::equal(cm->h_col_id.begin(), cm->h_col_id.begin() + c_nnz, pw.col_id.begin())) {
        set_error("csb_multigrid_setup: pairwise recursion and projected structure differ");
        delete cl; return;
      }
      if (!CSB_CUDA(cudaMemcpyAsync(cm->d_val.p, pw.d_val.data(), sizeof(double)*(size_t)c_n_rows,
                                    cudaMemcpyHostToDevice, c.stream))
          || !CSB_CUDA(cudaMemcpyAsync(cm->x_val.p, pw.x_val.data(), sizeof(double)*(size_t)c_nnz,
                                       cudaMemcpyHostToDevice, c.stream))
          || !CSB_CUDA(cudaStreamSynchronize(c.stream))) { delete cl; return; }
      cm->h_d_val.swap(pw.d_val);
      cm->h_x_val.swap(pw.x_val);
      cm->h_x_val.resize((size_t)c_nnz);
      cm->h_d_val.resize((size_t)c_n_rows);
      cm->h_vals_valid = true;
      t_asm += wall() - t0;
      matrix_plan_spmv(cm);
      mg->levels.push_back(cl);
      continue;
    }
    if (galerkin) {
      if (!launch_coarse_galerkin(ca)) { delete cl; return; }
      if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) { delete cl; return; }
      t_asm += wall() - t0;
      matrix_plan_spmv(cm);
      mg->levels.push_back(cl);
      continue;
    }
    if (native) {
      /* face-order sums (_compute_coarse_cell_quantities, _compute_coarse_quantities_native,
         _matrix_from_native), then the host copies the next face-list aggregation reads */
      NativeCoarseArgs na;
      na.f_level = f_level;
      na.f_n_rows = fm->n_rows; na.c_n_rows = c_n_rows; na.c_n_faces = cl->n_faces;
      na.f_row_index = fm->row_index.p; na.f_col_id = fm->col_id.p;
      na.f_d_val = fm->d_val.p; na.f_x_val = fm->x_val.p;
      na.f_c_row = f->coarse_row.p;
      na.c_f_row_index = f->c_f_row_index.p; na.c_f_row_ids = f->c_f_row_ids.p;
      na.c_row_index = cm->row_index.p; na.c_cell_face = cl->cell_face.p;
      na.c_face_cell = cl->face_cell.p;
      if (f_level == 0) {
        na.f_cell_face = fm->c2f.p; na.f_face_cell = fm->face_cell.p;
        na.f_cell_cen = fm->cell_cen.p; na.f_cell_vol = fm->cell_vol.p;
        na.f_face_u_normal = fm->face_u_normal.p; na.f_face_surf = fm->face_surf.p;
        na.f_face_normal = nullptr; na.f_xa0 = nullptr; na.f_xa0ij = nullptr;
      }
      else {
        na.f_cell_face = f->cell_face.p; na.f_face_cell = f->face_cell.p;
        na.f_cell_cen = f->cell_cen.p; na.f_cell_vol = f->cell_vol.p;
        na.f_face_u_normal = nullptr; na.f_face_surf = nullptr;
        na.f_face_normal = f->face_normal.p; na.f_xa0 = f->xa0.p; na.f_xa0ij = f->xa0ij.p;
      }
      na.c_cell_cen = cl->cell_cen.p; na.c_cell_vol = cl->cell_vol.p;
      na.c_face_normal = cl->face_normal.p; na.c_xa0 = cl->xa0.p; na.c_xa0ij = cl->xa0ij.p;
      na.c_xa = cl->xa.p;
      na.c_d_val = cm->d_val.p; na.c_x_val = cm->x_val.p;
      na.relax = mg->p0p1_relax;
      DevBuf<double> w1;
      if (!w1.alloc((size_t)fm->n_rows)) { delete cl; return; }
      na.w1 = w1.p;
      cl->h_xa.resize((size_t)std::max(cl->n_faces, 1));
      cm->h_d_val.resize((size_t)std::max(c_n_rows, 1));
      if (!launch_coarse_native(na)
          || (cl->n_faces > 0
              && !CSB_CUDA(cudaMemcpyAsync(cl->h_xa.data(), cl->xa.p,
                                           sizeof(double)*(size_t)cl->n_faces,
                                           cudaMemcpyDeviceToHost, c.stream)))
          || !CSB_CUDA(cudaMemcpyAsync(cm->h_d_val.data(), cm->d_val.p,
                                       sizeof(double)*(size_t)c_n_rows,
                                       cudaMemcpyDeviceToHost, c.stream))
          || !CSB_CUDA(cudaStreamSynchronize(c.stream))) { delete cl; return; }
      t_asm += wall() - t0;
      tr.mark("coarse values (device, face order)", f_level);
      matrix_plan_spmv(cm);
      mg->levels.push_back(cl);
      continue;
    }
    ca.c_cell_face = cl->cell_face.p; ca.c_cell_face_sgn = cl->cell_face_sgn.p;
    if (f_level == 0) {
      ca.f_cell_face = fm->c2f.p; ca.f_cell_face_sgn = fm->c2f_sgn.p;
      ca.f_cell_cen = fm->cell_cen.p; ca.f_cell_vol = fm->cell_vol.p;
      ca.f_face_u_normal = fm->face_u_normal.p; ca.f_face_surf = fm->face_surf.p;
      ca.f_face_normal = nullptr; ca.f_xa0 = nullptr; ca.f_xa0ij = nullptr;
    }
    else {
      ca.f_cell_face = f->cell_face.p; ca.f_cell_face_sgn = f->cell_face_sgn.p;
      ca.f_cell_cen = f->cell_cen.p; ca.f_cell_vol = f->cell_vol.p;
      ca.f_face_u_normal = nullptr; ca.f_face_surf = nullptr;
      ca.f_face_normal = f->face_normal.p; ca.f_xa0 = f->xa0.p; ca.f_xa0ij = f->xa0ij.p;
    }
    ca.c_d_val = cm->d_val.p; ca.c_x_val = cm->x_val.p;
    ca.c_cell_cen = cl->cell_cen.p; ca.c_cell_vol = cl->cell_vol.p;
    ca.c_face_normal = cl->face_normal.p; ca.c_xa0 = cl->xa0.p; ca.c_xa0ij = cl->xa0ij.p;
    ca.relax = mg->p0p1_relax;
    if (!launch_coarse_assembly(ca, 1)) { delete cl; return; }
    if (dist) {
      /* ghost cell centres and volumes (cs_grid.cpp:6561-6581) */
      HaloDev hv;
      if (!halo_dev_view(ch, hv)
          || !pcomm_window_exchange(pc, hv, 8, 3, cl->cell_cen.p, cl->cell_cen.p + 3*(size_t)c_n_rows)
          || !pcomm_window_exchange(pc, hv, 8, 1, cl->cell_vol.p, cl->cell_vol.p + c_n_rows)) {
        delete cl; return;
      }
    }
    if (!launch_coarse_assembly(ca, 2)) { delete cl; return; }
    if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) { delete cl; return; }
    t_asm += wall() - t0;
    tr.mark("coarse values (device)", f_level);

    matrix_plan_spmv(cm);
    mg->levels.push_back(cl);
    tr.mark("plan spmv", f_level);

    if (verbosity > 2)
      printf("   building level %2d grid\n     number of rows:      %10d\n"
             "     number of entries:   %10d\n",
             (int)mg->levels.size() - 1, c_n_rows, c_nnz + c_n_rows);
  }

  /* coarse levels merged onto rank 0 (cs_grid.cpp:8355-8372: mean rows per rank under
     the mean threshold, or global rows under the global one) */
  mg->merge_level = -1;
  if (dist && mg->merge_stride > 1) {
    const int R = pc->h.n_ranks;
    if (mg->merge_stride < R) {
      set_error("csb_multigrid_setup: merge rank_stride %d: only a merge of all %d ranks "
                "onto rank 0 (rank_stride >= number of ranks) is available", mg->merge_stride, R);
      return;
    }
    int Lm = -1;
    for (int l = 1; l < (int)mg->levels.size() && Lm < 0; l++) {
      const double g = mg->levels[l]->n_g_rows;
      if (g / R < (double)mg->merge_mean_threshold || g < (double)mg->merge_glob_threshold)
        Lm = l;
    }
    if (Lm > 0 && R > 1) {
      const double t0 = wall();
      if (!mg_merge_coarse_levels(mg, pc, Lm, galerkin)) return;
      t_xfer += wall() - t0;
    }
  }

  /* solve-time data */
  const int n_levels = (int)mg->levels.size();
  size_t wr_size = 1, aux_size = 1;
  for (int l = 0; l < n_levels; l++) {
    MgLevel *lv = mg->levels[l];
    const csb_matrix_t *m = lv->m;
    wr_size = std::max(wr_size, (size_t)m->n_cols_ext);
    aux_size = std::max(aux_size, 5*work_stride_of((size_t)m->n_cols_ext));
    if (!lv->ad_inv.alloc((size_t)m->n_rows)) return;
    if (!launch_ad_inv(m->n_rows, m->d_val.p, lv->ad_inv.p)) return;
    if (any_l1) {
      if (!lv->l1_inv.alloc((size_t)m->n_rows) || !launch_l1_inv(m, lv->l1_inv.p)) return;
    }
    if (kcycle) {
      lv->kstride = work_stride_of((size_t)m->n_cols_ext);
      if (!lv->kwork.alloc(10*lv->kstride)) return;
      if (l == 0 && !mg->kc_scal.alloc(8*(size_t)n_levels)) return;
    }
    if (l > 0) {
      if (!lv->rhs.alloc((size_t)m->n_cols_ext) || !lv->vx.alloc((size_t)m->n_cols_ext)) return;
    }
    if (dist && !lv->merged && !lv->stub
        && (!matrix_host_mirror_structure(const_cast<csb_matrix_t *>(m))
            || !build_dist_level(pc, lv->halo, m, lv->dist))) return;
  }
  tr.mark("per-level solve data");
  mg->work_stride = aux_size / 5;
  if (mg->merge_level >= 0 && !mg_merge_build_dev(mg, pc)) return;
  if (!mg->wr.alloc(wr_size) || !mg->aux.alloc(aux_size)) return;
  if (!mg->it_states.alloc(2*(size_t)n_levels)) return;
  {
    std::vector<MgLevelDesc> h_desc((size_t)n_levels);
    size_t max_entries = 0;
    for (int l = 0; l < n_levels; l++) {
      MgLevelDesc &d = h_desc[l];
      MgLevelDesc zero; memset(&zero, 0, sizeof(zero)); d = zero;
      MgLevel *lv = mg->levels[l];
      const csb_matrix_t *m = lv->m;
      size_t me = 0;
      plan_slabs(m->h_row_index, m->n_rows, d.slab_rows, d.n_slabs, me);
      d.gather_depth = plan_gather_depth(m->h_row_index, m->n_rows, &d.row_len_max);
      max_entries = std::max(max_entries, me);
      max_entries = std::max(max_entries, lv->r_max_entries);
      d.r_slab_rows = lv->r_slab_rows; d.r_n_slabs = lv->r_n_slabs;
      d.n_rows = m->n_rows;
      /* rows of the next level THIS rank restricts to: at the merge boundary its own part
         of the merged level */
      d.n_rows_next = (l + 1 < n_levels) ? mg->levels[l+1]->m->n_rows : 0;
      if (l + 1 == mg->merge_level) d.n_rows_next = mg->h_merge.n_local;
      d.level_id = l;
      d.row_index = m->row_index.p; d.col_id = m->col_id.p;
      d.d_val = m->d_val.p; d.x_val = m->x_val.p; d.ad_inv = lv->ad_inv.p;
      d.l1_inv = lv->l1_inv.p;
      d.rhs = (lv->arena_rhs != nullptr) ? lv->arena_rhs : lv->rhs.p;
      d.vx = (lv->arena_vx != nullptr) ? lv->arena_vx : lv->vx.p;
      d.coarse_row = lv->coarse_row.p;
      d.c_f_row_index = lv->c_f_row_index.p; d.c_f_row_ids = lv->c_f_row_ids.p;
      d.st_descent = mg->it_states.p + 2*l;
      d.st_ascent = mg->it_states.p + 2*l + 1;
      d.n_g_rows = dist ? lv->n_g_rows : (double)m->n_rows;
      d.dist = (dist && !lv->merged && !lv->stub) ? lv->dist.d.p : nullptr;
    }
    if (mg->merge_level >= 0) {
      MgLevel *ml = mg->levels[(size_t)mg->merge_level];
      mg->h_merge.stub_rhs = ml->rhs.p;
      mg->h_merge.stub_vx = ml->arena_vx;
      if (!mg->d_merge.upload(&mg->h_merge, 1)) return;
    }
    tr.mark("slab plans");
    mg->p_smem = persistent_smem_bytes(max_entries);
    mg->p_grid = persistent_grid(1, mg->p_smem);
    if (mg->p_grid < 1) return;
    if (!mg->descs.upload(h_desc.data(), (size_t)n_levels)) return;
    if (!mg->partials.alloc(12*(size_t)mg->p_grid) || !mg->barrier.alloc(2)) return;
    if (!CSB_CUDA(cudaMemsetAsync(mg->barrier.p, 0, 2*sizeof(unsigned), c.stream))) return;
    if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return;   /* h_desc dies here */
  }
  if (!mg->mg_state.p && !mg->mg_state.alloc(1)) return;
  if (mg->h_mg == nullptr)
    if (!CSB_CUDA(cudaMallocHost((void **)&mg->h_mg, sizeof(MgState)))) return;
  if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return;

  tr.mark("descriptors, state");
  mg->setup_done = true;
  mg->n_setups += 1;
  mg->t_setup[0] = wall() - t_start;
  mg->t_setup[1] = t_aggr; mg->t_setup[2] = t_asm; mg->t_setup[3] = t_xfer;
  mg->t_setup_tot += mg->t_setup[0];
}

/* cs_multigrid_solve (cs_multigrid.cpp:5362-5588) driving _multigrid_v_cycle
   (:3127-3542). One host sync per cycle (to read the gate); everything else
   is enqueued ahead. */
csb_sles_convergence_state_t
csb_multigrid_solve(void *context, const char *name, const csb_matrix_t *a, int verbosity,
                    double precision, double r_norm, int *n_iter, double *residual,
                    const csb_real_t *rhs, csb_real_t *vx_ini, csb_real_t *vx,
                    size_t aux_size, void *aux_vectors)
{
  (void)aux_size; (void)aux_vectors;
  csb_multigrid_t *mg = static_cast<csb_multigrid_t *>(context);
  Context &c = ctx();
  c.err.clear();
  if (n_iter) *n_iter = 0;
  if (residual) *residual = -1.;
  if (mg == nullptr || a == nullptr || rhs == nullptr || vx == nullptr) {
    set_error("csb_multigrid_solve: null argument");
    return CSB_SLES_BREAKDOWN;
  }
  if (vx_ini != nullptr && vx_ini != vx) {
    set_error("csb_multigrid_solve: vx_ini must be vx or NULL");
    return CSB_SLES_BREAKDOWN;
  }
  if (!mg->setup_done || mg->a != a)
    csb_multigrid_setup(mg, name, a, verbosity);
  if (!mg->setup_done) return CSB_SLES_BREAKDOWN;

  const double t0 = wall();
  const int n_levels = (int)mg->levels.size();
  const int coarsest = n_levels - 1;
  const double denom = 1.0 / (double)a->n_rows;   /* kernel-per-phase path: one rank only */

  StagedVec s_rhs, s_vx;
  if (!s_rhs.init(const_cast<double *>(rhs), (size_t)a->n_rows, true, mg->stage_rhs))
    return CSB_SLES_BREAKDOWN;
  if (!s_vx.init(vx, (size_t)a->n_cols_ext, vx_ini == vx, mg->stage_vx))
    return CSB_SLES_BREAKDOWN;

  MgState *ms = mg->mg_state.p;
  if (!launch_mg_state_init(ms, precision, r_norm, mg->n_max_cycles)) return CSB_SLES_BREAKDOWN;

  if (mg->n_max_cycles == 1 && a->halo == nullptr) {
    /* cs_multigrid_solve :5484-5499: a single cycle is the form without convergence
       tests (_multigrid_v_cycle_pc); the reference returns CS_SLES_ITERATING from it */
    if (!mg_cycle_pc_enqueue(mg, precision, r_norm, s_rhs.d, s_vx.d, vx_ini == vx))
      return CSB_SLES_BREAKDOWN;
    if (!CSB_CUDA(cudaMemcpyAsync(mg->h_mg, ms, MG_HEAD, cudaMemcpyDeviceToHost, c.stream)))
      return CSB_SLES_BREAKDOWN;
    if (!s_vx.copy_back((size_t)a->n_rows)) return CSB_SLES_BREAKDOWN;
    if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return CSB_SLES_BREAKDOWN;
    if (n_iter) *n_iter = mg->h_mg->n_equiv_iter;
    mg->n_cycles_last = 1;
    mg->n_cycles_tot += 1;
    mg->n_solves += 1;
    mg->t_solve_tot += wall() - t0;
    return (mg->h_mg->cvg < CSB_SLES_BREAKDOWN) ? (csb_sles_convergence_state_t)mg->h_mg->cvg
                                                 : CSB_SLES_ITERATING;
  }

  int cvg = CSB_SLES_ITERATING;
  int n_cycles = 0;
  double one_level_residual = -1.;
  const bool dist = (a->halo != nullptr);
  /* cs_multigrid_smoother smoothers run kernel per phase (the persistent kernel holds the
     PCG-smoothed cycle); a smoother reports no initial residual (:3270-3276) */
  const bool sm_descent = is_smoother_type((int)mg->solver_type[0]);
  const bool sm_ascent = is_smoother_type((int)mg->solver_type[1]);
  const bool persistent = dist || (env_int("CSB_MG_MODE_MULTI", 0) == 0 && !sm_descent
                                   && !sm_ascent);
  if (mg->merge_level >= 0) {
    set_error("csb_multigrid_solve: a hierarchy with merged coarse levels "
              "(csb_multigrid_set_merge_options) runs as the preconditioner of FCG "
              "(csb_sles_it_transfer_pc)");
    return CSB_SLES_BREAKDOWN;
  }
  if (dist && (sm_descent || sm_ascent)) {
    set_error("csb_multigrid_solve: on a partitioned matrix the cs_multigrid_smoother "
              "smoothers run as the preconditioner of FCG (csb_sles_it_transfer_pc); the "
              "V-cycle in solver mode takes PCG + Jacobi smoothers");
    return CSB_SLES_BREAKDOWN;
  }
  if (sm_descent && !launch_mg_set_initial_residual(ms, HUGE_VAL)) return CSB_SLES_BREAKDOWN;

  if (persistent) {
    /* the whole solve (all cycles, all levels) in one cooperative launch */
    MgSolveParams S;
    S.n_levels = n_levels; S.levels = mg->descs.p; S.ms = ms;
    S.rhs = s_rhs.d; S.vx = s_vx.d; S.wr = mg->wr.p;
    S.vx_ini_nonzero = (vx_ini == vx) ? 1 : 0;
    S.n_max_cycles = mg->n_max_cycles;
    for (int i = 0; i < 3; i++) {
      S.n_max_iter[i] = mg->n_max_iter[i];
      S.precision_mult[i] = mg->precision_mult[i];
    }
    S.precision = precision; S.r_norm = r_norm;
    S.single_rows = sb_rows_threshold();
    S.small_rows = env_int("CSB_PCG_SMALL_ROWS", 448);
    PersistentCommon P;
    P.work = mg->aux.p; P.work_stride = work_stride_of((size_t)a->n_cols_ext);
    P.partials = mg->partials.p; P.barrier = mg->barrier.p;
    P.comm = dist ? mg->pcomm->d : nullptr;
    if (dist) pcomm_host_barrier(mg->pcomm);   /* no rank is inside cudaFree while kernels spin */
    const bool launched_ok = launch_mg_persistent(S, P, mg->p_grid, mg->p_smem,
                                                  c.device_share <= 1);
    if (dist) pcomm_host_barrier(mg->pcomm);   /* all ranks launched before anyone blocks */
    if (!launched_ok) return CSB_SLES_BREAKDOWN;
    if (!CSB_CUDA(cudaMemcpyAsync(mg->h_mg, ms, MG_HEAD, cudaMemcpyDeviceToHost, c.stream)))
      return CSB_SLES_BREAKDOWN;
    if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return CSB_SLES_BREAKDOWN;
    cvg = mg->h_mg->cvg;
    n_cycles = mg->h_mg->cycle_id;
  }

  for (n_cycles = persistent ? n_cycles : 0; !persistent && cvg == CSB_SLES_ITERATING; n_cycles++) {
    const int cycle_id = n_cycles + 1;
    bool ok = launch_mg_cycle_begin(ms, cycle_id);

    /* descent */
    for (int level = 0; ok && level < coarsest; level++) {
      MgLevel *f = mg->levels[level], *cl = mg->levels[level+1];
      const double *rhs_lv = (level == 0) ? s_rhs.d : f->rhs.p;
      double *vx_lv = (level == 0) ? s_vx.d : f->vx.p;
      const bool ini = (level == 0) ? ((n_cycles > 0) || (vx_ini == vx)) : false;
      ItState *st = mg->it_states.p + 2*level;

      if (sm_descent)
        ok = ok && mg_level_smooth(mg, 0, level, rhs_lv, vx_lv, ini, precision, r_norm);
      else {
        ok = ok && pcg_enqueue(f->m, st, ms, f->ad_inv.p, precision*mg->precision_mult[0],
                               r_norm, mg->n_max_iter[0], rhs_lv, vx_lv, ini, true,
                               mg->aux.p, nullptr);
        ok = ok && launch_mg_accum_iter(ms, st, level, (double)f->m->n_rows * denom,
                                        level == 0 ? 1 : 0);
      }
      SpmvArgs sa;
      sa.m = f->m; sa.x = vx_lv; sa.y = mg->wr.p; sa.b = rhs_lv; sa.mg = ms;
      ok = ok && launch_spmv(level == 0 ? EPI_RESIDUAL_NORM : EPI_RESIDUAL, sa);
      ok = ok && launch_restrict(cl->m->n_rows, f->c_f_row_index.p, f->c_f_row_ids.p,
                                 mg->wr.p, cl->rhs.p, ms);
    }

    /* coarsest level solve */
    if (ok) {
      MgLevel *cl = mg->levels[coarsest];
      const double *rhs_lv = (coarsest == 0) ? s_rhs.d : cl->rhs.p;
      double *vx_lv = (coarsest == 0) ? s_vx.d : cl->vx.p;
      ItState *st = mg->it_states.p + 2*coarsest;
      ok = ok && pcg_enqueue(cl->m, st, ms, cl->ad_inv.p, precision*mg->precision_mult[2],
                             r_norm, mg->n_max_iter[2], rhs_lv, vx_lv, false, true,
                             mg->aux.p, nullptr);
      ok = ok && launch_mg_accum_iter(ms, st, coarsest, (double)cl->m->n_rows * denom, 0);
    }

    /* ascent */
    for (int level = coarsest - 1; ok && level > -1; level--) {
      MgLevel *f = mg->levels[level], *cl = mg->levels[level+1];
      double *vx_lv = (level == 0) ? s_vx.d : f->vx.p;
      ok = ok && launch_prolong_inc(f->m->n_rows, f->coarse_row.p, cl->vx.p, vx_lv, ms);
      if (level > 0 && sm_ascent)
        ok = ok && mg_level_smooth(mg, 1, level, f->rhs.p, vx_lv, true, precision, r_norm);
      else if (level > 0) {
        ItState *st = mg->it_states.p + 2*level + 1;
        ok = ok && pcg_enqueue(f->m, st, ms, f->ad_inv.p, precision*mg->precision_mult[1],
                               r_norm, mg->n_max_iter[1], f->rhs.p, vx_lv, true, true,
                               mg->aux.p, nullptr);
        ok = ok && launch_mg_accum_iter(ms, st, level, (double)f->m->n_rows * denom, 0);
      }
    }
    if (!ok) return CSB_SLES_BREAKDOWN;

    if (!CSB_CUDA(cudaMemcpyAsync(mg->h_mg, ms, MG_HEAD, cudaMemcpyDeviceToHost, c.stream)))
      return CSB_SLES_BREAKDOWN;
    if (coarsest == 0) {
      /* degenerate one-level hierarchy: no descent, hence no cycle test; the coarsest
         solver's result is the cycle's (as k_mg_persistent does) */
      ItState head;
      if (!CSB_CUDA(cudaMemcpyAsync(&head, mg->it_states.p, IT_HEAD, cudaMemcpyDeviceToHost,
                                    c.stream))
          || !CSB_CUDA(cudaStreamSynchronize(c.stream))) return CSB_SLES_BREAKDOWN;
      cvg = (head.cvg == CSB_SLES_ITERATING) ? CSB_SLES_MAX_ITERATION : head.cvg;
      one_level_residual = head.residual;
      n_cycles = 1;
      break;
    }
    if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return CSB_SLES_BREAKDOWN;
    cvg = mg->h_mg->cvg;
    if (cvg == CSB_SLES_ITERATING && cycle_id > mg->n_max_cycles)
      cvg = CSB_SLES_MAX_ITERATION;      /* never spin past the cycle budget */
    if (verbosity > 2)
      printf("     %4d  |               %6d  |  %12.4e\n", cycle_id, mg->h_mg->n_equiv_iter,
             mg->h_mg->residual / r_norm);
  }

  if (!CSB_CUDA(cudaMemcpyAsync(mg->h_mg, ms, sizeof(MgState), cudaMemcpyDeviceToHost,
                                c.stream))) return CSB_SLES_BREAKDOWN;
  if (!s_vx.copy_back((size_t)a->n_rows)) return CSB_SLES_BREAKDOWN;
  if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return CSB_SLES_BREAKDOWN;

  if (n_iter) *n_iter = mg->h_mg->n_equiv_iter;
  if (residual) *residual = (one_level_residual >= 0.) ? one_level_residual : mg->h_mg->residual;
  mg->n_cycles_last = n_cycles;
  mg->n_cycles_tot += n_cycles;
  mg->n_solves += 1;
  mg->t_solve_tot += wall() - t0;
  return (csb_sles_convergence_state_t)cvg;
}

/* host-only helpers (no device needed): the aggregation and coarse structure
   used by csb_multigrid_setup, exposed for CPU-side parity tests */
int csb_host_aggregate_dx_msr(int level, csb_lnum_t n_rows, const csb_lnum_t *row_index,
                              const csb_lnum_t *col_id, const csb_real_t *d_val,
                              const csb_real_t *x_val, int max_aggregation, int n_blocks,
                              csb_lnum_t *f_c_row)
{
  return aggregate_dx_msr(level, n_rows, row_index, col_id, d_val, x_val, max_aggregation,
                          n_blocks, f_c_row);
}

int csb_host_coarsen_pairwise(int level, csb_lnum_t n_rows, const csb_lnum_t *row_index,
                              const csb_lnum_t *col_id, const csb_real_t *d_val,
                              const csb_real_t *x_val, int aggregation_limit,
                              csb_lnum_t *f_c_row, csb_lnum_t *c_row_index,
                              csb_lnum_t *c_col_id, csb_real_t *c_d_val, csb_real_t *c_x_val)
{
  HostMsr c;
  const int nc = coarsen_pairwise_galerkin(level, n_rows, row_index, col_id, d_val, x_val,
                                           aggregation_limit, f_c_row, c);
  const int nnz = (nc > 0) ? c.row_index[nc] : 0;
  if (c_row_index) std::copy(c.row_index.begin(), c.row_index.begin() + nc + 1, c_row_index);
  if (c_col_id) std::copy(c.col_id.begin(), c.col_id.begin() + nnz, c_col_id);
  if (c_d_val) std::copy(c.d_val.begin(), c.d_val.begin() + nc, c_d_val);
  if (c_x_val) std::copy(c.x_val.begin(), c.x_val.begin() + nnz, c_x_val);
  return nc;
}

int csb_host_coarse_structure(csb_lnum_t f_n_rows, const csb_lnum_t *f_row_index,
                              const csb_lnum_t *f_col_id, const csb_lnum_t *f_c_row,
                              csb_lnum_t c_n_rows, csb_lnum_t *c_row_index,
                              csb_lnum_t *c_col_id, csb_lnum_t *c_cell_face,
                              short *c_cell_face_sgn, csb_lnum_t *c_n_faces)
{
  CoarseStructure cs;
  build_coarse_structure(f_n_rows, f_row_index, f_col_id, f_c_row, c_n_rows, cs);
  const int c_nnz = cs.c_row_index[c_n_rows];
  if (c_row_index) memcpy(c_row_index, cs.c_row_index.data(), sizeof(int)*((size_t)c_n_rows+1));
  if (c_col_id) memcpy(c_col_id, cs.c_col_id.data(), sizeof(int)*(size_t)c_nnz);
  if (c_cell_face) memcpy(c_cell_face, cs.c_cell_face.data(), sizeof(int)*(size_t)c_nnz);
  if (c_cell_face_sgn) memcpy(c_cell_face_sgn, cs.c_cell_face_sgn.data(), sizeof(short)*(size_t)c_nnz);
  if (c_n_faces) *c_n_faces = cs.c_n_faces;
  return c_nnz;
}

int csb_multigrid_get_n_levels(const csb_multigrid_t *mg)
{ return (mg != nullptr) ? (int)mg->levels.size() : 0; }

int csb_multigrid_is_setup(const csb_multigrid_t *mg)
{ return (mg != nullptr && mg->setup_done) ? 1 : 0; }

int csb_multigrid_get_n_cycles(const csb_multigrid_t *mg)
{ return (mg != nullptr) ? mg->n_cycles_last : 0; }

int csb_multigrid_get_level_info(const csb_multigrid_t *mg, int level, csb_lnum_t info[3])
{
  if (mg == nullptr || level < 0 || level >= (int)mg->levels.size()) return -1;
  const csb_matrix_t *m = mg->levels[level]->m;
  info[0] = m->n_rows; info[1] = m->n_cols_ext; info[2] = m->nnz;
  return 0;
}

int csb_multigrid_get_level_arrays(const csb_multigrid_t *mg, int level,
                                   csb_lnum_t *coarse_row, csb_lnum_t *row_index,
                                   csb_lnum_t *col_id, csb_real_t *d_val, csb_real_t *x_val)
{
  if (mg == nullptr || level < 0 || level >= (int)mg->levels.size()) return -1;
  const csb_matrix_t *m = mg->levels[level]->m;
  if (coarse_row != nullptr && level > 0) {
    const MgLevel *p = mg->levels[level-1];
    memcpy(coarse_row, p->h_coarse_row.data(), sizeof(int)*(size_t)p->m->n_cols_ext);
  }
  bool ok = true;
  if (row_index) ok = ok && m->row_index.download(row_index);
  if (col_id) ok = ok && m->col_id.download(col_id);
  if (d_val) ok = ok && m->d_val.download(d_val);
  if (x_val) ok = ok && m->x_val.download(x_val);
  return ok ? 0 : -1;
}

int csb_multigrid_get_level_halo(const csb_multigrid_t *mg, int level, int sizes[3],
                                 int *c_domain_rank, csb_lnum_t *send_index,
                                 csb_lnum_t *send_list, csb_lnum_t *index)
{
  if (mg == nullptr || level < 0 || level >= (int)mg->levels.size()) return -1;
  const csb_halo_t *h = mg->levels[level]->halo;
  if (h == nullptr) return -1;
  if (sizes) { sizes[0] = h->n_c_domains; sizes[1] = h->n_send; sizes[2] = h->n_ghost; }
  const int nd = h->n_c_domains;
  if (c_domain_rank) for (int i = 0; i < nd; i++) c_domain_rank[i] = h->c_domain_rank[i];
  if (send_index) for (int i = 0; i <= nd; i++) send_index[i] = h->send_index[i];
  if (index) for (int i = 0; i <= nd; i++) index[i] = h->index[i];
  if (send_list) for (int i = 0; i < h->n_send; i++) send_list[i] = h->h_send_list[i];
  return 0;
}

int csb_multigrid_get_cycle_residuals(const csb_multigrid_t *mg, double *res, int max_n)
{
  if (mg == nullptr || mg->h_mg == nullptr) return 0;
  int n = mg->h_mg->n_cycles_res;
  if (n > max_n) n = max_n;
  for (int i = 0; i < n; i++) res[i] = mg->h_mg->cycle_res[i];
  return n;
}

int csb_multigrid_get_phase_times(const csb_multigrid_t *mg, double *seconds, int max_levels)
{
  if (mg == nullptr || mg->h_mg == nullptr) return 0;
  int n = (int)mg->levels.size();
  if (n > max_levels) n = max_levels;
  if (n > PROF_LEVELS) n = PROF_LEVELS;
  for (int l = 0; l < n; l++)
    for (int p = 0; p < PROF_PHASES; p++)
      seconds[l*PROF_PHASES + p] = 1e-9 * (double)mg->h_mg->prof[l][p];
  return n;
}

int csb_multigrid_get_smoother_times(const csb_multigrid_t *mg, double *seconds, int max_levels)
{
  if (mg == nullptr || mg->h_mg == nullptr) return 0;
  int n = (int)mg->levels.size();
  if (n > max_levels) n = max_levels;
  if (n > PROF_LEVELS) n = PROF_LEVELS;
  for (int l = 0; l < n; l++)
    for (int p = 0; p < PROF_IT_PHASES; p++)
      seconds[l*PROF_IT_PHASES + p] = 1e-9 * (double)mg->h_mg->prof_it[l][p];
  return n;
}

int csb_multigrid_get_setup_times(const csb_multigrid_t *mg, double t[4])
{
  if (mg == nullptr) return -1;
  for (int i = 0; i < 4; i++) t[i] = mg->t_setup[i];
  return 0;
}

void csb_multigrid_log(const void *context, int log_type)
{
  (void)log_type;
  const csb_multigrid_t *mg = static_cast<const csb_multigrid_t *>(context);
  if (mg == nullptr) return;
  printf("  Solver type:                       Multigrid V-cycle (B200 device)\n"
         "  Number of setups:                  %12u\n"
         "  Number of calls:                   %12u\n"
         "  Number of levels (last):           %12d\n"
         "  Total number of cycles:            %12llu\n"
         "  Total setup time:                  %12.3f\n"
         "  Total solution time:               %12.3f\n",
         mg->n_setups, mg->n_solves, (int)mg->levels.size(), mg->n_cycles_tot,
         mg->t_setup_tot, mg->t_solve_tot);
}

void *csb_multigrid_copy(const void *context)
{
  const csb_multigrid_t *s = static_cast<const csb_multigrid_t *>(context);
  if (s == nullptr) return nullptr;
  csb_multigrid_t *d = csb_multigrid_create(s->type);
  if (d == nullptr) return nullptr;
  d->aggregation_limit = s->aggregation_limit;
  d->coarsening_type = s->coarsening_type;
  d->n_levels_max = s->n_levels_max;
  d->n_g_rows_min = s->n_g_rows_min;
  d->p0p1_relax = s->p0p1_relax;
  d->k_cycle_threshold = s->k_cycle_threshold;
  d->n_blocks = s->n_blocks;
  d->n_max_cycles = s->n_max_cycles;
  for (int i = 0; i < 3; i++) {
    d->solver_type[i] = s->solver_type[i];
    d->n_max_iter[i] = s->n_max_iter[i];
    d->poly_degree[i] = s->poly_degree[i];
    d->precision_mult[i] = s->precision_mult[i];
  }
  d->nccl_comm = s->nccl_comm;
  d->pcomm = s->pcomm;
  d->merge_stride = s->merge_stride;
  d->merge_mean_threshold = s->merge_mean_threshold;
  d->merge_glob_threshold = s->merge_glob_threshold;
  return d;
}

/*----------------------------------------------------------------------------
 * Multigrid as a preconditioner, and the multigrid smoothers as objects
 *----------------------------------------------------------------------------*/

/* cs_multigrid_pc_create (cs_multigrid.cpp:5682) -> _multigrid_pc_create (:1616-1640) with the
   device settings cs_multigrid_set_solver_options derives from them (:5070-5090): the
   symmetric Gauss-Seidel smoothers of the host path become RJ3 with twice the sweeps
   (1 -> 2, rounded up to the schedule length 3 at setup), PCG coarse solver, one cycle. */
csb_multigrid_t *csb_multigrid_pc_create(csb_multigrid_type_t mg_type)
{
  csb_multigrid_t *mg = csb_multigrid_create(mg_type);
  if (mg == nullptr) return nullptr;
  if (mg_type == CSB_MULTIGRID_K_CYCLE)
    /* :1642-1655 (truncated Gauss-Seidel) with the device substitution of :5093-5100:
       2 Jacobi sweeps; PCG coarse solver at 10 times the outer precision */
    csb_multigrid_set_solver_options(mg, CSB_SLES_JACOBI, CSB_SLES_JACOBI, CSB_SLES_PCG,
                                     1, 2, 2, 500, 0, 0, 0, -1., -1., 10.);
  else
    csb_multigrid_set_solver_options(mg, CSB_SLES_RJ3, CSB_SLES_RJ3, CSB_SLES_PCG,
                                     1, 2, 2, 500, 0, 0, 0, -1., -1., 1.);
  return mg;
}

/* cs_sles_it_transfer_pc (cs_sles_it.cpp:5612): the solver takes ownership, *pc is set to
   NULL.  Available for CSB_SLES_FCG created with poly_degree < 0. */
int csb_sles_it_transfer_pc(csb_sles_it_t *c, csb_multigrid_t **pc)
{
  if (c == nullptr || pc == nullptr) return -1;
  if ((c->type != CSB_SLES_FCG && c->type != CSB_SLES_PCG && c->type != CSB_SLES_IPCG)
      || c->poly_degree >= 0 || c->is_smoother) {
    set_error("csb_sles_it_transfer_pc: needs CSB_SLES_FCG, _IPCG or _PCG created with "
              "poly_degree < 0");
    return -1;
  }
  delete c->pc_mg;
  c->pc_mg = *pc;
  *pc = nullptr;
  c->setup_done = false;
  return 0;
}

csb_multigrid_t *csb_sles_it_get_pc(csb_sles_it_t *c)     /* cs_sles_it_get_pc :5590 */
{
  return (c != nullptr) ? c->pc_mg : nullptr;
}

/* cs_multigrid_smoother_create (cs_multigrid_smoother.cpp:1726-1812) */
csb_sles_it_t *
csb_multigrid_smoother_create(csb_sles_it_type_t smoother_type, int poly_degree, int n_iter)
{
  (void)poly_degree;
  if (!ctx().ready) { set_error("csb_multigrid_smoother_create: csb_init not called"); return nullptr; }
  if (!is_smoother_type((int)smoother_type)) {
    set_error("csb_multigrid_smoother_create: type %d is not a smoother available on this "
              "path (Jacobi, L1 Jacobi, relaxed Jacobi, RJ2, RJ3, local / symmetric / "
              "truncated Gauss-Seidel)", (int)smoother_type);
    return nullptr;
  }
  csb_sles_it_t *c = new csb_sles_it_t;
  c->type = smoother_type;
  c->n_max_iter = n_iter;
  c->is_smoother = true;
  return c;
}

/* cs_multigrid_smoother_setup (cs_multigrid_smoother.cpp:1825-1985) */
void csb_multigrid_smoother_setup(void *context, const char *name, const csb_matrix_t *a,
                                  int verbosity)
{
  (void)name; (void)verbosity;
  csb_sles_it_t *c = static_cast<csb_sles_it_t *>(context);
  if (c == nullptr || a == nullptr || !c->is_smoother) {
    set_error("csb_multigrid_smoother_setup: not a smoother context");
    return;
  }
  const double t0 = wall();
  c->a = a;
  c->setup_done = false;
  if (a->halo != nullptr) {
    set_error("csb_multigrid_smoother_setup: one rank on this path");
    return;
  }
  if (a->db != 1) {
    /* block-diagonal matrices: block Jacobi and block relaxed Jacobi
       (cs_multigrid_smoother_setup :1880-1960); L1 Jacobi has no block form there either */
    if (c->type != CSB_SLES_JACOBI && c->type != CSB_SLES_R_JACOBI && c->type != CSB_SLES_RJ2
        && c->type != CSB_SLES_RJ3) {
      set_error("csb_multigrid_smoother_setup: smoother type %d is not available for a diagonal "
                "block size of %d (Jacobi, relaxed Jacobi, RJ2, RJ3 are)", (int)c->type, a->db);
      return;
    }
    if (!c->ad_inv.alloc((size_t)a->n_rows*a->db*a->db)) return;
    if (!launch_block_nn_inverse(a->n_rows, a->db, a->d_val.p, c->ad_inv.p)) return;
    if (!c->work.alloc(2*work_stride_of((size_t)a->n_cols_ext*a->db))) return;
    if (!c->st.alloc(1)) return;
    c->n_max_iter = smoother_plan((int)c->type, c->n_max_iter).n_iter;
    c->setup_done = true;
    c->n_setups += 1;
    c->t_setup += wall() - t0;
    return;
  }
  if (!c->ad_inv.alloc((size_t)a->n_rows)) return;
  if (c->type == CSB_SLES_L1_JACOBI) { if (!launch_l1_inv(a, c->ad_inv.p)) return; }
  else if (!launch_ad_inv(a->n_rows, a->d_val.p, c->ad_inv.p)) return;
  if (!c->work.alloc(work_stride_of((size_t)a->n_cols_ext))) return;
  /* the scheduled variants run whole schedules (:1888-1898) */
  c->n_max_iter = smoother_plan((int)c->type, c->n_max_iter).n_iter;
  /* the truncated Gauss-Seidel sweeps are single iterations (:1497, :1627) */
  if (c->type == CSB_SLES_TS_F_GAUSS_SEIDEL || c->type == CSB_SLES_TS_B_GAUSS_SEIDEL)
    c->n_max_iter = 1;
  if (is_gs_type((int)c->type) && gs_plan_get(a) == nullptr) return;
  c->setup_done = true;
  c->n_setups += 1;
  c->t_setup += wall() - t0;
}

/* cs_multigrid_smoother_solve (cs_multigrid_smoother.cpp:2008-2090): fixed sweep count,
   *residual = -1 (not computed), returns CSB_SLES_MAX_ITERATION like the reference. */
csb_sles_convergence_state_t
csb_multigrid_smoother_solve(void *context, const char *name, const csb_matrix_t *a,
                             int verbosity, double precision, double r_norm, int *n_iter,
                             double *residual, const csb_real_t *rhs, csb_real_t *vx_ini,
                             csb_real_t *vx, size_t aux_size, void *aux_vectors)
{
  (void)precision; (void)r_norm; (void)aux_size; (void)aux_vectors;
  csb_sles_it_t *c = static_cast<csb_sles_it_t *>(context);
  Context &cx = ctx();
  cx.err.clear();
  if (n_iter) *n_iter = 0;
  if (residual) *residual = -1.;
  if (c == nullptr || !c->is_smoother || a == nullptr || rhs == nullptr || vx == nullptr) {
    set_error("csb_multigrid_smoother_solve: null argument or not a smoother context");
    return CSB_SLES_BREAKDOWN;
  }
  if (vx_ini != nullptr && vx_ini != vx) {
    set_error("csb_multigrid_smoother_solve: vx_ini must be vx or NULL");
    return CSB_SLES_BREAKDOWN;
  }
  if (!c->setup_done || c->a != a)
    csb_multigrid_smoother_setup(c, name, a, verbosity);
  if (!c->setup_done) return CSB_SLES_BREAKDOWN;

  const double t0 = wall();
  const bool ini = (vx_ini == vx);
  StagedVec s_rhs, s_vx;
  if (!s_rhs.init(const_cast<double *>(rhs), (size_t)a->n_rows*a->db, true, c->stage_rhs))
    return CSB_SLES_BREAKDOWN;
  if (!s_vx.init(vx, (size_t)a->n_cols_ext*a->db, ini, c->stage_vx)) return CSB_SLES_BREAKDOWN;
  if (a->db != 1) {
    if (!block_smoother_enqueue(a, (int)c->type, c->n_max_iter, c->ad_inv.p, s_rhs.d, s_vx.d,
                                ini, c->work.p, work_stride_of((size_t)a->n_cols_ext*a->db),
                                c->st.p)) return CSB_SLES_BREAKDOWN;
  }
  else if (!smoother_enqueue(a, (int)c->type, c->n_max_iter, c->ad_inv.p, s_rhs.d, s_vx.d, ini,
                             c->work.p, nullptr)) return CSB_SLES_BREAKDOWN;
  if (!s_vx.copy_back((size_t)a->n_rows*a->db)) return CSB_SLES_BREAKDOWN;
  if (!CSB_CUDA(cudaStreamSynchronize(cx.stream))) return CSB_SLES_BREAKDOWN;
  if (n_iter) *n_iter = c->n_max_iter;
  c->n_solves += 1;
  c->n_iterations_last = c->n_max_iter;
  c->n_iterations_tot += c->n_max_iter;
  c->t_solve += wall() - t0;
  return CSB_SLES_MAX_ITERATION;
}

} /* extern "C" */
