/* csb_internal.h -- internal C++ declarations of libcsb200 (not part of the C-ABI). */
#ifndef CSB_INTERNAL_H
#define CSB_INTERNAL_H

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "csb200.h"
#include "grid_host.h"

namespace csb {

/*--------------------------------------------------------------------------
 * Runtime context: one per process (one process per GPU).
 *--------------------------------------------------------------------------*/

struct Context {
  bool          ready = false;
  int           device = -1;
  int           sm_count = 148;
  cudaStream_t  own_stream = nullptr;
  cudaStream_t  stream = nullptr;       /* stream all kernels are launched on */
  cudaStream_t  comm_stream = nullptr;  /* halo exchange */
  cudaEvent_t   ev_a = nullptr, ev_b = nullptr;
  /* 2-stage deterministic reduction workspace */
  double       *partials = nullptr;     /* [RED_MAX_BLOCKS * RED_MAX_Q] */
  unsigned     *ticket = nullptr;       /* self-resetting atomicInc counter */
  double       *red_out = nullptr;      /* device scalars for plain csb_dot* */
  double       *red_out_h = nullptr;    /* pinned mirror */
  long long     launches = 0;
  int           partials_blocks = 0;    /* capacity of `partials` in CTAs */
  int           device_share = 1;       /* ranks sharing this device (in-process groups) */
  bool          pool_ok = false;        /* stream-ordered allocator usable (dev_alloc) */
  std::string   err;
};

Context &ctx();
void set_error(const char *fmt, ...);
bool cuda_ok(cudaError_t e, const char *what, const char *file, int line);

#define CSB_CUDA(call) ::csb::cuda_ok((call), #call, __FILE__, __LINE__)

constexpr int RED_MAX_BLOCKS = 4096;
constexpr int RED_MAX_Q = 4;
constexpr int HIST_MAX = 2048;
constexpr int CYCLE_MAX = 512;
constexpr int PROF_LEVELS = 32;
constexpr int PROF_PHASES = 6;
constexpr int PROF_IT_PHASES = 3;
enum { MG_PROF_DESCENT = 0, MG_PROF_RESIDUAL, MG_PROF_RESTRICT, MG_PROF_COARSEST,
       MG_PROF_PROLONG, MG_PROF_ASCENT };

/* true if p is a device (or managed) pointer */
bool is_device_ptr(const void *p);

/* Device memory of the library's objects: the stream-ordered allocator of the device's default
   memory pool (cudaMallocAsync / cudaFreeAsync on the library's stream, release threshold
   raised at csb_init so that freed blocks stay in the pool).  A hierarchy has ~12 arrays per
   level; with cudaMalloc / cudaFree a repeated setup spent ~0.25 s of its 0.55 s in the driver
   (profiles/r02/setup_trace_256_v1.log) -- the reference rebuilds its hierarchy at every
   pressure solve.  CSB_DEV_POOL=0 goes back to cudaMalloc.  NULL + set_error on failure. */
void *dev_alloc(size_t bytes);
void  dev_free(void *p);

/* RAII device buffer */
template <typename T>
struct DevBuf {
  T      *p = nullptr;
  size_t  n = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  bool alloc(size_t count) {
    release();
    n = count;
    if (count == 0) count = 1;
    /* 64 bytes of slack: bulk copies (cp.async.bulk) read whole 16-byte units */
    p = static_cast<T *>(dev_alloc(count * sizeof(T) + 64));
    return p != nullptr;
  }
  void release() { if (p) dev_free(p); p = nullptr; n = 0; }
  /* copy from host or device pointer */
  bool upload(const T *src, size_t count) {
    if (!alloc(count)) return false;
    if (count == 0) return true;
    return CSB_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyDefault,
                                    ctx().stream));
  }
  bool download(T *dst) const {
    if (n == 0) return true;
    bool ok = CSB_CUDA(cudaMemcpyAsync(dst, p, n * sizeof(T), cudaMemcpyDeviceToHost,
                                       ctx().stream));
    return ok && CSB_CUDA(cudaStreamSynchronize(ctx().stream));
  }
};

/*--------------------------------------------------------------------------
 * Device-resident solver state (no host round trip per iteration)
 *--------------------------------------------------------------------------*/

/* cs_sles_it convergence + Krylov scalars; lives in device memory. */
struct ItState {
  double precision, r_norm;
  double residual, initial_residual;
  double ro_0, ro_1, rk_gkm1, alpha, beta;
  double gamma, omega, betam1;           /* BiCGStab */
  int    n_iter, n_max_iter;
  int    cvg;                            /* csb_sles_convergence_state_t, 2 = skipped */
  int    n_hist;
  int    n_halo;                         /* ghost exchanges done by the last run */
  double hist[HIST_MAX];
};

/* cs_multigrid cycle state; lives in device memory. */
struct MgState {
  double precision, r_norm;
  double residual, initial_residual;
  int    cvg;              /* gate: != 0 turns the rest of the cycle into no-ops */
  int    cycle_id, n_max_cycles;
  int    n_equiv_iter;
  int    n_cycles_res;
  int    diverged_level;
  double cycle_res[CYCLE_MAX];
  /* device-side phase timers of the persistent V-cycle kernel: nanoseconds
     (%globaltimer, lead thread, after the phase's closing grid barrier) summed
     over the cycles of the last solve; [level][phase], phases = MG_PROF_* */
  unsigned long long prof[PROF_LEVELS][PROF_PHASES];
  /* inside the smoothers: [level][0 residual SpMV, 1 SpMV + 3 dots, 2 vector updates],
     each including the grid barrier / reduction that closes it */
  unsigned long long prof_it[PROF_LEVELS][PROF_IT_PHASES];
};

/*--------------------------------------------------------------------------
 * Peer-memory communicator (comm.cu): one arena per rank, mapped by every
 * other rank (cudaIpc between processes, plain pointers between host threads).
 *--------------------------------------------------------------------------*/

constexpr int COMM_MAX_RANKS = 16;
constexpr int COMM_MAX_DOM = 16;       /* neighbour ranks of one halo */

/* At arena offset 0 on every rank; written by peers. */
struct CommHeader {
  unsigned  xflag[2][COMM_MAX_RANKS];      /* all-reduce: epoch published by rank r */
  double    xsum[2][COMM_MAX_RANKS][8];    /* all-reduce: partial sums of rank r */
  unsigned  mflag[COMM_MAX_RANKS];         /* mailbox: message count from rank r */
  unsigned  mack[COMM_MAX_RANKS];          /* mailbox: my messages consumed by rank r */
  long long mbox[COMM_MAX_RANKS][8];
  unsigned  wflag[COMM_MAX_RANKS];         /* window exchange: message count from rank r */
  unsigned  wack[COMM_MAX_RANKS];
};

/* Rank-private device state. */
struct CommLocal {
  unsigned x_epoch;
  unsigned m_sent[COMM_MAX_RANKS], m_recv[COMM_MAX_RANKS];
  unsigned w_sent[COMM_MAX_RANKS], w_recv[COMM_MAX_RANKS];
};

/* What kernels need of the communicator (lives in device memory). */
struct CommDev {
  int        n_ranks, rank;
  char      *peer[COMM_MAX_RANKS];         /* arena base of every rank (own included) */
  CommLocal *local;
  size_t     window_off, window_bytes;     /* staging window inside the arena */
  unsigned long long timeout_ns;           /* a flag not seen for this long traps the kernel */
};

/* Halo as seen by the window-exchange kernel. */
struct HaloDev {
  int        n_dom;
  int        rank_of[COMM_MAX_DOM];
  int        send_index[COMM_MAX_DOM + 1];
  int        ghost_index[COMM_MAX_DOM + 1];
  int        peer_ghost_index[COMM_MAX_DOM];   /* where my section starts at the neighbour */
  const int *send_list;                        /* nullptr: source already in send order */
};

/* One level's ghost exchange inside the persistent kernels: CTAs store the send
   list straight into the neighbour's ghost buffer (double-buffered by exchange
   parity), the last CTA to finish publishes the exchange count in the
   neighbour's flag word. */
struct DistLevel {
  int        n_dom, n_send, n_ghost;
  int        send_index[COMM_MAX_DOM + 1];
  const int *send_list;
  double    *gbuf[2];                      /* local ghost values (arena) */
  double    *push_dst[COMM_MAX_DOM][2];    /* neighbour's gbuf[p] + its section for me */
  unsigned  *flag_local;                   /* [n_dom] arena words written by neighbours */
  unsigned  *flag_remote[COMM_MAX_DOM];    /* neighbour's flag word for me */
  unsigned  *count;                        /* exchanges done so far (rank-private) */
  unsigned  *ticket;                       /* CTA arrival counter (rank-private) */
  unsigned long long timeout_ns;
  /* overlap of the exchange with the rows that need no ghost value (persistent.cu,
     spmv_phase_g): bit r of bmask = row r has ghost columns; bsum[r] = its parked sum */
  const unsigned *bmask;
  double    *bsum;
  int        overlap;
};

/* Coarse levels merged onto rank 0 (cs_grid_merge_ranks, src/alge/cs_grid_rank_merge.cpp:1359;
   gather / scatter of cs_grid_restrict_row_var / _prolong_row_var, cs_grid.cpp:9058-9085,
   :9131-9155): what the persistent kernel needs at the merge boundary.  Every V-cycle has one
   gather (restricted right-hand sides of the merge level, every rank -> rank 0) and one
   scatter (the level's correction, rank 0 -> every rank); both are peer stores followed by
   a counter in a flag word of the destination's arena. */
struct MergeDev {
  int        is_root, n_ranks, rank;
  int        n_local;                          /* rows of the merge level owned by this rank */
  int        shift[COMM_MAX_RANKS];            /* first merged row of rank r */
  int        n_of[COMM_MAX_RANKS];             /* rows of rank r */
  double    *stub_rhs, *stub_vx;               /* this rank's part: local buffers */
  double    *root_rhs_dst;                     /* rank 0's merged rhs + my shift (peer) */
  unsigned  *root_flag;                        /* rank 0's flag word for me (peer) */
  unsigned  *flag_in;                          /* own arena: [rank of sender], [0] on others */
  double    *peer_vx[COMM_MAX_RANKS];          /* rank 0 only: stub_vx of rank r (peer) */
  unsigned  *peer_flag[COMM_MAX_RANKS];        /* rank 0 only: flag_in of rank r (peer) */
  unsigned  *count;                            /* exchanges done so far (rank-private) */
  unsigned long long timeout_ns;
};

#ifdef __CUDACC__
/* device-side helpers shared by comm.cu and persistent.cu */
__device__ __forceinline__ unsigned long long gtime_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

/* wait until *flag >= expect (flags only grow); trap instead of hanging forever */
__device__ __forceinline__ void spin_until_ge(const unsigned *flag, unsigned expect,
                                              unsigned long long timeout_ns, int what)
{
  const volatile unsigned *f = flag;
  if ((int)(*f - expect) >= 0) return;
  const unsigned long long t0 = gtime_ns();
  unsigned n = 0;
  while ((int)(*f - expect) < 0) {
    __nanosleep(64);
    if ((++n & 1023u) == 0 && gtime_ns() - t0 > timeout_ns) {
      printf("[csb200] peer flag timeout (wait %d, block %d thread %d): have %u, expect %u\n",
             what, blockIdx.x, threadIdx.x, *f, expect);
      __trap();
    }
  }
}
#endif

} // namespace csb

struct CommGroup;

struct _csb_pcomm_t {
  csb::CommDev   h;                 /* host copy */
  csb::CommDev  *d = nullptr;       /* device copy */
  char          *arena = nullptr;
  size_t         arena_bytes = 0, arena_used = 0;
  struct ArenaSeg { size_t start; bool alive; };
  std::vector<ArenaSeg> segs;       /* owner segments, bottom to top (comm.cu) */
  CommGroup     *group = nullptr;   /* in-process ranks (host threads) */
  bool           ipc_open[csb::COMM_MAX_RANKS] = {false};
  double        *scratch = nullptr; /* 16 doubles, device */
  double        *scratch_h = nullptr;
};
typedef struct _csb_pcomm_t csb_pcomm_t;

namespace csb {

void  *pcomm_arena_alloc(csb_pcomm_t *c, size_t bytes);   /* zero-filled, 256-byte aligned */
size_t pcomm_arena_mark(csb_pcomm_t *c);      /* opens an owner segment */
void   pcomm_arena_release(csb_pcomm_t *c, size_t mark);
void   pcomm_host_barrier(csb_pcomm_t *c);
/* sum of n (<= 8) doubles over all ranks, in place, result on the host */
bool   pcomm_allreduce(csb_pcomm_t *c, double *vals, int n);
/* send 8 values to / receive 8 values from each listed neighbour */
bool   pcomm_mailbox(csb_pcomm_t *c, int n_nbr, const int *nbr, const long long *send,
                     long long *recv);
/* every rank's value, on every rank (host) */
bool   pcomm_allgather_count(csb_pcomm_t *c, long long mine, long long *all);
/* counts[r] elements (4 or 8 bytes each) of every rank r, concatenated in rank order in
   dst_d on rank 0 (device arrays; dst_d is ignored elsewhere); chunked through the window */
bool   pcomm_gather_to_root(csb_pcomm_t *c, int elem_bytes, const long long *counts,
                            const void *src_d, void *dst_d);
/* ghost section of dst <- neighbours' send lists of src; device arrays */
bool   pcomm_window_exchange(csb_pcomm_t *c, const HaloDev &h, int elem_bytes, int stride,
                             const void *src_d, void *ghost_dst_d);

} // namespace csb

/*--------------------------------------------------------------------------
 * Halo
 *--------------------------------------------------------------------------*/

struct _csb_halo_t {
  void              *nccl_comm = nullptr;
  csb_pcomm_t       *pcomm = nullptr;
  std::vector<int>   peer_ghost_index;   /* set by csb_halo_set_pcomm */
  int                local_rank = 0;
  int                n_c_domains = 0;
  csb_lnum_t         n_local_elts = 0;
  csb_lnum_t         n_send = 0, n_ghost = 0;
  std::vector<int>         c_domain_rank;
  std::vector<csb_lnum_t>  send_index, index;   /* host copies */
  std::vector<csb_lnum_t>  h_send_list;         /* host copy (coarsening) */
  csb::DevBuf<csb_lnum_t>  send_list;
  csb::DevBuf<double>      send_buf;
};

namespace csb {
/* fill the kernel view of a halo; false if it has too many neighbours */
bool halo_dev_view(const _csb_halo_t *h, HaloDev &v);
}

/*--------------------------------------------------------------------------
 * Matrix
 *--------------------------------------------------------------------------*/

struct _csb_matrix_t {
  csb_lnum_t  n_rows = 0, n_cols_ext = 0, nnz = 0;
  int         db = 1;                  /* diagonal block size: 1, or 3 (CS_MATRIX_BLOCK_D) */
  bool        symmetric = true;
  csb::DevBuf<csb_lnum_t>  row_index, col_id;
  csb::DevBuf<double>      d_val, x_val;
  /* host mirrors of the structure and values (needed by the host-side
     aggregation, like the reference which always aggregates on the host,
     src/alge/cs_grid.cpp:8035-8041) */
  std::vector<csb_lnum_t>  h_row_index;
  csb::HostVec<csb_lnum_t> h_col_id;
  csb::HostVec<double>     h_d_val, h_x_val;
  bool        h_vals_valid = false;
  /* host arrays lent by the caller (csb_matrix_share_host_arrays): same content as the device
     arrays, valid until the values change or the matrix is destroyed; they stand in for the
     mirrors above, so that the setup of a hierarchy reads level 0 where it already is */
  const csb_lnum_t *b_col_id = nullptr;
  const double     *b_d_val = nullptr, *b_x_val = nullptr;
  const csb_lnum_t *host_col_id() const { return b_col_id != nullptr ? b_col_id : h_col_id.data(); }
  const double     *host_d_val() const { return b_d_val != nullptr ? b_d_val : h_d_val.data(); }
  const double     *host_x_val() const { return b_x_val != nullptr ? b_x_val : h_x_val.data(); }
  bool host_structure_ready() const {
    return b_col_id != nullptr || h_col_id.size() == (size_t)(nnz > 0 ? nnz : 1);
  }
  /* SpMV launch geometry */
  int         spmv_rows_per_block = 256;
  size_t      spmv_smem = 0;           /* bytes of staged products per block */
  bool        spmv_stream_ok = true;   /* false -> scalar fallback kernel */
  /* slab plan of the bulk-copy pipeline kernel (k_spmv_pipeline), made on first use */
  mutable int    pl_state = 0;         /* 0 not planned, 1 usable, -1 not usable */
  mutable int    pl_slab_rows = 0, pl_n_slabs = 0, pl_grid = 0, pl_gdepth = 2, pl_row_len_max = 0;
  mutable size_t pl_smem = 0;
  /* mesh association (device) */
  csb_lnum_t  n_faces = 0;
  csb::DevBuf<csb_lnum_t>  c2f;
  csb::DevBuf<short>       c2f_sgn;
  /* entries of every row in ascending face order (kernels.cu face_order(); 0: not built yet,
     1: usable, -1: rows too long, the kernels scan) */
  mutable csb::DevBuf<unsigned char> fperm;
  mutable int fperm_state = 0;
  csb::DevBuf<double>      cell_cen, cell_vol, face_u_normal, face_surf;
  bool        has_mesh = false;
  /* boundary faces by adjacent cell, ascending face ids (csb_matrix_set_boundary_faces):
     what cs_matrix_compute_coeffs needs beside the mesh association */
  csb_lnum_t  n_b_faces = 0;
  csb::DevBuf<csb_lnum_t>  cell_b_faces_idx, cell_b_faces;
  bool        has_b_faces = false;
  csb::DevBuf<double>      coef_stage;   /* host inputs of csb_matrix_compute_coeffs */
  _csb_halo_t *halo = nullptr;
  /* matrix mapped from native (face-based) arrays (csb_matrix_create_native): the faces and
     their values stay with the matrix, c2f / c2f_sgn above are the entry -> face map; the
     hierarchy is then built like the reference builds it for such a matrix (face-order
     aggregation and sums, cs_grid.cpp:3997, :5144) */
  /* face geometry of a non-orthogonal mesh (csb_matrix_set_face_geometry) and the cached
     least-squares cocg matrices of the cell gradient (single precision like the reference's,
     cs_gradient_priv.h:54) */
  csb::DevBuf<double>      i_dist, diipf, djjpf, b_face_u_normal, b_dist;
  bool        has_face_geo = false;
  csb::DevBuf<double>      diipb;        /* II' of the boundary faces (csb_matrix_set_boundary_iprime) */
  bool        has_diipb = false;
  /* Green-Gauss reconstruction (csb_matrix_set_green_geometry): face weights, OF vectors,
     boundary face surfaces, optional warped-face correction matrices; imrgra = the gradient
     csb_diffusion_potential_rc uses (1: least squares, 4: Green-Gauss with LSQ face values) */
  csb::DevBuf<double>      g_weight, dofij, b_face_surf, corr_grad_lin;
  /* face centres (csb_matrix_set_face_cog): the hydrostatic pressure treatment needs them */
  csb::DevBuf<double>      i_face_cog, b_face_cog;
  bool        has_face_cog = false;
  bool        has_green_geo = false, has_corr_grad_lin = false;
  int         imrgra = 1;
  /* iterative Green-Gauss gradient (imrgra = 0): eqp->nswrgr, eqp->epsrgr (defaults of
     cs_parameters.cpp:380, :394), cached cocg_it, sweeps of the last call */
  int         nswrgr = 100;
  double      epsrgr = 1.e-4;
  /* gradient limiter: eqp->imligr (-1: none), eqp->climgr, eqp->b_rc_clip_factor (the
     limiter of the values at I' of the boundary faces, on when imligr >= 0) */
  int         imligr = -1;
  double      climgr = -1., b_rc_clip_factor = 1.;
  mutable csb::DevBuf<double> cocg_it;
  mutable bool cocg_it_valid = false;
  mutable int  n_grad_sweeps = 0;
  mutable csb::DevBuf<float> cocg_lsq;
  mutable bool cocg_valid = false;
  bool        from_native = false;
  csb::DevBuf<csb_lnum_t>  face_cell;
  csb::DevBuf<double>      xa;
  csb::HostVec<csb_lnum_t> h_face_cell;
  csb::HostVec<double>     h_xa;
  /* matrix created from CSR arrays (cs_matrix_type_t CS_MATRIX_CSR): the CSR arrays are
     kept for cs_matrix_vector_multiply in CSR order (_mat_vec_p_l_csr); this object holds
     the MSR split (diagonal extracted) the solvers and the hierarchy work on */
  _csb_matrix_t *csr = nullptr;
  /* dependency levels of the Gauss-Seidel sweeps (sles.cu), built on first use */
  mutable struct GsPlan *gs_plan = nullptr;
  ~_csb_matrix_t();
};

namespace csb {

/*--------------------------------------------------------------------------
 * Kernel launch wrappers (kernels.cu)
 *--------------------------------------------------------------------------*/

enum SpmvEpilogue {
  EPI_STORE = 0,        /* y = A.x                                        */
  EPI_RESIDUAL,         /* y = b - A.x  (V-cycle residual, cs_multigrid.cpp:3296-3301) */
  EPI_RESIDUAL_NORM,    /* same + sum(y^2) -> mg finalize (level 0)       */
  EPI_PCG_INIT,         /* rk = A.x - b ; dk = rk*ad_inv  (cs_sles_it.cpp:1503-1508) */
  EPI_PCG_DOT3,         /* y = A.x ; sums (rk.rk, rk.x, x.y) -> pcg finalize */
  EPI_SMOOTH,           /* one fused Jacobi-family smoother sweep: y = S(x, b, A.x)
                           (cs_multigrid_smoother.cpp:205-213, :345-352, :880-887) */
  EPI_FCG_DOT4          /* y = A.x ; sums (rk.rk, rk.x, y.x, x.qk) -> fcg finalize
                           (_dot_products_vr_vw_vq_rr, cs_sles_it.cpp:693-724) */
};

/* EPI_SMOOTH variants (SpmvArgs::smooth_mode) */
enum SmoothMode {
  SMOOTH_JACOBI = 0,    /* y = (b - E.x) * ad_inv                            */
  SMOOTH_RELAXED,       /* y = w1*x + w0*(b - E.x)*ad_inv, w1 = 1 - w0      */
  SMOOTH_L1             /* y = x + (b - A.x) * l1_inv  (full product)        */
};

struct SpmvArgs {
  const _csb_matrix_t *m;
  bool          exclude_diag = false;
  const double *x = nullptr;
  double       *y = nullptr;
  const double *b = nullptr;       /* rhs for residual / init epilogues */
  const double *ad_inv = nullptr;
  double       *rk = nullptr;      /* PCG vectors for fused epilogues */
  double       *dk = nullptr;
  ItState      *st = nullptr;
  MgState      *mg = nullptr;
  int           phase = 0;         /* PCG_DOT3: 0 = first (ro_0, ro_1), 1 = iteration,
                                      2 = IPCG step length only */
  int           smooth_mode = 0;   /* EPI_SMOOTH */
  double        w0 = 1., w1 = 0.;  /* EPI_SMOOTH relaxation weights */
  const double *qk = nullptr;      /* EPI_FCG_DOT4 */
};

bool launch_spmv(SpmvEpilogue epi, const SpmvArgs &a);
const char *spmv_kernel_name();
bool launch_spmv_csr_exdiag(const _csb_matrix_t *m, const double *x, double *y);
bool launch_spmv_b3(const _csb_matrix_t *m, const double *x, double *y, bool exclude_diag);
bool launch_ad_inv_block(int n_rows, int db, const double *d_val, double *ad_inv);
/* block-diagonal MSR of any block size (thread per row component); mode -1 store, 0 / 1 the
   BiCGStab products with their dots */
bool launch_spmv_bn(int mode, const _csb_matrix_t *m, ItState *st, const double *x, double *y,
                    const double *b, bool exclude_diag);
bool launch_block_nn_inverse(int n_blocks, int db, const double *ad, double *ad_inv);
/* _block_relaxed_jacobi sweep (cs_multigrid_smoother.cpp:626-790); first: the sweep from zero */
bool launch_bn_relaxed_update(int n_blocks, int db, double wk, const double *ad_inv,
                              const double *rhs, const double *vxx, const double *vx_np,
                              double *vx_n, bool first);
bool launch_bn_jacobi_update(int n_blocks, int db, ItState *st, const double *ad,
                             const double *ad_inv, const double *rhs, const double *vxx,
                             double *vx, bool first);
bool launch_bicg_spmv_b3(int mode, const _csb_matrix_t *m, ItState *st, const double *x,
                         double *y, const double *b);

/* setup of SpMV launch geometry after (re)building a structure */
void matrix_plan_spmv(_csb_matrix_t *m);

/* MSR coefficients of a symmetric scalar diffusion system from face data
   (_sym_coeffs_msr, cs_matrix_building.cpp:136-206, and :2243-2262) */
bool launch_sym_coeffs_msr(const _csb_matrix_t *m, int idiffp, int ndircp, double thetap,
                           double relaxp, const double *cofbfp, const double *rovsdt,
                           const double *i_visc, const double *b_visc, double *d_val,
                           double *x_val);

/* same for a convection-diffusion scalar (_coeffs_msr, cs_matrix_building.cpp:407-520) */
bool launch_coeffs_msr(const _csb_matrix_t *m, int idiffp, int ndircp, double thetap,
                       double relaxp, const double *coefbp, const double *cofbfp,
                       const double *rovsdt, const double *i_massflux, const double *b_massflux,
                       const double *i_visc, const double *b_visc, const double *xcpp,
                       double *d_val, double *x_val);

/* vector variable, 3x3 diagonal blocks (_sym_coeffs_msr<3> / _coeffs_msr<3>, :250-353, :560-692) */
bool launch_coeffs_msr_v3(const _csb_matrix_t *m, int iconvp, int idiffp, int ndircp,
                          double thetap, double relaxp, const double *coefbp,
                          const double *cofbfp, const double *fimp, const double *i_massflux,
                          const double *b_massflux, const double *i_visc, const double *b_visc,
                          double *d_val, double *x_val);

/* ad_inv = 1/d  (cs_sles_pc.cpp:376) */
bool launch_ad_inv(int n, const double *d, double *ad_inv);
bool launch_fill(int n, double v, double *x);
bool launch_fill_gated(int n, double v, double *x, const MgState *gate);

/* plain reductions -> ctx().red_out[0..q) */
bool launch_dot(int n, const double *x, const double *y, const double *z, int mode);

/* device-state PCG pieces (cs_sles_it.cpp:1441-1645) */
bool launch_it_state_init(ItState *st, const MgState *gate, double precision,
                          double r_norm, int n_max_iter);
bool launch_pcg_init_zero(int n, ItState *st, const double *rhs, const double *ad_inv,
                          double *vx, double *rk, double *dk);
bool launch_pcg_update_first(int n, ItState *st, const double *ad_inv, const double *dk,
                             const double *zk, double *vx, double *rk, double *gk);
bool launch_pcg_update(int n, ItState *st, const double *ad_inv, const double *sk,
                       double *dk, double *gk, double *zk, double *rk, double *vx);
/* whole PCG in one thread block for small systems */
bool launch_pcg_single_block(const _csb_matrix_t *m, ItState *st, const MgState *gate,
                             double precision, double r_norm, int n_max_iter,
                             const double *ad_inv, const double *rhs, double *vx,
                             int vx_ini_nonzero, double *work /* 5*n */);

/* Jacobi (cs_sles_it.cpp:2098-2270) */
bool launch_jacobi_first(int n, ItState *st, const double *rhs, const double *ad_inv,
                         double *vx, double *rk);
bool launch_jacobi_sweep(const _csb_matrix_t *m, ItState *st, const double *rhs,
                         const double *ad_inv, double *vx, double *rk);
bool launch_copy(int n, const double *src, double *dst, const ItState *st);

/* BiCGStab (cs_sles_it.cpp:2693-2880) */
bool launch_bicg_init(int n, ItState *st, const double *rhs, double *vx, double *res0,
                      double *rk, double *pk, double *uk, bool have_ax);
bool launch_bicg_dots_a(int n, ItState *st, const double *rk, const double *res0);
bool launch_bicg_pk_zk(int n, ItState *st, const double *rk, const double *uk,
                       const double *ad_inv, double *pk, double *zk);
bool launch_bicg_spmv_b(const _csb_matrix_t *m, ItState *st, const double *zk, double *uk,
                        const double *res0);
bool launch_bicg_upd1(int n, ItState *st, const double *ad_inv, double *vx, double *rk,
                      double *zk, const double *uk);
bool launch_bicg_spmv_c(const _csb_matrix_t *m, ItState *st, const double *zk, double *vk,
                        const double *rk);
bool launch_bicg_upd2(int n, ItState *st, double *vx, double *rk, const double *zk,
                      const double *vk, const double *res0);

/* Multigrid smoothers without convergence test (cs_multigrid_smoother.cpp) */
bool launch_l1_inv(const _csb_matrix_t *m, double *l1_inv);   /* cs_sles_it_priv.cpp:231-262 */
/* first sweep from a zero start: vx = rhs*ad_inv, or w0*rhs*ad_inv when relaxed */
bool launch_smoother_first(int n, int smooth_mode, double w0, const double *rhs,
                           const double *ad_inv, double *vx, const MgState *gate);
bool launch_copy_gated(int n, const double *src, double *dst, const MgState *gate);
/* mirror the outer Krylov solver's state into the gate of its multigrid preconditioner */
bool launch_gate_from_it(MgState *gate, const ItState *st);

/* 3x3 block Jacobi (cs_sles_it.cpp:2284-2440): block inverses (cs_math.h:540-586), one
   sweep after vxx = E.vx with the residual of the previous iterate */
bool launch_gauss_inverse_33(int n_blocks, const double *ad, double *ad_inv);
bool launch_b3_jacobi_update(int n_blocks, ItState *st, const double *ad, const double *ad_inv,
                             const double *rhs, const double *vxx, double *vx, bool first);

/* single-reduction PCG with a general preconditioner (cs_sles_it.cpp:1229-1415) */
bool launch_pcg_pc_init(int n, const double *rhs, double *vx, double *rk, bool have_ax);
bool launch_pcg_pc_update_first(int n, ItState *st, const double *dk, const double *zk,
                                double *vx, double *rk);
bool launch_pcg_pc_update(int n, ItState *st, const double *gk, const double *sk, double *dk,
                          double *zk, double *rk, double *vx);

/* K-cycle vector steps (cs_multigrid.cpp:4300-4470): mode 0 y = b - a*x, 1 y = a*y,
   2 y = a*y + c*x, 3 y += x + b; overwrite-form prolongation (cs_grid.cpp:9180-9190) */
bool launch_kc_vec(int mode, int n, double a, double c, const double *x, const double *b,
                   double *y);
/* the modes 0..2 with a (and c) read from device memory; the K-cycle's inner Krylov
   coefficients computed on the device from ctx().red_out (kernels.cu k_kc_scalars) */
bool launch_kc_vec_d(int mode, int n, const double *pa, const double *pc, const double *x,
                     const double *b, double *y);
bool launch_kc_scalars(int mode, double *sc);
bool launch_prolong_set(int f_n_rows, const int *coarse_row, const double *c_var, double *f_var);

/* IPCG with a general preconditioner (cs_sles_it.cpp:1016-1203); the step length comes out of
   launch_spmv(EPI_PCG_DOT3) with phase = 2 */
bool launch_ipcg_dots(int n, ItState *st, bool first, const double *rk, const double *gk,
                      const double *rkm1, double *dk);
bool launch_ipcg_dir(int n, ItState *st, const double *gk, double *dk);
bool launch_ipcg_update(int n, ItState *st, bool first, const double *dk, const double *zk,
                        double *vx, double *rk, double *rkm1);

/* FCG with a general preconditioner, kernel per phase (cs_sles_it.cpp:832-985) */
bool launch_fcg_init(int n, ItState *st, const double *rhs, double *vx, double *rk,
                     double *qk, bool have_residual);
bool launch_fcg_update(int n, ItState *st, const double *vk, const double *wk, double *dk,
                       double *qk, double *rk, double *vx);

/* multigrid transfer operators (cs_grid.cpp:8987-9210) */
bool launch_restrict(int c_n_rows, const int *c_f_row_index, const int *c_f_row_ids,
                     const double *f_var, double *c_var, const MgState *gate);
bool launch_prolong_inc(int f_n_rows, const int *coarse_row, const double *c_var,
                        double *f_var, const MgState *gate);
bool launch_mg_state_init(MgState *mg, double precision, double r_norm, int n_max_cycles);
bool launch_mg_cycle_begin(MgState *mg, int cycle_id);
bool launch_mg_set_initial_residual(MgState *mg, double v);
bool launch_mg_accum_iter(MgState *mg, const ItState *st, int level, double n_rows_ratio,
                          int take_initial_residual);
/* same for a smoother whose sweep count is known: += n_iter * n_rows * denom */
bool launch_mg_accum_n(MgState *mg, int n_iter, int n_rows, double denom_n_g_rows_0);

/* coarse level assembly (cs_grid.cpp:6226-6590, :6841-6897) */
struct CoarseAssemblyArgs {
  int  f_level;
  int  f_n_rows, c_n_rows, c_n_faces;
  const int *f_row_index, *f_col_id;
  const double *f_d_val, *f_x_val;
  const int *f_c_row;                    /* fine -> coarse */
  const int *c_f_row_index, *c_f_row_ids;
  const int *c_row_index, *c_col_id;
  const int *f_cell_face; const short *f_cell_face_sgn;
  const int *c_cell_face; const short *c_cell_face_sgn;
  const double *f_cell_cen, *f_cell_vol;
  const double *f_face_u_normal, *f_face_surf;   /* level 0 */
  const double *f_face_normal, *f_xa0, *f_xa0ij; /* level > 0 */
  double *c_d_val, *c_x_val;
  double *c_cell_cen, *c_cell_vol, *c_face_normal, *c_xa0, *c_xa0ij;
  double relax;
};
bool launch_coarse_assembly(const CoarseAssemblyArgs &a, int stages /* 1 | 2 */);
/* relaxation parameter <= 0: plain Galerkin sums, no mesh quantities (cs_grid.cpp:5969-6090);
   uses the matrix, f_c_row and c_f_* members of the arguments only */
bool launch_coarse_galerkin(const CoarseAssemblyArgs &a);

/* coarse level of a matrix mapped from native (face-based) arrays: the sums of
   _compute_coarse_cell_quantities / _compute_coarse_quantities_native (cs_grid.cpp:4778,
   :5144-5558) run in FACE order there; here a thread owns a coarse face (a fine row, a
   coarse row) and takes its contributions in ascending face order */
struct NativeCoarseArgs {
  int  f_level;
  int  f_n_rows, c_n_rows, c_n_faces;
  const int *f_row_index, *f_col_id;
  const double *f_d_val, *f_x_val;
  const int *f_cell_face;                /* fine entry -> fine face */
  const int *f_face_cell;                /* fine faces, 2 cells each */
  const int *f_c_row;
  const int *c_f_row_index, *c_f_row_ids;
  const int *c_row_index, *c_cell_face;  /* coarse entry -> coarse face */
  const int *c_face_cell;
  const double *f_cell_cen, *f_cell_vol;
  const double *f_face_u_normal, *f_face_surf;   /* level 0 */
  const double *f_face_normal, *f_xa0, *f_xa0ij; /* level > 0 */
  double *c_cell_cen, *c_cell_vol, *c_face_normal, *c_xa0, *c_xa0ij, *c_xa;
  double *w1;                            /* f_n_rows values of scratch */
  double *c_d_val, *c_x_val;
  double relax;
};
bool launch_coarse_native(const NativeCoarseArgs &a);
/* x_val[k] = xa[c2f[k]] (cs_matrix_set_coefficients from edges, symmetric, unique faces) */
bool launch_native_to_msr(int nnz, const int *c2f, const double *xa, double *x_val);

bool launch_diffusion_potential(const _csb_matrix_t *m, int init, int inc, const double *pvar,
                                const double *cofaf, const double *cofbf, const double *i_visc,
                                const double *b_visc, double *diverg);
/* least-squares cell gradient (cs_gradient.cpp:2501-2680, :2695-2900) and the reconstructed
   form of cs_diffusion_potential (cs_convection_diffusion.cpp:8895-8982) */
bool launch_lsq_cocg(const _csb_matrix_t *m, float *cocg);
/* launch_b_iprime_lsq: values at I' of the boundary faces (cs_gradient_boundary.cpp:197-640)
   and the face values derived from them (val_f, may be nullptr); needed when the matrix holds
   II' of its boundary faces or when f_ext is given (hyd_p_flag = 1: exterior force
   [n_rows][3], needs the face centres of csb_matrix_set_face_cog).  The gradient kernels take
   val_f, or nullptr for inc*coefa + coefb*pvar. */
bool launch_b_iprime_lsq(const _csb_matrix_t *m, int inc, const double *pvar,
                         const double *coefa, const double *coefb, const double *f_ext,
                         double clip_coeff, double *val_ip, double *val_f);
/* gradient limiter (_scalar_gradient_clipping, cs_gradient.cpp:808-1215; clip_mode 0 / 1 / 2)
   applied to grad in place; val_f nullptr: inc*coefa + coefb*pvar; work: 2 n_rows doubles */
bool launch_gradient_clip(const _csb_matrix_t *m, int clip_mode, double climgp, int inc,
                          const double *pvar, const double *coefa, const double *coefb,
                          const double *val_f, double *grad, double *work);
bool launch_lsq_gradient(const _csb_matrix_t *m, const float *cocg, int inc, const double *pvar,
                         const double *coefa, const double *coefb, const double *val_f,
                         const double *f_ext, double *grad);
/* iterative Green-Gauss gradient (imrgra = 0): cocg_it [n_rows][3][3]; one sweep (r_grad
   nullptr: the initial non-reconstructed gradient into out, else rhs / V into out); the update
   grad += cocg_it . rhs */
bool launch_cocg_it(const _csb_matrix_t *m, double *cocg_it);
bool launch_green_iter_sweep(const _csb_matrix_t *m, int inc, const double *pvar,
                             const double *coefa, const double *coefb, const double *r_grad,
                             double *out);
bool launch_green_iter_update(const _csb_matrix_t *m, const double *cocg_it, const double *rhs,
                              double *grad);
/* cs_divergence (cs_divergence.cpp:1224) and cs_face_diffusion_potential
   (cs_convection_diffusion.cpp:8075; grad nullptr: no reconstruction) */
bool launch_divergence(const _csb_matrix_t *m, int init, const double *i_massflux,
                       const double *b_massflux, double *diverg);
bool launch_face_diffusion_potential(const _csb_matrix_t *m, int init, int inc,
                                     const double *pvar, const double *grad,
                                     const double *cofaf, const double *cofbf,
                                     const double *i_visc, const double *b_visc,
                                     const double *visel, const double *val_ip,
                                     double *i_massflux, double *b_massflux);
/* second half of the Green-Gauss gradient with least-squares face values (imrgra = 4,
   _reconstruct_scalar_gradient cs_gradient.cpp:4167): r_grad = the least-squares gradient */
bool launch_green_reconstruct(const _csb_matrix_t *m, int inc, const double *pvar,
                              const double *coefa, const double *coefb, const double *val_f,
                              const double *f_ext, const double *r_grad, double *grad);
bool launch_diffusion_potential_rc(const _csb_matrix_t *m, int irecmf, int init, int inc,
                                   const double *pvar, const double *grad, const double *cofaf,
                                   const double *cofbf, const double *i_visc,
                                   const double *b_visc, const double *visel,
                                   const double *val_ip, double *diverg);
/* matrix assembler (assembler.cu): slot s < n_rows is the diagonal of row s, slot
   n_rows + k the extra-diagonal entry k; every slot adds val[src[j]] for j in
   [index[s], index[s+1]) to its current value, in that order */
bool launch_assembler_fold(int n_rows, int nnz, const int *index, const int *src,
                           const double *val, double *d_val, double *x_val);

bool launch_halo_pack(int n_send, const int *send_list, const double *var, double *buf);
/* Gauss-Seidel update of the n rows listed in `rows` (one dependency level); mode 0 whole
   row, 1 lower part, 2 upper part (truncated sweeps); res != NULL: d_i (new - old) per row
   (solver form); launch_gs_chain: levels l0 .. l1-1 of a plan in one CTA */
bool launch_gs_rows(int n, const int *rows, const _csb_matrix_t *m, const double *ad_inv,
                    const double *rhs, double *vx, int mode, const MgState *gate,
                    const ItState *st, double *res);
bool launch_gs_chain(const int *rows, const int *level_start, int l0, int l1, bool backward,
                     const _csb_matrix_t *m, const double *ad_inv, const double *rhs,
                     double *vx, int mode, const MgState *gate, const ItState *st, double *res);
bool launch_gs_residual(int n, ItState *st, const double *res);
/* index helpers of the coarse-level merge: 0 out = shift + i, 1 out = map[in], 2 row
   lengths of a row index, 3 in + shift where in >= 0 */
bool launch_merge_index(int mode, int n, int shift, const int *in, const int *map, int *out);

} // namespace csb

#endif
