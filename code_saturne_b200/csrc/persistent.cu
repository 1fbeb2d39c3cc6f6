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

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace csb {

namespace {

/* CTA shape.  The kernels are HBM-latency bound unless every SM keeps ~45 KB of
   loads in flight (6.5 TB/s x ~1 us / 148 SMs).  Streams are kept in flight by
   bulk copies into a shared memory ring (no registers, no scoreboard), so the
   shape is "few fat CTAs": PT threads x MIN_CTAS per SM, 128 registers each for
   the two alternating sets of x[col] gathers, ~113 KB of shared memory per CTA. */
#ifndef CSB_PT
#define CSB_PT 256
#endif
#ifndef CSB_MIN_CTAS
#define CSB_MIN_CTAS 2
#endif
constexpr int PT = CSB_PT;         /* threads per CTA */
constexpr int NW = PT / 32;        /* warps per CTA */
#ifndef CSB_STAGES
#define CSB_STAGES 3
#endif
constexpr int NS = CSB_STAGES;     /* shared memory stages of the bulk-copy pipeline */
static_assert(NS >= 2 && NS <= 4, "pipeline depth");
/* In the streaming phases the last warp of the CTA is the producer (it only issues bulk
   copies); the other NWC warps consume: one row (SpMV) or two rows (elementwise) per lane. */
constexpr int NWC = NW - 1;
constexpr int SLAB_ROWS_MAX = NWC*32;
#ifndef CSB_SLAB_CAP
#define CSB_SLAB_CAP 2432
#endif
constexpr int SLAB_CAP = CSB_SLAB_CAP;   /* matrix entries of a slab (shared memory budget) */
static_assert(SLAB_CAP % 4 == 0, "slab capacity");
#ifndef CSB_GATHER
#define CSB_GATHER 12
#endif
constexpr int NG = CSB_GATHER;     /* x[col] gathers a lane keeps in flight for its next row */
static_assert(PT % 32 == 0 && NW <= 32, "CTA shape");
static_assert(4*NW + 4 <= 56 && 56 + 4*COMM_MAX_RANKS <= PERSISTENT_S_RED, "s_red layout");

#ifdef CSB_PHASE_TRACE
/* Debug build only (scripts/build_variants.sh "trace:-DCSB_PHASE_TRACE"): where the time of
   a small-level SpMV + dots phase goes, seen by lane 0 of CTA 0 on levels with
   TRACE_LO < n_rows <= TRACE_HI.  Sums of nanoseconds: [0] entry -> first stage full,
   [1] -> first gathers back, [2] -> own slabs consumed, [3] -> CTA joined, [4] reduction
   (all_sum), [5] whole phase; [15] number of phases. */
__device__ unsigned long long g_trace[16];
#ifndef TRACE_LO
#define TRACE_LO 5000
#endif
#ifndef TRACE_HI
#define TRACE_HI 10000
#endif
__device__ __forceinline__ unsigned long long trace_now()
{ unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

/* Layout of the head of a CTA's dynamic shared memory (the stages follow, see below). */
extern __shared__ __align__(128) unsigned char s_raw_base[];
constexpr int HALO_K_OFF = PERSISTENT_S_RED*8;        /* unsigned[PROF_LEVELS]: exchange counts */
constexpr int BAR_TGT_OFF = HALO_K_OFF + PROF_LEVELS*4;   /* unsigned: arrivals at the next barrier */
constexpr int PIPE_OFF = BAR_TGT_OFF + 16;            /* PipeState */

/*--------------------------------------------------------------------------
 * Grid-wide barrier (all CTAs are co-resident: cooperative launch).
 * bar[0] is an arrival counter that only grows (reset to 0 by the host before every
 * launch): barrier k is passed when it reaches k*gridDim.x.  One atomic per CTA and one
 * word to poll -- no "last CTA resets and bumps a generation" hop, no sleep in the poll.
 * Every CTA keeps the target of its next barrier in shared memory (all CTAs execute the
 * same sequence of barriers).
 *--------------------------------------------------------------------------*/

/* rows in flight per thread in the transfer operators (independent gather chains) */
#ifndef CSB_PROLONG_U
#define CSB_PROLONG_U 8
#endif
#ifndef CSB_RESTRICT_U
#define CSB_RESTRICT_U 6
#endif

/* Release / acquire fences of the message-passing patterns below.  __threadfence() is a
   sequentially-consistent fence (SASS MEMBAR.SC + ERRBAR); everything here is "publish data,
   then a flag / read the flag, then the data", for which acq_rel (MEMBAR.ALL) is enough. */
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

/* longer than any peer-flag timeout (CSB_SPIN_TIMEOUT_S, default 10 s): the peer wait reports first */
constexpr unsigned long long GRID_BARRIER_TIMEOUT_NS = 120ull*1000000000ull;

__device__ __forceinline__ void grid_barrier(unsigned *bar)
{
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned *tgt = (unsigned *)(s_raw_base + BAR_TGT_OFF);
    const unsigned target = *tgt + gridDim.x;
    *tgt = target;
    fence_gpu();                     /* release: this CTA's stores (bar.sync-cumulative) */
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(bar) : "memory");
    unsigned seen, n_poll = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if ((int)(seen - target) >= 0) break;
      /* a CTA that never arrives (a fault elsewhere, a peer rank that died inside its own
         wait) must not hang the device: after GRID_BARRIER_TIMEOUT_NS the kernel traps and
         the launch returns an error, like the peer-flag waits (spin_until_ge) */
      if ((++n_poll & 0x3FFFu) == 0) {
        const unsigned long long now = gtime_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > GRID_BARRIER_TIMEOUT_NS) {
          printf("[csb200] grid barrier timeout (block %d): have %u, expect %u\n",
                 blockIdx.x, seen, target);
          __trap();
        }
      }
    }
    fence_gpu();                     /* acquire */
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
__device__ __forceinline__ void cta_sum_bcast(double (&v)[Q], double *s_red /* Q*NW + Q */)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const double w = warp_sum(v[q]);
    if (lane == 0) s_red[q*NW + wid] = w;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      double w = (lane < NW) ? s_red[q*NW + lane] : 0.0;
      w = warp_sum(w);
      if (lane == 0) s_red[Q*NW + q] = w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < Q; q++) v[q] = s_red[Q*NW + q];
  __syncthreads();
}

/* CTA-wide sum of Q values, result in thread 0 only (one barrier). */
template <int Q>
__device__ __forceinline__ void cta_sum_t0(double (&v)[Q], double *s_red /* Q*NW */)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const double w = warp_sum(v[q]);
    if (lane == 0) s_red[q*NW + wid] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < Q; q++) {
      double w = 0.;
      for (int i = 0; i < NW; i++) w += s_red[q*NW + i];
      v[q] = w;
    }
  }
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
  if (SINGLE) cta_sum_bcast<Q>(v, s_red);
  else {
    const unsigned G = gridDim.x;
    cta_sum_t0<Q>(v, s_red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < Q; q++) partials[q*G + blockIdx.x] = v[q];
    }
    grid_barrier(bar);
    /* all loads of a thread are issued before the first add (one L2 latency, not
       one per term); thread t sums partials t, t + PT, ... in that order */
    {
      double w[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) w[q] = 0.0;
      for (unsigned i0 = threadIdx.x; i0 < G; i0 += 2*PT) {
        double a[Q], b[Q];
        const unsigned i1 = i0 + PT;
#pragma unroll
        for (int q = 0; q < Q; q++) {
          a[q] = __ldcg(partials + q*G + i0);
          b[q] = (i1 < G) ? __ldcg(partials + q*G + i1) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < Q; q++) { w[q] += a[q]; if (i1 < G) w[q] += b[q]; }
      }
#pragma unroll
      for (int q = 0; q < Q; q++) v[q] = w[q];
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
    fence_sys();
    for (int r = 0; r < R; r++)
      *(volatile unsigned *)&((CommHeader *)C->peer[r])->xflag[par][me] = e;
  }
  double *s_x = s_red + 56;        /* [Q][COMM_MAX_RANKS] */
  if (threadIdx.x < R) {
    spin_until_ge(&H->xflag[par][threadIdx.x], e, C->timeout_ns, 6);
    fence_gpu();
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

/* Exchanges done so far, per level: a copy private to the CTA (shared memory), loaded
   from DistLevel::count at kernel start (halo_counts_init) so that a CTA that reaches an
   exchange late cannot read a count that the packing CTAs have already advanced. */

template <bool SINGLE>
__device__ __forceinline__ const double *halo_exchange(const MgLevelDesc &L, const double *x)
{
  const DistLevel *D = L.dist;
  if (D == nullptr) return nullptr;
  const int n_dom = D->n_dom;
  if (n_dom == 0) return nullptr;
  unsigned *s_hk = (unsigned *)(s_raw_base + HALO_K_OFF) + L.level_id;
  const unsigned k = *s_hk + 1;
  const int par = k & 1;
  /* only as many CTAs as the send list needs pack, fence and take a ticket: on the
     coarse levels that is one or two, and the exchange costs one NVLink round trip
     instead of a grid-wide ticket */
  const int n_send = D->send_index[n_dom];
  const unsigned n_pack = SINGLE ? 1u
    : (unsigned)min((int)gridDim.x, max(1, (n_send + PT - 1) / PT));
  if (SINGLE || blockIdx.x < n_pack) {
    const int first = SINGLE ? threadIdx.x : blockIdx.x*PT + threadIdx.x;
    const int step = SINGLE ? PT : (int)n_pack*PT;
    const int *__restrict__ send_list = D->send_list;
    for (int i = 0; i < n_dom; i++) {
      const int s0 = D->send_index[i], s1 = D->send_index[i+1];
      double *dst = D->push_dst[i][par];
      for (int e = s0 + first; e < s1; e += step)
        dst[e - s0] = x[send_list[e]];
    }
    /* one system-scope fence per CTA, by the thread that takes the ticket: the CTA barrier
       orders the other threads' peer stores before it (cumulativity); 256 concurrent
       MEMBAR.SYS per CTA were a large part of the exchange's cost at 8 GPUs */
    __syncthreads();
    if (threadIdx.x == 0) {
      fence_sys();
      bool last = true;
      if (!SINGLE) last = (atomicInc(D->ticket, n_pack - 1) == n_pack - 1);
      if (last) {
        *(volatile unsigned *)D->count = k;
        fence_sys();
        for (int i = 0; i < n_dom; i++)
          *(volatile unsigned *)D->flag_remote[i] = k;
      }
    }
  }
  /* the wait for the neighbours' values is the SpMV phase's (halo_wait): with
     D->overlap it comes after the rows that need no ghost value have been summed */
  return D->gbuf[par];
}

/* second half of the exchange started by halo_exchange: wait until every neighbour has
   published this exchange's count */
__device__ __forceinline__ void halo_wait(const MgLevelDesc &L)
{
  const DistLevel *D = L.dist;
  unsigned *s_hk = (unsigned *)(s_raw_base + HALO_K_OFF) + L.level_id;
  const unsigned k = *s_hk + 1;
  if (threadIdx.x < D->n_dom) {
    spin_until_ge(D->flag_local + threadIdx.x, k, D->timeout_ns, 7);
    fence_gpu();
  }
  __syncthreads();
  if (threadIdx.x == 0) *s_hk = k;      /* read again only after further CTA barriers */
  __syncthreads();
}

__device__ __forceinline__ unsigned halo_count(const MgLevelDesc &L)
{
  return (L.dist != nullptr) ? ((const unsigned *)(s_raw_base + HALO_K_OFF))[L.level_id] : 0u;
}

/* After a level was run by CTA 0 alone (grid barrier passed): the other CTAs advance
   their exchange count of the level by what CTA 0 did. */
__device__ __forceinline__ void halo_count_follow(const MgLevelDesc &L, const ItState *st)
{
  if (L.dist == nullptr) return;
  if (blockIdx.x != 0 && threadIdx.x == 0)
    ((unsigned *)(s_raw_base + HALO_K_OFF))[L.level_id] += (unsigned)*((volatile int *)&st->n_halo);
  __syncthreads();
}

/* kernel start, levels with a ghost exchange: load the exchange counts, then a grid
   barrier (no CTA starts an exchange before every CTA holds its copy) */
__device__ __forceinline__ void halo_counts_init(const MgLevelDesc *levels, int n_levels,
                                                 unsigned *bar)
{
  unsigned *s_hk = (unsigned *)(s_raw_base + HALO_K_OFF);
  bool any = false;
  for (int l = 0; l < n_levels && l < PROF_LEVELS; l++) {
    const DistLevel *D = levels[l].dist;
    if (D != nullptr) {
      any = true;
      if (threadIdx.x == 0) s_hk[levels[l].level_id] = *(volatile unsigned *)D->count;
    }
  }
  if (any) grid_barrier(bar);
  else __syncthreads();
}

/*--------------------------------------------------------------------------
 * Shared memory of a CTA and the bulk-copy (TMA) pipeline.
 *
 * Everything that is *streamed* (matrix entries, row index, the vector
 * operands of an epilogue or of an elementwise update) reaches the SM through
 * cp.async.bulk copies issued by one thread into an NS-stage ring of shared
 * memory buffers, each guarded by an mbarrier (complete_tx byte counting): the
 * copies need no registers and no scoreboard, so NS-1 slabs per CTA are in
 * flight while one is consumed.  Only the x[col] gathers (L2 hits) go through
 * registers, issued one slab ahead of their use.
 *
 *   [ s_red: PERSISTENT_S_RED doubles | PipeState | NS x stage_bytes ]
 *   SpMV stage : col_id (cap+4 int) | x_val (cap+2 double) | row_index (PT+4 int)
 *                | 4 operand vectors (PT doubles each)
 *   row stage  : up to 7 operand vectors of ROW_CH doubles each
 * Bulk copies need 16-byte aligned addresses and sizes: a slab's entry range is
 * widened to the enclosing aligned range (k0 & ~3 ints, k0 & ~1 doubles) and
 * sizes are rounded up; device arrays are allocated with 64 bytes of slack.
 *--------------------------------------------------------------------------*/

struct PipeState {
  unsigned long long full[4];   /* stage filled: 1 arrival (expect_tx) + the copied bytes */
  unsigned long long empty[4];  /* stage consumed: one arrival per consumer warp */
  unsigned           base;      /* stages filled so far by this CTA (stage = count % NS) */
  unsigned           pad[15];
};

constexpr int SMEM_HDR = (PIPE_OFF + (int)sizeof(PipeState) + 127) / 128 * 128;
constexpr int ST_RI_BYTES = ((SLAB_ROWS_MAX + 1 + 3)*4 + 15) / 16 * 16;
constexpr int ST_OP_STRIDE = SLAB_ROWS_MAX;          /* doubles between staged operand vectors */
constexpr int ST_FIXED = 16 + 16 + ST_RI_BYTES + 4*ST_OP_STRIDE*8;
constexpr int ROW_CH = 2*SLAB_ROWS_MAX;              /* rows per chunk of an elementwise phase */
constexpr int ROW_STAGE_BYTES = 7*ROW_CH*8;

__device__ __forceinline__ unsigned smem_u32(const void *p)
{ return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  const unsigned addr = smem_u32(bar);
  unsigned done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}

/* global -> shared bulk copy; bytes > 0, multiple of 16; both addresses 16-byte aligned */
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes,
                                          unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* generic-proxy accesses (other CTAs' global stores seen through the grid barrier, this
   CTA's shared memory reads) are ordered before the async-proxy copies that follow */
__device__ __forceinline__ void fence_proxy_async()
{ asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned dyn_smem_bytes()
{
  unsigned v;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(v));
  return v;
}

struct SmemView {
  PipeState     *pipe;
  unsigned char *stages;
  int            stage_bytes;
  int            cap;            /* entries of a SpMV stage */
};

__device__ __forceinline__ SmemView smem_view(unsigned char *s_raw)
{
  SmemView v;
  v.pipe = (PipeState *)(s_raw + PIPE_OFF);
  v.stages = s_raw + SMEM_HDR;
  v.stage_bytes = (int)((dyn_smem_bytes() - SMEM_HDR) / NS) & ~127;
  v.cap = ((v.stage_bytes - ST_FIXED) / 12) & ~3;
  return v;
}

/* once per kernel */
__device__ __forceinline__ void pipe_init(unsigned char *s_raw)
{
  PipeState *pipe = (PipeState *)(s_raw + PIPE_OFF);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; i++) { mbar_init(&pipe->full[i], 1); mbar_init(&pipe->empty[i], NWC); }
    pipe->base = 0;
    *(unsigned *)(s_raw + BAR_TGT_OFF) = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

/* Fill number q (counted per CTA over the whole kernel) goes to stage q % NS and is the
   (q / NS)-th use of it: the producer first waits until use q/NS - 1 has been released
   by all consumer warps; consumers wait for the fill, read, then release. */
/* The release (mbarrier arrive by the consumer warps) / acquire (mbarrier wait here) pair
   orders the consumers' reads of a stage before the bulk copy that refills it: no proxy
   fence per stage.  Round 1 had a fence.proxy.async here; it also waits for the producer's
   own row_index loads in flight -- one DRAM latency per slab on the one thread every
   consumer waits for (ncu, profiles/r02: half of the producer warp's samples on MEMBAR /
   FENCE.VIEW.ASYNC, consumers 35 % of theirs waiting for a stage to fill) and capped the
   SpMV phases at ~0.75 of the HBM rate; without it they run at 0.86-0.92.  One proxy fence
   per phase, before the first bulk copy, still orders what other CTAs stored ahead of the
   grid barrier before this phase's bulk reads. */
__device__ __forceinline__ void producer_acquire(const SmemView &sv, unsigned q)
{
  const unsigned u = q / NS;
  if (u > 0) mbar_wait(&sv.pipe->empty[q % NS], (u - 1) & 1u);
}
__device__ __forceinline__ void consumer_wait(const SmemView &sv, unsigned q)
{ mbar_wait(&sv.pipe->full[q % NS], (q / NS) & 1u); }
__device__ __forceinline__ void consumer_release(const SmemView &sv, unsigned q)
{
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&sv.pipe->empty[q % NS]);
}

/*--------------------------------------------------------------------------
 * SpMV over the slabs owned by this CTA, fused with an epilogue.
 * A slab = L.slab_rows (<= SLAB_ROWS_MAX) consecutive rows; its off-diagonal
 * entries are contiguous in memory.  Warp-specialised, no CTA-wide barrier
 * inside the loop:
 *   producer warp  - for every slab: wait until the stage is released, then bulk-copy
 *                    col_id / x_val / row_index / the epilogue's vector operands;
 *   consumer warps - lane = row.  The x[col] gathers of the lane's row in the NEXT
 *                    slab are issued (L2 hits; consecutive rows read consecutive x on
 *                    a structured mesh: coalesced) before the row of THIS slab is
 *                    summed **in CSR order** (the reference's order,
 *                    cs_matrix_spmv.cpp:2138-2142) and passed to the epilogue.
 * F (by value) supplies NV (staged vectors), src(j) (their base pointers) and
 *   void epi(int row, double off_diagonal_sum, const double *op, double (&acc)[4])
 * where op[j*ST_OP_STRIDE] is the value of staged vector j at this row.
 * acc_out[0..3] receives this thread's partial sums.
 *--------------------------------------------------------------------------*/

/* number of staged operand vectors of an epilogue: F::NV, or f.nv() when it depends on a
   run-time mode (EpiSmooth) */
template <class F>
__device__ __forceinline__ auto epi_nv(const F &f, int) -> decltype(f.nv()) { return f.nv(); }
template <class F>
__device__ __forceinline__ int epi_nv(const F &, long) { return F::NV; }

/* GD = gather depth: register sets of x[col] values a lane keeps in flight.  2 (NG values
   each): the gathers of the next slab overlap the sums of this one.  3 (NG3 values each,
   levels whose rows have at most NG3 entries, e.g. the 7-point fine level): gathers are
   issued TWO slabs ahead, so that a loaded L2 latency of ~2 slab times is covered; needs a
   fourth stage (-DCSB_STAGES=4 -DCSB_SLAB_CAP=1648).  Measured (profiles/r02): level 0
   0.91 against 0.88 with two sets, but the smaller stages cost the denser levels more --
   the default build keeps two sets and three stages. */
constexpr int NG3 = 8;

/* GHOST: 0 no ghost columns; 1 the exchange is waited for first and ghost columns are
   gathered in line, from the ghost buffer (the default: measured faster, see sles.cu
   build_dist_level); 2 rows with ghost columns are parked and finished after the wait
   (the overlapped form, CSB_HALO_OVERLAP=1) */
/* NOVAL: the matrix is a 0/1 matrix given by its structure alone (the restriction: rows of the
   next level, columns = this level's rows, cs_grid.cpp:9025-9030): no value array is staged
   and a row sum is the sum of the gathered x values in CSR order. */
template <bool SINGLE, int GHOST, int GD, bool NOVAL, class F>
__device__ __noinline__ void spmv_phase_gv(const MgLevelDesc &L, const double *x,
                                           const double *ghost, unsigned char *s_raw,
                                           const F f, double *acc_out)
{
  constexpr int NGS = (GD == 3) ? NG3 : NG;      /* gathers per register set */
  constexpr int PARKED = 0x40000000;             /* flag carried in the row's start offset */
  double *bsum = (GHOST == 2) ? L.dist->bsum : nullptr; /* parked row sums, by row */
  /* in-line form: column c < n_rows reads x[c], a ghost column ghost[c - n_rows] */
  const double *xg = (GHOST == 1) ? ghost - L.n_rows : x;
  const int tid = threadIdx.x;
  const int first = SINGLE ? 0 : blockIdx.x;
  const int step = SINGLE ? 1 : gridDim.x;
  const int n_slabs = L.n_slabs, slab_rows = L.slab_rows, n_rows = L.n_rows;
  const SmemView sv = smem_view(s_raw);
  const int n_my = (first < n_slabs) ? (n_slabs - first + step - 1) / step : 0;
  const unsigned base = sv.pipe->base;
  const int off_val = 4*(sv.cap + 4), off_ri = off_val + 8*(sv.cap + 2),
            off_ops = off_ri + ST_RI_BYTES;
  double acc[4] = {0., 0., 0., 0.};

  /* GHOST: the exchange of x has been STARTED (halo_exchange); its values are needed by the
     rows that have ghost columns only (a few per cent: the faces of the brick).  Those rows
     are summed over their local columns here -- ghost columns come last in a row, so this is
     the head of the row's CSR-order sum -- and parked in bsum; everything else runs to its
     epilogue while the neighbours' stores cross NVLink.  The wait comes after the main
     pass, then the parked rows get their ghost terms and their epilogue (the reference's
     overlap of halo exchange and local product, _mat_vec_p_l_dist
     cs_matrix_spmv.cpp:3092-3112).  That is GHOST == 2; GHOST == 1 waits first and reads the
     ghost columns in line. */
  if (GHOST == 1) halo_wait(L);

  if (n_my > 0) {
    if (tid >= NWC*32) {
      /* ---------------- producer ---------------- */
      if (tid == NWC*32) {
        fence_proxy_async();
        const int    *__restrict__ row_index = L.row_index;
        const int    *col_id = L.col_id;
        const double *x_val = L.x_val;
        const int n_ops = epi_nv(f, 0);
        int k0 = __ldg(row_index + first*slab_rows);
        int k1 = __ldg(row_index + min(first*slab_rows + slab_rows, n_rows));
        for (int j = 0; j < n_my; j++) {
          const int r0 = (first + j*step) * slab_rows;
          int k0n = 0, k1n = 0;
          if (j + 1 < n_my) {
            const int r0n = r0 + step*slab_rows;
            k0n = __ldg(row_index + r0n);
            k1n = __ldg(row_index + min(r0n + slab_rows, n_rows));
          }
          const unsigned q = base + (unsigned)j;
          unsigned char *sb = sv.stages + (q % NS)*sv.stage_bytes;
          unsigned long long *bar = &sv.pipe->full[q % NS];
          const int rows = min(slab_rows, n_rows - r0);
          const int ka = k0 & ~3, nc = (k1 - ka + 3) & ~3;
          const int kb = k0 & ~1, nv = (k1 - kb + 1) & ~1;
          const unsigned ri_b = (unsigned)(((rows + 1)*4 + 15) & ~15);
          const unsigned v_b = (unsigned)((rows*8 + 15) & ~15);
          producer_acquire(sv, q);
          mbar_expect_tx(bar, 4u*nc + (NOVAL ? 0u : 8u*nv) + ri_b + (unsigned)n_ops*v_b);
          if (nc > 0) bulk_load(sb, col_id + ka, 4u*nc, bar);
          if (!NOVAL && nv > 0) bulk_load(sb + off_val, x_val + kb, 8u*nv, bar);
          bulk_load(sb + off_ri, row_index + r0, ri_b, bar);
#pragma unroll
          for (int v = 0; v < F::NV; v++)
            if (v < n_ops) bulk_load(sb + off_ops + v*ST_OP_STRIDE*8, f.src(v) + r0, v_b, bar);
          k0 = k0n; k1 = k1n;
        }
      }
      __syncwarp();
    }
    else {
      /* ---------------- consumers: lane = row ---------------- */
      /* Two register sets (A, B) alternate between "being gathered for the next slab" and
         "being consumed": the loop body handles two slabs, so no set is ever copied and
         every element is indexed by a compile-time constant (they stay in registers). */
      double gA[NGS], gB[NGS], gC[NGS];   /* gC: GD == 3 only */
      int ksA = 0, keA = 0, ksB = 0, keB = 0, ksC = 0, keC = 0;   /* entries of the lane's row, relative to the
                                                   start of the staged x_val */
      /* row bounds of local iteration j and the first NG gathers of the lane's row */
      auto prefetch = [&](int j, double (&gg)[NGS], int &pks, int &pke) {
        const unsigned q = base + (unsigned)j;
        const unsigned char *sb = sv.stages + (q % NS)*sv.stage_bytes;
        consumer_wait(sv, q);
        const int *s_ri = (const int *)(sb + off_ri);
        const int rows = min(slab_rows, n_rows - (first + j*step)*slab_rows);
        const int k0 = s_ri[0];
        int ks = 0, ke = 0;
        if (tid < rows) { ks = s_ri[tid] - k0; ke = s_ri[tid + 1] - k0; }
        const int *s_col = (const int *)sb + (k0 & 3);
        /* GHOST: the row's local entries end where its ghost columns (sorted last) begin */
        int hg = 0;
        if (GHOST == 2) {
          const int ke_all = ke;
          while (ke > ks && s_col[ke - 1] >= n_rows) ke--;
          hg = (ke < ke_all) ? PARKED : 0;
        }
#pragma unroll
        for (int e = 0; e < NGS; e++) {
          double v = 0.;
          if (ks + e < ke) {
            const int c = s_col[ks + e];
            const double *px = (GHOST != 1 || c < n_rows) ? x : xg;   /* one load, selected base */
            v = px[c];
          }
          gg[e] = v;
        }
        pks = (ks + (k0 & 1)) | hg;  /* x_val is staged from an even entry */
        pke = ke + (k0 & 1);
      };
      auto consume = [&](int j, const double (&gg)[NGS], int ks, int ke) {
        const unsigned q = base + (unsigned)j;
        const unsigned char *sb = sv.stages + (q % NS)*sv.stage_bytes;
        const int r0 = (first + j*step)*slab_rows;
        const bool parked = (ks & PARKED) != 0;
        ks &= ~PARKED;
        if (tid < min(slab_rows, n_rows - r0)) {
          const double *s_val = (const double *)(sb + off_val);
          double sii = 0.0;
#pragma unroll
          for (int e = 0; e < NGS; e++)
            if (ks + e < ke) sii += NOVAL ? gg[e] : s_val[ks + e] * gg[e];
          if (ks + NGS < ke) {
            /* row longer than the gather pipeline: the rest is gathered NT at a time
               (one L2 latency per batch, not per entry), in registers the first NG
               values no longer need */
            constexpr int NT = 8;
            const int k0 = ((const int *)(sb + off_ri))[0];
            const int *s_col = (const int *)sb + (k0 & 3) - (k0 & 1);
#pragma unroll 1
            for (int kb = ks + NGS; kb < ke; kb += NT) {
              double gt[NT];
#pragma unroll
              for (int e = 0; e < NT; e++) {
                double v = 0.;
                if (kb + e < ke) {
                  const int c = s_col[kb + e];
                  const double *px = (GHOST != 1 || c < n_rows) ? x : xg;
                  v = px[c];
                }
                gt[e] = v;
              }
#pragma unroll
              for (int e = 0; e < NT; e++)
                if (kb + e < ke) sii += NOVAL ? gt[e] : s_val[kb + e] * gt[e];
            }
          }
          if (GHOST == 2 && parked) bsum[r0 + tid] = sii;
          else f.template epi<ST_OP_STRIDE>(r0 + tid, sii, (const double *)(sb + off_ops) + tid, acc);
        }
        consumer_release(sv, q);
      };

      /* (a prefetch past the last slab repeats the last one: its stage is still full) */
      const int j_last = n_my - 1;
#ifdef CSB_PHASE_TRACE
      const bool tr = (!SINGLE && blockIdx.x == 0 && tid == 0 && n_rows > TRACE_LO
                       && n_rows <= TRACE_HI);
      const unsigned long long tr0 = trace_now();
      if (tr) consumer_wait(sv, base);
      const unsigned long long tr1 = trace_now();
#endif
      prefetch(0, gA, ksA, keA);
#ifdef CSB_PHASE_TRACE
      if (tr && gA[0] == 1.2345e300) g_trace[14] += 1;      /* wait for the gathers */
      const unsigned long long tr2 = trace_now();
#endif
      if constexpr (GD == 3) {
        /* three sets rotate: the set consumed at step j was gathered at step j - 2 */
        prefetch(min(1, j_last), gB, ksB, keB);
        for (int j = 0; j < n_my; j += 3) {
          prefetch(min(j + 2, j_last), gC, ksC, keC);
          consume(j, gA, ksA, keA);
          if (j + 1 < n_my) {
            prefetch(min(j + 3, j_last), gA, ksA, keA);
            consume(j + 1, gB, ksB, keB);
          }
          if (j + 2 < n_my) {
            prefetch(min(j + 4, j_last), gB, ksB, keB);
            consume(j + 2, gC, ksC, keC);
          }
        }
      }
      else {
        for (int j = 0; j < n_my; j += 2) {
          prefetch(min(j + 1, j_last), gB, ksB, keB);
          consume(j, gA, ksA, keA);
          if (j + 1 < n_my) {
            prefetch(min(j + 2, j_last), gA, ksA, keA);
            consume(j + 1, gB, ksB, keB);
          }
        }
      }
#ifdef CSB_PHASE_TRACE
      const unsigned long long tr3 = trace_now();
      if (tr) { g_trace[0] += tr1 - tr0; g_trace[1] += tr2 - tr1; g_trace[2] += tr3 - tr2;
                g_trace[13] = tr3; }
#endif
    }
    __syncthreads();
#ifdef CSB_PHASE_TRACE
    if (!SINGLE && blockIdx.x == 0 && tid == 0 && n_rows > TRACE_LO && n_rows <= TRACE_HI)
      g_trace[3] += trace_now() - g_trace[13];
#endif
    if (tid == 0) sv.pipe->base = base + (unsigned)n_my;
    __syncthreads();
  }

  if (GHOST == 2) {
    /* the neighbours' values */
    halo_wait(L);
    /* parked rows of this CTA's slabs, in a fixed order per thread: + the ghost terms (the
       tail of the row in CSR order), then the epilogue with operands read in place */
    const unsigned *__restrict__ bmask = L.dist->bmask;
    const int n_ops = epi_nv(f, 0);
    const int tot = n_my * slab_rows;
    for (int idx = tid; idx < tot; idx += PT) {
      const int j = idx / slab_rows;
      const int r = (first + j*step)*slab_rows + (idx - j*slab_rows);
      if (r < n_rows && ((bmask[r >> 5] >> (r & 31)) & 1u)) {
        double sii = bsum[r];
        const int ke = L.row_index[r+1];
        for (int k = L.row_index[r]; k < ke; k++) {
          const int c = L.col_id[k];
          if (c >= n_rows) sii += L.x_val[k] * ghost[c - n_rows];
        }
        double opv[F::NV > 0 ? F::NV : 1];
#pragma unroll
        for (int v = 0; v < F::NV; v++) opv[v] = (v < n_ops) ? f.src(v)[r] : 0.;
        f.template epi<1>(r, sii, opv, acc);
      }
    }
  }
  if (acc_out != nullptr) {
    acc_out[0] = acc[0]; acc_out[1] = acc[1]; acc_out[2] = acc[2]; acc_out[3] = acc[3];
  }
}

template <bool SINGLE, int GHOST, int GD, class F>
__device__ __forceinline__ void spmv_phase_g(const MgLevelDesc &L, const double *x,
                                             const double *ghost, unsigned char *s_raw,
                                             const F &f, double *acc_out)
{
  spmv_phase_gv<SINGLE, GHOST, GD, false>(L, x, ghost, s_raw, f, acc_out);
}

template <bool SINGLE, class F>
__device__ __forceinline__ void spmv_phase(const MgLevelDesc &L, const double *x,
                                           const double *ghost, unsigned char *s_raw,
                                           const F &f, double *acc_out)
{
  const int gm = (ghost == nullptr) ? 0 : (L.dist->overlap ? 2 : 1);
  if (!SINGLE && L.gather_depth == 3) {
    if (gm == 2) spmv_phase_g<false, 2, 3>(L, x, ghost, s_raw, f, acc_out);
    else if (gm == 1) spmv_phase_g<false, 1, 3>(L, x, ghost, s_raw, f, acc_out);
    else spmv_phase_g<false, 0, 3>(L, x, ghost, s_raw, f, acc_out);
  }
  else if (gm == 2) spmv_phase_g<SINGLE, 2, 2>(L, x, ghost, s_raw, f, acc_out);
  else if (gm == 1) spmv_phase_g<SINGLE, 1, 2>(L, x, ghost, s_raw, f, acc_out);
  else spmv_phase_g<SINGLE, 0, 2>(L, x, ghost, s_raw, f, acc_out);
}

/* r = (A.vx - rhs), dk = r * ad_inv            (cs_sles_it.cpp:1502-1516) */
struct EpiInitResidual {
  static constexpr int NV = 4;
  const double *d_val, *vx, *rhs, *ad_inv;
  double *rk, *dk;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? d_val : (v == 1) ? vx : (v == 2) ? rhs : ad_inv; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&)[4]) const
  {
    const double rr = (sii + op[0]*op[S]) - op[2*S];
    rk[r] = rr;
    dk[r] = rr * op[3*S];
  }
};

/* out = A.p and the three dots rk.rk, rk.p, p.out      (:1521-1535, :1591-1600) */
struct EpiDots {
  static constexpr int NV = 3;
  const double *d_val, *p, *rk;
  double *out;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? d_val : (v == 1) ? p : rk; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&acc)[4]) const
  {
    const double d = op[S];
    const double z = sii + op[0]*d;
    out[r] = z;
    const double rr = op[2*S];
    acc[0] += rr*rr; acc[1] += rr*d; acc[2] += d*z;
  }
};

/* wr = rhs - A.vx and its squared norm          (cs_multigrid.cpp:3296-3301) */
struct EpiResidual {
  static constexpr int NV = 3;
  const double *d_val, *vx, *rhs;
  double *wr;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? d_val : (v == 1) ? vx : rhs; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&acc)[4]) const
  {
    const double y = sii + op[0]*op[S];
    const double t = op[2*S] - y;
    wr[r] = t;
    acc[0] += t*t;
  }
};

/*--------------------------------------------------------------------------
 * Elementwise phases over chunks of ROW_CH rows, same producer / consumer
 * split: the F::NV input vectors of a chunk are bulk-copied into a stage; a
 * consumer lane handles rows t and t + ROW_CH/2 of the chunk:
 * F::st(row, op, acc) with op[v*ROW_CH] = input vector v at this row; acc_out[0..3]
 * receives this thread's partial sums.
 *--------------------------------------------------------------------------*/

template <bool SINGLE, class F>
__device__ __noinline__ void row_phase(int n, unsigned char *s_raw, const F f, double *acc_out)
{
  const int tid = threadIdx.x;
  const int first = SINGLE ? 0 : blockIdx.x;
  const int step = SINGLE ? 1 : gridDim.x;
  const SmemView sv = smem_view(s_raw);
  const int n_chunks = (n + ROW_CH - 1) / ROW_CH;
  const int n_my = (first < n_chunks) ? (n_chunks - first + step - 1) / step : 0;
  const unsigned base = sv.pipe->base;
  double acc[4] = {0., 0., 0., 0.};

  if (n_my > 0) {
    if (tid >= NWC*32) {
      if (tid == NWC*32) {
        fence_proxy_async();
        for (int j = 0; j < n_my; j++) {
          const unsigned q = base + (unsigned)j;
          unsigned char *sb = sv.stages + (q % NS)*sv.stage_bytes;
          unsigned long long *bar = &sv.pipe->full[q % NS];
          const int c0 = (first + j*step) * ROW_CH;
          const int rows = min(ROW_CH, n - c0);
          const unsigned v_b = (unsigned)((rows*8 + 15) & ~15);
          producer_acquire(sv, q);
          mbar_expect_tx(bar, (unsigned)F::NV*v_b);
#pragma unroll
          for (int v = 0; v < F::NV; v++)
            bulk_load(sb + v*ROW_CH*8, f.src(v) + c0, v_b, bar);
        }
      }
      __syncwarp();
    }
    else {
      for (int j = 0; j < n_my; j++) {
        const unsigned q = base + (unsigned)j;
        const double *sb = (const double *)(sv.stages + (q % NS)*sv.stage_bytes);
        consumer_wait(sv, q);
        const int c0 = (first + j*step) * ROW_CH;
        const int rows = min(ROW_CH, n - c0);
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int i = tid + u*(ROW_CH/2);
          if (i < rows) f.st(c0 + i, sb + i, acc);
        }
        consumer_release(sv, q);
      }
    }
    __syncthreads();
    if (tid == 0) sv.pipe->base = base + (unsigned)n_my;
    __syncthreads();
  }
  if (acc_out != nullptr) {
    acc_out[0] = acc[0]; acc_out[1] = acc[1]; acc_out[2] = acc[2]; acc_out[3] = acc[3];
  }
}

/* vx = 0, rk = -rhs, dk = rk * ad_inv                     (:1502-1516, vx = 0) */
struct RowInitZero {
  static constexpr int NV = 2;
  const double *rhs, *ad_inv;
  double *vx, *rk, *dk;
  __device__ __forceinline__ const double *src(int v) const { return (v == 0) ? rhs : ad_inv; }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  {
    vx[i] = 0;
    const double rr = -op[0];
    rk[i] = rr;
    dk[i] = rr * op[ROW_CH];
  }
};

/* first update of the single-reduction variant                   (:1550-1558) */
struct RowFirstUpdate {
  static constexpr int NV = 5;
  double alpha;
  const double *dk, *zk, *ad_inv;
  double *rk, *vx, *gk;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? dk : (v == 1) ? zk : (v == 2) ? rk : (v == 3) ? vx : ad_inv; }
  __device__ __forceinline__ void st(int i, const double *op, double (&acc)[4]) const
  {
    vx[i] = op[3*ROW_CH] + (alpha * op[0]);
    const double rr = op[2*ROW_CH] + (alpha * op[ROW_CH]);
    rk[i] = rr;
    gk[i] = rr * op[4*ROW_CH];
    acc[0] += rr*rr;
  }
};

/* the 5-vector update of an iteration                            (:1623-1631) */
struct RowUpdate {
  static constexpr int NV = 7;
  double alpha, beta;
  const double *sk, *ad_inv;
  double *gk, *dk, *vx, *zk, *rk;
  __device__ __forceinline__ const double *src(int v) const
  {
    return (v == 0) ? gk : (v == 1) ? dk : (v == 2) ? vx : (v == 3) ? sk : (v == 4) ? zk
         : (v == 5) ? rk : ad_inv;
  }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  {
    const double d = op[0] + (beta * op[ROW_CH]);
    dk[i] = d;
    vx[i] = op[2*ROW_CH] + alpha * d;
    const double z = op[3*ROW_CH] + (beta * op[4*ROW_CH]);
    zk[i] = z;
    const double rr = op[5*ROW_CH] + alpha * z;
    rk[i] = rr;
    gk[i] = rr * op[6*ROW_CH];
  }
};

/* last sweep of a smoother: only vx is used afterwards */
struct RowLastUpdate {
  static constexpr int NV = 3;
  double alpha, beta;
  const double *gk, *dk;
  double *vx;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? gk : (v == 1) ? dk : vx; }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  {
    const double d = op[0] + (beta * op[ROW_CH]);
    vx[i] = op[2*ROW_CH] + alpha * d;
  }
};

/* Prolongation f[i] += c[coarse_row[i]] (cs_grid.cpp:9156-9166): the coarse
   values are gathered for all U rows before the first store. */
template <bool SINGLE>
__device__ __noinline__ void prolong_rows(int n, const int *__restrict__ coarse_row,
                                          const double *c_vx, double *vx)
{
  constexpr int U = CSB_PROLONG_U;
  const int first = SINGLE ? threadIdx.x : blockIdx.x*PT + threadIdx.x;
  const int step = SINGLE ? PT : gridDim.x*PT;
  for (int base = first; base < n; base += U*step) {
    int ic[U];
    double v[U], c[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = base + u*step;
      ic[u] = -1;
      if (i < n) { ic[u] = __ldg(coarse_row + i); v[u] = vx[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (ic[u] >= 0) c[u] = c_vx[ic[u]];
#pragma unroll
    for (int u = 0; u < U; u++)
      if (ic[u] >= 0) vx[base + u*step] = v[u] + c[u];
  }
}

/* Restriction: coarse row = sum of its fine rows in ascending fine order
   (== the sequential loop of cs_grid.cpp:9025-9030).  Three dependent loads
   (row range -> fine row ids -> values); each level is issued for U coarse
   rows x up to 4 fine rows before the next one is touched. */
template <bool SINGLE>
__device__ __noinline__ void restrict_rows(int nc, const int *__restrict__ cfi,
                                           const int *__restrict__ cfr,
                                           const double *wr, double *c_rhs)
{
  constexpr int U = CSB_RESTRICT_U, J = 4;
  const int first = SINGLE ? threadIdx.x : blockIdx.x*PT + threadIdx.x;
  const int step = SINGLE ? PT : gridDim.x*PT;
  for (int base = first; base < nc; base += U*step) {
    int rs[U], cnt[U], id[U][J];
    double w[U][J];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int ic = base + u*step;
      rs[u] = 0; cnt[u] = 0;
      if (ic < nc) { rs[u] = __ldg(cfi + ic); cnt[u] = __ldg(cfi + ic + 1) - rs[u]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < J; j++)
        if (j < cnt[u]) id[u][j] = __ldg(cfr + rs[u] + j);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < J; j++)
        if (j < cnt[u]) w[u][j] = wr[id[u][j]];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int ic = base + u*step;
      if (ic < nc) {
        double s = 0.;
#pragma unroll
        for (int j = 0; j < J; j++)
          if (j < cnt[u]) s += w[u][j];
        for (int q = J; q < cnt[u]; q++) s += wr[__ldg(cfr + rs[u] + q)];
        c_rhs[ic] = s;
      }
    }
  }
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
                             ItState *st, unsigned char *s_raw,
                             unsigned long long *prof_it = nullptr)
{
  /* lead thread: time spent per kind of phase, closing barrier / reduction included */
  const bool timed = (prof_it != nullptr) && blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long t_mark = timed ? global_ns() : 0ull;
  auto mark = [&](int kind) {
    if (timed) { const unsigned long long t = global_ns(); prof_it[kind] += t - t_mark; t_mark = t; }
  };
  const int n = L.n_rows;
  const size_t ws = P.work_stride;
  double *rk = P.work, *dk = P.work + ws, *gk = P.work + 2*ws, *zk = P.work + 3*ws,
         *sk = P.work + 4*ws;
  const double *__restrict__ ad_inv = L.ad_inv;
  const double *__restrict__ d_val = L.d_val;
  const bool rec = (st != nullptr) && (SINGLE || blockIdx.x == 0) && threadIdx.x == 0;
  double *part0 = P.partials, *part1 = P.partials + 4*(size_t)gridDim.x;
  double *s_red = (double *)s_raw;

  PcgResult R;
  R.n_iter = 0; R.cvg = CSB_SLES_ITERATING; R.residual = -1.; R.initial_residual = -1.;
  if (rec) st->n_hist = 0;
  const unsigned hk0 = halo_count(L);

  /* residual and descent direction (:1502-1516) */
  const CommDev *comm = P.comm;
  if (vx_ini_nonzero) {
    const double *gh = halo_exchange<SINGLE>(L, vx);
    EpiInitResidual f;
    f.d_val = d_val; f.vx = vx; f.rhs = rhs; f.ad_inv = ad_inv; f.rk = rk; f.dk = dk;
    spmv_phase<SINGLE>(L, vx, gh, s_raw, f, nullptr);
  }
  else {
    RowInitZero f;
    f.rhs = rhs; f.ad_inv = ad_inv; f.vx = vx; f.rk = rk; f.dk = dk;
    row_phase<SINGLE>(n, s_raw, f, nullptr);
  }
  phase_barrier<SINGLE>(P.barrier);
  mark(vx_ini_nonzero ? 0 : 2);

  /* zk = A.dk and the first 3 dots (:1521-1535) */
  double v[4] = {0., 0., 0., 0.};
  const double *gh_d = halo_exchange<SINGLE>(L, dk);
  {
    EpiDots f;
    f.d_val = d_val; f.p = dk; f.rk = rk; f.out = zk;
    spmv_phase<SINGLE>(L, dk, gh_d, s_raw, f, v);
  }
  { double v3[3] = {v[0], v[1], v[2]}; all_sum<SINGLE, 3>(v3, part0, P.barrier, s_red, comm); v[0] = v3[0]; v[1] = v3[1]; v[2] = v3[2]; }
  mark(1);

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
    double w[4] = {0., 0., 0., 0.};
    {
      RowFirstUpdate f;                                /* :1550-1558 */
      f.alpha = alpha; f.dk = dk; f.zk = zk; f.ad_inv = ad_inv; f.rk = rk; f.vx = vx; f.gk = gk;
      row_phase<SINGLE>(n, s_raw, f, w);
    }
    { double w1[1] = {w[0]}; all_sum<SINGLE, 1>(w1, part1, P.barrier, s_red, comm); w[0] = w1[0]; }
    if (SINGLE) __syncthreads();
    mark(2);
    R.residual = sqrt(w[0]);
    R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, 1, R.residual);
    if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;
  }

  while (R.cvg == CSB_SLES_ITERATING) {
    if (smoother && R.n_iter >= n_max_iter) break;
    v[0] = 0.; v[1] = 0.; v[2] = 0.;
#ifdef CSB_PHASE_TRACE
    const bool tr = (!SINGLE && blockIdx.x == 0 && threadIdx.x == 0 && L.n_rows > TRACE_LO
                     && L.n_rows <= TRACE_HI);
    const unsigned long long trp0 = trace_now();
#endif
    const double *gh_g = halo_exchange<SINGLE>(L, gk);
    {
      EpiDots f;                                                       /* :1591-1600 */
      f.d_val = d_val; f.p = gk; f.rk = rk; f.out = sk;
      spmv_phase<SINGLE>(L, gk, gh_g, s_raw, f, v);
    }
#ifdef CSB_PHASE_TRACE
    const unsigned long long trp1 = trace_now();
#endif
    { double v3[3] = {v[0], v[1], v[2]}; all_sum<SINGLE, 3>(v3, part0, P.barrier, s_red, comm); v[0] = v3[0]; v[1] = v3[1]; v[2] = v3[2]; }
#ifdef CSB_PHASE_TRACE
    if (tr) { const unsigned long long t = trace_now();
              g_trace[4] += t - trp1; g_trace[5] += t - trp0; g_trace[15] += 1; }
#endif
    mark(1);
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
      RowLastUpdate f;
      f.alpha = alpha; f.beta = beta; f.gk = gk; f.dk = dk; f.vx = vx;
      row_phase<SINGLE>(n, s_raw, f, nullptr);
    }
    else {
      RowUpdate f;                              /* :1623-1631 */
      f.alpha = alpha; f.beta = beta; f.sk = sk; f.ad_inv = ad_inv;
      f.gk = gk; f.dk = dk; f.vx = vx; f.zk = zk; f.rk = rk;
      row_phase<SINGLE>(n, s_raw, f, nullptr);
    }
    phase_barrier<SINGLE>(P.barrier);
    mark(2);
  }

  if (rec) {
    st->precision = precision; st->r_norm = r_norm;
    st->residual = R.residual; st->initial_residual = R.initial_residual;
    st->n_iter = R.n_iter; st->n_max_iter = n_max_iter; st->cvg = R.cvg;
    st->alpha = alpha; st->beta = beta; st->ro_0 = ro_0; st->ro_1 = ro_1;
    st->rk_gkm1 = rk_gkm1;
    st->n_halo = (int)(halo_count(L) - hk0);
  }
  return R;
}

/*--------------------------------------------------------------------------
 * Flexible CG + Jacobi (_flexible_conjugate_gradient, cs_sles_it.cpp:832-985):
 * per iteration ONE SpMV phase (wk = A.vk with the four dots of
 * _dot_products_vr_vw_vq_rr, :693-724) and ONE vector phase (the two update
 * loops :939-951 fused with the next preconditioning vk = rk * ad_inv, :906).
 * Work vectors: rk vk wk dk qk.  Residual sign r = b - A.x.
 *--------------------------------------------------------------------------*/

/* rk = rhs - A.vx, qk = 0, vk = rk * ad_inv                    (:874-906) */
struct EpiFcgInit {
  static constexpr int NV = 4;
  const double *d_val, *vx, *rhs, *ad_inv;
  double *rk, *qk, *vk;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? d_val : (v == 1) ? vx : (v == 2) ? rhs : ad_inv; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&)[4]) const
  {
    const double rr = op[2*S] - (sii + op[0]*op[S]);
    rk[r] = rr;
    qk[r] = 0;
    vk[r] = rr * op[3*S];
  }
};

/* wk = A.vk; alpha = vk.rk, beta = vk.wk, gamma = vk.qk, rr = rk.rk     (:908-918) */
struct EpiFcg {
  static constexpr int NV = 4;
  const double *d_val, *vk, *rk, *qk;
  double *wk;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? d_val : (v == 1) ? vk : (v == 2) ? rk : qk; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&acc)[4]) const
  {
    const double v = op[S];
    const double w = sii + op[0]*v;
    wk[r] = w;
    const double rr = op[2*S];
    acc[0] += rr*v; acc[1] += w*v; acc[2] += v*op[3*S]; acc[3] += rr*rr;
  }
};

/* vx = 0, rk = rhs, qk = 0, vk = rk * ad_inv                     (:880-893) */
struct RowFcgInitZero {
  static constexpr int NV = 2;
  const double *rhs, *ad_inv;
  double *vx, *rk, *qk, *vk;
  __device__ __forceinline__ const double *src(int v) const { return (v == 0) ? rhs : ad_inv; }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  {
    vx[i] = 0;
    rk[i] = op[0];
    qk[i] = 0;
    vk[i] = op[0] * op[ROW_CH];
  }
};

/* dk, vx, qk, rk updates (:939-951; first iteration :960-972) + vk = rk * ad_inv */
struct RowFcgUpdate {
  static constexpr int NV = 7;
  double gk_rk1, ak_rk;
  int first;
  const double *wk, *ad_inv;
  double *vk, *dk, *vx, *qk, *rk;
  __device__ __forceinline__ const double *src(int v) const
  {
    return (v == 0) ? vk : (v == 1) ? dk : (v == 2) ? vx : (v == 3) ? wk : (v == 4) ? qk
         : (v == 5) ? rk : ad_inv;
  }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  {
    const double d = first ? op[0] : op[0] - gk_rk1 * op[ROW_CH];
    dk[i] = d;
    vx[i] = op[2*ROW_CH] + ak_rk * d;
    const double q = first ? op[3*ROW_CH] : op[3*ROW_CH] - gk_rk1 * op[4*ROW_CH];
    qk[i] = q;
    const double rr = op[5*ROW_CH] - ak_rk * q;
    rk[i] = rr;
    vk[i] = rr * op[6*ROW_CH];
  }
};

template <bool SINGLE>
__device__ __noinline__ PcgResult fcg_run(const MgLevelDesc &L, const double *rhs, double *vx,
                             bool vx_ini_nonzero, int n_max_iter, double precision,
                             double r_norm, const PersistentCommon &P, ItState *st,
                             unsigned char *s_raw)
{
  const int n = L.n_rows;
  const size_t ws = P.work_stride;
  double *rk = P.work, *vk = P.work + ws, *wk = P.work + 2*ws, *dk = P.work + 3*ws,
         *qk = P.work + 4*ws;
  const bool rec = (st != nullptr) && (SINGLE || blockIdx.x == 0) && threadIdx.x == 0;
  double *s_red = (double *)s_raw;
  const CommDev *comm = P.comm;

  PcgResult R;
  R.n_iter = 0; R.cvg = CSB_SLES_ITERATING; R.residual = -1.; R.initial_residual = -1.;
  if (rec) st->n_hist = 0;

  if (vx_ini_nonzero) {
    const double *gh = halo_exchange<SINGLE>(L, vx);
    EpiFcgInit f;
    f.d_val = L.d_val; f.vx = vx; f.rhs = rhs; f.ad_inv = L.ad_inv; f.rk = rk; f.qk = qk; f.vk = vk;
    spmv_phase<SINGLE>(L, vx, gh, s_raw, f, nullptr);
  }
  else {
    RowFcgInitZero f;
    f.rhs = rhs; f.ad_inv = L.ad_inv; f.vx = vx; f.rk = rk; f.qk = qk; f.vk = vk;
    row_phase<SINGLE>(n, s_raw, f, nullptr);
  }
  phase_barrier<SINGLE>(P.barrier);

  double rho_km1 = 0., ak_rk = 0., gk_rk1 = 0.;
  while (R.cvg == CSB_SLES_ITERATING) {
    double v[4] = {0., 0., 0., 0.};
    const double *gh = halo_exchange<SINGLE>(L, vk);
    {
      EpiFcg f;
      f.d_val = L.d_val; f.vk = vk; f.rk = rk; f.qk = qk; f.wk = wk;
      spmv_phase<SINGLE>(L, vk, gh, s_raw, f, v);
    }
    all_sum<SINGLE, 4>(v, P.partials, P.barrier, s_red, comm);
    const double alpha_k = v[0], beta_k = v[1], gamma_k = v[2];
    R.residual = sqrt(v[3]);
    if (R.n_iter > 0)
      R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, R.n_iter,
                          R.residual);
    else
      R.initial_residual = R.residual;
    if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;
    if (R.cvg != CSB_SLES_ITERATING) break;

    double rho_k;
    if (R.n_iter > 0) {
      gk_rk1 = (fabs(rho_km1) > DBL_MIN) ? gamma_k / rho_km1 : 0.;
      rho_k = beta_k - gamma_k * gk_rk1;
    }
    else {
      gk_rk1 = 0.;
      rho_k = beta_k;
    }
    ak_rk = (fabs(rho_k) > DBL_MIN) ? alpha_k / rho_k : 0.;
    {
      RowFcgUpdate f;
      f.gk_rk1 = gk_rk1; f.ak_rk = ak_rk; f.first = (R.n_iter == 0) ? 1 : 0;
      f.wk = wk; f.ad_inv = L.ad_inv; f.vk = vk; f.dk = dk; f.vx = vx; f.qk = qk; f.rk = rk;
      row_phase<SINGLE>(n, s_raw, f, nullptr);
    }
    phase_barrier<SINGLE>(P.barrier);
    rho_km1 = rho_k;
    R.n_iter += 1;
  }

  if (rec) {
    st->precision = precision; st->r_norm = r_norm;
    st->residual = R.residual; st->initial_residual = R.initial_residual;
    st->n_iter = R.n_iter; st->n_max_iter = n_max_iter; st->cvg = R.cvg;
    st->alpha = ak_rk; st->beta = gk_rk1; st->ro_0 = rho_km1; st->ro_1 = 0.; st->rk_gkm1 = 0.;
  }
  return R;
}

/*--------------------------------------------------------------------------
 * PCG + Jacobi on a system small enough to LIVE in shared memory (the coarsest
 * level of a hierarchy: tens to a few hundred rows, solved to convergence in
 * every cycle).  pcg_run<true> streams the matrix through the bulk-copy ring
 * again for each of the ~50 iterations and pays two pipeline fills per
 * iteration (~3 us); here matrix and vectors are loaded once and an iteration
 * is a few __syncthreads.  Same recurrences, same per-entry arithmetic and --
 * thread t owns rows t, t + SLAB_ROWS_MAX, ... exactly as the lanes of the
 * streamed phases do -- the same summation order: results are bit-identical
 * to pcg_run<true> (checked by tests/test_gpu_r2.py).  One CTA, zero start,
 * no ghost columns.
 *--------------------------------------------------------------------------*/

__device__ __forceinline__ bool pcg_small_fits(const MgLevelDesc &L, int small_rows)
{
  if (L.dist != nullptr || L.n_rows > small_rows || L.n_rows < 1) return false;
  const int n = L.n_rows;
  const long long nnz_max = (long long)n * L.row_len_max;
  const long long bytes = 8ll*((n + 2)/2 + (nnz_max + 1)/2) + 8ll*nnz_max + 72ll*n + 64;
  return bytes <= (long long)dyn_smem_bytes() - SMEM_HDR;
}

__device__ __noinline__ void pcg_small(const MgLevelDesc &L, const double *rhs, double *vx,
                                       int n_max_iter, double precision, double r_norm,
                                       bool smoother, ItState *st, unsigned char *s_raw)
{
  const int n = L.n_rows, tid = threadIdx.x, slab_rows = L.slab_rows;
  double *s_red = (double *)s_raw;
  unsigned char *area = s_raw + SMEM_HDR;
  int *s_ri = (int *)area;
  for (int i = tid; i <= n; i += PT) s_ri[i] = L.row_index[i];
  __syncthreads();
  const int nnz = s_ri[n];
  int *s_col = s_ri + ((n + 2) & ~1);
  double *s_val = (double *)(s_col + ((nnz + 1) & ~1));
  double *s_d = s_val + nnz, *s_inv = s_d + n, *s_rhs = s_inv + n, *s_vx = s_rhs + n,
         *rk = s_vx + n, *dk = rk + n, *gk = dk + n, *zk = gk + n, *sk = zk + n;
  for (int k = tid; k < nnz; k += PT) { s_col[k] = L.col_id[k]; s_val[k] = L.x_val[k]; }
  for (int i = tid; i < n; i += PT) {
    s_d[i] = L.d_val[i]; s_inv[i] = L.ad_inv[i]; s_rhs[i] = rhs[i];
    /* vx = 0, rk = -rhs, dk = rk * ad_inv   (cs_sles_it.cpp:1502-1516, vx = 0) */
    s_vx[i] = 0;
    const double rr = -rhs[i];
    rk[i] = rr;
    dk[i] = rr * L.ad_inv[i];
  }
  __syncthreads();

  const bool rec = (st != nullptr) && tid == 0;
  int n_iter = 0, cvg = CSB_SLES_ITERATING;
  double residual = -1., initial_residual = -1.;
  if (rec) st->n_hist = 0;

  /* out = A.p in CSR order, diagonal last; dots rk.rk, rk.p, p.out (:1521-1535, :1591-1600) */
  auto spmv_dots = [&](const double *p_, double *out, double (&v)[3]) {
    v[0] = 0.; v[1] = 0.; v[2] = 0.;
    if (tid < slab_rows)          /* the lanes of the streamed SpMV phase own rows t + k*slab_rows */
      for (int i = tid; i < n; i += slab_rows) {
        double sii = 0.0;
        for (int k = s_ri[i]; k < s_ri[i+1]; k++) sii += s_val[k] * p_[s_col[k]];
        const double d = p_[i];
        const double z = sii + s_d[i]*d;
        out[i] = z;
        const double rr = rk[i];
        v[0] += rr*rr; v[1] += rr*d; v[2] += d*z;
      }
    cta_sum_bcast<3>(v, s_red);
  };

  double v[3];
  spmv_dots(dk, zk, v);
  double ro_0 = v[1], ro_1 = v[2], rk_gkm1 = 0., alpha = 0., beta = 0.;
  residual = sqrt(v[0]);
  initial_residual = residual;
  cvg = it_cvg_eval(precision, r_norm, n_max_iter, initial_residual, 0, residual);
  if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = residual;

  if (cvg == CSB_SLES_ITERATING) {
    n_iter = 1;
    const double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    rk_gkm1 = ro_0;
    double w[1] = {0.};
    if (tid < SLAB_ROWS_MAX)
      for (int i = tid; i < n; i += SLAB_ROWS_MAX) {            /* :1550-1558 */
        s_vx[i] = s_vx[i] + (alpha * dk[i]);
        const double rr = rk[i] + (alpha * zk[i]);
        rk[i] = rr;
        gk[i] = rr * s_inv[i];
        w[0] += rr*rr;
      }
    cta_sum_bcast<1>(w, s_red);
    residual = sqrt(w[0]);
    cvg = it_cvg_eval(precision, r_norm, n_max_iter, initial_residual, 1, residual);
    if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = residual;
  }

  while (cvg == CSB_SLES_ITERATING) {
    if (smoother && n_iter >= n_max_iter) break;
    spmv_dots(gk, sk, v);
    const double rk_gk = v[1], gk_sk = v[2];
    residual = sqrt(v[0]);
    if (n_iter > 1) {
      cvg = it_cvg_eval(precision, r_norm, n_max_iter, initial_residual, n_iter, residual);
      if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = residual;
    }
    if (cvg != CSB_SLES_ITERATING) break;
    n_iter += 1;
    beta = (fabs(rk_gkm1) > DBL_MIN) ? rk_gk / rk_gkm1 : 0.;    /* :1614-1621 */
    rk_gkm1 = rk_gk;
    ro_1 = gk_sk - beta*beta*ro_1;
    ro_0 = rk_gk;
    const double d_ro_1 = (fabs(ro_1) > DBL_MIN) ? 1. / ro_1 : 0.;
    alpha = - ro_0 * d_ro_1;
    const bool last = (smoother && n_iter >= n_max_iter);
    if (tid < SLAB_ROWS_MAX)
      for (int i = tid; i < n; i += SLAB_ROWS_MAX) {            /* :1623-1631 */
        const double d = gk[i] + (beta * dk[i]);
        s_vx[i] = s_vx[i] + alpha * d;
        if (!last) {
          dk[i] = d;
          const double z = sk[i] + (beta * zk[i]);
          zk[i] = z;
          const double rr = rk[i] + alpha * z;
          rk[i] = rr;
          gk[i] = rr * s_inv[i];
        }
      }
    __syncthreads();
  }

  for (int i = tid; i < n; i += PT) vx[i] = s_vx[i];
  if (rec) {
    st->precision = precision; st->r_norm = r_norm;
    st->residual = residual; st->initial_residual = initial_residual;
    st->n_iter = n_iter; st->n_max_iter = n_max_iter; st->cvg = cvg;
    st->alpha = alpha; st->beta = beta; st->ro_0 = ro_0; st->ro_1 = ro_1;
    st->rk_gkm1 = rk_gkm1;
    st->n_halo = 0;
  }
  /* the stage area goes back to the bulk-copy rings */
  __syncthreads();
  fence_proxy_async();
}

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_fcg_persistent(const MgLevelDesc L, const double *rhs, double *vx, int vx_ini_nonzero,
                 int n_max_iter, double precision, double r_norm,
                 const PersistentCommon P, ItState *st, int single)
{
  unsigned char *s_raw = s_raw_base;
  pipe_init(s_raw);
  halo_counts_init(&L, 1, P.barrier);
  if (single) {
    if (blockIdx.x == 0)
      fcg_run<true>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, P, st,
                    s_raw);
  }
  else
    fcg_run<false>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, P, st,
                   s_raw);
}

/*--------------------------------------------------------------------------
 * Inexact-preconditioning CG + Jacobi (_conjugate_gradient_ip,
 * cs_sles_it.cpp:1016-1203).  Work vectors: rk rkm1 dk gk zk.  Per iteration:
 * vector phase (gk = rk*ad_inv + 3 dots), vector phase (dk = gk + beta*dk),
 * SpMV phase (zk = A.dk + 2 dots), vector phase (vx, rkm1, rk).
 *--------------------------------------------------------------------------*/

/* rk = A.vx - rhs, gk = rk*ad_inv, dk = gk; dots rk.rk, rk.gk            (:1060-1092) */
struct EpiIpcgInit {
  static constexpr int NV = 4;
  const double *d_val, *vx, *rhs, *ad_inv;
  double *rk, *gk, *dk;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? d_val : (v == 1) ? vx : (v == 2) ? rhs : ad_inv; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&acc)[4]) const
  {
    const double rr = (sii + op[0]*op[S]) - op[2*S];
    const double g = rr * op[3*S];
    rk[r] = rr; gk[r] = g; dk[r] = g;
    acc[0] += rr*rr; acc[1] += rr*g;
  }
};

struct RowIpcgInitZero {     /* vx = 0, rk = -rhs, gk = dk = rk*ad_inv */
  static constexpr int NV = 2;
  const double *rhs, *ad_inv;
  double *vx, *rk, *gk, *dk;
  __device__ __forceinline__ const double *src(int v) const { return (v == 0) ? rhs : ad_inv; }
  __device__ __forceinline__ void st(int i, const double *op, double (&acc)[4]) const
  {
    vx[i] = 0;
    const double rr = -op[0];
    const double g = rr * op[ROW_CH];
    rk[i] = rr; gk[i] = g; dk[i] = g;
    acc[0] += rr*rr; acc[1] += rr*g;
  }
};

struct RowIpcgPc {           /* gk = rk*ad_inv; dots rk.rk, rk.gk, rkm1.gk     (:1133-1139) */
  static constexpr int NV = 3;
  const double *rk, *ad_inv, *rkm1;
  double *gk;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? rk : (v == 1) ? ad_inv : rkm1; }
  __device__ __forceinline__ void st(int i, const double *op, double (&acc)[4]) const
  {
    const double g = op[0] * op[ROW_CH];
    gk[i] = g;
    acc[0] += op[0]*op[0]; acc[1] += op[0]*g; acc[2] += g*op[2*ROW_CH];
  }
};

struct RowIpcgDir {          /* dk = gk + beta*dk                               (:1158-1160) */
  static constexpr int NV = 2;
  double beta;
  const double *gk;
  double *dk;
  __device__ __forceinline__ const double *src(int v) const { return (v == 0) ? gk : dk; }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  { dk[i] = op[0] + (beta * op[ROW_CH]); }
};

struct RowIpcgUpdate {       /* vx += alpha*dk; rkm1 = rk; rk += alpha*zk; dot rk.rk */
  static constexpr int NV = 4;
  double alpha;
  const double *dk, *zk;
  double *vx, *rk, *rkm1;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? dk : (v == 1) ? zk : (v == 2) ? vx : rk; }
  __device__ __forceinline__ void st(int i, const double *op, double (&acc)[4]) const
  {
    vx[i] = op[2*ROW_CH] + (alpha * op[0]);
    rkm1[i] = op[3*ROW_CH];
    const double rr = op[3*ROW_CH] + (alpha * op[ROW_CH]);
    rk[i] = rr;
    acc[0] += rr*rr;
  }
};

template <bool SINGLE>
__device__ __noinline__ PcgResult ipcg_run(const MgLevelDesc &L, const double *rhs, double *vx,
                             bool vx_ini_nonzero, int n_max_iter, double precision,
                             double r_norm, const PersistentCommon &P, ItState *st,
                             unsigned char *s_raw)
{
  const int n = L.n_rows;
  const size_t ws = P.work_stride;
  double *rk = P.work, *rkm1 = P.work + ws, *dk = P.work + 2*ws, *gk = P.work + 3*ws,
         *zk = P.work + 4*ws;
  const bool rec = (st != nullptr) && (SINGLE || blockIdx.x == 0) && threadIdx.x == 0;
  double *s_red = (double *)s_raw;
  const CommDev *comm = P.comm;

  PcgResult R;
  R.n_iter = 0; R.cvg = CSB_SLES_ITERATING; R.residual = -1.; R.initial_residual = -1.;
  if (rec) st->n_hist = 0;
  double v[4] = {0., 0., 0., 0.};
  double alpha = 0., beta = 0., rk_gk_m1 = 0.;
  /* three regions of per-CTA partials: two reductions that follow each other without a
     grid barrier in between never share one (a fast CTA would overwrite what a slow one
     is still reading) */
  double *part0 = P.partials, *part1 = P.partials + 4*(size_t)gridDim.x,
         *part2 = P.partials + 8*(size_t)gridDim.x;

  auto spmv_dk = [&]() {      /* zk = A.dk; v[1] = rk.dk, v[2] = dk.zk */
    const double *gh = halo_exchange<SINGLE>(L, dk);
    EpiDots f;
    f.d_val = L.d_val; f.p = dk; f.rk = rk; f.out = zk;
    spmv_phase<SINGLE>(L, dk, gh, s_raw, f, v);
    all_sum<SINGLE, 4>(v, part0, P.barrier, s_red, comm);
  };
  auto update = [&]() {       /* v[0] = rk.rk after the update */
    RowIpcgUpdate f;
    f.alpha = alpha; f.dk = dk; f.zk = zk; f.vx = vx; f.rk = rk; f.rkm1 = rkm1;
    row_phase<SINGLE>(n, s_raw, f, v);
  };

  if (vx_ini_nonzero) {
    const double *gh = halo_exchange<SINGLE>(L, vx);
    EpiIpcgInit f;
    f.d_val = L.d_val; f.vx = vx; f.rhs = rhs; f.ad_inv = L.ad_inv; f.rk = rk; f.gk = gk; f.dk = dk;
    spmv_phase<SINGLE>(L, vx, gh, s_raw, f, v);
  }
  else {
    RowIpcgInitZero f;
    f.rhs = rhs; f.ad_inv = L.ad_inv; f.vx = vx; f.rk = rk; f.gk = gk; f.dk = dk;
    row_phase<SINGLE>(n, s_raw, f, v);
  }
  all_sum<SINGLE, 4>(v, part2, P.barrier, s_red, comm);
  R.residual = sqrt(v[0]);
  rk_gk_m1 = v[1];
  R.initial_residual = R.residual;
  R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, 0, R.residual);
  if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;

  if (R.cvg == CSB_SLES_ITERATING) {
    R.n_iter = 1;
    spmv_dk();
    const double d_ro_1 = (fabs(v[2]) > DBL_MIN) ? 1. / v[2] : 0.;
    alpha = - v[1] * d_ro_1;
    update();
    all_sum<SINGLE, 4>(v, part1, P.barrier, s_red, comm);
    R.residual = sqrt(v[0]);
    R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, 1, R.residual);
    if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;
  }

  while (R.cvg == CSB_SLES_ITERATING) {
    {
      RowIpcgPc f;
      f.rk = rk; f.ad_inv = L.ad_inv; f.rkm1 = rkm1; f.gk = gk;
      row_phase<SINGLE>(n, s_raw, f, v);
    }
    all_sum<SINGLE, 4>(v, part2, P.barrier, s_red, comm);
    const double rk_gk = v[1], rkm1_gk = v[2];
    R.residual = sqrt(v[0]);
    if (R.n_iter > 1) {
      R.cvg = it_cvg_eval(precision, r_norm, n_max_iter, R.initial_residual, R.n_iter,
                          R.residual);
      if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;
    }
    if (R.cvg != CSB_SLES_ITERATING) break;
    R.n_iter += 1;
    beta = (fabs(rk_gk_m1) > DBL_MIN) ? (rk_gk - rkm1_gk) / rk_gk_m1 : 0.;
    rk_gk_m1 = rk_gk;
    {
      RowIpcgDir f;
      f.beta = beta; f.gk = gk; f.dk = dk;
      row_phase<SINGLE>(n, s_raw, f, nullptr);
    }
    phase_barrier<SINGLE>(P.barrier);
    spmv_dk();
    const double d_ro_1 = (fabs(v[2]) > DBL_MIN) ? 1. / v[2] : 0.;
    alpha = - v[1] * d_ro_1;
    update();
    phase_barrier<SINGLE>(P.barrier);
  }

  if (rec) {
    st->precision = precision; st->r_norm = r_norm;
    st->residual = R.residual; st->initial_residual = R.initial_residual;
    st->n_iter = R.n_iter; st->n_max_iter = n_max_iter; st->cvg = R.cvg;
    st->alpha = alpha; st->beta = beta; st->ro_0 = 0.; st->ro_1 = 0.; st->rk_gkm1 = rk_gk_m1;
  }
  return R;
}

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_ipcg_persistent(const MgLevelDesc L, const double *rhs, double *vx, int vx_ini_nonzero,
                  int n_max_iter, double precision, double r_norm,
                  const PersistentCommon P, ItState *st, int single)
{
  unsigned char *s_raw = s_raw_base;
  pipe_init(s_raw);
  halo_counts_init(&L, 1, P.barrier);
  if (single) {
    if (blockIdx.x == 0)
      ipcg_run<true>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, P, st,
                     s_raw);
  }
  else
    ipcg_run<false>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, P, st,
                    s_raw);
}

/*--------------------------------------------------------------------------
 * Stand-alone SpMV (cs_matrix_vector_multiply, _mat_vec_p_l_msr
 * cs_matrix_spmv.cpp:2103-2165) through the same pipeline.
 *--------------------------------------------------------------------------*/

struct EpiStore {            /* y = E.x + d*x  (diagonal last, :2138-2142) */
  static constexpr int NV = 2;
  const double *d_val, *x;
  double *y;
  __device__ __forceinline__ const double *src(int v) const { return (v == 0) ? d_val : x; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&)[4]) const
  { y[r] = sii + op[0]*op[S]; }
};

struct EpiStoreExDiag {      /* y = E.x */
  static constexpr int NV = 0;
  double *y;
  __device__ __forceinline__ const double *src(int) const { return nullptr; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *, double (&)[4]) const
  { y[r] = sii; }
};

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_spmv_pipeline(const MgLevelDesc L, const double *x, double *y, int exclude_diag)
{
  unsigned char *s_raw = s_raw_base;
  pipe_init(s_raw);
  if (exclude_diag) {
    EpiStoreExDiag f;
    f.y = y;
    spmv_phase<false>(L, x, nullptr, s_raw, f, nullptr);
  }
  else {
    EpiStore f;
    f.d_val = L.d_val; f.x = x; f.y = y;
    spmv_phase<false>(L, x, nullptr, s_raw, f, nullptr);
  }
}

/*--------------------------------------------------------------------------
 * Kernel 1: one PCG solve (cs_sles_it_solve with CS_SLES_PCG + Jacobi)
 *--------------------------------------------------------------------------*/

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_pcg_persistent(const MgLevelDesc L, const double *rhs, double *vx, int vx_ini_nonzero,
                 int n_max_iter, double precision, double r_norm,
                 const PersistentCommon P, ItState *st, int single)
{
  unsigned char *s_raw = s_raw_base;
  pipe_init(s_raw);
  halo_counts_init(&L, 1, P.barrier);
  if (single) {
    if (blockIdx.x == 0)
      pcg_run<true>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, false,
                    P, st, s_raw);
  }
  else
    pcg_run<false>(L, rhs, vx, vx_ini_nonzero != 0, n_max_iter, precision, r_norm, false,
                   P, st, s_raw);
}

/*--------------------------------------------------------------------------
 * Kernel 2: whole multigrid solve
 *--------------------------------------------------------------------------*/

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_mg_persistent(const MgSolveParams S, const PersistentCommon P)
{
  unsigned char *s_raw = s_raw_base;
  double *s_red = (double *)s_raw;
  pipe_init(s_raw);
  halo_counts_init(S.levels, S.n_levels, P.barrier);

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
                            L.st_descent, s_raw, ms->prof_it[level < PROF_LEVELS ? level : PROF_LEVELS-1]);
        grid_barrier(P.barrier);
        /* other CTAs need the iteration count for n_equiv_iter */
        R.n_iter = *((volatile int *)&L.st_descent->n_iter);
        R.cvg = *((volatile int *)&L.st_descent->cvg);
        R.initial_residual = *((volatile double *)&L.st_descent->initial_residual);
        halo_count_follow(L, L.st_descent);
      }
      else {
        R = pcg_run<false>(L, rhs_lv, vx_lv, ini, S.n_max_iter[0],
                           S.precision*S.precision_mult[0], S.r_norm, true, P,
                           L.st_descent, s_raw, ms->prof_it[level < PROF_LEVELS ? level : PROF_LEVELS-1]);
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
        EpiResidual f;
        f.d_val = d_val; f.vx = vx_lv; f.rhs = rhs_lv; f.wr = wr;
        double acc[4];
        spmv_phase<false>(L, vx_lv, gh, s_raw, f, acc);
        w[0] = acc[0];
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
      restrict_rows<false>(C.n_rows, L.c_f_row_index, L.c_f_row_ids, S.wr, C.rhs);
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
        if (blockIdx.x == 0) {
          if (P.comm == nullptr && pcg_small_fits(L, S.small_rows))
            pcg_small(L, rhs_lv, vx_lv, S.n_max_iter[2], S.precision*S.precision_mult[2],
                      S.r_norm, true, L.st_descent, s_raw);
          else
            R = pcg_run<true>(L, rhs_lv, vx_lv, false, S.n_max_iter[2],
                              S.precision*S.precision_mult[2], S.r_norm, true, P,
                              L.st_descent, s_raw, ms->prof_it[coarsest < PROF_LEVELS ? coarsest : PROF_LEVELS-1]);
        }
        grid_barrier(P.barrier);
        R.n_iter = *((volatile int *)&L.st_descent->n_iter);
        R.cvg = *((volatile int *)&L.st_descent->cvg);
        halo_count_follow(L, L.st_descent);
      }
      else {
        R = pcg_run<false>(L, rhs_lv, vx_lv, false, S.n_max_iter[2],
                           S.precision*S.precision_mult[2], S.r_norm, true, P,
                           L.st_descent, s_raw, ms->prof_it[coarsest < PROF_LEVELS ? coarsest : PROF_LEVELS-1]);
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
          prolong_rows<false>(L.n_rows, coarse_row, c_vx, vx_lv);
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
                                L.st_ascent, s_raw, ms->prof_it[level < PROF_LEVELS ? level : PROF_LEVELS-1]);
            grid_barrier(P.barrier);
            R.n_iter = *((volatile int *)&L.st_ascent->n_iter);
            R.cvg = *((volatile int *)&L.st_ascent->cvg);
            halo_count_follow(L, L.st_ascent);
          }
          else {
            R = pcg_run<false>(L, L.rhs, vx_lv, true, S.n_max_iter[1],
                               S.precision*S.precision_mult[1], S.r_norm, true, P,
                               L.st_ascent, s_raw, ms->prof_it[level < PROF_LEVELS ? level : PROF_LEVELS-1]);
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

/*--------------------------------------------------------------------------
 * Kernel 3: flexible CG preconditioned by one V-cycle per iteration -- the
 * reference's production pressure-solver layout (cs_sles_it FCG,
 * _flexible_conjugate_gradient cs_sles_it.cpp:832-985, with cs_multigrid_pc,
 * _multigrid_v_cycle_pc cs_multigrid.cpp:3569-4020) and the Jacobi-family
 * cs_multigrid_smoother smoothers its device settings select (:5070-5090;
 * _jacobi / _relaxed_jacobi / _l1_jacobi, cs_multigrid_smoother.cpp:143-373,
 * :804-905).  The whole solve -- every FCG iteration, every level of every
 * cycle -- is ONE cooperative launch.
 *
 * A smoother sweep is one pass over the level's matrix: the (diagonal-
 * excluded) product in CSR order with the update in the epilogue of the row
 * sum.  Sweeps ping-pong between the level's vx and one shared buffer (tmp);
 * the buffer a smoother starts in is chosen so that its last sweep lands in
 * vx (no copy back): the prolongation writes vx + P.c_vx into that buffer, the
 * first sweep from zero (elementwise) is the tail of the phase that produced
 * its right-hand side (the FCG update at level 0, the restriction below).
 * No reduction anywhere in the cycle except the coarsest-level PCG: per level
 * 4 phases down (2 sweeps, residual, restriction) and 4 up (prolongation, 3
 * sweeps) with RJ3 x 3; the levels small enough for one CTA are run by CTA 0
 * as one block (descent, coarsest solve, ascent) between two grid barriers.
 *--------------------------------------------------------------------------*/

/* first sweep of a smoother from a zero start
   (cs_multigrid_smoother.cpp:196-199, :322-327, :858-861) */
__device__ __forceinline__ double smooth_first_value(int mode, double w0, double rhs, double inv)
{
  return (mode == SMOOTH_RELAXED) ? w0 * rhs*inv : rhs*inv;
}

/* one sweep: y = S(x, rhs, E.x), or S(x, rhs, A.x) for L1 (:205-213, :345-352, :880-887) */
struct EpiSmooth {
  static constexpr int NV = 4;
  int mode;
  double w0, w1;
  const double *rhs, *inv, *xold, *d_val;
  double *y;
  __device__ __forceinline__ int nv() const
  { return (mode == SMOOTH_JACOBI) ? 2 : (mode == SMOOTH_RELAXED) ? 3 : 4; }
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? rhs : (v == 1) ? inv : (v == 2) ? xold : d_val; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&)[4]) const
  {
    const double b = op[0], iv = op[S];
    if (mode == SMOOTH_JACOBI)
      y[r] = (b - sii) * iv;
    else if (mode == SMOOTH_RELAXED) {
      const double t = b - sii;
      y[r] = w1*op[2*S] + w0*t*iv;
    }
    else {
      const double xr = op[2*S];
      const double yr = sii + op[3*S]*xr;
      const double t = b - yr;
      y[r] = xr + t*iv;
    }
  }
};

struct RowSmoothFirst {
  static constexpr int NV = 2;
  int mode;
  double w0;
  const double *rhs, *inv;
  double *y;
  __device__ __forceinline__ const double *src(int v) const { return (v == 0) ? rhs : inv; }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  { y[i] = smooth_first_value(mode, w0, op[0], op[ROW_CH]); }
};

/* rk = rhs - A.vx, qk = 0 (cs_sles_it.cpp:874-900) + the preconditioner's first sweep */
struct EpiFcgPcInit {
  static constexpr int NV = 4;
  int mode;
  double w0;
  const double *d_val, *vx, *rhs, *inv;
  double *rk, *qk, *z0;
  __device__ __forceinline__ const double *src(int v) const
  { return (v == 0) ? d_val : (v == 1) ? vx : (v == 2) ? rhs : inv; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&)[4]) const
  {
    const double rr = op[2*S] - (sii + op[0]*op[S]);
    rk[r] = rr;
    qk[r] = 0;
    z0[r] = smooth_first_value(mode, w0, rr, op[3*S]);
  }
};

struct RowFcgPcInitZero {     /* vx = 0, rk = rhs, qk = 0 (:880-893) + first sweep */
  static constexpr int NV = 2;
  int mode;
  double w0;
  const double *rhs, *inv;
  double *vx, *rk, *qk, *z0;
  __device__ __forceinline__ const double *src(int v) const { return (v == 0) ? rhs : inv; }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  {
    vx[i] = 0;
    rk[i] = op[0];
    qk[i] = 0;
    z0[i] = smooth_first_value(mode, w0, op[0], op[ROW_CH]);
  }
};

/* dk, vx, qk, rk updates (:939-951; first iteration :960-972) + the preconditioner's first
   sweep from the new rk.  z0 may be vk itself: operands are read from the staged copy. */
struct RowFcgPcUpdate {
  static constexpr int NV = 7;
  int mode;
  double w0;
  double gk_rk1, ak_rk;
  int first;
  const double *vk, *wk, *inv;
  double *dk, *vx, *qk, *rk, *z0;
  __device__ __forceinline__ const double *src(int v) const
  {
    return (v == 0) ? vk : (v == 1) ? dk : (v == 2) ? vx : (v == 3) ? wk : (v == 4) ? qk
         : (v == 5) ? rk : inv;
  }
  __device__ __forceinline__ void st(int i, const double *op, double (&)[4]) const
  {
    const double d = first ? op[0] : op[0] - gk_rk1 * op[ROW_CH];
    dk[i] = d;
    vx[i] = op[2*ROW_CH] + ak_rk * d;
    const double q = first ? op[3*ROW_CH] : op[3*ROW_CH] - gk_rk1 * op[4*ROW_CH];
    qk[i] = q;
    const double rr = op[5*ROW_CH] - ak_rk * q;
    rk[i] = rr;
    z0[i] = smooth_first_value(mode, w0, rr, op[6*ROW_CH]);
  }
};

/* dst = vx + P.c_vx (cs_grid.cpp:9156-9166; rows without a coarse row keep vx) */
template <bool SINGLE>
__device__ __noinline__ void prolong_rows_to(int n, const int *__restrict__ coarse_row,
                                             const double *c_vx, const double *vx, double *dst)
{
  constexpr int U = CSB_PROLONG_U;
  const int first = SINGLE ? threadIdx.x : blockIdx.x*PT + threadIdx.x;
  const int step = SINGLE ? PT : gridDim.x*PT;
  for (int base = first; base < n; base += U*step) {
    int ic[U];
    double v[U], c[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = base + u*step;
      ic[u] = -2;
      if (i < n) { ic[u] = __ldg(coarse_row + i); v[u] = vx[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (ic[u] >= 0) c[u] = c_vx[ic[u]];
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (ic[u] >= 0) dst[base + u*step] = v[u] + c[u];
      else if (ic[u] == -1) dst[base + u*step] = v[u];
    }
  }
}

/* restriction (as restrict_rows) + the next level's first sweep from the restricted value */
template <bool SINGLE>
__device__ __noinline__ void restrict_rows_first(int nc, const int *__restrict__ cfi,
                                                 const int *__restrict__ cfr, const double *wr,
                                                 double *c_rhs, int mode, double w0,
                                                 const double *c_inv, double *c_start)
{
  constexpr int U = CSB_RESTRICT_U, J = 4;
  const int first = SINGLE ? threadIdx.x : blockIdx.x*PT + threadIdx.x;
  const int step = SINGLE ? PT : gridDim.x*PT;
  for (int base = first; base < nc; base += U*step) {
    int rs[U], cnt[U], id[U][J];
    double w[U][J], iv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int ic = base + u*step;
      rs[u] = 0; cnt[u] = 0; iv[u] = 0.;
      if (ic < nc) {
        rs[u] = __ldg(cfi + ic); cnt[u] = __ldg(cfi + ic + 1) - rs[u];
        if (c_start != nullptr) iv[u] = c_inv[ic];
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < J; j++)
        if (j < cnt[u]) id[u][j] = __ldg(cfr + rs[u] + j);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < J; j++)
        if (j < cnt[u]) w[u][j] = wr[id[u][j]];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int ic = base + u*step;
      if (ic < nc) {
        double s = 0.;
#pragma unroll
        for (int j = 0; j < J; j++)
          if (j < cnt[u]) s += w[u][j];
        for (int q = J; q < cnt[u]; q++) s += wr[__ldg(cfr + rs[u] + q)];
        c_rhs[ic] = s;
        if (c_start != nullptr) c_start[ic] = smooth_first_value(mode, w0, s, iv[u]);
      }
    }
  }
}

/* restriction through the slab pipeline: the 0/1 matrix (c_f_row_index, c_f_row_ids) times the
   residual, + the next level's first sweep from the restricted value (operand: its inverse
   diagonal) */
struct EpiRestrictFirst {
  static constexpr int NV = 1;
  int mode;
  double w0;
  const double *c_inv;
  double *c_rhs, *c_start;
  __device__ __forceinline__ int nv() const { return (c_start != nullptr) ? 1 : 0; }
  __device__ __forceinline__ const double *src(int) const { return c_inv; }
  template <int S>
  __device__ __forceinline__ void epi(int r, double sii, const double *op, double (&)[4]) const
  {
    c_rhs[r] = sii;
    if (c_start != nullptr) c_start[r] = smooth_first_value(mode, w0, sii, op[0]);
  }
};

struct VcProf {
  MgState *ms;
  unsigned long long t_last;
  bool lead;
  __device__ __forceinline__ void mark(int level, int phase)
  { if (lead) prof_mark(ms, t_last, level, phase); }
  __device__ __forceinline__ void mark_it(int level, int kind)
  {
    if (lead) {
      const unsigned long long t = global_ns();
      if (level < PROF_LEVELS) ms->prof_it[level][kind] += t - t_last;
      t_last = t;
    }
  }
};

__device__ __forceinline__ const double *smooth_inv(const MgLevelDesc &L, int mode)
{ return (mode == SMOOTH_L1) ? L.l1_inv : L.ad_inv; }

/* buffer a smoother of n sweeps must start in for its last sweep to land in vx */
__device__ __forceinline__ double *sweep_start(int n_toggles, double *vx_lv, double *tmp)
{ return (n_toggles & 1) ? tmp : vx_lv; }

/* sweeps it0 .. n_iter-1: each reads the other buffer and writes the current one */
template <bool SINGLE>
__device__ __forceinline__ void vc_sweeps(const MgLevelDesc &L, const SmoothPlanDev &sp,
                                          const double *rhs_lv, double *cur, double *oth, int it0,
                                          const PersistentCommon &P, unsigned char *s_raw)
{
  const double *inv = smooth_inv(L, sp.mode);
  for (int it = it0; it < sp.n_iter; it++) {
    double *t = cur; cur = oth; oth = t;
    const double wk = sp.w[it % sp.wk_m];
    const double *gh = halo_exchange<SINGLE>(L, oth);
    EpiSmooth f;
    f.mode = sp.mode; f.w0 = wk; f.w1 = 1.0 - wk;
    f.rhs = rhs_lv; f.inv = inv; f.xold = oth; f.d_val = L.d_val; f.y = cur;
    spmv_phase<SINGLE>(L, oth, gh, s_raw, f, nullptr);
    phase_barrier<SINGLE>(P.barrier);
  }
}

/* one level of the descent (cs_multigrid.cpp:3660-3790): smoother from zero, residual,
   restriction (+ first sweep of the next level unless it is the coarsest) */
template <bool SINGLE>
__device__ __forceinline__ void vc_descend(const FcgMgParams &S, const PersistentCommon &P,
                                           int level, const double *rhs_lv, double *vx_lv,
                                           bool first_done, unsigned char *s_raw, VcProf &pf)
{
  const MgLevelDesc &L = S.levels[level];
  const MgLevelDesc &C = S.levels[level+1];
  const SmoothPlanDev &sp = S.sm[0];
  double *start = sweep_start(sp.n_iter - 1, vx_lv, S.tmp);
  double *other = (start == vx_lv) ? S.tmp : vx_lv;
  if (!first_done) {
    RowSmoothFirst f;
    f.mode = sp.mode; f.w0 = sp.w[0]; f.rhs = rhs_lv; f.inv = smooth_inv(L, sp.mode); f.y = start;
    row_phase<SINGLE>(L.n_rows, s_raw, f, nullptr);
    phase_barrier<SINGLE>(P.barrier);
  }
  vc_sweeps<SINGLE>(L, sp, rhs_lv, start, other, 1, P, s_raw);
  pf.mark(level, MG_PROF_DESCENT);
  {
    const double *gh = halo_exchange<SINGLE>(L, vx_lv);
    EpiResidual f;
    f.d_val = L.d_val; f.vx = vx_lv; f.rhs = rhs_lv; f.wr = S.wr;
    double acc[4];
    spmv_phase<SINGLE>(L, vx_lv, gh, s_raw, f, acc);
    phase_barrier<SINGLE>(P.barrier);
  }
  pf.mark(level, MG_PROF_RESIDUAL);
  {
    /* not into the coarsest level (PCG from zero) nor across the merge boundary (the
       level's rows of the other ranks arrive later) */
    const bool fuse = (level + 1 < S.n_levels - 1) && (level + 1 != S.merge_level);
    double *c_start = fuse ? sweep_start(sp.n_iter - 1, C.vx, S.tmp) : nullptr;
    if (L.r_n_slabs > 0) {
      MgLevelDesc R;
      R.n_rows = L.n_rows_next; R.n_slabs = L.r_n_slabs; R.slab_rows = L.r_slab_rows;
      R.row_index = L.c_f_row_index; R.col_id = L.c_f_row_ids; R.x_val = nullptr;
      R.dist = nullptr;
      EpiRestrictFirst f;
      f.mode = sp.mode; f.w0 = sp.w[0]; f.c_inv = smooth_inv(C, sp.mode);
      f.c_rhs = C.rhs; f.c_start = c_start;
      spmv_phase_gv<SINGLE, 0, 2, true>(R, S.wr, nullptr, s_raw, f, nullptr);
    }
    else
      restrict_rows_first<SINGLE>(L.n_rows_next, L.c_f_row_index, L.c_f_row_ids, S.wr, C.rhs,
                                  sp.mode, sp.w[0], smooth_inv(C, sp.mode), c_start);
    phase_barrier<SINGLE>(P.barrier);
  }
  pf.mark(level, MG_PROF_RESTRICT);
}

/* one level of the ascent (:3880-3990): prolongation, smoother from the prolonged iterate */
template <bool SINGLE>
__device__ __forceinline__ void vc_ascend(const FcgMgParams &S, const PersistentCommon &P,
                                          int level, const double *rhs_lv, double *vx_lv,
                                          unsigned char *s_raw, VcProf &pf)
{
  const MgLevelDesc &L = S.levels[level];
  const MgLevelDesc &C = S.levels[level+1];
  const SmoothPlanDev &sp = S.sm[1];
  double *start = sweep_start(sp.n_iter, vx_lv, S.tmp);
  double *other = (start == vx_lv) ? S.tmp : vx_lv;
  prolong_rows_to<SINGLE>(L.n_rows, L.coarse_row, C.vx, vx_lv, start);
  phase_barrier<SINGLE>(P.barrier);
  pf.mark(level, MG_PROF_PROLONG);
  vc_sweeps<SINGLE>(L, sp, rhs_lv, start, other, 0, P, s_raw);
  pf.mark(level, MG_PROF_ASCENT);
}

/*--------------------------------------------------------------------------
 * Merge boundary (coarse levels merged onto rank 0; cs_grid_restrict_row_var /
 * cs_grid_prolong_row_var with merge_sub_size > 1, cs_grid.cpp:9058-9085,
 * :9131-9155): exchange number k of the launch sequence.
 *   gather : every other rank stores its rows of the merge level's rhs into
 *            rank 0's merged rhs and publishes k; rank 0 waits for all of them
 *   scatter: rank 0 stores each rank's rows of the level's vx into that rank's
 *            buffer and publishes k; the others wait for it
 * One buffer each way is enough: a rank sends gather k+1 only after scatter k
 * has told it that rank 0 is done with the level.
 *--------------------------------------------------------------------------*/

__device__ __forceinline__ void merge_gather(const MergeDev *M, unsigned k, unsigned *bar)
{
  if (M->is_root) {
    if (blockIdx.x == 0 && threadIdx.x > 0 && threadIdx.x < M->n_ranks) {
      spin_until_ge(M->flag_in + threadIdx.x, k, M->timeout_ns, 10);
      fence_sys();
    }
  }
  else {
    const double *src = M->stub_rhs;
    double *dst = M->root_rhs_dst;
    for (int i = blockIdx.x*PT + threadIdx.x; i < M->n_local; i += gridDim.x*PT) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) fence_sys();
  }
  grid_barrier(bar);
  if (!M->is_root && blockIdx.x == 0 && threadIdx.x == 0) {
    fence_sys();
    *(volatile unsigned *)M->root_flag = k;
  }
}

__device__ __forceinline__ void merge_scatter(const MergeDev *M, const double *vx_merged,
                                              unsigned k, unsigned *bar)
{
  if (M->is_root) {
    for (int r = 1; r < M->n_ranks; r++) {
      const double *src = vx_merged + M->shift[r];
      double *dst = M->peer_vx[r];
      for (int i = blockIdx.x*PT + threadIdx.x; i < M->n_of[r]; i += gridDim.x*PT) dst[i] = src[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) fence_sys();
    grid_barrier(bar);
    if (blockIdx.x == 0 && threadIdx.x > 0 && threadIdx.x < M->n_ranks) {
      fence_sys();
      *(volatile unsigned *)M->peer_flag[threadIdx.x] = k;
    }
  }
  else {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      spin_until_ge(M->flag_in, k, M->timeout_ns, 11);
      fence_sys();
    }
    grid_barrier(bar);
  }
}

__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_fcg_mg_persistent(const FcgMgParams S, const PersistentCommon P)
{
  unsigned char *s_raw = s_raw_base;
  double *s_red = (double *)s_raw;
  pipe_init(s_raw);
  halo_counts_init(S.levels, S.n_levels, P.barrier);

  const MgLevelDesc *__restrict__ LV = S.levels;
  const MgLevelDesc &L0 = LV[0];
  const int n = L0.n_rows;
  const int coarsest = S.n_levels - 1;
  double *rk = S.fwork, *vk = S.fwork + S.fstride, *wk = S.fwork + 2*S.fstride,
         *dk = S.fwork + 3*S.fstride, *qk = S.fwork + 4*S.fstride;
  ItState *st = S.st;
  const bool rec = (blockIdx.x == 0 && threadIdx.x == 0);
  VcProf pf;
  pf.ms = S.ms; pf.lead = rec; pf.t_last = rec ? global_ns() : 0ull;

  /* coarse levels merged onto rank 0: the other ranks' hierarchy ends with their part of
     the merge level (rhs out, vx back) */
  const MergeDev *MD = S.merge;
  const int Lm = (MD != nullptr) ? S.merge_level : -1;
  const bool merged_away = (MD != nullptr && !MD->is_root);
  unsigned mk = (MD != nullptr) ? *(volatile unsigned *)MD->count : 0u;

  /* first level small enough for one CTA: it and everything below is CTA 0's (never the
     levels down to the merge boundary, whose exchanges are grid-wide) */
  int ls = S.n_levels;
  for (int l = (Lm >= 0) ? Lm + 1 : 0; l < S.n_levels; l++)
    if (LV[l].n_rows <= S.single_rows) { ls = l; break; }
  if (merged_away) ls = S.n_levels;
  const int n_grid_desc = merged_away ? Lm : ((ls < coarsest) ? ls : coarsest);

  const SmoothPlanDev &sp0 = S.sm[0];
  const double *inv0 = smooth_inv(L0, sp0.mode);
  double *z0 = sweep_start(sp0.n_iter - 1, vk, S.tmp);

  PcgResult R;
  R.n_iter = 0; R.cvg = CSB_SLES_ITERATING; R.residual = -1.; R.initial_residual = -1.;
  if (rec) st->n_hist = 0;

  if (S.vx_ini_nonzero) {
    const double *gh = halo_exchange<false>(L0, S.vx);
    EpiFcgPcInit f;
    f.mode = sp0.mode; f.w0 = sp0.w[0];
    f.d_val = L0.d_val; f.vx = S.vx; f.rhs = S.rhs; f.inv = inv0; f.rk = rk; f.qk = qk; f.z0 = z0;
    spmv_phase<false>(L0, S.vx, gh, s_raw, f, nullptr);
  }
  else {
    RowFcgPcInitZero f;
    f.mode = sp0.mode; f.w0 = sp0.w[0];
    f.rhs = S.rhs; f.inv = inv0; f.vx = S.vx; f.rk = rk; f.qk = qk; f.z0 = z0;
    row_phase<false>(n, s_raw, f, nullptr);
  }
  grid_barrier(P.barrier);
  pf.mark_it(0, 0);

  double rho_km1 = 0., ak_rk = 0., gk_rk1 = 0.;
  while (R.cvg == CSB_SLES_ITERATING) {

    /* ---- vk = M^-1 rk: one V-cycle from zero (_multigrid_pc_apply, :1500-1600) ---- */
    if (MD != nullptr) mk += 1;
    for (int level = 0; level < n_grid_desc; level++) {
      if (level == Lm) merge_gather(MD, mk, P.barrier);       /* rank 0: rows of the others */
      vc_descend<false>(S, P, level, (level == 0) ? rk : LV[level].rhs,
                        (level == 0) ? vk : LV[level].vx, level != Lm, s_raw, pf);
    }
    if (merged_away) {
      merge_gather(MD, mk, P.barrier);
      merge_scatter(MD, nullptr, mk, P.barrier);
      pf.mark(Lm, MG_PROF_COARSEST);          /* waiting for rank 0's coarse levels */
    }
    else if (ls <= coarsest) {
      if (Lm == coarsest && coarsest >= n_grid_desc) merge_gather(MD, mk, P.barrier);
      if (blockIdx.x == 0) {
        for (int level = ls; level < coarsest; level++)
          vc_descend<true>(S, P, level, (level == 0) ? rk : LV[level].rhs,
                           (level == 0) ? vk : LV[level].vx, true, s_raw, pf);
        const MgLevelDesc &L = LV[coarsest];
        PersistentCommon Pl = P;
        if (L.dist == nullptr) Pl.comm = nullptr;   /* merged level: this rank alone */
        if (Pl.comm == nullptr && pcg_small_fits(L, S.small_rows))
          pcg_small(L, (coarsest == 0) ? rk : L.rhs, (coarsest == 0) ? vk : L.vx,
                    S.coarse_n_max_iter, S.precision*S.coarse_precision_mult, S.r_norm, true,
                    L.st_descent, s_raw);
        else
          pcg_run<true>(L, (coarsest == 0) ? rk : L.rhs, (coarsest == 0) ? vk : L.vx, false,
                        S.coarse_n_max_iter, S.precision*S.coarse_precision_mult, S.r_norm, true,
                        Pl, L.st_descent, s_raw);
        __syncthreads();
        pf.mark(coarsest, MG_PROF_COARSEST);
        for (int level = coarsest - 1; level >= ls; level--)
          vc_ascend<true>(S, P, level, (level == 0) ? rk : LV[level].rhs,
                          (level == 0) ? vk : LV[level].vx, s_raw, pf);
      }
      grid_barrier(P.barrier);
      /* exchanges CTA 0 did alone: every CTA takes the levels' counts from the source */
      if (threadIdx.x == 0)
        for (int l = ls; l <= coarsest && l < PROF_LEVELS; l++)
          if (LV[l].dist != nullptr)
            ((unsigned *)(s_raw_base + HALO_K_OFF))[LV[l].level_id]
              = *(volatile unsigned *)LV[l].dist->count;
      __syncthreads();
    }
    else {
      const MgLevelDesc &L = LV[coarsest];
      if (Lm == coarsest) merge_gather(MD, mk, P.barrier);
      PersistentCommon Pl = P;
      if (L.dist == nullptr) Pl.comm = nullptr;     /* merged level: this rank alone */
      pcg_run<false>(L, L.rhs, L.vx, false, S.coarse_n_max_iter,
                     S.precision*S.coarse_precision_mult, S.r_norm, true, Pl, L.st_descent,
                     s_raw);
      grid_barrier(P.barrier);
      pf.mark(coarsest, MG_PROF_COARSEST);
    }
    if (!merged_away && Lm == coarsest) merge_scatter(MD, LV[coarsest].vx, mk, P.barrier);
    for (int level = n_grid_desc - 1; level >= 0; level--) {
      vc_ascend<false>(S, P, level, (level == 0) ? rk : LV[level].rhs,
                       (level == 0) ? vk : LV[level].vx, s_raw, pf);
      if (!merged_away && level == Lm) merge_scatter(MD, LV[level].vx, mk, P.barrier);
    }

    /* ---- wk = A.vk and the four dot products (:908-918) ---- */
    double v[4] = {0., 0., 0., 0.};
    {
      const double *gh = halo_exchange<false>(L0, vk);
      EpiFcg f;
      f.d_val = L0.d_val; f.vk = vk; f.rk = rk; f.qk = qk; f.wk = wk;
      spmv_phase<false>(L0, vk, gh, s_raw, f, v);
    }
    all_sum<false, 4>(v, P.partials + 8*(size_t)gridDim.x, P.barrier, s_red, P.comm);
    pf.mark_it(0, 1);
    const double alpha_k = v[0], beta_k = v[1], gamma_k = v[2];
    R.residual = sqrt(v[3]);
    if (R.n_iter > 0)
      R.cvg = it_cvg_eval(S.precision, S.r_norm, S.n_max_iter, R.initial_residual, R.n_iter,
                          R.residual);
    else
      R.initial_residual = R.residual;
    if (rec && st->n_hist < HIST_MAX) st->hist[st->n_hist++] = R.residual;
    if (R.cvg != CSB_SLES_ITERATING) break;

    double rho_k;
    if (R.n_iter > 0) {
      gk_rk1 = (fabs(rho_km1) > DBL_MIN) ? gamma_k / rho_km1 : 0.;
      rho_k = beta_k - gamma_k * gk_rk1;
    }
    else {
      gk_rk1 = 0.;
      rho_k = beta_k;
    }
    ak_rk = (fabs(rho_k) > DBL_MIN) ? alpha_k / rho_k : 0.;
    {
      RowFcgPcUpdate f;
      f.mode = sp0.mode; f.w0 = sp0.w[0];
      f.gk_rk1 = gk_rk1; f.ak_rk = ak_rk; f.first = (R.n_iter == 0) ? 1 : 0;
      f.vk = vk; f.wk = wk; f.inv = inv0;
      f.dk = dk; f.vx = S.vx; f.qk = qk; f.rk = rk; f.z0 = z0;
      row_phase<false>(n, s_raw, f, nullptr);
    }
    grid_barrier(P.barrier);
    pf.mark_it(0, 2);
    rho_km1 = rho_k;
    R.n_iter += 1;
  }

  if (rec) {
    st->precision = S.precision; st->r_norm = S.r_norm;
    st->residual = R.residual; st->initial_residual = R.initial_residual;
    st->n_iter = R.n_iter; st->n_max_iter = S.n_max_iter; st->cvg = R.cvg;
    st->alpha = ak_rk; st->beta = gk_rk1; st->ro_0 = rho_km1; st->ro_1 = 0.; st->rk_gkm1 = 0.;
    if (MD != nullptr) *(volatile unsigned *)MD->count = mk;
  }
}


} // anonymous namespace

int plan_gather_depth(const std::vector<csb_lnum_t> &ri, int n_rows, int *row_len_max)
{
  static int want = -1;
  if (want < 0) {
    const char *s = getenv("CSB_GATHER_DEPTH");
    want = (s != nullptr) ? atoi(s) : ((NS >= 4) ? 3 : 2);
  }
  int mx = 1 << 30;
  if ((size_t)n_rows + 1 <= ri.size()) {
    mx = 0;
    for (int i = 0; i < n_rows; i++) mx = std::max(mx, (int)(ri[i+1] - ri[i]));
  }
  if (row_len_max != nullptr) *row_len_max = mx;
  /* three register sets need a fourth stage in flight (the stage two slabs ahead must have
     landed while the oldest one is still being summed) and rows of at most NG3 entries */
  return (want == 3 && NS >= 4 && mx <= NG3) ? 3 : 2;
}

void plan_slabs(const std::vector<csb_lnum_t> &ri, int n_rows, int &slab_rows, int &n_slabs,
                size_t &max_entries)
{
  /* Tallest slab (whole consumer warps: SLAB_ROWS_MAX, -32, ...) whose widest instance
     fits a pipeline stage of SLAB_CAP entries. */
  slab_rows = SLAB_ROWS_MAX;
  while (true) {
    max_entries = 0;
    for (int r0 = 0; r0 < n_rows; r0 += slab_rows) {
      const int r1 = (r0 + slab_rows < n_rows) ? r0 + slab_rows : n_rows;
      const size_t nk = (size_t)(ri[r1] - ri[r0]);
      if (nk > max_entries) max_entries = nk;
    }
    if (max_entries <= (size_t)SLAB_CAP || slab_rows <= 4) break;
    slab_rows = (slab_rows > 32) ? slab_rows - 32 : slab_rows / 2;
  }
  n_slabs = (n_rows + slab_rows - 1) / slab_rows;
}

size_t persistent_smem_bytes(size_t max_slab_entries)
{
  /* header (reduction scratch + mbarriers) + NS stages; a stage holds the widest slab
     (entries rounded up to 4; see the layout in the kernel) or a chunk of an
     elementwise phase, whichever is larger */
  size_t cap = (max_slab_entries + 3) / 4 * 4;
  if (cap < 4) cap = 4;
  size_t stage = 12*cap + ST_FIXED;
  if (stage < (size_t)ROW_STAGE_BYTES) stage = ROW_STAGE_BYTES;
  stage = (stage + 127) / 128 * 128;
  return SMEM_HDR + NS*stage;
}

/* number of co-resident CTAs for the given dynamic shared memory size */
int persistent_grid(int which, size_t smem)
{
  Context &c = ctx();
  int occ = 0;
  /* opt in to the device maximum once, not to this call's size: the attribute belongs to
     the (shared) context, and ranks living in one process ask for different sizes */
  int optin = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device);
  if ((size_t)optin < smem) {
    set_error("persistent kernel needs %zu bytes of shared memory, the device offers %d",
              smem, optin);
    return 0;
  }
  const void *k = (which == 0) ? (const void *)k_pcg_persistent
                : (which == 2) ? (const void *)k_fcg_persistent
                : (which == 3) ? (const void *)k_ipcg_persistent
                : (which == 4) ? (const void *)k_fcg_mg_persistent : (const void *)k_mg_persistent;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
  if (which == 0)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pcg_persistent, PT, smem);
  else if (which == 2)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fcg_persistent, PT, smem);
  else if (which == 3)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ipcg_persistent, PT, smem);
  else if (which == 4)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fcg_mg_persistent, PT, smem);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mg_persistent, PT, smem);
  if (occ < 1) {
    set_error("persistent kernel does not fit on an SM with %zu bytes of shared memory", smem);
    return 0;
  }
  static int cap = -1;
  if (cap < 0) {
    const char *s = getenv("CSB_CTAS_PER_SM");
    cap = (s != nullptr) ? atoi(s) : CSB_MIN_CTAS;
  }
  if (occ > cap && cap > 0) occ = cap;
  int g = occ * c.sm_count / (c.device_share > 1 ? c.device_share : 1);
  return (g > 0) ? g : 1;
}

#ifdef CSB_PHASE_TRACE
namespace {
/* Debug build only: latency of n grid barriers / n grid reductions in a row. */
__global__ void __launch_bounds__(PT, CSB_MIN_CTAS)
k_barrier_bench(unsigned *bar, double *partials, int n, int what, unsigned long long *out)
{
  unsigned char *s_raw = s_raw_base;
  pipe_init(s_raw);
  double *s_red = (double *)s_raw;
  grid_barrier(bar);
  const unsigned long long t0 = trace_now();
  double acc = 0.;
  for (int i = 0; i < n; i++) {
    if (what == 0) grid_barrier(bar);
    else {
      double v[3] = {1.0 + threadIdx.x, 2.0, 3.0 + i};
      all_sum<false, 3>(v, partials + ((i & 1) ? 4*(size_t)gridDim.x : 0), bar, s_red, nullptr);
      acc += v[0];
    }
  }
  const unsigned long long t1 = trace_now();
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (unsigned long long)acc; }
}
}
#endif

} // namespace csb
extern "C" double csb_debug_barrier_bench(int n, int what)
{
#ifdef CSB_PHASE_TRACE
  using namespace csb;
  Context &c = ctx();
  int optin = 0, occ = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device);
  const size_t smem = persistent_smem_bytes(1344);
  cudaFuncSetAttribute(k_barrier_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_barrier_bench, PT, smem);
  if (occ > CSB_MIN_CTAS) occ = CSB_MIN_CTAS;
  const int grid = occ * c.sm_count;
  unsigned *bar; double *partials; unsigned long long *out;
  cudaMalloc(&bar, 8); cudaMalloc(&partials, sizeof(double)*12*grid); cudaMalloc(&out, 16);
  cudaMemset(bar, 0, 8);
  void *args[] = {(void *)&bar, (void *)&partials, (void *)&n, (void *)&what, (void *)&out};
  cudaLaunchCooperativeKernel((void *)k_barrier_bench, dim3(grid), dim3(PT), args, smem, c.stream);
  unsigned long long h[2] = {0, 0};
  cudaStreamSynchronize(c.stream);
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  cudaFree(bar); cudaFree(partials); cudaFree(out);
  return (double)h[0] / n * 1e-3;      /* microseconds each */
#else
  (void)n; (void)what;
  return -1.;
#endif
}
extern "C" int csb_debug_phase_trace(double *out16, int reset)
{
#ifdef CSB_PHASE_TRACE
  unsigned long long h[16];
  if (cudaMemcpyFromSymbol(h, csb::g_trace, sizeof(h)) != cudaSuccess) return -1;
  for (int i = 0; i < 16; i++) out16[i] = (double)h[i];
  if (reset) { memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(csb::g_trace, h, sizeof(h)); }
  return 1;
#else
  (void)out16; (void)reset;
  return 0;
#endif
}
namespace csb {

/* the arrival counter of the grid barrier starts every launch at 0 */
static bool reset_barrier(const PersistentCommon &P)
{
  return CSB_CUDA(cudaMemsetAsync(P.barrier, 0, 2*sizeof(unsigned), ctx().stream));
}

bool launch_spmv_pipeline(const _csb_matrix_t *m, const double *x, double *y, bool exclude_diag,
                          bool *launched)
{
  *launched = false;
  Context &c = ctx();
  if (m->pl_state == 0) {
    m->pl_state = -1;
    if ((size_t)m->h_row_index.size() == (size_t)m->n_rows + 1 && m->n_rows > 0) {
      size_t max_entries = 0;
      plan_slabs(m->h_row_index, m->n_rows, m->pl_slab_rows, m->pl_n_slabs, max_entries);
      m->pl_gdepth = plan_gather_depth(m->h_row_index, m->n_rows, &m->pl_row_len_max);
      m->pl_smem = persistent_smem_bytes(max_entries);
      int optin = 0, occ = 0;
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device);
      if ((size_t)optin >= m->pl_smem && max_entries <= (size_t)SLAB_CAP) {
        cudaFuncSetAttribute(k_spmv_pipeline, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_pipeline, PT, m->pl_smem);
        if (occ > CSB_MIN_CTAS) occ = CSB_MIN_CTAS;
        if (occ >= 1) {
          m->pl_grid = occ * c.sm_count;
          if (m->pl_grid > m->pl_n_slabs) m->pl_grid = m->pl_n_slabs;
          m->pl_state = 1;
        }
      }
    }
  }
  /* the diagonal epilogue bulk-copies x: 16-byte aligned, whole 16-byte units */
  if (m->pl_state != 1) return true;
  if (!exclude_diag && ((((uintptr_t)x) & 15) != 0
                        || ((m->n_rows & 1) != 0 && m->n_cols_ext == m->n_rows))) return true;
  MgLevelDesc L;
  memset(&L, 0, sizeof(L));
  L.n_rows = m->n_rows; L.n_slabs = m->pl_n_slabs; L.slab_rows = m->pl_slab_rows;
  L.gather_depth = m->pl_gdepth; L.row_len_max = m->pl_row_len_max;
  L.row_index = m->row_index.p; L.col_id = m->col_id.p;
  L.d_val = m->d_val.p; L.x_val = m->x_val.p;
  k_spmv_pipeline<<<m->pl_grid, PT, m->pl_smem, c.stream>>>(L, x, y, exclude_diag ? 1 : 0);
  c.launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel k_spmv_pipeline launch failed: %s", cudaGetErrorString(e));
    return false;
  }
  *launched = true;
  return true;
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
  if (!reset_barrier(P)) return false;
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

bool launch_fcg_persistent(const MgLevelDesc &L, const double *rhs, double *vx,
                           bool vx_ini_nonzero, int n_max_iter, double precision,
                           double r_norm, const PersistentCommon &P, ItState *st,
                           int grid, size_t smem, bool single, bool cooperative, bool ipcg)
{
  int ini = vx_ini_nonzero ? 1 : 0;
  int sgl = single ? 1 : 0;
  void *args[] = {(void *)&L, (void *)&rhs, (void *)&vx, (void *)&ini, (void *)&n_max_iter,
                  (void *)&precision, (void *)&r_norm, (void *)&P, (void *)&st, (void *)&sgl};
  void *kern = ipcg ? (void *)k_ipcg_persistent : (void *)k_fcg_persistent;
  if (!reset_barrier(P)) return false;
  cudaError_t e = cooperative
    ? cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(PT), args, smem, ctx().stream)
    : cudaLaunchKernel(kern, dim3(grid), dim3(PT), args, smem, ctx().stream);
  ctx().launches += 1;
  if (e != cudaSuccess) {
    set_error("cooperative launch of k_%s_persistent failed: %s", ipcg ? "ipcg" : "fcg",
              cudaGetErrorString(e));
    return false;
  }
  return true;
}

bool launch_fcg_mg_persistent(const FcgMgParams &S, const PersistentCommon &P, int grid,
                              size_t smem, bool cooperative)
{
  void *args[] = {(void *)&S, (void *)&P};
  if (!reset_barrier(P)) return false;
  cudaError_t e = cooperative
    ? cudaLaunchCooperativeKernel((void *)k_fcg_mg_persistent, dim3(grid), dim3(PT), args, smem,
                                  ctx().stream)
    : cudaLaunchKernel((void *)k_fcg_mg_persistent, dim3(grid), dim3(PT), args, smem,
                       ctx().stream);
  ctx().launches += 1;
  if (e != cudaSuccess) {
    set_error("cooperative launch of k_fcg_mg_persistent failed: %s", cudaGetErrorString(e));
    return false;
  }
  return true;
}

bool launch_mg_persistent(const MgSolveParams &S, const PersistentCommon &P, int grid,
                          size_t smem, bool cooperative)
{
  void *args[] = {(void *)&S, (void *)&P};
  if (!reset_barrier(P)) return false;
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
