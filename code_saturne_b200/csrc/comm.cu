/*============================================================================
 * comm.cu -- peer-memory communicator between the GPUs of one NVSwitch domain.
 *
 * Replaces, for the solver path, the reference's MPI/NCCL plumbing
 * (src/base/cs_halo.cpp:518-587 grouped ncclSend/ncclRecv;
 * src/alge/cs_sles_it_cuda.cu:1000-1006 ncclAllReduce of 1-5 doubles;
 * src/alge/cs_grid.cpp:802-910 MPI exchange of coarsening ids).
 *
 * Every rank owns one "arena" (a single cudaMalloc) that all other ranks map:
 *   - one process per GPU: cudaIpcGetMemHandle / cudaIpcOpenMemHandle
 *   - several ranks as host threads of one process (tests on one GPU, or one
 *     process driving 8 GPUs): plain pointers + cudaDeviceEnablePeerAccess.
 * All communication is a kernel storing into a peer's arena over NVLink and
 * then publishing a monotonically increasing counter in a flag word of that
 * arena (payload stores, __threadfence_system(), flag store); the consumer
 * polls its own (local) flag word.  Nothing is ever reset, so there is no
 * reset race; buffers that are reused are protected either by an explicit
 * acknowledgement counter (mailbox, window) or by double buffering plus the
 * lock-step structure of the algorithm (all-reduce, see persistent.cu).
 *
 * Host-callable collectives here are used at setup time (hierarchy
 * construction) and by csb_halo_sync_var; the solve-time exchange and
 * all-reduce are fused into the persistent solver kernels (persistent.cu).
 *============================================================================*/

#include "csb_internal.h"

#include <condition_variable>
#include <mutex>

/* ranks living in one process as host threads */
struct CommGroup {
  int                      n_ranks = 0;
  std::mutex               m;
  std::condition_variable  cv;
  int                      waiting = 0;
  unsigned                 generation = 0;
  std::vector<char *>      arena;
  std::vector<int>         device;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = generation;
    if (++waiting == n_ranks) { waiting = 0; generation++; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != g; });
  }
};

namespace csb {

namespace {

constexpr size_t ARENA_ALIGN = 256;

inline size_t align_up(size_t v) { return (v + ARENA_ALIGN - 1) / ARENA_ALIGN * ARENA_ALIGN; }

/*--------------------------------------------------------------------------
 * all-reduce (sum) of n <= 8 doubles; one warp
 *--------------------------------------------------------------------------*/

__global__ void k_comm_allreduce(const CommDev *C, double *vals, int n)
{
  __shared__ double s_v[COMM_MAX_RANKS][8];
  const int R = C->n_ranks, me = C->rank, t = threadIdx.x;
  CommLocal *L = C->local;
  CommHeader *H = (CommHeader *)C->peer[me];
  const unsigned e = L->x_epoch + 1;
  const int par = e & 1;
  __syncwarp();
  if (t == 0) L->x_epoch = e;
  if (t < R) {
    CommHeader *P = (CommHeader *)C->peer[t];
    for (int q = 0; q < n; q++) P->xsum[par][me][q] = vals[q];
    __threadfence_system();
    *(volatile unsigned *)&P->xflag[par][me] = e;
    spin_until_ge(&H->xflag[par][t], e, C->timeout_ns, 1);
    __threadfence();
    for (int q = 0; q < n; q++) s_v[t][q] = __ldcg(&H->xsum[par][t][q]);
  }
  __syncwarp();
  if (t < n) {
    double w = 0.;
    for (int r = 0; r < R; r++) w += s_v[r][t];      /* rank order: same on every rank */
    vals[t] = w;
  }
}

/*--------------------------------------------------------------------------
 * mailbox: 8 values to / from each neighbour; one thread per neighbour
 *--------------------------------------------------------------------------*/

__global__ void k_comm_mailbox(const CommDev *C, int n_nbr, const int *nbr,
                               const long long *send, long long *recv)
{
  const int t = threadIdx.x;
  if (t >= n_nbr) return;
  const int me = C->rank, q = nbr[t];
  CommLocal *L = C->local;
  CommHeader *H = (CommHeader *)C->peer[me];
  CommHeader *P = (CommHeader *)C->peer[q];
  const unsigned s = L->m_sent[q] + 1;
  spin_until_ge(&H->mack[q], s - 1, C->timeout_ns, 2);           /* my previous message was consumed */
  for (int k = 0; k < 8; k++) P->mbox[me][k] = send[t*8 + k];
  __threadfence_system();
  *(volatile unsigned *)&P->mflag[me] = s;
  L->m_sent[q] = s;
  const unsigned r = L->m_recv[q] + 1;
  spin_until_ge(&H->mflag[q], r, C->timeout_ns, 3);
  __threadfence();
  for (int k = 0; k < 8; k++) recv[t*8 + k] = __ldcg(&H->mbox[q][k]);
  __threadfence_system();
  *(volatile unsigned *)&P->mack[me] = r;
  L->m_recv[q] = r;
}

/*--------------------------------------------------------------------------
 * window exchange: gather the send list into each neighbour's staging
 * window, then copy the own window to the ghost section.  W = 4 or 8 bytes.
 *--------------------------------------------------------------------------*/

template <typename T>
__global__ void __launch_bounds__(1024)
k_comm_window(const CommDev *C, const HaloDev h, int stride, const T *src, T *ghost_dst)
{
  const int t = threadIdx.x, nt = blockDim.x, me = C->rank;
  CommLocal *L = C->local;
  CommHeader *H = (CommHeader *)C->peer[me];
  if (t < h.n_dom)
    spin_until_ge(&H->wack[h.rank_of[t]], L->w_sent[h.rank_of[t]], C->timeout_ns, 4);
  __syncthreads();
  for (int i = 0; i < h.n_dom; i++) {
    const int q = h.rank_of[i];
    T *win = (T *)(C->peer[q] + C->window_off) + (size_t)h.peer_ghost_index[i] * stride;
    const int s0 = h.send_index[i], n = (h.send_index[i+1] - s0) * stride;
    for (int k = t; k < n; k += nt) {
      const int e = s0 + k / stride, c = k % stride;
      const int id = (h.send_list != nullptr) ? h.send_list[e] : e;
      win[k] = src[(size_t)id * stride + c];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (t < h.n_dom) {
    const int q = h.rank_of[t];
    CommHeader *P = (CommHeader *)C->peer[q];
    const unsigned s = L->w_sent[q] + 1;
    *(volatile unsigned *)&P->wflag[me] = s;
    L->w_sent[q] = s;
    spin_until_ge(&H->wflag[q], L->w_recv[q] + 1, C->timeout_ns, 5);
    __threadfence();
  }
  __syncthreads();
  const T *own = (const T *)(C->peer[me] + C->window_off);
  const int ng = h.ghost_index[h.n_dom] * stride;
  for (int k = t; k < ng; k += nt) ghost_dst[k] = __ldcg(own + k);
  __threadfence_system();
  __syncthreads();
  if (t < h.n_dom) {
    const int q = h.rank_of[t];
    CommHeader *P = (CommHeader *)C->peer[q];
    const unsigned r = L->w_recv[q] + 1;
    *(volatile unsigned *)&P->wack[me] = r;
    L->w_recv[q] = r;
  }
}

/*--------------------------------------------------------------------------
 * gather to rank 0 through its staging window, in chunks: sender q owns slot
 * q - 1 of the window.  Same per-pair counters as the window exchange (a chunk
 * is a message from q to 0: w_sent / wflag on the way, w_recv / wack back).
 * count[q] elements of rank q land at dst + offset[q] on rank 0.
 *--------------------------------------------------------------------------*/

struct GatherArgs {
  int       count[COMM_MAX_RANKS];
  long long offset[COMM_MAX_RANKS];
  int       slot_elems;                 /* elements per window slot */
};

template <typename T>
__global__ void __launch_bounds__(1024)
k_comm_gather(const CommDev *C, const GatherArgs g, const T *src, T *dst)
{
  const int t = threadIdx.x, nt = blockDim.x, me = C->rank, R = C->n_ranks;
  CommLocal *L = C->local;
  CommHeader *H = (CommHeader *)C->peer[me];
  if (me != 0) {
    CommHeader *P = (CommHeader *)C->peer[0];
    T *slot = (T *)(C->peer[0] + C->window_off) + (size_t)(me - 1) * g.slot_elems;
    for (int c0 = 0; c0 < g.count[me]; c0 += g.slot_elems) {
      const int n = min(g.slot_elems, g.count[me] - c0);
      if (t == 0) spin_until_ge(&H->wack[0], L->w_sent[0], C->timeout_ns, 8);
      __syncthreads();
      for (int k = t; k < n; k += nt) slot[k] = src[c0 + k];
      __threadfence_system();
      __syncthreads();
      if (t == 0) {
        const unsigned s = L->w_sent[0] + 1;
        *(volatile unsigned *)&P->wflag[me] = s;
        L->w_sent[0] = s;
      }
      __syncthreads();
    }
  }
  else {
    for (int k = t; k < g.count[0]; k += nt) dst[g.offset[0] + k] = src[k];
    for (int q = 1; q < R; q++) {
      CommHeader *P = (CommHeader *)C->peer[q];
      const T *slot = (const T *)(C->peer[0] + C->window_off) + (size_t)(q - 1) * g.slot_elems;
      for (int c0 = 0; c0 < g.count[q]; c0 += g.slot_elems) {
        const int n = min(g.slot_elems, g.count[q] - c0);
        if (t == 0) spin_until_ge(&H->wflag[q], L->w_recv[q] + 1, C->timeout_ns, 9);
        __syncthreads();
        __threadfence();
        for (int k = t; k < n; k += nt) dst[g.offset[q] + c0 + k] = __ldcg(slot + k);
        __threadfence_system();
        __syncthreads();
        if (t == 0) {
          const unsigned r = L->w_recv[q] + 1;
          *(volatile unsigned *)&P->wack[0] = r;
          L->w_recv[q] = r;
        }
        __syncthreads();
      }
    }
  }
}

bool launched(const char *name)
{
  ctx().launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel %s launch failed: %s", name, cudaGetErrorString(e));
    return false;
  }
  return true;
}

} // anonymous namespace

void *pcomm_arena_alloc(csb_pcomm_t *c, size_t bytes)
{
  const size_t off = align_up(c->arena_used);
  const size_t end = off + align_up(bytes > 0 ? bytes : 1);
  if (end > c->arena_bytes) {
    set_error("peer arena exhausted: %zu bytes requested, %zu of %zu in use "
              "(create the communicator with a larger arena)", bytes, c->arena_used,
              c->arena_bytes);
    return nullptr;
  }
  void *p = c->arena + off;
  if (!CSB_CUDA(cudaMemsetAsync(p, 0, end - off, ctx().stream))) return nullptr;
  c->arena_used = end;
  return p;
}

/* The arena is a stack of owner segments.  An owner (a solver or a hierarchy) opens a
   segment with pcomm_arena_mark before its allocations and closes it with
   pcomm_arena_release.  Owners may be freed in any order: a released segment is only
   handed out again once every segment above it has been released too, so a live
   owner's ghost buffers and flag words (which peers keep writing) are never reused. */
size_t pcomm_arena_mark(csb_pcomm_t *c)
{
  c->segs.push_back({c->arena_used, true});
  return c->arena_used;
}

void pcomm_arena_release(csb_pcomm_t *c, size_t mark)
{
  for (size_t i = c->segs.size(); i-- > 0; )
    if (c->segs[i].alive && c->segs[i].start == mark) { c->segs[i].alive = false; break; }
  while (!c->segs.empty() && !c->segs.back().alive) {
    c->arena_used = c->segs.back().start;
    c->segs.pop_back();
  }
}

void pcomm_host_barrier(csb_pcomm_t *c)
{
  if (c != nullptr && c->group != nullptr) c->group->barrier();
}

bool pcomm_allreduce(csb_pcomm_t *c, double *vals, int n)
{
  if (c == nullptr || c->h.n_ranks == 1) return true;
  if (n > 8) { set_error("pcomm_allreduce: at most 8 values"); return false; }
  Context &cx = ctx();
  for (int i = 0; i < n; i++) c->scratch_h[i] = vals[i];
  pcomm_host_barrier(c);
  if (!CSB_CUDA(cudaMemcpyAsync(c->scratch, c->scratch_h, sizeof(double)*n,
                                cudaMemcpyHostToDevice, cx.stream))) return false;
  k_comm_allreduce<<<1, 32, 0, cx.stream>>>(c->d, c->scratch, n);
  const bool ok = launched("comm_allreduce");
  /* in-process ranks: nobody blocks in the driver before every rank has
     launched (a blocking call may hold a lock the peers' launches need) */
  pcomm_host_barrier(c);
  if (!ok) return false;
  if (!CSB_CUDA(cudaMemcpyAsync(c->scratch_h, c->scratch, sizeof(double)*n,
                                cudaMemcpyDeviceToHost, cx.stream))) return false;
  if (!CSB_CUDA(cudaStreamSynchronize(cx.stream))) return false;
  for (int i = 0; i < n; i++) vals[i] = c->scratch_h[i];
  return true;
}

bool pcomm_mailbox(csb_pcomm_t *c, int n_nbr, const int *nbr, const long long *send,
                   long long *recv)
{
  if (c == nullptr) { set_error("pcomm_mailbox: no communicator"); return false; }
  if (n_nbr > COMM_MAX_DOM) {
    set_error("pcomm_mailbox: %d neighbours, at most %d supported", n_nbr, COMM_MAX_DOM);
    return false;
  }
  Context &cx = ctx();
  pcomm_host_barrier(c);
  if (n_nbr == 0) { pcomm_host_barrier(c); return true; }   /* keep the barrier count */
  DevBuf<int> d_nbr;
  DevBuf<long long> d_send, d_recv;
  bool ok = d_nbr.upload(nbr, n_nbr) && d_send.upload(send, (size_t)n_nbr*8)
         && d_recv.alloc((size_t)n_nbr*8);
  if (ok) {
    k_comm_mailbox<<<1, 32, 0, cx.stream>>>(c->d, n_nbr, d_nbr.p, d_send.p, d_recv.p);
    ok = launched("comm_mailbox");
  }
  /* every rank has launched before anyone blocks in the driver (copies to
     pageable memory, cudaFree: device-wide syncs that may hold driver locks) */
  pcomm_host_barrier(c);
  if (!ok) return false;
  ok = CSB_CUDA(cudaMemcpyAsync(recv, d_recv.p, sizeof(long long)*8*(size_t)n_nbr,
                                cudaMemcpyDeviceToHost, cx.stream))
    && CSB_CUDA(cudaStreamSynchronize(cx.stream));
  return ok;
}

bool pcomm_allgather_count(csb_pcomm_t *c, long long mine, long long *all)
{
  const int R = (c != nullptr) ? c->h.n_ranks : 1, me = (c != nullptr) ? c->h.rank : 0;
  for (int r0 = 0; r0 < R; r0 += 8) {           /* exact in doubles below 2^53 */
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (me >= r0 && me < r0 + 8) v[me - r0] = (double)mine;
    if (!pcomm_allreduce(c, v, 8)) return false;
    for (int k = 0; k < 8 && r0 + k < R; k++) all[r0 + k] = (long long)v[k];
  }
  return true;
}

bool pcomm_gather_to_root(csb_pcomm_t *c, int elem_bytes, const long long *counts,
                          const void *src_d, void *dst_d)
{
  if (c == nullptr) { set_error("pcomm_gather_to_root: no communicator"); return false; }
  const int R = c->h.n_ranks;
  GatherArgs g;
  long long off = 0;
  for (int r = 0; r < COMM_MAX_RANKS; r++) {
    g.count[r] = (r < R) ? (int)counts[r] : 0;
    g.offset[r] = off;
    off += g.count[r];
  }
  const size_t slot_bytes = (R > 1) ? (c->h.window_bytes / (size_t)(R - 1)) / 256 * 256 : 0;
  g.slot_elems = (int)(slot_bytes / (size_t)elem_bytes);
  if (R > 1 && g.slot_elems < 1) {
    set_error("pcomm_gather_to_root: staging window of %zu bytes is too small for %d ranks",
              c->h.window_bytes, R);
    return false;
  }
  Context &cx = ctx();
  /* The window is shared with the neighbour exchanges: no rank may still be reading its
     window (or a neighbour writing it) when the slots are used, nor the other way round
     afterwards.  The all-reduce is a device-side barrier over all ranks (a rank leaves it
     only when every rank has reached it in stream order). */
  double sync = 0.;
  if (!pcomm_allreduce(c, &sync, 1)) return false;
  pcomm_host_barrier(c);
  bool ok = true;
  if (elem_bytes == 8)
    k_comm_gather<double><<<1, 1024, 0, cx.stream>>>(c->d, g, (const double *)src_d,
                                                     (double *)dst_d);
  else if (elem_bytes == 4)
    k_comm_gather<int><<<1, 1024, 0, cx.stream>>>(c->d, g, (const int *)src_d, (int *)dst_d);
  else { set_error("pcomm_gather_to_root: element size %d", elem_bytes); ok = false; }
  if (ok) ok = launched("comm_gather");
  pcomm_host_barrier(c);
  return ok && pcomm_allreduce(c, &sync, 1);
}

bool pcomm_window_exchange(csb_pcomm_t *c, const HaloDev &h, int elem_bytes, int stride,
                           const void *src_d, void *ghost_dst_d)
{
  if (c == nullptr) { set_error("pcomm_window_exchange: no communicator"); return false; }
  const size_t need = (size_t)h.ghost_index[h.n_dom] * stride * elem_bytes;
  if (need > c->h.window_bytes) {
    set_error("pcomm_window_exchange: %zu bytes of ghosts exceed the %zu-byte window",
              need, c->h.window_bytes);
    return false;
  }
  Context &cx = ctx();
  pcomm_host_barrier(c);
  bool ok = true;
  if (h.n_dom > 0) {
    if (elem_bytes == 8)
      k_comm_window<double><<<1, 1024, 0, cx.stream>>>(c->d, h, stride, (const double *)src_d,
                                                       (double *)ghost_dst_d);
    else if (elem_bytes == 4)
      k_comm_window<int><<<1, 1024, 0, cx.stream>>>(c->d, h, stride, (const int *)src_d,
                                                    (int *)ghost_dst_d);
    else { set_error("pcomm_window_exchange: element size %d", elem_bytes); ok = false; }
    if (ok) ok = launched("comm_window");
  }
  pcomm_host_barrier(c);    /* all launched before the caller's next blocking call */
  return ok;
}

bool halo_dev_view(const _csb_halo_t *h, HaloDev &v)
{
  if (h->n_c_domains > COMM_MAX_DOM) {
    set_error("halo with %d neighbour ranks: at most %d supported", h->n_c_domains,
              COMM_MAX_DOM);
    return false;
  }
  if ((int)h->peer_ghost_index.size() != h->n_c_domains) {
    set_error("halo is not attached to a peer communicator (csb_halo_set_pcomm)");
    return false;
  }
  v.n_dom = h->n_c_domains;
  for (int i = 0; i < v.n_dom; i++) {
    v.rank_of[i] = h->c_domain_rank[i];
    v.send_index[i] = h->send_index[i];
    v.ghost_index[i] = h->index[i];
    v.peer_ghost_index[i] = h->peer_ghost_index[i];
  }
  v.send_index[v.n_dom] = h->send_index[v.n_dom];
  v.ghost_index[v.n_dom] = h->index[v.n_dom];
  v.send_list = h->send_list.p;
  return true;
}

} // namespace csb

using namespace csb;

namespace {

csb_pcomm_t *pcomm_new(int n_ranks, int rank, size_t arena_bytes, size_t window_bytes)
{
  if (!ctx().ready) { set_error("csb_pcomm_create: csb_init not called"); return nullptr; }
  if (n_ranks < 1 || n_ranks > COMM_MAX_RANKS || rank < 0 || rank >= n_ranks) {
    set_error("csb_pcomm_create: %d ranks (rank %d); at most %d supported", n_ranks, rank,
              COMM_MAX_RANKS);
    return nullptr;
  }
  csb_pcomm_t *c = new csb_pcomm_t;
  memset(&c->h, 0, sizeof(c->h));
  const size_t hdr = (sizeof(CommHeader) + ARENA_ALIGN - 1) / ARENA_ALIGN * ARENA_ALIGN;
  window_bytes = (window_bytes + ARENA_ALIGN - 1) / ARENA_ALIGN * ARENA_ALIGN;
  c->arena_bytes = hdr + window_bytes + arena_bytes;
  bool ok = CSB_CUDA(cudaMalloc((void **)&c->arena, c->arena_bytes))
         && CSB_CUDA(cudaMemset(c->arena, 0, hdr))
         && CSB_CUDA(cudaMalloc((void **)&c->h.local, sizeof(CommLocal)))
         && CSB_CUDA(cudaMemset(c->h.local, 0, sizeof(CommLocal)))
         && CSB_CUDA(cudaMalloc((void **)&c->d, sizeof(CommDev)))
         && CSB_CUDA(cudaMalloc((void **)&c->scratch, sizeof(double)*16))
         && CSB_CUDA(cudaMallocHost((void **)&c->scratch_h, sizeof(double)*16))
         && CSB_CUDA(cudaDeviceSynchronize());   /* memsets done before any peer stores */
  if (!ok) { delete c; return nullptr; }
  {
    const char *e = getenv("CSB_SPIN_TIMEOUT_S");
    const double t = (e != nullptr) ? atof(e) : 60.;
    c->h.timeout_ns = (unsigned long long)((t > 0.1 ? t : 0.1) * 1e9);
  }
  c->h.n_ranks = n_ranks;
  c->h.rank = rank;
  c->h.window_off = hdr;
  c->h.window_bytes = window_bytes;
  c->h.peer[rank] = c->arena;
  c->arena_used = hdr + window_bytes;
  return c;
}

bool pcomm_publish(csb_pcomm_t *c)
{
  return CSB_CUDA(cudaMemcpy(c->d, &c->h, sizeof(CommDev), cudaMemcpyHostToDevice));
}

} // anonymous namespace

extern "C" {

csb_pcomm_t *csb_pcomm_create(int n_ranks, int rank, size_t arena_bytes, size_t window_bytes)
{
  csb_pcomm_t *c = pcomm_new(n_ranks, rank, arena_bytes, window_bytes);
  if (c != nullptr && n_ranks == 1 && !pcomm_publish(c)) { csb_pcomm_destroy(&c); }
  return c;
}

int csb_pcomm_get_ipc_handle(csb_pcomm_t *c, void *handle_64_bytes)
{
  if (c == nullptr) return -1;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t hd;
  if (!CSB_CUDA(cudaIpcGetMemHandle(&hd, c->arena))) return -1;
  memcpy(handle_64_bytes, &hd, 64);
  return 0;
}

int csb_pcomm_open_peers(csb_pcomm_t *c, const void *all_handles)
{
  if (c == nullptr) return -1;
  for (int r = 0; r < c->h.n_ranks; r++) {
    if (r == c->h.rank) continue;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char *)all_handles + 64*(size_t)r, 64);
    void *p = nullptr;
    if (!CSB_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess))) return -1;
    c->h.peer[r] = (char *)p;
    c->ipc_open[r] = true;
  }
  return pcomm_publish(c) ? 0 : -1;
}

csb_pcomm_group_t *csb_pcomm_group_create(int n_ranks)
{
  if (n_ranks < 1 || n_ranks > COMM_MAX_RANKS) return nullptr;
  CommGroup *g = new CommGroup;
  g->n_ranks = n_ranks;
  g->arena.assign(n_ranks, nullptr);
  g->device.assign(n_ranks, -1);
  return (csb_pcomm_group_t *)g;
}

void csb_pcomm_group_destroy(csb_pcomm_group_t **group)
{
  if (group == nullptr || *group == nullptr) return;
  delete (CommGroup *)*group;
  *group = nullptr;
}

csb_pcomm_t *csb_pcomm_create_in_group(csb_pcomm_group_t *group, int rank, size_t arena_bytes,
                                       size_t window_bytes)
{
  CommGroup *g = (CommGroup *)group;
  if (g == nullptr) { set_error("csb_pcomm_create_in_group: null group"); return nullptr; }
  csb_pcomm_t *c = pcomm_new(g->n_ranks, rank, arena_bytes, window_bytes);
  /* every thread must reach the barriers, even on failure */
  g->arena[rank] = (c != nullptr) ? c->arena : nullptr;
  g->device[rank] = ctx().device;
  g->barrier();
  bool ok = (c != nullptr);
  int share = 0;
  for (int r = 0; r < g->n_ranks; r++) {
    if (g->arena[r] == nullptr) ok = false;
    if (g->device[r] == ctx().device) share++;
  }
  if (ok) {
    for (int r = 0; r < g->n_ranks; r++) {
      c->h.peer[r] = g->arena[r];
      if (g->device[r] != ctx().device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(g->device[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          set_error("cudaDeviceEnablePeerAccess(%d) failed: %s", g->device[r],
                    cudaGetErrorString(e));
          ok = false;
        }
        cudaGetLastError();
      }
    }
    c->group = g;
    ctx().device_share = share;
    ok = ok && pcomm_publish(c);
  }
  g->barrier();
  if (!ok && c != nullptr) { c->group = nullptr; csb_pcomm_destroy(&c); }
  return c;
}

void csb_pcomm_destroy(csb_pcomm_t **comm)
{
  if (comm == nullptr || *comm == nullptr) return;
  csb_pcomm_t *c = *comm;
  if (c->group != nullptr) c->group->barrier();   /* nobody still stores into this arena */
  cudaDeviceSynchronize();
  for (int r = 0; r < COMM_MAX_RANKS; r++)
    if (c->ipc_open[r]) cudaIpcCloseMemHandle(c->h.peer[r]);
  if (c->arena) cudaFree(c->arena);
  if (c->h.local) cudaFree(c->h.local);
  if (c->d) cudaFree(c->d);
  if (c->scratch) cudaFree(c->scratch);
  if (c->scratch_h) cudaFreeHost(c->scratch_h);
  delete c;
  *comm = nullptr;
}

int csb_pcomm_rank(const csb_pcomm_t *c) { return c ? c->h.rank : -1; }
int csb_pcomm_n_ranks(const csb_pcomm_t *c) { return c ? c->h.n_ranks : 0; }

int csb_pcomm_allreduce_sum(csb_pcomm_t *c, double *vals, int n)
{
  return pcomm_allreduce(c, vals, n) ? 0 : -1;
}

int csb_pcomm_barrier(csb_pcomm_t *c)
{
  double v[1] = {0.};
  return pcomm_allreduce(c, v, 1) ? 0 : -1;
}

} /* extern "C" */
