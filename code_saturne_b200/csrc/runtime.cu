/*============================================================================
 * runtime.cu -- context, memory helpers, matrix object, BLAS-1 entry points.
 *============================================================================*/

#include "csb_internal.h"

#include <cstdarg>
#include <cstdint>
#include <map>
#include <mutex>
#include <new>
#include <unordered_map>

namespace csb {

/* one context per host thread: a rank is a process, or a thread of one */
static thread_local Context g_ctx;

Context &ctx() { return g_ctx; }

void set_error(const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_ctx.err = buf;
  if (getenv("CSB_VERBOSE_ERRORS") != nullptr)
    fprintf(stderr, "[csb200] error: %s\n", buf);
}

bool cuda_ok(cudaError_t e, const char *what, const char *file, int line)
{
  if (e == cudaSuccess) return true;
  set_error("%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
  return false;
}

bool is_device_ptr(const void *p)
{
  if (p == nullptr) return false;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
}

void reset_partials();

} // namespace csb

using namespace csb;

namespace csb { void gs_plan_free(GsPlan *p); }
_csb_matrix_t::~_csb_matrix_t() { delete csr; if (gs_plan) csb::gs_plan_free(gs_plan); }

/*--------------------------------------------------------------------------
 * Host pool: page-locked blocks for the host side of the hierarchy setup (HostVec,
 * grid_host.h).  Blocks of at least POOL_MIN bytes are cudaHostAlloc'ed, rounded up to 2 MiB,
 * and cached on free (best fit on reuse, a block at most 1.5 times the request); smaller
 * requests go to malloc.  The cache is bounded (CSB_HOST_POOL_MAX_MB, default 16384): beyond
 * it freed blocks are returned to the driver.  Before csb_init (or without a device) the pool
 * falls back to malloc -- it is host memory either way.
 *--------------------------------------------------------------------------*/

namespace {
struct HostPool {
  std::mutex mtx;
  std::multimap<size_t, void *> cache;                 /* free pinned blocks by size */
  std::unordered_map<void *, size_t> live;             /* pinned blocks in use -> block size */
  size_t cached_bytes = 0, max_cached = 0;
  bool   enabled = true;
  HostPool() {
    const char *e = getenv("CSB_HOST_POOL_MAX_MB");
    max_cached = (size_t)(e != nullptr ? atoll(e) : 16384) << 20;
    const char *d = getenv("CSB_HOST_POOL");
    if (d != nullptr && atoi(d) == 0) enabled = false;
  }
  ~HostPool() { /* blocks are released with the process: the CUDA context may be gone */ }
};
HostPool &host_pool() { static HostPool *p = new HostPool; return *p; }
const size_t POOL_MIN = 1u << 20, POOL_GRAIN = 2u << 20;
}

namespace csb {

namespace {
bool dev_pool_enabled()
{
  static int v = -1;
  if (v < 0) { const char *e = getenv("CSB_DEV_POOL"); v = (e != nullptr && atoi(e) == 0) ? 0 : 1; }
  return v == 1;
}
/* blocks that came from the stream-ordered allocator, with the stream they belong to */
std::mutex g_dev_mtx;
std::unordered_map<void *, cudaStream_t> g_dev_async;
}

void *dev_alloc(size_t bytes)
{
  Context &c = ctx();
  void *p = nullptr;
#ifndef CUDASIM
  if (dev_pool_enabled() && c.ready && c.pool_ok) {
    if (cudaMallocAsync(&p, bytes, c.stream) == cudaSuccess && p != nullptr
        && (reinterpret_cast<uintptr_t>(p) & 255u) == 0) {
      std::lock_guard<std::mutex> lk(g_dev_mtx);
      g_dev_async[p] = c.stream;
      return p;
    }
    cudaGetLastError();
    if (p != nullptr) { cudaFreeAsync(p, c.stream); p = nullptr; }
  }
#endif
  if (!CSB_CUDA(cudaMalloc(&p, bytes))) return nullptr;
  return p;
}

void dev_free(void *p)
{
  if (p == nullptr) return;
  cudaStream_t s = nullptr;
  bool async = false;
  {
    std::lock_guard<std::mutex> lk(g_dev_mtx);
    auto it = g_dev_async.find(p);
    if (it != g_dev_async.end()) { s = it->second; async = true; g_dev_async.erase(it); }
  }
#ifndef CUDASIM
  if (async) {
    /* on the stream the block was allocated on; work enqueued on another stream since
       (csb_set_stream) is waited for first.  If the allocation stream is gone (csb_finalize
       before the object's destruction) cudaFree takes the block */
    Context &c = ctx();
    if (c.ready && c.stream != s) cudaStreamSynchronize(c.stream);
    if (cudaFreeAsync(p, s) == cudaSuccess) return;
    cudaGetLastError();
  }
#endif
  (void)async; (void)s;
  cudaFree(p);
}

} // namespace csb

extern "C" void *csb_host_pool_alloc(size_t bytes)
{
  if (bytes == 0) bytes = 1;
  HostPool &hp = host_pool();
  if (bytes < POOL_MIN || !hp.enabled || !csb::ctx().ready) {
    void *p = malloc(bytes);
    if (p == nullptr) throw std::bad_alloc();
    return p;
  }
  const size_t want = (bytes + POOL_GRAIN - 1)/POOL_GRAIN*POOL_GRAIN;
  {
    std::lock_guard<std::mutex> lk(hp.mtx);
    auto it = hp.cache.lower_bound(want);
    if (it != hp.cache.end() && it->first <= want + want/2) {
      void *p = it->second;
      const size_t sz = it->first;
      hp.cache.erase(it);
      hp.cached_bytes -= sz;
      hp.live[p] = sz;
      return p;
    }
  }
  void *p = nullptr;
  if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess || p == nullptr) {
    cudaGetLastError();
    p = malloc(bytes);                                 /* not registered in live: freed by free */
    if (p == nullptr) throw std::bad_alloc();
    return p;
  }
  std::lock_guard<std::mutex> lk(hp.mtx);
  hp.live[p] = want;
  return p;
}

extern "C" void csb_host_pool_free(void *p, size_t bytes)
{
  (void)bytes;
  if (p == nullptr) return;
  HostPool &hp = host_pool();
  size_t sz = 0;
  {
    std::lock_guard<std::mutex> lk(hp.mtx);
    auto it = hp.live.find(p);
    if (it == hp.live.end()) { free(p); return; }
    sz = it->second;
    hp.live.erase(it);
    if (hp.cached_bytes + sz <= hp.max_cached) {
      hp.cache.emplace(sz, p);
      hp.cached_bytes += sz;
      return;
    }
  }
  cudaFreeHost(p);
}

/* release the cached blocks (csb_finalize; also callable to trim host memory) */
extern "C" void csb_host_pool_trim(void)
{
  HostPool &hp = host_pool();
  std::lock_guard<std::mutex> lk(hp.mtx);
  for (auto &kv : hp.cache) cudaFreeHost(kv.second);
  hp.cache.clear();
  hp.cached_bytes = 0;
}

extern "C" {

const char *csb_version(void) { return "csb200 0.1 (sm_100a, FP64)"; }
const char *csb_spmv_kernel_name(void) { return csb::spmv_kernel_name(); }

const char *csb_last_error(void) { return ctx().err.c_str(); }

int csb_get_device(void) { return ctx().ready ? ctx().device : -1; }

int csb_init(int device_id)
{
  Context &c = ctx();
  if (c.ready && c.device == device_id) return 0;
  if (c.ready) csb_finalize();
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) {
    cudaGetLastError();
    set_error("csb_init: no CUDA device available (this library has no CPU fallback)");
    return -1;
  }
  if (device_id < 0 || device_id >= n_dev) {
    set_error("csb_init: device %d out of range (%d devices)", device_id, n_dev);
    return -1;
  }
  if (!CSB_CUDA(cudaSetDevice(device_id))) return -1;
  cudaDeviceProp prop;
  if (!CSB_CUDA(cudaGetDeviceProperties(&prop, device_id))) return -1;
  c.sm_count = prop.multiProcessorCount;
  c.device = device_id;
  if (!CSB_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking))) return -1;
  if (!CSB_CUDA(cudaStreamCreateWithFlags(&c.comm_stream, cudaStreamNonBlocking))) return -1;
  c.stream = c.own_stream;
#ifndef CUDASIM
  {
    /* stream-ordered allocator (dev_alloc): keep freed blocks in the device's default pool */
    int pools = 0;
    cudaMemPool_t mp;
    if (cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, device_id) == cudaSuccess
        && pools != 0 && cudaDeviceGetDefaultMemPool(&mp, device_id) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      if (cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr) == cudaSuccess)
        c.pool_ok = true;
    }
    cudaGetLastError();
  }
#endif
  if (!CSB_CUDA(cudaEventCreateWithFlags(&c.ev_a, cudaEventDisableTiming))) return -1;
  if (!CSB_CUDA(cudaEventCreateWithFlags(&c.ev_b, cudaEventDisableTiming))) return -1;
  if (!CSB_CUDA(cudaMalloc((void **)&c.ticket, sizeof(unsigned)))) return -1;
  if (!CSB_CUDA(cudaMemset(c.ticket, 0, sizeof(unsigned)))) return -1;
  if (!CSB_CUDA(cudaMalloc((void **)&c.red_out, sizeof(double)*8))) return -1;
  if (!CSB_CUDA(cudaMallocHost((void **)&c.red_out_h, sizeof(double)*8))) return -1;
  c.partials = nullptr;
  reset_partials();
  c.launches = 0;
  c.ready = true;
  c.err.clear();
  return 0;
}

void csb_finalize(void)
{
  Context &c = ctx();
  if (!c.ready) return;
  cudaDeviceSynchronize();
  if (c.partials) cudaFree(c.partials);
  if (c.ticket) cudaFree(c.ticket);
  if (c.red_out) cudaFree(c.red_out);
  if (c.red_out_h) cudaFreeHost(c.red_out_h);
  if (c.ev_a) cudaEventDestroy(c.ev_a);
  if (c.ev_b) cudaEventDestroy(c.ev_b);
  if (c.own_stream) cudaStreamDestroy(c.own_stream);
  if (c.comm_stream) cudaStreamDestroy(c.comm_stream);
  c = Context();
  reset_partials();
}

int csb_set_stream(void *cuda_stream)
{
  Context &c = ctx();
  if (!c.ready) { set_error("csb_set_stream: csb_init not called"); return -1; }
  /* blocks of the stream-ordered allocator are handed from the old stream to the new one */
  if (c.stream != nullptr) cudaStreamSynchronize(c.stream);
  c.stream = (cuda_stream != nullptr) ? (cudaStream_t)cuda_stream : c.own_stream;
  return 0;
}

int csb_stream_synchronize(void)
{
  return CSB_CUDA(cudaStreamSynchronize(ctx().stream)) ? 0 : -1;
}

void *csb_malloc_device(size_t n_bytes)
{
  void *p = nullptr;
  if (!CSB_CUDA(cudaMalloc(&p, n_bytes > 0 ? n_bytes : 1))) return nullptr;
  return p;
}

void csb_free_device(void *p) { if (p) cudaFree(p); }

int csb_copy_h2d(void *dst_d, const void *src_h, size_t n_bytes)
{
  if (!CSB_CUDA(cudaMemcpyAsync(dst_d, src_h, n_bytes, cudaMemcpyHostToDevice, ctx().stream)))
    return -1;
  return CSB_CUDA(cudaStreamSynchronize(ctx().stream)) ? 0 : -1;
}

int csb_copy_d2h(void *dst_h, const void *src_d, size_t n_bytes)
{
  if (!CSB_CUDA(cudaMemcpyAsync(dst_h, src_d, n_bytes, cudaMemcpyDeviceToHost, ctx().stream)))
    return -1;
  return CSB_CUDA(cudaStreamSynchronize(ctx().stream)) ? 0 : -1;
}

void *csb_malloc_pinned(size_t n_bytes)
{
  void *p = nullptr;
  if (!CSB_CUDA(cudaMallocHost(&p, n_bytes > 0 ? n_bytes : 1))) return nullptr;
  return p;
}

void csb_free_pinned(void *p) { if (p) cudaFreeHost(p); }

long long csb_launch_count(int reset)
{
  long long n = ctx().launches;
  if (reset) ctx().launches = 0;
  return n;
}

/*----------------------------------------------------------------------------
 * Matrix
 *----------------------------------------------------------------------------*/

csb_matrix_t *
csb_matrix_create_msr(csb_lnum_t n_rows, csb_lnum_t n_cols_ext,
                      const csb_lnum_t *row_index, const csb_lnum_t *col_id,
                      const csb_real_t *d_val, const csb_real_t *x_val, int symmetric)
{
  Context &c = ctx();
  if (!c.ready) { set_error("csb_matrix_create_msr: csb_init not called"); return nullptr; }
  if (n_rows < 0 || n_cols_ext < n_rows || row_index == nullptr) {
    set_error("csb_matrix_create_msr: invalid sizes"); return nullptr;
  }
  csb_matrix_t *m = new csb_matrix_t;
  m->n_rows = n_rows;
  m->n_cols_ext = n_cols_ext;
  m->symmetric = (symmetric != 0);

  m->h_row_index.resize((size_t)n_rows + 1);
  if (is_device_ptr(row_index)) {
    if (!CSB_CUDA(cudaMemcpy(m->h_row_index.data(), row_index,
                             sizeof(csb_lnum_t)*((size_t)n_rows+1), cudaMemcpyDeviceToHost))) {
      delete m; return nullptr;
    }
  }
  else
    memcpy(m->h_row_index.data(), row_index, sizeof(csb_lnum_t)*((size_t)n_rows+1));
  m->nnz = m->h_row_index[n_rows];

  bool ok = m->row_index.upload(row_index, (size_t)n_rows + 1)
         && m->col_id.upload(col_id, (size_t)m->nnz)
         && m->d_val.upload(d_val, (size_t)n_rows)
         && m->x_val.upload(x_val, (size_t)m->nnz)
         && CSB_CUDA(cudaStreamSynchronize(c.stream));
  if (!ok) { delete m; return nullptr; }
  matrix_plan_spmv(m);
  return m;
}

/* Matrix mapped from native (face-based) arrays: cs_matrix_structure_create(CS_MATRIX_MSR,
   n_rows, n_cols_ext, n_edges, edges) + cs_matrix_set_coefficients(m, symmetric, 1, 1,
   n_edges, edges, da, xa) (src/alge/cs_matrix.cpp:4328, :5315).  The MSR structure is built
   on the host (sorted columns), the values are placed on the device; the faces and their
   values stay with the matrix: csb_multigrid_setup then follows the reference's path for
   such a matrix (cs_grid_create_from_shared, cs_grid.cpp:7485-7491). */
csb_matrix_t *
csb_matrix_create_native(csb_lnum_t n_rows, csb_lnum_t n_cols_ext, csb_lnum_t n_edges,
                         const csb_lnum_t *edges, const csb_real_t *da, const csb_real_t *xa,
                         int symmetric)
{
  Context &c = ctx();
  if (!c.ready) { set_error("csb_matrix_create_native: csb_init not called"); return nullptr; }
  if (n_rows < 0 || n_cols_ext < n_rows || n_edges < 0 || da == nullptr
      || (n_edges > 0 && (edges == nullptr || xa == nullptr))) {
    set_error("csb_matrix_create_native: invalid arguments"); return nullptr;
  }
  if (!symmetric) {
    set_error("csb_matrix_create_native: symmetric matrices only (one value per face)");
    return nullptr;
  }
  csb_matrix_t *m = new csb_matrix_t;
  m->n_rows = n_rows;
  m->n_cols_ext = n_cols_ext;
  m->symmetric = true;
  m->from_native = true;
  m->n_faces = n_edges;
  m->h_face_cell.resize(2*(size_t)std::max(n_edges, 1));
  m->h_xa.resize((size_t)std::max(n_edges, 1));
  m->h_d_val.resize((size_t)std::max(n_rows, 1));
  bool ok = true;
  if (n_edges > 0)
    ok = CSB_CUDA(cudaMemcpy(m->h_face_cell.data(), edges, sizeof(csb_lnum_t)*2*(size_t)n_edges,
                             cudaMemcpyDefault))
      && CSB_CUDA(cudaMemcpy(m->h_xa.data(), xa, sizeof(double)*(size_t)n_edges,
                             cudaMemcpyDefault));
  if (ok && n_rows > 0)
    ok = CSB_CUDA(cudaMemcpy(m->h_d_val.data(), da, sizeof(double)*(size_t)n_rows,
                             cudaMemcpyDefault));
  if (!ok) { delete m; return nullptr; }
  for (csb_lnum_t f = 0; f < n_edges; f++) {
    const csb_lnum_t a = m->h_face_cell[2*(size_t)f], b = m->h_face_cell[2*(size_t)f+1];
    if (a < 0 || b < 0 || a >= n_cols_ext || b >= n_cols_ext || a == b) {
      set_error("csb_matrix_create_native: edge %d joins cells %d and %d", f, a, b);
      delete m; return nullptr;
    }
  }
  csb::HostVec<csb_lnum_t> c2f;
  csb::HostVec<short> c2f_sgn;
  if (!csb::msr_structure_from_faces(n_rows, n_edges, m->h_face_cell.data(), m->h_row_index,
                                     m->h_col_id, c2f, c2f_sgn)) {
    set_error("csb_matrix_create_native: two edges join the same pair of cells (values to be "
              "added by the caller first)");
    delete m; return nullptr;
  }
  m->nnz = m->h_row_index[n_rows];
  ok = m->row_index.upload(m->h_row_index.data(), (size_t)n_rows + 1)
    && m->col_id.upload(m->h_col_id.data(), (size_t)m->nnz)
    && m->c2f.upload(c2f.data(), (size_t)m->nnz)
    && m->c2f_sgn.upload(c2f_sgn.data(), (size_t)m->nnz)
    && m->face_cell.upload(m->h_face_cell.data(), 2*(size_t)n_edges)
    && m->xa.upload(m->h_xa.data(), (size_t)n_edges)
    && m->d_val.upload(m->h_d_val.data(), (size_t)n_rows)
    && m->x_val.alloc((size_t)m->nnz)
    && csb::launch_native_to_msr(m->nnz, m->c2f.p, m->xa.p, m->x_val.p)
    && CSB_CUDA(cudaStreamSynchronize(c.stream));
  if (!ok) { delete m; return nullptr; }
  matrix_plan_spmv(m);
  return m;
}

/* Host arrays lent to the matrix: the reference's matrices share the caller's arrays
   (cs_matrix_set_coefficients keeps pointers unless told to copy, src/alge/cs_matrix.cpp:5451);
   here the device holds the working copy and the host arrays, when lent, serve the host side
   of the hierarchy setup (aggregation of level 0) without a 1.3 GB device-to-host copy. */
int csb_matrix_share_host_arrays(csb_matrix_t *m, const csb_lnum_t *col_id,
                                 const csb_real_t *d_val, const csb_real_t *x_val)
{
  if (m == nullptr) { set_error("csb_matrix_share_host_arrays: null matrix"); return -1; }
  if (m->db != 1 || m->csr != nullptr) {
    set_error("csb_matrix_share_host_arrays: scalar MSR matrices only");
    return -1;
  }
  if ((col_id != nullptr && is_device_ptr(col_id)) || (d_val != nullptr && is_device_ptr(d_val))
      || (x_val != nullptr && is_device_ptr(x_val))) {
    set_error("csb_matrix_share_host_arrays: host pointers expected");
    return -1;
  }
  m->b_col_id = col_id;
  m->b_d_val = d_val;
  m->b_x_val = x_val;
  return 0;
}

csb_matrix_t *
csb_matrix_create_msr_block(csb_lnum_t n_rows, csb_lnum_t n_cols_ext, int db_size,
                            const csb_lnum_t *row_index, const csb_lnum_t *col_id,
                            const csb_real_t *d_val, const csb_real_t *x_val, int symmetric)
{
  if (db_size == 1)
    return csb_matrix_create_msr(n_rows, n_cols_ext, row_index, col_id, d_val, x_val, symmetric);
  Context &c = ctx();
  if (!c.ready) { set_error("csb_matrix_create_msr_block: csb_init not called"); return nullptr; }
  if (db_size < 2 || db_size > 9) {     /* DB_SIZE_MAX, cs_sles_it_priv.h:92 */
    set_error("csb_matrix_create_msr_block: diagonal block size %d (1 to 9 are available)",
              db_size);
    return nullptr;
  }
  if (n_rows < 0 || n_cols_ext < n_rows || row_index == nullptr) {
    set_error("csb_matrix_create_msr_block: invalid sizes"); return nullptr;
  }
  csb_matrix_t *m = new csb_matrix_t;
  m->n_rows = n_rows;
  m->n_cols_ext = n_cols_ext;
  m->db = db_size;
  m->symmetric = (symmetric != 0);
  m->h_row_index.resize((size_t)n_rows + 1);
  if (is_device_ptr(row_index)) {
    if (!CSB_CUDA(cudaMemcpy(m->h_row_index.data(), row_index,
                             sizeof(csb_lnum_t)*((size_t)n_rows+1), cudaMemcpyDeviceToHost))) {
      delete m; return nullptr;
    }
  }
  else
    memcpy(m->h_row_index.data(), row_index, sizeof(csb_lnum_t)*((size_t)n_rows+1));
  m->nnz = m->h_row_index[n_rows];
  bool ok = m->row_index.upload(row_index, (size_t)n_rows + 1)
         && m->col_id.upload(col_id, (size_t)m->nnz)
         && m->d_val.upload(d_val, (size_t)n_rows*db_size*db_size)
         && m->x_val.upload(x_val, (size_t)m->nnz)
         && CSB_CUDA(cudaStreamSynchronize(c.stream));
  if (!ok) { delete m; return nullptr; }
  matrix_plan_spmv(m);
  return m;
}

int csb_matrix_get_diag_block_size(const csb_matrix_t *m) { return m ? m->db : 0; }

csb_matrix_t *
csb_matrix_create_csr(csb_lnum_t n_rows, csb_lnum_t n_cols_ext,
                      const csb_lnum_t *row_index, const csb_lnum_t *col_id,
                      const csb_real_t *val, int symmetric)
{
  Context &c = ctx();
  if (!c.ready) { set_error("csb_matrix_create_csr: csb_init not called"); return nullptr; }
  if (n_rows < 0 || n_cols_ext < n_rows || row_index == nullptr) {
    set_error("csb_matrix_create_csr: invalid sizes"); return nullptr;
  }
  if (is_device_ptr(row_index) || is_device_ptr(col_id) || is_device_ptr(val)) {
    set_error("csb_matrix_create_csr: host arrays expected (the MSR split is made on the host)");
    return nullptr;
  }
  /* MSR split: diagonal out, the other entries keep their order */
  const size_t nnz = (size_t)row_index[n_rows];
  std::vector<csb_lnum_t> ri((size_t)n_rows + 1, 0), ci;
  std::vector<double> dv((size_t)n_rows, 0.0), xv;
  ci.reserve(nnz); xv.reserve(nnz);
  for (csb_lnum_t i = 0; i < n_rows; i++) {
    for (csb_lnum_t k = row_index[i]; k < row_index[i+1]; k++) {
      if (col_id[k] == i) dv[i] += val[k];
      else { ci.push_back(col_id[k]); xv.push_back(val[k]); }
    }
    ri[(size_t)i + 1] = (csb_lnum_t)ci.size();
  }
  csb_lnum_t dummy_i = 0; double dummy_d = 0.;
  csb_matrix_t *m = csb_matrix_create_msr(n_rows, n_cols_ext, ri.data(),
                                          ci.empty() ? &dummy_i : ci.data(), dv.data(),
                                          xv.empty() ? &dummy_d : xv.data(), symmetric);
  if (m == nullptr) return nullptr;
  /* the CSR arrays as an "off-diagonal only" matrix: E.x of it is the CSR product */
  std::vector<double> zero((size_t)n_rows, 0.0);
  m->csr = csb_matrix_create_msr(n_rows, n_cols_ext, row_index, col_id, zero.data(), val,
                                 symmetric);
  if (m->csr == nullptr) { delete m; return nullptr; }
  return m;
}

int csb_matrix_set_coefficients(csb_matrix_t *m, int symmetric,
                                const csb_real_t *d_val, const csb_real_t *x_val)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (d_val == nullptr || (x_val == nullptr && m->nnz > 0)) {
    set_error("csb_matrix_set_coefficients: null coefficient array"); return -1;
  }
  if (m->csr != nullptr) {
    /* the CSR value array kept for cs_matrix_vector_multiply is not expressible in the MSR
       arguments of this call: refuse rather than multiply with stale values */
    set_error("csb_matrix_set_coefficients: matrix was created from CSR arrays; create it "
              "again with csb_matrix_create_csr to change its values");
    return -1;
  }
  if (m->from_native) {
    set_error("csb_matrix_set_coefficients: matrix was created from native arrays; use "
              "csb_matrix_set_coefficients_native(da, xa)");
    return -1;
  }
  m->symmetric = (symmetric != 0);
  Context &c = ctx();
  /* d_val holds db x db blocks per row (CS_MATRIX_BLOCK_D: n_rows*9 values) */
  const size_t n_d = (size_t)m->n_rows * (size_t)m->db * (size_t)m->db;
  bool ok = CSB_CUDA(cudaMemcpyAsync(m->d_val.p, d_val, sizeof(double)*n_d,
                                     cudaMemcpyDefault, c.stream))
         && CSB_CUDA(cudaMemcpyAsync(m->x_val.p, x_val, sizeof(double)*(size_t)m->nnz,
                                     cudaMemcpyDefault, c.stream))
         && CSB_CUDA(cudaStreamSynchronize(c.stream));
  m->h_vals_valid = false;
  m->b_d_val = nullptr; m->b_x_val = nullptr;   /* lent host values are stale now */
  return ok ? 0 : -1;
}

/* New values on the same edges (cs_matrix_set_coefficients called again with the edges the
   structure was built from) */
int csb_matrix_set_coefficients_native(csb_matrix_t *m, const csb_real_t *da,
                                       const csb_real_t *xa)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!m->from_native || da == nullptr || (xa == nullptr && m->n_faces > 0)) {
    set_error("csb_matrix_set_coefficients_native: needs a matrix made by "
              "csb_matrix_create_native and both arrays");
    return -1;
  }
  Context &c = ctx();
  bool ok = CSB_CUDA(cudaMemcpy(m->h_d_val.data(), da, sizeof(double)*(size_t)m->n_rows,
                                cudaMemcpyDefault))
         && (m->n_faces == 0
             || CSB_CUDA(cudaMemcpy(m->h_xa.data(), xa, sizeof(double)*(size_t)m->n_faces,
                                    cudaMemcpyDefault)))
         && CSB_CUDA(cudaMemcpyAsync(m->d_val.p, m->h_d_val.data(),
                                     sizeof(double)*(size_t)m->n_rows, cudaMemcpyHostToDevice,
                                     c.stream))
         && (m->n_faces == 0
             || CSB_CUDA(cudaMemcpyAsync(m->xa.p, m->h_xa.data(),
                                         sizeof(double)*(size_t)m->n_faces,
                                         cudaMemcpyHostToDevice, c.stream)))
         && csb::launch_native_to_msr(m->nnz, m->c2f.p, m->xa.p, m->x_val.p)
         && CSB_CUDA(cudaStreamSynchronize(c.stream));
  m->h_vals_valid = false;
  return ok ? 0 : -1;
}

int csb_matrix_set_mesh_association(csb_matrix_t *m, csb_lnum_t n_faces,
                                    const csb_lnum_t *c2f, const short *c2f_sgn,
                                    const csb_real_t *cell_cen, const csb_real_t *cell_vol,
                                    const csb_real_t *face_u_normal,
                                    const csb_real_t *face_surf)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  /* a matrix mapped from native arrays already knows the face of every entry */
  const bool own_c2f = (m->from_native && c2f == nullptr && c2f_sgn == nullptr);
  if (m->from_native && (!own_c2f || n_faces != m->n_faces)) {
    set_error("csb_matrix_set_mesh_association: matrix created from %d edges: pass that face "
              "count and NULL for c2f / c2f_sgn", m->n_faces);
    return -1;
  }
  if ((!own_c2f && (c2f == nullptr || c2f_sgn == nullptr)) || cell_cen == nullptr
      || cell_vol == nullptr || face_u_normal == nullptr || face_surf == nullptr) {
    set_error("csb_matrix_set_mesh_association: all arrays are required");
    return -1;
  }
  if (m->has_mesh && n_faces != m->n_faces)      /* per-face geometry no longer matches */
    m->has_face_geo = m->has_green_geo = m->has_diipb = m->has_face_cog = false;
  m->n_faces = n_faces;
  bool ok = (own_c2f || (m->c2f.upload(c2f, (size_t)m->nnz)
                         && m->c2f_sgn.upload(c2f_sgn, (size_t)m->nnz)))
         && m->cell_cen.upload(cell_cen, 3*(size_t)m->n_cols_ext)
         && m->cell_vol.upload(cell_vol, (size_t)m->n_cols_ext)
         && m->face_u_normal.upload(face_u_normal, 3*(size_t)n_faces)
         && m->face_surf.upload(face_surf, (size_t)n_faces)
         && CSB_CUDA(cudaStreamSynchronize(ctx().stream));
  m->has_mesh = ok;
  m->cocg_valid = false;
  m->cocg_it_valid = false;
  m->fperm_state = 0;
  return ok ? 0 : -1;
}

/* cs_diffusion_potential (src/alge/cs_convection_diffusion.cpp:8743-8990), no flux
   reconstruction: diverg (+)= divergence of the diffusive flux of pvar over the mesh the matrix
   is associated with.  Arrays may be host or device pointers; diverg on the host is copied in
   (init = 0) and out. */
int csb_diffusion_potential(const csb_matrix_t *m, int init, int inc, const csb_real_t *pvar,
                            const csb_real_t *cofaf, const csb_real_t *cofbf,
                            const csb_real_t *i_visc, const csb_real_t *b_visc,
                            csb_real_t *diverg)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!m->has_mesh || m->db != 1 || m->halo != nullptr) {
    set_error("csb_diffusion_potential: needs a scalar matrix with its mesh association "
              "(csb_matrix_set_mesh_association) on one rank");
    return -1;
  }
  if (init < 0) { set_error("csb_diffusion_potential: invalid value of init"); return -1; }
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  if (pvar == nullptr || diverg == nullptr || (m->n_faces > 0 && i_visc == nullptr)
      || (bnd && (cofaf == nullptr || cofbf == nullptr || b_visc == nullptr))) {
    set_error("csb_diffusion_potential: null array");
    return -1;
  }
  Context &c = ctx();
  const size_t nr = (size_t)m->n_rows, nf = (size_t)m->n_faces, nb = bnd ? (size_t)m->n_b_faces : 0;
  const double *src[6] = {pvar, i_visc, cofaf, cofbf, b_visc, diverg};
  const size_t len[6] = {nr, nf, nb, nb, nb, nr};
  const double *dev[6];
  csb::DevBuf<double> stage;
  size_t need = 0;
  for (int i = 0; i < 6; i++)
    if (len[i] > 0 && !is_device_ptr(src[i])) need += len[i];
  if (need > 0 && !stage.alloc(need)) return -1;
  size_t off = 0;
  for (int i = 0; i < 6; i++) {
    dev[i] = src[i];
    if (len[i] > 0 && !is_device_ptr(src[i])) {
      if ((i < 5 || init == 0)
          && !CSB_CUDA(cudaMemcpyAsync(stage.p + off, src[i], sizeof(double)*len[i],
                                       cudaMemcpyHostToDevice, c.stream))) return -1;
      dev[i] = stage.p + off;
      off += len[i];
    }
  }
  double *d_div = const_cast<double *>(dev[5]);
  if (!csb::launch_diffusion_potential(m, init, inc, dev[0], dev[2], dev[3], dev[1], dev[4], d_div))
    return -1;
  if (d_div != diverg
      && !CSB_CUDA(cudaMemcpyAsync(diverg, d_div, sizeof(double)*nr, cudaMemcpyDeviceToHost,
                                   c.stream))) return -1;
  return CSB_CUDA(cudaStreamSynchronize(c.stream)) ? 0 : -1;
}

namespace {
/* device views of arrays that may live on the host: host ones are copied into `stage` */
bool stage_arrays(int n, const double *const *src, const size_t *len, const bool *copy_in,
                  const double **dev, csb::DevBuf<double> &stage)
{
  size_t need = 0;
  for (int i = 0; i < n; i++)
    if (len[i] > 0 && src[i] != nullptr && !is_device_ptr(src[i])) need += len[i];
  if (need > 0 && !stage.alloc(need)) return false;
  size_t off = 0;
  for (int i = 0; i < n; i++) {
    dev[i] = src[i];
    if (len[i] > 0 && src[i] != nullptr && !is_device_ptr(src[i])) {
      if (copy_in[i]
          && !CSB_CUDA(cudaMemcpyAsync(stage.p + off, src[i], sizeof(double)*len[i],
                                       cudaMemcpyHostToDevice, ctx().stream))) return false;
      dev[i] = stage.p + off;
      off += len[i];
    }
  }
  return true;
}

bool ensure_cocg(const csb_matrix_t *m)
{
  if (m->cocg_valid) return true;
  if (!m->cocg_lsq.alloc(6*(size_t)std::max(m->n_rows, 1))) return false;
  if (!csb::launch_lsq_cocg(m, m->cocg_lsq.p)) return false;
  m->cocg_valid = true;
  return true;
}
} // anonymous namespace

/* Face geometry of a non-orthogonal mesh, as cs_mesh_quantities_t holds it
   (src/mesh/cs_mesh_quantities.h:101-125): i_dist[n_faces], diipf / djjpf [n_faces][3] (II',
   JJ'), b_face_u_normal[n_b_faces][3], b_dist[n_b_faces].  II' of the boundary faces is zero
   until csb_matrix_set_boundary_iprime gives it.  Needs csb_matrix_set_mesh_association and
   csb_matrix_set_boundary_faces first. */
int csb_matrix_set_face_geometry(csb_matrix_t *m, const csb_real_t *i_dist,
                                 const csb_real_t *diipf, const csb_real_t *djjpf,
                                 const csb_real_t *b_face_u_normal, const csb_real_t *b_dist)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!m->has_mesh) {
    set_error("csb_matrix_set_face_geometry: csb_matrix_set_mesh_association first");
    return -1;
  }
  const size_t nf = (size_t)m->n_faces, nb = m->has_b_faces ? (size_t)m->n_b_faces : 0;
  if ((nf > 0 && (i_dist == nullptr || diipf == nullptr || djjpf == nullptr))
      || (nb > 0 && (b_face_u_normal == nullptr || b_dist == nullptr))) {
    set_error("csb_matrix_set_face_geometry: null array");
    return -1;
  }
  const bool ok = m->i_dist.upload(i_dist, nf) && m->diipf.upload(diipf, 3*nf)
               && m->djjpf.upload(djjpf, 3*nf) && m->b_face_u_normal.upload(b_face_u_normal, 3*nb)
               && m->b_dist.upload(b_dist, nb) && CSB_CUDA(cudaStreamSynchronize(ctx().stream));
  m->has_face_geo = ok;
  m->has_diipb = false;          /* II' of the boundary faces is given after the geometry */
  m->has_green_geo = false;      /* and so are the Green-Gauss quantities */
  m->imrgra = 1;
  m->cocg_valid = false;
  return ok ? 0 : -1;
}

/* II' of the boundary faces (cs_mesh_quantities_t::diipb, [n_b_faces][3]); NULL takes it back
   to zero.  With it the boundary face values of the gradient and of the reconstructed flux use
   the value at I' (cs_gradient_boundary_iprime_lsq_s, cs_gradient_boundary.cpp:197-640, through
   cs_boundary_conditions_update_bc_coeff_face_values,
   src/base/cs_boundary_conditions_set_coeffs.cpp:125-195), and the boundary directions of the
   least-squares system are n + II'/b_dist (cs_gradient.cpp:2632-2636, :2880-2885). */
int csb_matrix_set_boundary_iprime(csb_matrix_t *m, const csb_real_t *diipb)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!m->has_face_geo) {
    set_error("csb_matrix_set_boundary_iprime: csb_matrix_set_face_geometry first");
    return -1;
  }
  m->cocg_valid = false;
  const size_t nb = m->has_b_faces ? (size_t)m->n_b_faces : 0;
  if (diipb == nullptr || nb == 0) { m->has_diipb = false; return 0; }
  m->has_diipb = m->diipb.upload(diipb, 3*nb) && CSB_CUDA(cudaStreamSynchronize(ctx().stream));
  return m->has_diipb ? 0 : -1;
}

/* Quantities of the Green-Gauss reconstruction (_reconstruct_scalar_gradient,
   cs_gradient.cpp:4167-4540), as cs_mesh_quantities_t holds them: weight[n_faces] (interior
   face weighting factor), dofij[n_faces][3] (OF vectors), b_face_surf[n_b_faces] and,
   optionally, corr_grad_lin[n_rows][3][3] (the warped-face correction the reference applies
   when CS_BAD_CELLS_WARPED_CORRECTION is set, its default; NULL: no correction).  Interior
   normals, surfaces and cell volumes are those of the mesh association. */
int csb_matrix_set_green_geometry(csb_matrix_t *m, const csb_real_t *weight,
                                  const csb_real_t *dofij, const csb_real_t *b_face_surf,
                                  const csb_real_t *corr_grad_lin)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!m->has_face_geo) {
    set_error("csb_matrix_set_green_geometry: csb_matrix_set_face_geometry first");
    return -1;
  }
  const size_t nf = (size_t)m->n_faces, nb = m->has_b_faces ? (size_t)m->n_b_faces : 0;
  if ((nf > 0 && (weight == nullptr || dofij == nullptr)) || (nb > 0 && b_face_surf == nullptr)) {
    set_error("csb_matrix_set_green_geometry: null array");
    return -1;
  }
  m->cocg_it_valid = false;
  m->has_green_geo = m->g_weight.upload(weight, nf) && m->dofij.upload(dofij, 3*nf)
                  && m->b_face_surf.upload(b_face_surf, nb);
  m->has_corr_grad_lin = false;
  if (m->has_green_geo && corr_grad_lin != nullptr) {
    m->has_corr_grad_lin = m->corr_grad_lin.upload(corr_grad_lin, 9*(size_t)m->n_rows);
    if (!m->has_corr_grad_lin) m->has_green_geo = false;
  }
  if (!CSB_CUDA(cudaStreamSynchronize(ctx().stream))) m->has_green_geo = false;
  return m->has_green_geo ? 0 : -1;
}

/* Gradient csb_diffusion_potential_rc reconstructs with: eqp->imrgra, whose default is
   cs_glob_space_disc->imrgra = 4 in the reference (cs_parameters.cpp:470);
   cs_gradient_type_by_imrgra (cs_gradient.cpp:10317): 1 = CS_GRADIENT_LSQ (the initial value
   here), 4 = CS_GRADIENT_GREEN_LSQ, 0 = CS_GRADIENT_GREEN_ITER; the other types are not built. */
int csb_matrix_set_gradient_type(csb_matrix_t *m, int imrgra)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (imrgra != 0 && imrgra != 1 && imrgra != 4) {
    set_error("csb_matrix_set_gradient_type: imrgra = %d is not built (0: iterative "
              "Green-Gauss, 1: least squares, 4: Green-Gauss with least-squares face values)",
              imrgra);
    return -1;
  }
  if (imrgra != 1 && !m->has_green_geo) {
    set_error("csb_matrix_set_gradient_type: imrgra = %d needs csb_matrix_set_green_geometry",
              imrgra);
    return -1;
  }
  m->imrgra = imrgra;
  return 0;
}

/* eqp->nswrgr and eqp->epsrgr of the iterative gradient (imrgra = 0): at most nswrgr - 1
   reconstruction sweeps, stopped when the increment falls under epsrgr times the norm of the
   initial gradient (defaults 100, 1e-4: cs_parameters.cpp:380, :394) */
int csb_matrix_set_gradient_sweeps(csb_matrix_t *m, int nswrgr, double epsrgr)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  m->nswrgr = nswrgr;
  m->epsrgr = epsrgr;
  return 0;
}

/* Gradient limiter: eqp->imligr (-1 none -- the default --, 0 CS_GRADIENT_LIMIT_CELL, 1 _FACE,
   2 _CELL_RC; _scalar_gradient_clipping, cs_gradient.cpp:808-1215, standard neighbourhood),
   eqp->climgr (its factor; < 0: none) and eqp->b_rc_clip_factor (limiter of the values at I'
   of the boundary faces, cs_gradient_boundary.cpp:612-628, active when imligr >= 0; the
   reference's default is 1) */
int csb_matrix_set_gradient_limiter(csb_matrix_t *m, int imligr, double climgr,
                                    double b_rc_clip_factor)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (imligr < -1 || imligr > 2) {
    set_error("csb_matrix_set_gradient_limiter: imligr = %d (-1: none, 0: cell, 1: face, "
              "2: cell, values at I' / J')", imligr);
    return -1;
  }
  m->imligr = imligr;
  m->climgr = climgr;
  m->b_rc_clip_factor = b_rc_clip_factor;
  return 0;
}

/* sweeps the last iterative gradient on this matrix took (the reference's n_sweeps) */
int csb_matrix_get_gradient_sweeps(const csb_matrix_t *m)
{
  return (m != nullptr) ? m->n_grad_sweeps : -1;
}

namespace {
/* work arrays of one gradient: values at I' and face values of the boundary faces (when the
   matrix holds II' of its boundary faces or the hydrostatic treatment is on), the
   least-squares gradient the Green-Gauss pass reconstructs from, the staged exterior force */
struct GradWork {
  csb::DevBuf<double> val_ip, val_f, work, fext, clip;
  const double *d_ip = nullptr;     /* for the boundary fluxes; nullptr: the cell value */
};

/* hyd_p_flag as the reference reads it: 1 = on, 0 and 2 = off (cs_gradient.cpp:9678-9679) */
bool hyd_check(const csb_matrix_t *m, const char *who, int hyd_p_flag, const double *f_ext)
{
  if (hyd_p_flag < 0 || hyd_p_flag > 2) {
    set_error("%s: invalid hydrostatic pressure flag %d", who, hyd_p_flag);
    return false;
  }
  if (hyd_p_flag == 1 && (f_ext == nullptr || !m->has_face_cog)) {
    set_error("%s: the hydrostatic pressure treatment needs the exterior force and "
              "csb_matrix_set_face_cog", who);
    return false;
  }
  return true;
}

/* sum of the squares of x on the device, read back */
bool sum_sq_to_host(size_t n, const double *x, double *r)
{
  Context &c = ctx();
  if (!csb::launch_dot((int)n, x, nullptr, nullptr, 1)
      || !CSB_CUDA(cudaMemcpyAsync(c.red_out_h, c.red_out, sizeof(double), cudaMemcpyDeviceToHost,
                                   c.stream))
      || !CSB_CUDA(cudaStreamSynchronize(c.stream))) return false;
  *r = c.red_out_h[0];
  return true;
}

bool gradient_compute(const csb_matrix_t *m, int imrgra, int inc, const double *d_fext,
                      const double *d_vf, const double *pvar, const double *coefa,
                      const double *coefb, double *out, GradWork &w);

/* boundary values, then the least-squares gradient into `out`, or -- imrgra 4 -- into a work
   array followed by the Green-Gauss reconstruction into `out`; f_ext (host or device,
   [n_rows][3]) non-null = hyd_p_flag 1 */
bool gradient_run(const csb_matrix_t *m, int imrgra, int inc, const double *f_ext,
                  const double *pvar, const double *coefa, const double *coefb, double *out,
                  GradWork &w)
{
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  const size_t nr = (size_t)m->n_rows, nb = bnd ? (size_t)m->n_b_faces : 0;
  const double *d_fext = f_ext;
  if (f_ext != nullptr && !is_device_ptr(f_ext)) {
    if (!w.fext.alloc(3*std::max(nr, (size_t)1))
        || !CSB_CUDA(cudaMemcpyAsync(w.fext.p, f_ext, sizeof(double)*3*nr, cudaMemcpyHostToDevice,
                                     ctx().stream))) return false;
    d_fext = w.fext.p;
  }
  const double *d_vf = nullptr;
  w.d_ip = nullptr;
  if (bnd && (m->has_diipb || d_fext != nullptr)) {
    if (!w.val_ip.alloc(nb) || !w.val_f.alloc(nb)
        || !csb::launch_b_iprime_lsq(m, inc, pvar, coefa, coefb, d_fext,
                                     (m->imligr < 0) ? -1. : m->b_rc_clip_factor, w.val_ip.p,
                                     w.val_f.p))
      return false;
    d_vf = w.val_f.p;
    if (m->has_diipb) w.d_ip = w.val_ip.p;
  }
  if (!gradient_compute(m, imrgra, inc, d_fext, d_vf, pvar, coefa, coefb, out, w)) return false;
  /* limiter of the cell gradient, whatever its type (cs_gradient.cpp:7812-7822) */
  if (m->imligr >= 0 && m->climgr >= 0 && nr > 0)
    return w.clip.alloc(2*nr)
        && csb::launch_gradient_clip(m, m->imligr, m->climgr, inc, pvar, coefa, coefb, d_vf, out,
                                     w.clip.p);
  return true;
}

bool gradient_compute(const csb_matrix_t *m, int imrgra, int inc, const double *d_fext,
                      const double *d_vf, const double *pvar, const double *coefa,
                      const double *coefb, double *out, GradWork &w)
{
  const size_t nr = (size_t)m->n_rows;
  if (imrgra == 0) {
    /* iterative Green-Gauss (_initialize_scalar_gradient + _iterative_scalar_gradient,
       cs_gradient.cpp:1329, :1891): one reduction comes back to the host per sweep, like the
       reference's global sum, for the convergence test */
    if (!m->cocg_it_valid) {
      if (!m->cocg_it.alloc(9*std::max(nr, (size_t)1)) || !csb::launch_cocg_it(m, m->cocg_it.p))
        return false;
      m->cocg_it_valid = true;
    }
    m->n_grad_sweeps = 0;
    if (!csb::launch_green_iter_sweep(m, inc, pvar, coefa, coefb, nullptr, out)) return false;
    if (m->nswrgr < 1 || nr < 1) return true;
    double ss = 0.;
    if (!sum_sq_to_host(3*nr, out, &ss)) return false;
    const double rnorm = std::sqrt(ss);
    if (rnorm <= 1.e-12) return true;                                  /* cs_math_epzero */
    if (!w.work.alloc(3*nr)) return false;
    int n_sweeps;
    for (n_sweeps = 1; n_sweeps < m->nswrgr; n_sweeps++) {
      if (!csb::launch_green_iter_sweep(m, inc, pvar, coefa, coefb, out, w.work.p)
          || !sum_sq_to_host(3*nr, w.work.p, &ss)
          || !csb::launch_green_iter_update(m, m->cocg_it.p, w.work.p, out)) return false;
      if (std::sqrt(ss) < m->epsrgr*rnorm) break;
    }
    m->n_grad_sweeps = n_sweeps;
    return true;
  }
  if (imrgra != 4)
    return csb::launch_lsq_gradient(m, m->cocg_lsq.p, inc, pvar, coefa, coefb, d_vf, d_fext, out);
  return w.work.alloc(3*std::max(nr, (size_t)1))
      && csb::launch_lsq_gradient(m, m->cocg_lsq.p, inc, pvar, coefa, coefb, d_vf, d_fext,
                                  w.work.p)
      && csb::launch_green_reconstruct(m, inc, pvar, coefa, coefb, d_vf, d_fext, w.work.p, out);
}
} // anonymous namespace

/* Face centres (cs_mesh_quantities_t::i_face_cog [n_faces][3], b_face_cog [n_b_faces][3]): read
   by the hydrostatic pressure treatment only (hyd_p_flag = 1). */
int csb_matrix_set_face_cog(csb_matrix_t *m, const csb_real_t *i_face_cog,
                            const csb_real_t *b_face_cog)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!m->has_mesh) {
    set_error("csb_matrix_set_face_cog: csb_matrix_set_mesh_association first");
    return -1;
  }
  const size_t nf = (size_t)m->n_faces, nb = m->has_b_faces ? (size_t)m->n_b_faces : 0;
  if ((nf > 0 && i_face_cog == nullptr) || (nb > 0 && b_face_cog == nullptr)) {
    set_error("csb_matrix_set_face_cog: null array");
    return -1;
  }
  m->has_face_cog = m->i_face_cog.upload(i_face_cog, 3*nf) && m->b_face_cog.upload(b_face_cog, 3*nb)
                 && CSB_CUDA(cudaStreamSynchronize(ctx().stream));
  return m->has_face_cog ? 0 : -1;
}

/* cs_gradient_scalar_synced_input with CS_GRADIENT_LSQ, standard neighbourhood, no weighting,
   no limiter (src/alge/cs_gradient.cpp:9651, _lsq_scalar_gradient :2695-2900): grad[n_rows][3].
   Boundary face values are inc*coefa + coefb*(value at I'), the cell value when II' = 0 at the
   boundary.  The cocg
   matrices are computed on first use and kept with the matrix. */
int csb_gradient_scalar_lsq(const csb_matrix_t *m, int inc, const csb_real_t *pvar,
                            const csb_real_t *coefa, const csb_real_t *coefb, csb_real_t *grad)
{
  return csb_gradient_scalar(m, 1, inc, pvar, coefa, coefb, grad);
}

/* cs_gradient_scalar_synced_input for the gradient type of imrgra (1: CS_GRADIENT_LSQ,
   4: CS_GRADIENT_GREEN_LSQ = the least-squares gradient, then _reconstruct_scalar_gradient,
   cs_gradient.cpp:7697-7797) */
int csb_gradient_scalar(const csb_matrix_t *m, int imrgra, int inc, const csb_real_t *pvar,
                        const csb_real_t *coefa, const csb_real_t *coefb, csb_real_t *grad)
{
  return csb_gradient_scalar_hyd(m, imrgra, inc, 0, nullptr, pvar, coefa, coefb, grad);
}

/* the same with the hydrostatic pressure treatment of cs_gradient_scalar_synced_input
   (hyd_p_flag = 1, exterior force f_ext [n_rows][3]: _lsq_scalar_gradient_hyd_p
   cs_gradient.cpp:3230-3560, the hydrostatic branch of _reconstruct_scalar_gradient
   :4314-4424) and of the boundary face values (cs_gradient_boundary.cpp:285-640 with
   hyd_p_flag, cs_boundary_conditions_set_coeffs.cpp:176-192) */
int csb_gradient_scalar_hyd(const csb_matrix_t *m, int imrgra, int inc, int hyd_p_flag,
                            const csb_real_t *f_ext, const csb_real_t *pvar,
                            const csb_real_t *coefa, const csb_real_t *coefb, csb_real_t *grad)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!hyd_check(m, "csb_gradient_scalar", hyd_p_flag, f_ext)) return -1;
  if (!m->has_face_geo || m->db != 1 || m->halo != nullptr) {
    set_error("csb_gradient_scalar: needs a scalar matrix on one rank with its mesh "
              "association, boundary faces and csb_matrix_set_face_geometry");
    return -1;
  }
  if (imrgra != 0 && imrgra != 1 && imrgra != 4) {
    set_error("csb_gradient_scalar: imrgra = %d is not built (0: iterative Green-Gauss, 1: least "
              "squares, 4: Green-Gauss with least-squares face values)", imrgra);
    return -1;
  }
  if (imrgra != 1 && !m->has_green_geo) {
    set_error("csb_gradient_scalar: imrgra = %d needs csb_matrix_set_green_geometry", imrgra);
    return -1;
  }
  if (imrgra == 0 && hyd_p_flag == 1) {
    set_error("csb_gradient_scalar: the hydrostatic pressure treatment of the iterative "
              "gradient (imrgra = 0) is not built");
    return -1;
  }
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  if (pvar == nullptr || grad == nullptr || (bnd && (coefa == nullptr || coefb == nullptr))) {
    set_error("csb_gradient_scalar: null array");
    return -1;
  }
  Context &c = ctx();
  const size_t nr = (size_t)m->n_rows, nb = bnd ? (size_t)m->n_b_faces : 0;
  const double *src[4] = {pvar, coefa, coefb, grad};
  const size_t len[4] = {nr, nb, nb, 3*nr};
  const bool in[4] = {true, true, true, false};
  const double *dev[4];
  csb::DevBuf<double> stage;
  GradWork gw;
  if (!ensure_cocg(m) || !stage_arrays(4, src, len, in, dev, stage)) return -1;
  double *d_grad = const_cast<double *>(dev[3]);
  if (!gradient_run(m, imrgra, inc, (hyd_p_flag == 1) ? f_ext : nullptr, dev[0], dev[1], dev[2],
                    d_grad, gw)) return -1;
  if (d_grad != grad
      && !CSB_CUDA(cudaMemcpyAsync(grad, d_grad, sizeof(double)*3*nr, cudaMemcpyDeviceToHost,
                                   c.stream))) return -1;
  return CSB_CUDA(cudaStreamSynchronize(c.stream)) ? 0 : -1;
}

/* cs_diffusion_potential WITH flux reconstruction (eqp->ircflu = 1, imrgra = 1;
   src/alge/cs_convection_diffusion.cpp:8895-8982): the least-squares gradient of pvar, then the
   face fluxes with the non-orthogonality term -- irecmf = 0: 0.5 (visel grad)_I+J . (II' - JJ')
   S/d, irecmf = 1: i_visc (grad_I . II' - grad_J . JJ'). */
int csb_diffusion_potential_rc(const csb_matrix_t *m, int irecmf, int init, int inc,
                               const csb_real_t *pvar, const csb_real_t *coefa,
                               const csb_real_t *coefb, const csb_real_t *cofaf,
                               const csb_real_t *cofbf, const csb_real_t *i_visc,
                               const csb_real_t *b_visc, const csb_real_t *visel,
                               csb_real_t *diverg)
{
  return csb_diffusion_potential_hyd(m, irecmf, init, inc, 0, nullptr, pvar, coefa, coefb, cofaf,
                                     cofbf, i_visc, b_visc, visel, diverg);
}

/* the same with iphydp and frcxt [n_rows][3] in the reference's positions
   (cs_diffusion_potential's argument list): iphydp = 1 is what cs_pressure_correction hands
   over by default (iphydr = 1, src/base/cs_pressure_correction.cpp:2113-2127) */
int csb_diffusion_potential_hyd(const csb_matrix_t *m, int irecmf, int init, int inc,
                                int iphydp, const csb_real_t *frcxt, const csb_real_t *pvar,
                                const csb_real_t *coefa, const csb_real_t *coefb,
                                const csb_real_t *cofaf, const csb_real_t *cofbf,
                                const csb_real_t *i_visc, const csb_real_t *b_visc,
                                const csb_real_t *visel, csb_real_t *diverg)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!hyd_check(m, "csb_diffusion_potential", iphydp, frcxt)) return -1;
  if (m->imrgra == 0 && iphydp == 1) {
    set_error("csb_diffusion_potential: the hydrostatic pressure treatment of the iterative "
              "gradient (imrgra = 0) is not built");
    return -1;
  }
  if (!m->has_face_geo || m->db != 1 || m->halo != nullptr) {
    set_error("csb_diffusion_potential_rc: needs a scalar matrix on one rank with its mesh "
              "association, boundary faces and csb_matrix_set_face_geometry");
    return -1;
  }
  if (init < 0) { set_error("csb_diffusion_potential_rc: invalid value of init"); return -1; }
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  if (pvar == nullptr || diverg == nullptr || visel == nullptr
      || (m->n_faces > 0 && i_visc == nullptr)
      || (bnd && (coefa == nullptr || coefb == nullptr || cofaf == nullptr || cofbf == nullptr
                  || b_visc == nullptr))) {
    set_error("csb_diffusion_potential_rc: null array");
    return -1;
  }
  Context &c = ctx();
  const size_t nr = (size_t)m->n_rows, nf = (size_t)m->n_faces, nb = bnd ? (size_t)m->n_b_faces : 0;
  const double *src[9] = {pvar, coefa, coefb, cofaf, cofbf, i_visc, b_visc, visel, diverg};
  const size_t len[9] = {nr, nb, nb, nb, nb, nf, nb, nr, nr};
  const bool in[9] = {true, true, true, true, true, true, true, true, init == 0};
  const double *dev[9];
  csb::DevBuf<double> stage, grad;
  GradWork gw;
  if (!ensure_cocg(m) || !stage_arrays(9, src, len, in, dev, stage) || !grad.alloc(3*nr)) return -1;
  double *d_div = const_cast<double *>(dev[8]);
  if (!gradient_run(m, m->imrgra, inc, (iphydp == 1) ? frcxt : nullptr, dev[0], dev[1], dev[2],
                    grad.p, gw)
      || !csb::launch_diffusion_potential_rc(m, irecmf, init, inc, dev[0], grad.p, dev[3], dev[4],
                                             dev[5], dev[6], dev[7], gw.d_ip, d_div)) return -1;
  if (d_div != diverg
      && !CSB_CUDA(cudaMemcpyAsync(diverg, d_div, sizeof(double)*nr, cudaMemcpyDeviceToHost,
                                   c.stream))) return -1;
  return CSB_CUDA(cudaStreamSynchronize(c.stream)) ? 0 : -1;
}

/* cs_divergence (src/alge/cs_divergence.cpp:1224-1290): init >= 1 zeroes diverg, then every
   cell adds the mass fluxes of its interior faces (signed by the face's orientation) in
   ascending face order and those of its boundary faces: bit-identical to the reference's
   one-thread result.  Host or device pointers. */
int csb_divergence(const csb_matrix_t *m, int init, const csb_real_t *i_massflux,
                   const csb_real_t *b_massflux, csb_real_t *diverg)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (!m->has_mesh || m->db != 1 || m->halo != nullptr) {
    set_error("csb_divergence: needs a scalar matrix on one rank with its mesh association");
    return -1;
  }
  if (init < 0) { set_error("csb_divergence: invalid value of init"); return -1; }
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  if (diverg == nullptr || (m->n_faces > 0 && i_massflux == nullptr)
      || (bnd && b_massflux == nullptr)) {
    set_error("csb_divergence: null array");
    return -1;
  }
  Context &c = ctx();
  const size_t nr = (size_t)m->n_rows, nf = (size_t)m->n_faces, nb = bnd ? (size_t)m->n_b_faces : 0;
  const double *src[3] = {i_massflux, b_massflux, diverg};
  const size_t len[3] = {nf, nb, nr};
  const bool in[3] = {true, true, init == 0};
  const double *dev[3];
  csb::DevBuf<double> stage;
  if (!stage_arrays(3, src, len, in, dev, stage)) return -1;
  double *d_div = const_cast<double *>(dev[2]);
  if (!csb::launch_divergence(m, init, dev[0], dev[1], d_div)) return -1;
  if (d_div != diverg
      && !CSB_CUDA(cudaMemcpyAsync(diverg, d_div, sizeof(double)*nr, cudaMemcpyDeviceToHost,
                                   c.stream))) return -1;
  return CSB_CUDA(cudaStreamSynchronize(c.stream)) ? 0 : -1;
}

/* cs_face_diffusion_potential (src/alge/cs_convection_diffusion.cpp:8075-8310): the face mass
   fluxes of the diffusion of pvar -- what the pressure correction adds to the mass flux after
   the solve.  ircflu = 0: i_massflux[f] += i_visc (pvar_I - pvar_J), b_massflux[f] += b_visc
   (inc cofaf + cofbf pvar_I); coefa, coefb, visel are not read.  ircflu = 1: the reconstructed
   form with the gradient of csb_matrix_set_gradient_type and the value at I' of the boundary
   faces.  init >= 1 starts from zero.  Bit-identical to the reference. */
int csb_face_diffusion_potential(const csb_matrix_t *m, int ircflu, int init, int inc,
                                 const csb_real_t *pvar, const csb_real_t *coefa,
                                 const csb_real_t *coefb, const csb_real_t *cofaf,
                                 const csb_real_t *cofbf, const csb_real_t *i_visc,
                                 const csb_real_t *b_visc, const csb_real_t *visel,
                                 csb_real_t *i_massflux, csb_real_t *b_massflux)
{
  return csb_face_diffusion_potential_hyd(m, ircflu, init, inc, 0, nullptr, pvar, coefa, coefb,
                                          cofaf, cofbf, i_visc, b_visc, visel, i_massflux,
                                          b_massflux);
}

/* the same with iphydp and frcxt in the reference's positions; they only matter when
   ircflu = 1 (gradient and values at I') */
int csb_face_diffusion_potential_hyd(const csb_matrix_t *m, int ircflu, int init, int inc,
                                     int iphydp, const csb_real_t *frcxt,
                                     const csb_real_t *pvar, const csb_real_t *coefa,
                                     const csb_real_t *coefb, const csb_real_t *cofaf,
                                     const csb_real_t *cofbf, const csb_real_t *i_visc,
                                     const csb_real_t *b_visc, const csb_real_t *visel,
                                     csb_real_t *i_massflux, csb_real_t *b_massflux)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (ircflu && !hyd_check(m, "csb_face_diffusion_potential", iphydp, frcxt)) return -1;
  if (ircflu && m->imrgra == 0 && iphydp == 1) {
    set_error("csb_face_diffusion_potential: the hydrostatic pressure treatment of the "
              "iterative gradient (imrgra = 0) is not built");
    return -1;
  }
  if (!m->has_mesh || m->db != 1 || m->halo != nullptr || (ircflu && !m->has_face_geo)) {
    set_error("csb_face_diffusion_potential: needs a scalar matrix on one rank with its mesh "
              "association%s", ircflu ? ", boundary faces and csb_matrix_set_face_geometry" : "");
    return -1;
  }
  if (init != 0 && init < 1) {
    set_error("csb_face_diffusion_potential: invalid value of init");
    return -1;
  }
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  if (pvar == nullptr || (m->n_faces > 0 && (i_visc == nullptr || i_massflux == nullptr))
      || (bnd && (cofaf == nullptr || cofbf == nullptr || b_visc == nullptr
                  || b_massflux == nullptr))
      || (ircflu && (visel == nullptr || (bnd && (coefa == nullptr || coefb == nullptr))))) {
    set_error("csb_face_diffusion_potential: null array");
    return -1;
  }
  Context &c = ctx();
  const size_t nr = (size_t)m->n_rows, nf = (size_t)m->n_faces, nb = bnd ? (size_t)m->n_b_faces : 0;
  const double *src[10] = {pvar, ircflu ? coefa : nullptr, ircflu ? coefb : nullptr, cofaf, cofbf,
                           i_visc, b_visc, ircflu ? visel : nullptr, i_massflux, b_massflux};
  const size_t len[10] = {nr, nb, nb, nb, nb, nf, nb, nr, nf, nb};
  const bool in[10] = {true, true, true, true, true, true, true, true, init == 0, init == 0};
  const double *dev[10];
  csb::DevBuf<double> stage, grad;
  GradWork gw;
  if (!stage_arrays(10, src, len, in, dev, stage)) return -1;
  const double *d_grad = nullptr, *d_ip = nullptr;
  if (ircflu) {
    if (!ensure_cocg(m) || !grad.alloc(3*nr)) return -1;
    if (!gradient_run(m, m->imrgra, inc, (iphydp == 1) ? frcxt : nullptr, dev[0], dev[1], dev[2],
                      grad.p, gw)) return -1;
    d_ip = gw.d_ip;
    d_grad = grad.p;
  }
  double *d_im = const_cast<double *>(dev[8]), *d_bm = const_cast<double *>(dev[9]);
  if (!csb::launch_face_diffusion_potential(m, init, inc, dev[0], d_grad, dev[3], dev[4], dev[5],
                                            dev[6], dev[7], d_ip, d_im, d_bm)) return -1;
  if (d_im != i_massflux && nf > 0
      && !CSB_CUDA(cudaMemcpyAsync(i_massflux, d_im, sizeof(double)*nf, cudaMemcpyDeviceToHost,
                                   c.stream))) return -1;
  if (d_bm != b_massflux && nb > 0
      && !CSB_CUDA(cudaMemcpyAsync(b_massflux, d_bm, sizeof(double)*nb, cudaMemcpyDeviceToHost,
                                   c.stream))) return -1;
  return CSB_CUDA(cudaStreamSynchronize(c.stream)) ? 0 : -1;
}

/* Boundary faces of the mesh by adjacent cell (cs_mesh_t::b_face_cells); builds the
   cell -> boundary faces adjacency the reference keeps in cs_mesh_adjacencies
   (cell_b_faces_idx / cell_b_faces, src/mesh/cs_mesh_adjacencies.cpp:705-760). */
int csb_matrix_set_boundary_faces(csb_matrix_t *m, csb_lnum_t n_b_faces,
                                  const csb_lnum_t *b_face_cells)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (n_b_faces < 0 || (n_b_faces > 0 && b_face_cells == nullptr)) {
    set_error("csb_matrix_set_boundary_faces: bad arguments");
    return -1;
  }
  std::vector<csb_lnum_t> bfc((size_t)n_b_faces);
  if (n_b_faces > 0
      && !CSB_CUDA(cudaMemcpy(bfc.data(), b_face_cells, sizeof(csb_lnum_t)*(size_t)n_b_faces,
                              cudaMemcpyDefault))) return -1;
  std::vector<csb_lnum_t> idx((size_t)m->n_rows + 1, 0), ids((size_t)std::max(n_b_faces, 1));
  for (csb_lnum_t f = 0; f < n_b_faces; f++) {
    if (bfc[f] < 0 || bfc[f] >= m->n_rows) {
      set_error("csb_matrix_set_boundary_faces: face %d is adjacent to cell %d", f, bfc[f]);
      return -1;
    }
    idx[(size_t)bfc[f] + 1]++;
  }
  for (csb_lnum_t i = 0; i < m->n_rows; i++) idx[(size_t)i + 1] += idx[i];
  {
    std::vector<csb_lnum_t> fill(idx.begin(), idx.end() - 1);
    for (csb_lnum_t f = 0; f < n_b_faces; f++) ids[(size_t)fill[bfc[f]]++] = f;   /* ascending */
  }
  if (n_b_faces != m->n_b_faces || !m->has_b_faces) {
    /* per-boundary-face geometry given earlier no longer matches: to be given again */
    m->has_face_geo = m->has_green_geo = m->has_diipb = m->has_face_cog = false;
    m->cocg_valid = false;
  }
  m->n_b_faces = n_b_faces;
  bool ok = m->cell_b_faces_idx.upload(idx.data(), idx.size())
         && m->cell_b_faces.upload(ids.data(), (size_t)n_b_faces)
         && CSB_CUDA(cudaStreamSynchronize(ctx().stream));
  m->has_b_faces = ok;
  return ok ? 0 : -1;
}

/* cs_matrix_compute_coeffs (src/alge/cs_matrix_building.cpp:1948-2283), scalar cases: the
   MSR coefficients are computed in HBM from per-face and per-cell arrays (host or device). */
int csb_matrix_compute_coeffs(csb_matrix_t *m, int iconvp, int idiffp, int ndircp,
                              double thetap, double relaxp, int imucpp,
                              const csb_real_t *coefbp, const csb_real_t *cofbfp,
                              const csb_real_t *rovsdt, const csb_real_t *i_massflux,
                              const csb_real_t *b_massflux, const csb_real_t *i_visc,
                              const csb_real_t *b_visc, const csb_real_t *xcpp)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (m->db != 1) {
    set_error("csb_matrix_compute_coeffs: scalar matrices only on this path");
    return -1;
  }
  if (m->csr != nullptr) {
    set_error("csb_matrix_compute_coeffs: matrix was created from CSR arrays (its CSR values "
              "would go stale); use an MSR matrix");
    return -1;
  }
  if (m->from_native) {
    set_error("csb_matrix_compute_coeffs: matrix was created from native arrays (its face "
              "values would go stale); use csb_matrix_set_coefficients_native");
    return -1;
  }
  if (!m->has_mesh) {
    set_error("csb_matrix_compute_coeffs: needs the mesh association "
              "(csb_matrix_set_mesh_association)");
    return -1;
  }
  const bool conv = (iconvp != 0);
  const bool diff = (idiffp != 0);
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  if (imucpp == 0) xcpp = nullptr;
  if (rovsdt == nullptr || ((diff || conv) && i_visc == nullptr)
      || (conv && i_massflux == nullptr) || (imucpp != 0 && conv && xcpp == nullptr)
      || (bnd && (conv || diff) && (b_visc == nullptr || cofbfp == nullptr))
      || (bnd && conv && (b_massflux == nullptr || coefbp == nullptr))) {
    set_error("csb_matrix_compute_coeffs: null coefficient array");
    return -1;
  }
  Context &c = ctx();
  /* host arrays are staged in one buffer that lives with the matrix */
  const size_t nf = (size_t)m->n_faces, nb = (size_t)m->n_b_faces, nr = (size_t)m->n_rows;
  const size_t nce = (size_t)m->n_cols_ext;
  const bool face_terms = conv || diff;
  const double *src[8] = {rovsdt, i_visc, b_visc, cofbfp, coefbp, i_massflux, b_massflux, xcpp};
  const size_t len[8] = {nr, face_terms ? nf : 0, face_terms ? nb : 0, face_terms ? nb : 0,
                         conv ? nb : 0, conv ? nf : 0, conv ? nb : 0, conv ? nce : 0};
  const double *dev[8];
  size_t need = 0;
  for (int i = 0; i < 8; i++)
    if (len[i] > 0 && src[i] != nullptr && !is_device_ptr(src[i])) need += len[i];
  if (need > m->coef_stage.n && !m->coef_stage.alloc(need)) return -1;
  size_t off = 0;
  for (int i = 0; i < 8; i++) {
    dev[i] = src[i];
    if (len[i] > 0 && src[i] != nullptr && !is_device_ptr(src[i])) {
      if (!CSB_CUDA(cudaMemcpyAsync(m->coef_stage.p + off, src[i], sizeof(double)*len[i],
                                    cudaMemcpyHostToDevice, c.stream))) return -1;
      dev[i] = m->coef_stage.p + off;
      off += len[i];
    }
  }
  const bool ok = conv
    ? launch_coeffs_msr(m, idiffp, ndircp, thetap, relaxp, dev[4], dev[3], dev[0], dev[5],
                        dev[6], dev[1], dev[2], dev[7], m->d_val.p, m->x_val.p)
    : launch_sym_coeffs_msr(m, idiffp, ndircp, thetap, relaxp, dev[3], dev[0], dev[1], dev[2],
                            m->d_val.p, m->x_val.p);
  if (!ok) return -1;
  if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return -1;   /* host inputs may go away */
  m->symmetric = !conv;
  m->h_vals_valid = false;
  m->b_d_val = nullptr; m->b_x_val = nullptr;   /* lent host values are stale now */
  return 0;
}

/* cs_matrix_compute_coeffs<3> (src/alge/cs_matrix_building.cpp:2284-2600) for a vector
   variable with scalar diffusion: 3x3 diagonal blocks, scalar extra-diagonal terms. */
int csb_matrix_compute_coeffs_vector(csb_matrix_t *m, int iconvp, int idiffp, int ndircp,
                                     double thetap, double relaxp, const csb_real_t *coefbp,
                                     const csb_real_t *cofbfp, const csb_real_t *fimp,
                                     const csb_real_t *i_massflux,
                                     const csb_real_t *b_massflux, const csb_real_t *i_visc,
                                     const csb_real_t *b_visc)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  if (m->db != 3) {
    set_error("csb_matrix_compute_coeffs_vector: needs a matrix with 3x3 diagonal blocks");
    return -1;
  }
  if (!m->has_mesh) {
    set_error("csb_matrix_compute_coeffs_vector: needs the mesh association "
              "(csb_matrix_set_mesh_association)");
    return -1;
  }
  const bool conv = (iconvp != 0), diff = (idiffp != 0);
  const bool bnd = (m->has_b_faces && m->n_b_faces > 0);
  if (((conv || diff) && i_visc == nullptr) || (conv && i_massflux == nullptr)
      || (bnd && (conv || diff) && (b_visc == nullptr || cofbfp == nullptr))
      || (bnd && conv && (b_massflux == nullptr || coefbp == nullptr))) {
    set_error("csb_matrix_compute_coeffs_vector: null coefficient array");
    return -1;
  }
  Context &c = ctx();
  const size_t nf = (size_t)m->n_faces, nb = (size_t)m->n_b_faces, nr = (size_t)m->n_rows;
  const bool face_terms = conv || diff;
  const double *src[7] = {fimp, i_visc, b_visc, cofbfp, coefbp, i_massflux, b_massflux};
  const size_t len[7] = {9*nr, face_terms ? nf : 0, face_terms ? nb : 0,
                         face_terms ? 9*nb : 0, conv ? 9*nb : 0, conv ? nf : 0, conv ? nb : 0};
  const double *dev[7];
  size_t need = 0;
  for (int i = 0; i < 7; i++)
    if (len[i] > 0 && src[i] != nullptr && !is_device_ptr(src[i])) need += len[i];
  if (need > m->coef_stage.n && !m->coef_stage.alloc(need)) return -1;
  size_t off = 0;
  for (int i = 0; i < 7; i++) {
    dev[i] = src[i];
    if (len[i] > 0 && src[i] != nullptr && !is_device_ptr(src[i])) {
      if (!CSB_CUDA(cudaMemcpyAsync(m->coef_stage.p + off, src[i], sizeof(double)*len[i],
                                    cudaMemcpyHostToDevice, c.stream))) return -1;
      dev[i] = m->coef_stage.p + off;
      off += len[i];
    }
  }
  if (!launch_coeffs_msr_v3(m, iconvp, idiffp, ndircp, thetap, relaxp, dev[4], dev[3], dev[0],
                            dev[5], dev[6], dev[1], dev[2], m->d_val.p, m->x_val.p)) return -1;
  if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return -1;
  m->symmetric = !conv;
  m->h_vals_valid = false;
  m->b_d_val = nullptr; m->b_x_val = nullptr;   /* lent host values are stale now */
  return 0;
}

/* The MSR coefficients as they are in HBM, copied to d_val[n_rows*db*db] / x_val[nnz]
   (host or device); cs_matrix_get_msr_arrays (cs_matrix.cpp:6199) lends them, here they are
   copied because they live on the device. */
int csb_matrix_get_coefficients(const csb_matrix_t *m, csb_real_t *d_val, csb_real_t *x_val)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  Context &c = ctx();
  bool ok = true;
  if (d_val != nullptr)
    ok = CSB_CUDA(cudaMemcpyAsync(d_val, m->d_val.p, sizeof(double)*(size_t)m->n_rows*m->db*m->db,
                                  cudaMemcpyDefault, c.stream));
  if (ok && x_val != nullptr && m->nnz > 0)
    ok = CSB_CUDA(cudaMemcpyAsync(x_val, m->x_val.p, sizeof(double)*(size_t)m->nnz,
                                  cudaMemcpyDefault, c.stream));
  return (ok && CSB_CUDA(cudaStreamSynchronize(c.stream))) ? 0 : -1;
}

int csb_matrix_set_halo(csb_matrix_t *m, csb_halo_t *halo)
{
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  m->halo = halo;
  return 0;
}

void csb_matrix_destroy(csb_matrix_t **m)
{
  if (m == nullptr || *m == nullptr) return;
  delete *m;
  *m = nullptr;
}

csb_lnum_t csb_matrix_get_n_rows(const csb_matrix_t *m) { return m ? m->n_rows : 0; }
csb_lnum_t csb_matrix_get_n_columns(const csb_matrix_t *m) { return m ? m->n_cols_ext : 0; }
csb_lnum_t csb_matrix_get_n_entries(const csb_matrix_t *m) { return m ? m->nnz + m->n_rows : 0; }

double csb_matrix_spmv_bytes(const csb_matrix_t *m)
{
  if (m == nullptr) return 0.;
  return 12.0*m->nnz + 4.0*(m->n_rows + 1.0) + 8.0*m->n_rows + 8.0*m->n_cols_ext
         + 8.0*m->n_rows;
}

static int _spmv_api(const csb_matrix_t *m, csb_real_t *x, csb_real_t *y, bool exdiag)
{
  Context &c = ctx();
  if (m == nullptr) { set_error("The matrix is not defined."); return -1; }
  const bool xd = is_device_ptr(x), yd = is_device_ptr(y);
  DevBuf<double> tx, ty;
  double *x_d = x, *y_d = y;
  if (!xd) {
    if (!tx.upload(x, (size_t)m->n_cols_ext*m->db)) return -1;
    x_d = tx.p;
  }
  if (!yd) {
    if (!ty.alloc((size_t)m->n_rows*m->db)) return -1;
    y_d = ty.p;
  }
  if (m->db > 1) {
    if (m->halo != nullptr) { set_error("block matrices: no halo on this path yet"); return -1; }
    if (m->db == 3) { if (!launch_spmv_b3(m, x_d, y_d, exdiag)) return -1; }
    else if (!launch_spmv_bn(-1, m, nullptr, x_d, y_d, nullptr, exdiag)) return -1;
    if (!yd && !CSB_CUDA(cudaMemcpyAsync(y, y_d, sizeof(double)*(size_t)m->n_rows*m->db,
                                         cudaMemcpyDeviceToHost, c.stream))) return -1;
    if ((!xd || !yd) && !CSB_CUDA(cudaStreamSynchronize(c.stream))) return -1;
    return 0;
  }
  if (m->halo != nullptr)
    if (csb_halo_sync_var(m->halo, x_d) != 0) return -1;
  if (m->csr != nullptr) {
    /* CSR order: every entry (diagonal included) in place; exclude_diag skips col == row */
    if (exdiag) { if (!launch_spmv_csr_exdiag(m->csr, x_d, y_d)) return -1; }
    else {
      SpmvArgs a;
      a.m = m->csr; a.exclude_diag = true; a.x = x_d; a.y = y_d;
      if (!launch_spmv(EPI_STORE, a)) return -1;
    }
  }
  else {
    SpmvArgs a;
    a.m = m; a.exclude_diag = exdiag; a.x = x_d; a.y = y_d;
    if (!launch_spmv(EPI_STORE, a)) return -1;
  }
  if (!yd) {
    if (!CSB_CUDA(cudaMemcpyAsync(y, y_d, sizeof(double)*(size_t)m->n_rows,
                                  cudaMemcpyDeviceToHost, c.stream))) return -1;
  }
  if (!xd || !yd)
    if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return -1;
  return 0;
}

int csb_matrix_vector_multiply(const csb_matrix_t *m, csb_real_t *x, csb_real_t *y)
{ return _spmv_api(m, x, y, false); }

int csb_matrix_vector_multiply_exdiag(const csb_matrix_t *m, csb_real_t *x, csb_real_t *y)
{ return _spmv_api(m, x, y, true); }

/*----------------------------------------------------------------------------
 * BLAS-1
 *----------------------------------------------------------------------------*/

static int _dot_api(csb_lnum_t n, const double *x, const double *y, const double *z,
                    int mode, double *out, int n_out)
{
  Context &c = ctx();
  DevBuf<double> tx, ty, tz;
  const double *xd = x, *yd = y, *zd = z;
  if (x && !is_device_ptr(x)) { if (!tx.upload(x, n)) return -1; xd = tx.p; }
  if (y && !is_device_ptr(y)) { if (!ty.upload(y, n)) return -1; yd = ty.p; }
  if (z && !is_device_ptr(z)) { if (!tz.upload(z, n)) return -1; zd = tz.p; }
  if (!launch_dot(n, xd, yd, zd, mode)) return -1;
  if (!CSB_CUDA(cudaMemcpyAsync(c.red_out_h, c.red_out, sizeof(double)*3,
                                cudaMemcpyDeviceToHost, c.stream))) return -1;
  if (!CSB_CUDA(cudaStreamSynchronize(c.stream))) return -1;
  for (int i = 0; i < n_out; i++) out[i] = c.red_out_h[i];
  return 0;
}

double csb_dot(csb_lnum_t n, const csb_real_t *x, const csb_real_t *y)
{
  double r[1] = {NAN};
  _dot_api(n, x, y, nullptr, 0, r, 1);
  return r[0];
}

double csb_dot_xx(csb_lnum_t n, const csb_real_t *x)
{
  double r[1] = {NAN};
  _dot_api(n, x, nullptr, nullptr, 1, r, 1);
  return r[0];
}

int csb_dot_xx_xy_yz(csb_lnum_t n, const csb_real_t *x, const csb_real_t *y,
                     const csb_real_t *z, double *xx, double *xy, double *yz)
{
  double r[3];
  if (_dot_api(n, x, y, z, 2, r, 3) != 0) return -1;
  *xx = r[0]; *xy = r[1]; *yz = r[2];
  return 0;
}

} /* extern "C" */
