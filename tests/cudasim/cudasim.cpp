/* tests/cudasim/cudasim.cpp -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
 * Block executor: a pool of host threads runs the threads of one block at a time. */
#include "cuda_runtime.h"

#include <condition_variable>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <vector>

namespace cudasim {

thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {

constexpr unsigned MAX_THREADS = 1024;

struct Pool {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  unsigned long long generation = 0;
  unsigned active = 0, remaining = 0;
  unsigned block_id = 0;
  const std::function<void()> *body = nullptr;
  pthread_barrier_t block_barrier;
  pthread_barrier_t warp_barrier[MAX_THREADS/32];
  unsigned barrier_count = 0;
  std::vector<unsigned char> smem;
  double xchg[MAX_THREADS];
  bool stop = false;

  void worker(unsigned tid)
  {
    unsigned long long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu);
      cv_go.wait(lk, [&] { return stop || (generation != seen && tid < active); });
      if (stop) return;
      seen = generation;
      const std::function<void()> *b = body;
      t_threadIdx.x = tid;
      t_blockIdx.x = block_id;
      lk.unlock();
      (*b)();
      lk.lock();
      if (--remaining == 0) cv_done.notify_one();
    }
  }

  void ensure(unsigned n)
  {
    while (workers.size() < n) {
      unsigned tid = (unsigned)workers.size();
      workers.emplace_back([this, tid] { worker(tid); });
    }
  }

  void set_barriers(unsigned n)
  {
    if (n == barrier_count) return;
    if (barrier_count) {
      pthread_barrier_destroy(&block_barrier);
      for (unsigned w = 0; w < MAX_THREADS/32; w++) pthread_barrier_destroy(&warp_barrier[w]);
    }
    pthread_barrier_init(&block_barrier, nullptr, n);
    for (unsigned w = 0; w < MAX_THREADS/32; w++) {
      unsigned lanes = (n >= (w + 1)*32) ? 32 : (n > w*32 ? n - w*32 : 1);
      pthread_barrier_init(&warp_barrier[w], nullptr, lanes);
    }
    barrier_count = n;
  }

  ~Pool()
  {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv_go.notify_all();
    for (auto &t : workers) t.join();
  }
};

Pool &pool() { static Pool *p = new Pool; return *p; }   /* leaked on purpose: no exit-order issues */

} // namespace

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body)
{
  Pool &P = pool();
  if (grid == 0 || block == 0) return;
  if (block > MAX_THREADS) abort();
  g_blockDim.x = block; g_gridDim.x = grid;
  if (block == 1) {              /* scalar kernels: run inline */
    for (unsigned b = 0; b < grid; b++) {
      t_threadIdx.x = 0; t_blockIdx.x = b;
      body();
    }
    return;
  }
  P.ensure(block);
  P.set_barriers(block);
  if (P.smem.size() < smem + 16) P.smem.resize(smem + 16);
  for (unsigned b = 0; b < grid; b++) {
    std::unique_lock<std::mutex> lk(P.mu);
    P.body = &body; P.block_id = b; P.active = block; P.remaining = block;
    P.generation++;
    P.cv_go.notify_all();
    P.cv_done.wait(lk, [&] { return P.remaining == 0; });
  }
}

void *dyn_smem() { return pool().smem.data(); }

void syncthreads()
{
  if (g_blockDim.x > 1) pthread_barrier_wait(&pool().block_barrier);
}

double shfl_down(double v, unsigned delta)
{
  Pool &P = pool();
  const unsigned tid = t_threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (g_blockDim.x == 1) return v;
  P.xchg[tid] = v;
  pthread_barrier_wait(&P.warp_barrier[w]);
  const unsigned src = (lane + delta < 32 && tid + delta < g_blockDim.x) ? tid + delta : tid;
  const double r = P.xchg[src];
  pthread_barrier_wait(&P.warp_barrier[w]);
  return r;
}

} // namespace cudasim
