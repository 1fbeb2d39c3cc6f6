/* tests/cudasim/cudasim.cpp -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
 * Block executor: the threads of one block are fibers (ucontext) of the calling OS thread,
 * run round-robin in ascending thread id; blocks one after the other.
 *   __syncthreads   : the fiber yields and is resumed once every live fiber of the block
 *                     waits at a barrier;
 *   __shfl_down_sync: the fiber publishes its value, yields once, then reads the source
 *                     lane -- with ascending round-robin every lane of the warp has published
 *                     before any lane reads, and a lane only ever reads higher lanes, which
 *                     have not yet overwritten their slot in the current pass.
 * Deterministic, no OS synchronisation; a block whose threads disagree on the number of
 * barriers is reported as a deadlock instead of hanging. */
#include "cuda_runtime.h"

#include <cstdio>
#include <ucontext.h>
#include <vector>

namespace cudasim {

thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;

namespace {

constexpr unsigned MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = 128*1024;

enum State { READY, AT_BARRIER, AT_SHFL, DONE };

struct Fiber {
  ucontext_t ctx;
  char      *stack = nullptr;
  State      state = DONE;
};

struct Exec {
  std::vector<Fiber> fibers;
  ucontext_t main_ctx;
  const std::function<void()> *body = nullptr;
  unsigned current = 0;
  std::vector<unsigned char> smem;
  double xchg[MAX_THREADS];
  bool in_kernel = false;
};

Exec &ex() { static Exec *e = new Exec; return *e; }

void fiber_entry()
{
  Exec &E = ex();
  (*E.body)();
  E.fibers[E.current].state = DONE;
  swapcontext(&E.fibers[E.current].ctx, &E.main_ctx);
}

void yield(State s)
{
  Exec &E = ex();
  Fiber &f = E.fibers[E.current];
  f.state = s;
  swapcontext(&f.ctx, &E.main_ctx);
}

void run_block(unsigned block)
{
  Exec &E = ex();
  if (E.fibers.size() < block) E.fibers.resize(block);
  for (unsigned t = 0; t < block; t++) {
    Fiber &f = E.fibers[t];
    if (f.stack == nullptr) f.stack = (char *)malloc(STACK_BYTES);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STACK_BYTES;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, fiber_entry, 0);
    f.state = READY;
  }
  for (;;) {
    unsigned alive = 0, at_barrier = 0;
    for (unsigned t = 0; t < block; t++) {
      if (E.fibers[t].state != DONE) alive++;
      if (E.fibers[t].state == AT_BARRIER) at_barrier++;
    }
    if (alive == 0) break;
    const bool release = (at_barrier == alive);
    unsigned resumed = 0;
    for (unsigned t = 0; t < block; t++) {
      Fiber &f = E.fibers[t];
      if (f.state == DONE || (f.state == AT_BARRIER && !release)) continue;
      E.current = t;
      t_threadIdx.x = t;
      f.state = READY;
      resumed++;
      swapcontext(&E.main_ctx, &f.ctx);
    }
    if (resumed == 0) {
      fprintf(stderr, "cudasim: deadlock -- %u of %u live threads wait at __syncthreads\n",
              at_barrier, alive);
      abort();
    }
  }
}

} // namespace

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body)
{
  Exec &E = ex();
  if (grid == 0 || block == 0) return;
  if (block > MAX_THREADS || E.in_kernel) abort();
  g_blockDim.x = block; g_gridDim.x = grid;
  if (E.smem.size() < smem + 16) E.smem.resize(smem + 16);
  E.body = &body;
  E.in_kernel = true;
  for (unsigned b = 0; b < grid; b++) {
    t_blockIdx.x = b;
    if (block == 1) {            /* scalar kernels: no fiber needed */
      t_threadIdx.x = 0;
      body();
    }
    else
      run_block(block);
  }
  E.in_kernel = false;
}

void *dyn_smem() { return ex().smem.data(); }

void syncthreads()
{
  if (g_blockDim.x > 1) yield(AT_BARRIER);
}

void warp_yield()
{
  if (g_blockDim.x > 1) yield(AT_SHFL);
}

double shfl_down(double v, unsigned delta)
{
  if (g_blockDim.x == 1) return v;
  Exec &E = ex();
  const unsigned tid = E.current, lane = tid & 31;
  E.xchg[tid] = v;
  yield(AT_SHFL);
  const unsigned src = (lane + delta < 32 && tid + delta < g_blockDim.x) ? tid + delta : tid;
  return E.xchg[src];
}

} // namespace cudasim
