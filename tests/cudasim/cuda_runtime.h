/* tests/cudasim/cuda_runtime.h -- TEST INFRASTRUCTURE, never part of the product.
 *
 * A stand-in for the CUDA runtime that lets the UNMODIFIED kernel sources of libcsb200
 * (csrc/kernels.cu, the kernel-per-phase path) and its host logic (csrc/sles.cu,
 * csrc/runtime.cu) be compiled by g++ and executed on host threads, so that kernels and
 * launch sequences can be desk-checked against the oracle in a container without a GPU
 * (tests/test_cudasim.py).  The CUDA threads of a block are fibers run round-robin, blocks
 * one after the other, __syncthreads = a barrier among the fibers, __shfl_down_sync = exchange
 * through a per-block array (cudasim.cpp).  It proves nothing about performance or about real concurrency on the device;
 * the `-m gpu` tests remain the parity tests proper.  The product library is never built
 * from this header and has no CPU path.
 */
#ifndef CUDASIM_CUDA_RUNTIME_H
#define CUDASIM_CUDA_RUNTIME_H

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

#define CUDASIM 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a) : x(a) {} };

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorSim = 1 };
typedef struct cudasim_stream *cudaStream_t;
typedef struct cudasim_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost,
                      cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost, cudaMemoryTypeDevice,
                      cudaMemoryTypeManaged };
struct cudaPointerAttributes { cudaMemoryType type; };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64];
                        int major, minor; int cooperativeLaunch; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace cudasim {
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void  launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body);
void *dyn_smem();
void  syncthreads();
double shfl_down(double v, unsigned delta);
void  warp_yield();
}
#define threadIdx (cudasim::t_threadIdx)
#define blockIdx  (cudasim::t_blockIdx)
#define blockDim  (cudasim::g_blockDim)
#define gridDim   (cudasim::g_gridDim)

inline void __syncthreads() { cudasim::syncthreads(); }
inline void __threadfence() { __sync_synchronize(); }
inline double __shfl_down_sync(unsigned, double v, unsigned delta) { return cudasim::shfl_down(v, delta); }
/* __syncwarp: one yield -- with ascending round-robin every lane of the warp has reached the
   same yield before any lane passes it (warps keep uniform control flow around their syncs) */
inline void __syncwarp(unsigned = 0xffffffffu) { cudasim::warp_yield(); }
inline int atomicAdd(int *p, int v) { const int o = *p; *p += v; return o; }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *(const volatile T *)p; }
inline unsigned atomicInc(unsigned *p, unsigned limit)
{
  /* blocks run one after the other and only one thread of a block calls this */
  unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  for (;;) {
    unsigned nv = (old >= limit) ? 0u : old + 1u;
    if (__atomic_compare_exchange_n(p, &old, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST))
      return old;
  }
}
using std::isnan; using std::isinf; using std::sqrt; using std::fabs;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

/* runtime: "device" memory is host memory */
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorSim; }
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
#define cudaHostAllocPortable 1u
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "cudasim error"; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
  memset(p, 0, sizeof(*p));
  p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227*1024; p->major = 10;
  strcpy(p->name, "cudasim (host threads)");
  return cudaSuccess;
}
/* every pointer is reported as host memory: callers' buffers are staged, like host arrays
   on the real device */
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeHost; return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

#endif
