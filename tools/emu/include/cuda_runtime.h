// tools/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the CUDA runtime that lets g++ compile the SOURCE of the split-path kernels
// (beom_b200/csrc/gpu/split.cuh, diag.cuh) and of the library's host code (beom_gpu.cu) for the CPU, every kernel
// launch becoming a serial loop over its grid (tools/emu/build_emu.py rewrites the <<< >>> launches).  Purpose: check
// the LOGIC of kernels and host plumbing against the oracle on a machine without a GPU -- indexing, masks, option
// switches, upload/download, periodic images, open-boundary segments.  It cannot say anything about races, barriers,
// TMA or performance, the fused step is not part of it (stubbed: every case runs the split path), and kernels that
// cooperate through __syncthreads / shuffles abort if they are ever launched.
//
// The product never builds, ships or loads this: beom_b200/build.py does not know it, beom_b200/_lib.py loads
// beom_b200/lib/libbeom_gpu.so only, and the library built here reports itself as "cpu-emulation" (tests/test_emulation.py).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <vector>

#define BEOM_CUDA_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
namespace emu {
struct Idx { unsigned x, y, z; };
extern thread_local Idx block_idx, thread_idx;
extern thread_local dim3 block_dim, grid_dim;
// A stream capture in progress (cudaStreamBeginCapture below): launches are recorded, not run.  One capture at a time, streams
// are not told apart (the library captures on its compute stream only).
struct Graph { std::vector<std::function<void()>> nodes; };
inline Graph *capturing = nullptr;
template <class F>
void run_grid(dim3 grid, dim3 block, const F &body);
template <class F>
void launch(dim3 grid, dim3 block, F &&body) {
  if (capturing) {
    capturing->nodes.push_back([grid, block, body]() { run_grid(grid, block, body); });
    return;
  }
  run_grid(grid, block, body);
}
template <class F>
void run_grid(dim3 grid, dim3 block, const F &body) {
  grid_dim = grid;
  block_dim = block;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++)
        for (unsigned tz = 0; tz < block.z; tz++)
          for (unsigned ty = 0; ty < block.y; ty++)
            for (unsigned tx = 0; tx < block.x; tx++) {
              block_idx = Idx{bx, by, bz};
              thread_idx = Idx{tx, ty, tz};
              body();
            }
}
[[noreturn]] inline void not_emulated(const char *what) {
  std::fprintf(stderr, "cuda emulation: %s is not emulated (threads run one after the other)\n", what);
  std::abort();
}
}  // namespace emu
#define blockIdx (emu::block_idx)
#define threadIdx (emu::thread_idx)
#define blockDim (emu::block_dim)
#define gridDim (emu::grid_dim)

#include "simt.h"  // shuffles, votes, __syncwarp, __syncthreads, mbarriers: only inside emu::launch_simt
namespace emu {
inline void launch_simt_c(dim3 grid, dim3 block, size_t shmem, std::function<void()> body) {
  if (capturing) {
    capturing->nodes.push_back([grid, block, shmem, body]() { launch_simt(grid, block, shmem, body); });
    return;
  }
  launch_simt(grid, block, shmem, body);
}
}  // namespace emu
template <class T> inline void __stcs(T *p, T v) { *p = v; }
using std::max;
using std::min;

enum cudaError_t { cudaSuccess = 0, cudaErrorEmulation = 1 };
inline const char *cudaGetErrorString(cudaError_t) { return "cuda emulation"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
typedef struct emuStream *cudaStream_t;
struct emuEvent { std::chrono::steady_clock::time_point t; };
typedef emuEvent *cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
// (multiProcessorCount = 1 and one resident block per SM: a persistent kernel sized by the occupancy query -- the rigid-lid
// solver -- becomes ONE block here, whose warps are concurrent OS threads; blocks run one after the other)
struct cudaDeviceProp { int major = 10, minor = 0; int multiProcessorCount = 1; char name[64] = "cpu-emulation of sm_100a"; };
template <class F> inline int cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return 0; }

enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount, cudaDevAttrMaxSharedMemoryPerBlockOptin };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
enum { cudaSharedmemCarveoutMaxShared = 100 };
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
// CUDA graphs: a capture records the launches (with their arguments as they were when issued) instead of running them, a graph
// launch runs them in order.  BEOM_EMU_NO_CAPTURE=1 makes the capture fail, which is how the library's fallback is tested.
typedef emu::Graph *cudaGraph_t;
typedef emu::Graph *cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
inline cudaError_t cudaDeviceGetPCIBusId(char *b, int n, int) { if (n > 0) b[0] = 0; return cudaErrorEmulation; }  // no PCI device: no NUMA placement
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) { *v = a == cudaDevAttrMultiProcessorCount ? 148 : 227 * 1024; return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
  if (emu::capturing || (getenv("BEOM_EMU_NO_CAPTURE") && atoi(getenv("BEOM_EMU_NO_CAPTURE")) > 0)) return cudaErrorEmulation;
  emu::capturing = new emu::Graph;
  return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) {
  *g = emu::capturing;
  emu::capturing = nullptr;
  return *g ? cudaSuccess : cudaErrorEmulation;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long = 0) { *e = new emu::Graph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) {
  if (emu::capturing) return cudaErrorEmulation;
  for (auto &n : e->nodes) n();
  return cudaSuccess;
}
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
// fresh device memory is not zero: fill it with a byte pattern (as doubles: -6.2e+66), so that a kernel relying on cudaMalloc'ed
// memory it never wrote shows up in the comparison with the oracle
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) {
  *p = static_cast<T *>(std::malloc(n ? n : 1));
  if (*p) std::memset(*p, 0xCD, n ? n : 1);
  return *p ? cudaSuccess : cudaErrorEmulation;
}
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return emu::capturing ? cudaErrorEmulation : cudaMemcpy(d, s, n, k); }  // (not recordable here)
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { return emu::capturing ? cudaErrorEmulation : cudaMemset(d, v, n); }
