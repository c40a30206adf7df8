// tools/emu/stubs.cc -- TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): what the emulated library does not have.
#include <cuda_runtime.h>

#include <string>

#include "../../beom_b200/csrc/gpu/comm.h"
#include "../../beom_b200/csrc/gpu/fused.cuh"

namespace emu {
thread_local Idx block_idx, thread_idx;
thread_local dim3 block_dim, grid_dim;
}  // namespace emu

namespace beom {
// the FMA-contracted copies of the fused step are a hardware experiment: not built here
int fused_launch_lean1_fma(const FusedLaunch &, bool) { return -1; }
int fused_launch_lean2_fma(const FusedLaunch &, bool) { return -1; }
int fused_launch_lean3_fma(const FusedLaunch &, bool) { return -1; }
int fused_launch_lean4_fma(const FusedLaunch &, bool) { return -1; }
// Ranks: each rank is its own copy of the emulated library (loaded from its own file, so with its own globals) driven by
// its own thread of the test process; the halo exchange is handed to a callback of the test, which pairs the messages
// the way NCCL does (what a rank sends to its lower neighbour is what that neighbour receives from above).
typedef int (*emu_exchange_fn)(const double *send_lo, double *recv_lo, int peer_lo, const double *send_hi, double *recv_hi, int peer_hi,
                               size_t count);
static emu_exchange_fn g_exchange = nullptr;
static int g_rank = 0, g_size = 1;
int comm_unique_id(char[128], std::string *err) { if (err) *err = "cuda emulation: ranks are wired up by emu_comm_set"; return -40; }
int comm_init(const char[128], int, int, int, std::string *err) { if (err) *err = "cuda emulation: ranks are wired up by emu_comm_set"; return -40; }
int comm_finalize() { g_exchange = nullptr; g_rank = 0; g_size = 1; return 0; }
bool comm_ready() { return g_exchange != nullptr; }
int comm_rank() { return g_rank; }
int comm_size() { return g_size; }
int comm_exchange(const double *send_lo, double *recv_lo, int peer_lo, const double *send_hi, double *recv_hi, int peer_hi, size_t count,
                  cudaStream_t, std::string *err) {
  if (!g_exchange) { if (err) *err = "cuda emulation: no exchange callback"; return -44; }
  const int rc = g_exchange(send_lo, recv_lo, peer_lo, send_hi, recv_hi, peer_hi, count);
  if (rc && err) *err = "cuda emulation: the exchange callback failed";
  return rc;
}
int comm_allreduce_sum(double *, size_t, cudaStream_t, std::string *) { return 0; }
int comm_allreduce_max(double *, size_t, cudaStream_t, std::string *) { return 0; }
}  // namespace beom

extern "C" void emu_comm_set(int rank, int size, beom::emu_exchange_fn fn) {
  beom::g_rank = rank;
  beom::g_size = size;
  beom::g_exchange = fn;
}
