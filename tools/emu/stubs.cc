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
// the fused step (TMA, mbarriers, warp shuffles) is hardware: every case runs the split path here
int fused_configure(const Dev &, const beom_params &, int, int, bool *enabled) { *enabled = false; return 0; }
bool fused_supports(bool, bool) { return false; }
int fused_step(const Dev &, const Dev &, int, bool, cudaStream_t, int *, int, int) { return -1; }
void fused_release() {}
// one rank only
int comm_unique_id(char[128], std::string *err) { if (err) *err = "cuda emulation: no communicator"; return -40; }
int comm_init(const char[128], int, int, int, std::string *err) { if (err) *err = "cuda emulation: no communicator"; return -40; }
int comm_finalize() { return 0; }
bool comm_ready() { return false; }
int comm_rank() { return 0; }
int comm_size() { return 1; }
int comm_exchange(const double *, double *, int, const double *, double *, int, size_t, cudaStream_t, std::string *err) {
  if (err) *err = "cuda emulation: no communicator";
  return -44;
}
int comm_allreduce_sum(double *, size_t, cudaStream_t, std::string *) { return 0; }
int comm_allreduce_max(double *, size_t, cudaStream_t, std::string *) { return 0; }
}  // namespace beom
