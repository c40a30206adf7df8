// fused_plan.cu -- host-only: prints the shared-memory plan of the fused step (fused_kernel.cuh smem_plan) for a
// layer count / column-group count / stream count, so that a test can check on a machine without a GPU that every
// specialised configuration still fits one SM.  Build: nvcc -std=c++17 -arch=sm_100a -o tools/fused_plan tools/fused_plan.cu
#include <cstdio>
#include <cstdlib>

#include "../beom_b200/csrc/gpu/fused_kernel.cuh"

int main(int argc, char **argv) {
  using namespace beom::fusedk;
  if (argc < 6) {
    std::printf("usage: fused_plan nlay groups n_all n_nowind wind_layers\n");
    return 2;
  }
  const int nlay = std::atoi(argv[1]), groups = std::atoi(argv[2]), n_all = std::atoi(argv[3]), n_nowind = std::atoi(argv[4]);
  const int wl = std::atoi(argv[5]);
  const SmemPlan p = smem_plan(nlay, groups, n_all, n_nowind, wl);
  std::printf("{\"total\": %zu, \"off_bars\": %zu, \"off_wring\": %zu, \"off_ring\": %zu, \"seg_bytes\": %zu, \"max_warps\": %d, \"mandatory\": %d, "
              "\"streams\": %d}\n",
              p.total, p.off_bars, p.off_wring, p.off_ring, p.seg_bytes, kMaxWarps, (int)kMandatory, (int)S_COUNT);
  return 0;
}
