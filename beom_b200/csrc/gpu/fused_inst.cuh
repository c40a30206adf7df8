// fused_inst.cuh -- launch helper shared by the instantiation units of the fused step kernel.
#ifndef BEOM_FUSED_INST_CUH
#define BEOM_FUSED_INST_CUH
#include <cstdio>
#include <cstdlib>

#include "fused.cuh"
#include "fused_kernel.cuh"

namespace beom {
template <bool UF, bool VI, int NL, int FEAT, int GROUPS, int FLAVOR = 0, bool G0 = false>
int fused_launch_one(const FusedLaunch &a) {
  static size_t configured = 0;
  auto kern = fusedk::k_fused_step<UF, VI, NL, FEAT, GROUPS, FLAVOR, G0>;
  if (a.shmem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.shmem) != cudaSuccess) return -61;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);  // (two CTAs per SM need it)
    configured = a.shmem;
    if (getenv("BEOM_FUSED_VERBOSE")) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, (int)a.block.x, a.shmem);
      fprintf(stderr, "[fused] %d layers, %d column groups: %u threads, %zu B of shared memory per CTA, %d CTA(s) per SM\n", NL, GROUPS, a.block.x, a.shmem, nb);
    }
  }
  kern<<<a.grid, a.block, a.shmem, a.stream>>>(*a.in, *a.out, *a.tab, a.open, a.open4, a.open4_words, a.groups, a.rows_per_chunk, a.wind_layers);
  return 0;
}
}  // namespace beom
#endif
