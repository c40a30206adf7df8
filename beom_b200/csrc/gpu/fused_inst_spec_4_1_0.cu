// fused_inst_spec_4_1_0.cu -- the fused step specialised for the option set 4 (fused_kernel.cuh: FB_NUDG = 1, FB_OCRP = 2,
// FB_BDRG = 4), 1 layer(s), 15 column groups, Leith/constant viscosity false: gene = 1 and the gene = 0 start-up copy.
#include "fused_inst.cuh"
namespace beom {
int fused_launch_spec_4_1_0(const FusedLaunch &a, bool ufirst, bool gene0) {
  if (gene0) return ufirst ? fused_launch_one<true, false, 1, 4, 15, 0, true>(a) : fused_launch_one<false, false, 1, 4, 15, 0, true>(a);
  return ufirst ? fused_launch_one<true, false, 1, 4, 15>(a) : fused_launch_one<false, false, 1, 4, 15>(a);
}
}  // namespace beom
