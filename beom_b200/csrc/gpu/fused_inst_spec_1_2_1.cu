// fused_inst_spec_1_2_1.cu -- the fused step specialised for the option set 1 (fused_kernel.cuh: FB_NUDG = 1, FB_OCRP = 2,
// FB_BDRG = 4), 2 layer(s), 7 column groups, Leith/constant viscosity true: gene = 1 and the gene = 0 start-up copy.
#include "fused_inst.cuh"
namespace beom {
int fused_launch_spec_1_2_1(const FusedLaunch &a, bool ufirst, bool gene0) {
  if (gene0) return ufirst ? fused_launch_one<true, true, 2, 1, 7, 0, true>(a) : fused_launch_one<false, true, 2, 1, 7, 0, true>(a);
  return ufirst ? fused_launch_one<true, true, 2, 1, 7>(a) : fused_launch_one<false, true, 2, 1, 7>(a);
}
}  // namespace beom
