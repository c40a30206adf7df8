// fused_inst_spec_7_4_1.cu -- the fused step specialised for the option set 7 (fused_kernel.cuh: FB_NUDG = 1, FB_OCRP = 2,
// FB_BDRG = 4), 4 layer(s), 3 column groups, Leith/constant viscosity true: gene = 1 and the gene = 0 start-up copy.
#include "fused_inst.cuh"
namespace beom {
int fused_launch_spec_7_4_1(const FusedLaunch &a, bool ufirst, bool gene0) {
  if (gene0) return ufirst ? fused_launch_one<true, true, 4, 7, 3, 0, true>(a) : fused_launch_one<false, true, 4, 7, 3, 0, true>(a);
  return ufirst ? fused_launch_one<true, true, 4, 7, 3>(a) : fused_launch_one<false, true, 4, 7, 3>(a);
}
}  // namespace beom
