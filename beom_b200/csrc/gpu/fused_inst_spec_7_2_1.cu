// fused_inst_spec_7_2_1.cu -- the fused step specialised for the option set 7 (fused_kernel.cuh: FB_NUDG = 1, FB_OCRP = 2,
// FB_BDRG = 4), 2 layer(s), 6 column groups, Leith/constant viscosity true: gene = 1 and the gene = 0 start-up copy.
#include "fused_inst.cuh"
namespace beom {
int fused_launch_spec_7_2_1(const FusedLaunch &a, bool ufirst, bool gene0) {
  if (gene0) return ufirst ? fused_launch_one<true, true, 2, 7, 6, 0, true>(a) : fused_launch_one<false, true, 2, 7, 6, 0, true>(a);
  return ufirst ? fused_launch_one<true, true, 2, 7, 6>(a) : fused_launch_one<false, true, 2, 7, 6>(a);
}
}  // namespace beom
