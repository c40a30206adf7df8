// fused_inst_lean2.cu -- the specialised (4x unrolled, rotation-free) fused step for 2 layer(s).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_lean2(const FusedLaunch &a, bool ufirst) {
  return ufirst ? fused_launch_one<true, true, 2, true, fusedk::kMaxWarps / 2>(a) : fused_launch_one<false, true, 2, true, fusedk::kMaxWarps / 2>(a);
}
}  // namespace beom
