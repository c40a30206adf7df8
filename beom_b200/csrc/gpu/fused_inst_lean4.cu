// fused_inst_lean4.cu -- the specialised (4x unrolled, rotation-free) fused step for 4 layer(s).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_lean4(const FusedLaunch &a, bool ufirst) {
  return ufirst ? fused_launch_one<true, true, 4, true, fusedk::kMaxWarps / 4>(a) : fused_launch_one<false, true, 4, true, fusedk::kMaxWarps / 4>(a);
}
}  // namespace beom
