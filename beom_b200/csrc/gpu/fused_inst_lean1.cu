// fused_inst_lean1.cu -- the specialised (4x unrolled, rotation-free) fused step for 1 layer(s).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_lean1(const FusedLaunch &a, bool ufirst) {
  return ufirst ? fused_launch_one<true, true, 1, true, fusedk::kMaxWarps / 1>(a) : fused_launch_one<false, true, 1, true, fusedk::kMaxWarps / 1>(a);
}
}  // namespace beom
