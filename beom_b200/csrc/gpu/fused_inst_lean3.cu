// fused_inst_lean3.cu -- the specialised (4x unrolled, rotation-free) fused step for 3 layer(s).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_lean3(const FusedLaunch &a, bool ufirst) {
  return ufirst ? fused_launch_one<true, true, 3, true, fusedk::kMaxWarps / 3>(a) : fused_launch_one<false, true, 3, true, fusedk::kMaxWarps / 3>(a);
}
}  // namespace beom
