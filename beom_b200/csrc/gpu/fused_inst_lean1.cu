// fused_inst_lean1.cu -- the specialised fused step for 1 layer(s): generalized forward-backward (gene = 1) and the
// plain forward-backward start-up steps (gene = 0).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_lean1(const FusedLaunch &a, bool ufirst, bool gene0) {
  constexpr int GR = fusedk::kMaxWarps / 1;
  if (gene0) return ufirst ? fused_launch_one<true, true, 1, 0, GR, 0, true>(a) : fused_launch_one<false, true, 1, 0, GR, 0, true>(a);
  return ufirst ? fused_launch_one<true, true, 1, 0, GR>(a) : fused_launch_one<false, true, 1, 0, GR>(a);
}
}  // namespace beom
