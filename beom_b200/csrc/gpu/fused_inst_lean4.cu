// fused_inst_lean4.cu -- the specialised fused step for 4 layer(s): generalized forward-backward (gene = 1) and the
// plain forward-backward start-up steps (gene = 0).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_lean4(const FusedLaunch &a, bool ufirst, bool gene0) {
  constexpr int GR = BEOM_LEAN4_GROUPS;
  if (gene0) return ufirst ? fused_launch_one<true, true, 4, 0, GR, 0, true>(a) : fused_launch_one<false, true, 4, 0, GR, 0, true>(a);
  return ufirst ? fused_launch_one<true, true, 4, 0, GR>(a) : fused_launch_one<false, true, 4, 0, GR>(a);
}
}  // namespace beom
