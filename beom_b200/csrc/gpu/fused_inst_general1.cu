// fused_inst_general1.cu -- the general fused step (every option a runtime switch) for 1 layer(s).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_general1(const FusedLaunch &a, bool ufirst, bool visc) {
  return visc ? (ufirst ? fused_launch_one<true, true, 1, -1, 0>(a) : fused_launch_one<false, true, 1, -1, 0>(a))
              : (ufirst ? fused_launch_one<true, false, 1, -1, 0>(a) : fused_launch_one<false, false, 1, -1, 0>(a));
}
}  // namespace beom
