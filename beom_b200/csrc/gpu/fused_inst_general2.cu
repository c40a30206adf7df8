// fused_inst_general2.cu -- the general fused step (every option a runtime switch) for 2 layer(s).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_general2(const FusedLaunch &a, bool ufirst, bool visc) {
  return visc ? (ufirst ? fused_launch_one<true, true, 2, -1, 0>(a) : fused_launch_one<false, true, 2, -1, 0>(a))
              : (ufirst ? fused_launch_one<true, false, 2, -1, 0>(a) : fused_launch_one<false, false, 2, -1, 0>(a));
}
}  // namespace beom
