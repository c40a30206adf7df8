// fused_inst_general4.cu -- the general fused step (every option a runtime switch) for 4 layer(s).
#include "fused_inst.cuh"
namespace beom {
int fused_launch_general4(const FusedLaunch &a, bool ufirst, bool visc) {
  return visc ? (ufirst ? fused_launch_one<true, true, 4, -1, 0>(a) : fused_launch_one<false, true, 4, -1, 0>(a))
              : (ufirst ? fused_launch_one<true, false, 4, -1, 0>(a) : fused_launch_one<false, false, 4, -1, 0>(a));
}
}  // namespace beom
