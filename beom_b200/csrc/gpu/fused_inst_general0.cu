// fused_inst_general0.cu -- the general fused step (every option a runtime switch) for any layer count <= 8.
#include "fused_inst.cuh"
namespace beom {
int fused_launch_general0(const FusedLaunch &a, bool ufirst, bool visc) {
  return visc ? (ufirst ? fused_launch_one<true, true, 0, -1, 0>(a) : fused_launch_one<false, true, 0, -1, 0>(a))
              : (ufirst ? fused_launch_one<true, false, 0, -1, 0>(a) : fused_launch_one<false, false, 0, -1, 0>(a));
}
}  // namespace beom
