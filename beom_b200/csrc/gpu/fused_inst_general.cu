// fused_inst_general.cu -- the general fused step: every option a runtime switch, any layer count <= 8.
#include "fused_inst.cuh"
namespace beom {
template <int NL>
static int pick(const FusedLaunch &a, bool ufirst, bool visc) {
  return visc ? (ufirst ? fused_launch_one<true, true, NL, false, 0>(a) : fused_launch_one<false, true, NL, false, 0>(a))
              : (ufirst ? fused_launch_one<true, false, NL, false, 0>(a) : fused_launch_one<false, false, NL, false, 0>(a));
}
int fused_launch_general(const FusedLaunch &a, bool ufirst, bool visc, int nlay) {
  switch (nlay) {
    case 1: return pick<1>(a, ufirst, visc);
    case 2: return pick<2>(a, ufirst, visc);
    case 3: return pick<3>(a, ufirst, visc);
    case 4: return pick<4>(a, ufirst, visc);
    default: return pick<0>(a, ufirst, visc);
  }
}
}  // namespace beom
