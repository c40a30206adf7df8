// fused_inst_general.cu -- dispatch of the general fused step by layer count (instantiations: fused_inst_general<N>.cu).
#include "fused.cuh"
namespace beom {
int fused_launch_general0(const FusedLaunch &a, bool ufirst, bool visc);
int fused_launch_general1(const FusedLaunch &a, bool ufirst, bool visc);
int fused_launch_general2(const FusedLaunch &a, bool ufirst, bool visc);
int fused_launch_general3(const FusedLaunch &a, bool ufirst, bool visc);
int fused_launch_general4(const FusedLaunch &a, bool ufirst, bool visc);
int fused_launch_general(const FusedLaunch &a, bool ufirst, bool visc, int nlay) {
  switch (nlay) {
    case 1: return fused_launch_general1(a, ufirst, visc);
    case 2: return fused_launch_general2(a, ufirst, visc);
    case 3: return fused_launch_general3(a, ufirst, visc);
    case 4: return fused_launch_general4(a, ufirst, visc);
    default: return fused_launch_general0(a, ufirst, visc);
  }
}
}  // namespace beom
