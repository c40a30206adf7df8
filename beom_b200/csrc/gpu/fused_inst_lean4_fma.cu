// fused_inst_lean4_fma.cu -- the specialised fused step for 4 layer(s), compiled WITH FMA contraction (-fmad=true, see
// beom_b200/build.py).  Opt-in (BEOM_FMA=1): results then agree with the strict build to rounding (tests: max relative
// field difference <= 1e-10 after N steps), like the reference's own -Ofast build, instead of bit for bit.
#include "fused_inst.cuh"
namespace beom {
int fused_launch_lean4_fma(const FusedLaunch &a, bool ufirst) {
  return ufirst ? fused_launch_one<true, true, 4, 0, fusedk::kMaxWarps / 4, 1>(a) : fused_launch_one<false, true, 4, 0, fusedk::kMaxWarps / 4, 1>(a);
}
}  // namespace beom
