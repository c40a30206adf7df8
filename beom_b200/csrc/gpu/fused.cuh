// fused.cuh -- the fused single-pass step (DESIGN.md section "fused step").
#ifndef BEOM_FUSED_CUH
#define BEOM_FUSED_CUH
#include <cuda_runtime.h>

#include "dev.cuh"

namespace beom {
// Decides whether the case can run on the fused path and prepares it.
int fused_configure(const Dev &D, const beom_params &P, int nmir, int nranks, bool *enabled);
bool fused_supports(bool first_three, bool upst);
int fused_step(const Dev &in, const Dev &out, int tstp, bool first_three, cudaStream_t s, int *nlaunch);
}  // namespace beom
#endif
