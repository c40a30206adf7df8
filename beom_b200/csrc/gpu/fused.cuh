// fused.cuh -- the fused single-pass step (DESIGN.md section "fused step").
#ifndef BEOM_FUSED_CUH
#define BEOM_FUSED_CUH
#include <cuda_runtime.h>

#include "dev.cuh"

namespace beom {
// Decides whether the case can run on the fused path and prepares it.
int fused_configure(const Dev &D, const beom_params &P, int nmir, int nranks, bool *enabled);
bool fused_supports(bool first_three, bool upst);
int fused_step(const Dev &in, const Dev &out, int tstp, bool first_three, cudaStream_t s, int *nlaunch, int part = 0, int edge = 0);
void fused_release();
const char *fused_variant();  // which instantiation runs: "specialised (options .., .. layers, .. column groups)" / "general (..)"

// ---- dispatch into the instantiation translation units (fused_inst_*.cu, compiled in parallel) ----
namespace fusedk { struct StreamTab; }
struct FusedLaunch {
  dim3 grid, block;
  size_t shmem;
  const Dev *in, *out;
  const fusedk::StreamTab *tab;
  const uint8_t *open;
  const unsigned *open4;
  int open4_words;
  int groups, rows_per_chunk, wind_layers;
  cudaStream_t stream;
};
int fused_launch_lean1(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_lean2(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_lean3(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_lean4(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_lean1_fma(const FusedLaunch &a, bool ufirst);  // FMA-contracted copies (BEOM_FMA=1)
int fused_launch_lean2_fma(const FusedLaunch &a, bool ufirst);
int fused_launch_lean3_fma(const FusedLaunch &a, bool ufirst);
int fused_launch_lean4_fma(const FusedLaunch &a, bool ufirst);
int fused_launch_general(const FusedLaunch &a, bool ufirst, bool visc, int nlay);
// specialised option sets (fused_inst_spec_<feat>_<nlay>_<visc>.cu; fused_kernel.cuh: FB_NUDG = 1, FB_OCRP = 2, FB_BDRG = 4)
int fused_launch_spec_3_2_1(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_spec_3_4_1(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_spec_1_2_1(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_spec_1_4_1(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_spec_7_2_1(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_spec_7_4_1(const FusedLaunch &a, bool ufirst, bool gene0);
int fused_launch_spec_4_1_0(const FusedLaunch &a, bool ufirst, bool gene0);
}  // namespace beom
#endif
