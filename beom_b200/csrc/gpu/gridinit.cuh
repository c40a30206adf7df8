// gridinit.cuh -- read_input_data's grid-shaped work on the device (beom_gpu_init_grids, beom_gpu.cu): the masks and the
// vector numbering of index_grid_points (private_mod.f95:567-764) as a prefix sum, the rest thickness
// (:154-183: layers stacked from the bottom, or the per-column Newton solve of get_equilibrium_thickness_h_0, :309-502),
// and read_input_file's conversions of the raw float32 input grids (:840-964) -- written straight into the dense planes
// of this rank, with the reference's operations in the reference's order (the host restatement init.cc produces the same
// bits: tests/test_grid_init.py).  On periodic domains the numbering, the flags and the image lists come from the host's layout
// analysis instead of k_gi_index; what is not covered at all makes beom_gpu_init_grids say so and the caller falls back to
// read_input_data + beom_gpu_init.
#ifndef BEOM_GRIDINIT_CUH
#define BEOM_GRIDINIT_CUH
#include "dev.cuh"
#include "rest_solver.h"

namespace beom {

struct GridIn {
  int lm, mm, nlay;
  int NX, NY, j_off;          // dense layout of this rank
  int jlo, jhi;               // grid rows held (owned + halo rows, clipped to 0 .. mm+1)
  double hdry, flat;          // dry threshold; default depth cext**2/grav (private_mod.f95:119-121)
  const float *h_bo;          // raw files (device copies), Fortran order (0:lm+1, 0:mm+1, ...); null = absent
  const float *init, *nudg, *taus, *fcor, *hdot, *tide;
  double tauw[2], f0, dmax;
  double topl[BEOM_MAXLAY];
};

// depth of cell (i, j), i in -1 .. lm+2: margins dry, below hdry dry (private_mod.f95:827-839)
__device__ __forceinline__ double gi_depth(const GridIn &A, int i, int j) {
  if (i < 1 || i > A.lm || j < 1 || j > A.mm) return 0.0;
  if (!A.h_bo) return A.flat;
  const double d = (double)A.h_bo[(size_t)j * (A.lm + 2) + i];
  return d < A.hdry ? 0.0 : d;
}
__device__ __forceinline__ bool gi_wet(const GridIn &A, int i, int j) { return gi_depth(A, i, j) > A.hdry; }
// a grid point enters the vector if it carries an eta, u, v or psi point that touches water (pm:692-730)
__device__ __forceinline__ bool gi_carries(const GridIn &A, int i, int j) {
  return gi_wet(A, i, j) || gi_wet(A, i - 1, j) || gi_wet(A, i, j - 1) || gi_wet(A, i - 1, j - 1);
}

// vector points per grid row j = 0 .. mm+1 (one block per row), and the depth extremes of the row (pm:134-135)
__global__ void k_gi_row_counts(const __grid_constant__ GridIn A, int *__restrict__ rowcnt, double *__restrict__ rowmin, double *__restrict__ rowmax) {
  const int j = blockIdx.x;
  int n = 0;
  double lo = INFINITY, hi = 0.0;
  for (int i = threadIdx.x; i <= A.lm + 1; i += blockDim.x) {
    n += gi_carries(A, i, j) ? 1 : 0;
    const double d = gi_depth(A, i, j);
    if (d > A.hdry) lo = fmin(lo, d);
    hi = fmax(hi, d);
  }
  __shared__ int sn[32];
  __shared__ double slo[32], shi[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, o);
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { sn[w] = n; slo[w] = lo; shi[w] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < nw; k++) { n += sn[k]; lo = fmin(lo, slo[k]); hi = fmax(hi, shi[k]); }
    rowcnt[j] = n;
    rowmin[j] = lo;
    rowmax[j] = hi;
  }
}

// the vector numbering (ipnt ascends with i inside j, pm:692-730): rowoff[j] = points in the rows below; one block per held row
// writes cell_of_point, the flag byte (the five masks of pm:700-714 + "is a vector point") and h_th (pm:753-757)
__global__ void k_gi_index(const __grid_constant__ GridIn A, const int *__restrict__ rowoff, int *__restrict__ cell_of_point,
                           uint8_t *__restrict__ flags, double *__restrict__ h_th) {
  const int j = A.jlo + blockIdx.x;
  __shared__ int wsum[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int base = rowoff[j];
  for (int i0 = 0; i0 <= A.lm + 1; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool c = i <= A.lm + 1 && gi_carries(A, i, j);
    const unsigned b = __ballot_sync(0xffffffffu, c);
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    int before = carry;
    for (int k = 0; k < w; k++) before += wsum[k];
    if (c) {
      const int p = 1 + base + before + __popc(b & ((1u << lane) - 1u));
      const int cell = (j + A.j_off) * A.NX + (i + GX0);
      cell_of_point[p] = cell;
      uint8_t f = F_ACT | F_PI;  // mkpi: carries() already says one of the four cells is wet (pm:712-714)
      const bool w00 = gi_wet(A, i, j), wW = gi_wet(A, i - 1, j), wS = gi_wet(A, i, j - 1), wSW = gi_wet(A, i - 1, j - 1);
      if (w00) f |= F_N;
      if (wW && w00) f |= F_U;
      if (wS && w00) f |= F_V;
      if (wSW && wS && wW && w00) f |= F_PE;
      flags[cell] = f;
      h_th[cell] = gi_depth(A, i, j);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < nw; k++) t += wsum[k];
      carry += t;
    }
    __syncthreads();
  }
}

// periodic domains (the cell map and the flags come from the host's layout analysis): h_th of the vector points' cells (pm:753-757)
__global__ void k_gi_hth(const __grid_constant__ GridIn A, const uint8_t *__restrict__ flags, double *__restrict__ h_th) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= A.NX) return;
  const size_t c = (size_t)y * A.NX + x;
  if (flags[c] & F_ACT) h_th[c] = gi_depth(A, x - GX0, y - A.j_off);
}

// rest thickness without outcropping: layers stacked from the bottom (pm:154-175); h_0 stays 0 at dry points
__global__ void k_gi_h0_stack(const __grid_constant__ GridIn A, const uint8_t *__restrict__ flags, const double *__restrict__ h_th,
                              double *__restrict__ h_0, size_t plane) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= plane || (flags[c] & (F_ACT | F_N)) != (F_ACT | F_N)) return;  // wet vector points (not their periodic images)
  const double depth = h_th[c];
  for (int l = A.nlay - 1; l >= 0; l--) {
    const double above = l > 0 ? A.dmax * A.topl[l] : 0.0;
    double below = 0.0;
    for (int k = l + 1; k < A.nlay; k++) below += h_0[(size_t)k * plane + c];
    h_0[(size_t)l * plane + c] = depth - above - below;
  }
}

// get_equilibrium_thickness_h_0 (pm:309-502): the Newton iteration of rest_solver.h (the host's own source), one thread per
// wet cell; *bad receives the largest cell index whose column did not converge within itmx iterations (pm:395-401), else stays -1
__global__ void k_gi_h0_newton(const __grid_constant__ RestSolver S, const uint8_t *__restrict__ flags, const double *__restrict__ h_th,
                               double *__restrict__ h_0, size_t plane, int *__restrict__ bad) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= plane || (flags[c] & (F_ACT | F_N)) != (F_ACT | F_N)) return;  // wet vector points (not their periodic images)
  double col[BEOM_MAXLAY];
  if (!S.column(h_th[c], col)) {
    atomicMax(bad, (int)c);
    return;
  }
  for (int l = 0; l < S.nlay; l++) h_0[(size_t)l * plane + c] = col[l];
}

// forcing files -> dense planes (read_input_file, pm:840-964), one thread per dense cell that is a vector point
struct GridOut {
  size_t plane;
  const uint8_t *flags;
  const double *h_0;                 // [nlay]
  double *hlay, *u, *v;              // [nlay] state, level 0
  double *nudg;                      // [3] or null
  double *fnud;                      // [3][nlay] or null
  double *taus;                      // [2] or null
  double *fcor;                      // [1]
  double *hdot;                      // [nlay] or null
  double *tide;                      // [3][2] (amplitude, phase of eta, u, v) or null
  unsigned *any;                     // bit 0: nudg live (> 1e-9), 1: nudg non-zero, 2: |taus| > 1e-7, 3: hdot non-zero
  int set_state;                     // rsta < 0.5: init.bin is the initial state too
};
__global__ void k_gi_forcing(const __grid_constant__ GridIn A, const __grid_constant__ GridOut O) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= A.NX) return;
  const size_t c = (size_t)y * A.NX + x, pl = O.plane;
  const uint8_t f = O.flags[c];
  if (!(f & F_ACT)) return;
  const int i = x - GX0, j = y - A.j_off, lm2 = A.lm + 2, mm2 = A.mm + 2, nlay = A.nlay;
  const size_t gp = (size_t)lm2 * mm2, k0 = (size_t)j * lm2 + i;
  const double mkn = (f & F_N) ? 1.0 : 0.0;
  unsigned any = 0;
  // rest state (pm:198-200)
  for (int l = 0; l < nlay; l++) O.hlay[(size_t)l * pl + c] = O.h_0[(size_t)l * pl + c] * mkn;
  if (A.nudg) {  // pm:843-881: eta coefficient as it is, u / v averaged to the faces where both cells are nudged
    const float *n0 = A.nudg, *n1 = A.nudg + gp, *n2 = A.nudg + 2 * gp;
    const double ne = (double)n0[k0];
    double nu = 0.0, nv = 0.0;
    if (i >= 1 && n1[k0 - 1] > 1.e-9f && n1[k0] > 1.e-9f) nu = (double)n1[k0] * 0.5 + (double)n1[k0 - 1] * 0.5;
    if (j >= 1 && n2[k0 - lm2] > 1.e-9f && n2[k0] > 1.e-9f) nv = (double)n2[k0] * 0.5 + (double)n2[k0 - lm2] * 0.5;
    O.nudg[c] = ne; O.nudg[pl + c] = nu; O.nudg[2 * pl + c] = nv;
    if (ne > 1.e-9 || nu > 1.e-9 || nv > 1.e-9) any |= 1u;
    if (ne != 0.0 || nu != 0.0 || nv != 0.0) any |= 2u;
    for (int l = 0; l < nlay; l++) O.fnud[(size_t)l * pl + c] = O.hlay[(size_t)l * pl + c];  // targets default to the rest state
  }
  if (A.init) {  // pm:882-910: interface elevations -> thickness targets (and the initial state)
    for (int l = 0; l < nlay; l++) {
      const size_t k = (size_t)l * pl + c;
      double t = O.hlay[k] + (double)A.init[((size_t)0 * nlay + l) * gp + k0];
      if (l < nlay - 1) t = t - (double)A.init[((size_t)0 * nlay + l + 1) * gp + k0];
      const double fn = t * mkn;
      const double fu = (double)A.init[((size_t)1 * nlay + l) * gp + k0], fv = (double)A.init[((size_t)2 * nlay + l) * gp + k0];
      if (O.fnud) { O.fnud[k] = fn; O.fnud[(size_t)nlay * pl + k] = fu; O.fnud[(size_t)2 * nlay * pl + k] = fv; }
      if (O.set_state) { O.u[k] = fu; O.v[k] = fv; }
      // (hlay of layer l + 1 is still the rest state when it is read above: the layers ascend)
      if (O.set_state) O.hlay[k] = fn * mkn;
    }
  }
  if (O.taus) {  // pm:920-931 (default: the uniform tauw, pm:302-303)
    double tx = A.tauw[0], ty = A.tauw[1];
    if (A.taus) { tx = (double)A.taus[k0]; ty = (double)A.taus[gp + k0]; }
    O.taus[c] = tx; O.taus[pl + c] = ty;
    if (fabs(tx) > 1.e-7 || fabs(ty) > 1.e-7) any |= 4u;
  }
  if (A.fcor) {  // pm:932-950: psi-point average taken in float32
    if (i > 0 && j > 0) {
      float t = A.fcor[k0] * 0.25f;
      t = t + A.fcor[k0 - 1] * 0.25f;
      t = t + A.fcor[k0 - lm2] * 0.25f;
      t = t + A.fcor[k0 - lm2 - 1] * 0.25f;
      O.fcor[c] = (double)t;
    } else {
      O.fcor[c] = (double)A.fcor[k0];
    }
  } else {
    O.fcor[c] = A.f0;
  }
  if (A.hdot) {  // pm:911-919
    for (int l = 0; l < nlay; l++) {
      const double hd = (double)A.hdot[(size_t)l * gp + k0];
      O.hdot[(size_t)l * pl + c] = hd;
      if (hd != 0.0) any |= 8u;
    }
  }
  if (A.tide && O.tide) {  // pm:951-964: tide(2, 1, 0:lm+1, 0:mm+1, 3), amplitude and phase de-interleaved into planes
    for (int c3 = 0; c3 < 3; c3++)
      for (int a = 0; a < 2; a++) O.tide[((size_t)c3 * 2 + a) * pl + c] = (double)A.tide[((size_t)c3 * gp + k0) * 2 + a];
  }
  if (any) atomicOr(O.any, any);
}

// h_0.bin's content (float32, [nlay][ndeg], pm:185-194) and the grid coordinates of the vector points, from the dense planes
__global__ void k_gi_h0r4(const double *__restrict__ h_0, size_t plane, const int *__restrict__ cell, int p0, int n, int ndeg, int nlay,
                          float *__restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int cc = cell[p0 + k];
  for (int l = 0; l < nlay; l++) out[(size_t)l * ndeg + (p0 + k - 1)] = cc >= 0 ? (float)h_0[(size_t)l * plane + cc] : 0.0f;
}
__global__ void k_gi_subc(const int *__restrict__ cell, int p0, int n, int NX, int j_off, int *__restrict__ si, int *__restrict__ sj) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int cc = cell[p0 + k];
  si[k] = cc >= 0 ? cc % NX - GX0 : 0;
  sj[k] = cc >= 0 ? cc / NX - j_off : 0;
}

// grid.bin's five records (pm:732-749): posc = i + 1 + j (lm + 2) and the masks mk_n, mk_u, mk_v, mkpi as integers, per vector point
__global__ void k_gi_grid_record(const int *__restrict__ cell, const uint8_t *__restrict__ flags, int p0, int n, int NX, int j_off, int lm,
                                 int *__restrict__ out /* [5][n] */) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int cc = cell[p0 + k];
  const uint8_t f = cc >= 0 ? flags[cc] : 0;
  const int i = cc >= 0 ? cc % NX - GX0 : 0, j = cc >= 0 ? cc / NX - j_off : 0;
  out[k] = i + 1 + j * (lm + 2);
  out[n + k] = (f & F_N) ? 1 : 0;
  out[2 * n + k] = (f & F_U) ? 1 : 0;
  out[3 * n + k] = (f & F_V) ? 1 : 0;
  out[4 * n + k] = (f & F_PI) ? 1 : 0;
}

}  // namespace beom
#endif
