// diag.cuh -- output-side kernels: the diagnostic records of write_array (float32 'pvor', 'mont', 'v_cc';
// private_mod.f95:2884-2974) and the conservation integrals of testcases/conservation.m:116-211.
// Same conventions as split.cuh: dense planes, E/W = +-1, N/S = +-NX, expressions in the reference's order
// (here with its '/ dl' divisions and its float32 roundings, so the records are bit-identical to the oracle's).
#ifndef BEOM_DIAG_CUH
#define BEOM_DIAG_CUH
#include "dev.cuh"

namespace beom {

#define BEOM_CELL_PAD(D)                                                    \
  const int x = (D).x_lo - 1 + blockIdx.x * blockDim.x + threadIdx.x;       \
  const int y = (D).y_lo - 1 + blockIdx.y * blockDim.y + threadIdx.y;       \
  if (x > (D).x_hi + 1 || y > (D).y_hi + 1) return;                         \
  const int NX = (D).NX;                                                    \
  const size_t c = (size_t)y * NX + x;                                      \
  const uint8_t f = (D).flags[c];

// wrk1 (relative vorticity) and wrk2 (divergence) of the 'v_cc' record, pm:2893-2905; 0 outside the vector points
__global__ void k_rec_vort_dive(const __grid_constant__ Dev D, double *__restrict__ w1, double *__restrict__ w2) {
  BEOM_CELL_PAD(D)
  const size_t L = (size_t)blockIdx.z * D.plane;
  double a = 0.0, b = 0.0;
  if (f & F_ACT) {
    const double *u = D.u + L, *v = D.v + L;
    a = ((v[c] - v[c - 1]) / D.dl - (u[c] - u[c - NX]) / D.dl) * m_pe(f);
    b = (u[c + 1] - u[c]) / D.dl + (v[c + NX] - v[c]) / D.dl;
  }
  w1[L + c] = a;
  w2[L + c] = b;
}
// 'v_cc' record, pm:2907-2926
__global__ void k_rec_vcc(const __grid_constant__ Dev D, const double *__restrict__ w1, const double *__restrict__ w2, float *__restrict__ out) {
  BEOM_CELL(D)
  const size_t L = (size_t)blockIdx.z * D.plane;
  const double *p = w1 + L, *q = w2 + L;
  const double a = p[c + 1] - p[c], b = p[c + NX + 1] - p[c + NX], cc = p[c + NX] - p[c], d = p[c + NX + 1] - p[c + 1];
  const double e = q[c + 1] - q[c], g2 = q[c] - q[c - 1], h = q[c + NX] - q[c], k = q[c] - q[c - NX];
  out[L + c] = (float)(D.bvis + D.dvis * (D.dl * D.dl) * sqrt(a * a + b * b + cc * cc + d * d + e * e + g2 * g2 + h * h + k * k));
}
// 'mont' record, pm:2930-2950 (float32 accumulation of the baroclinic terms)
__global__ void k_rec_mont(const __grid_constant__ Dev D, float *__restrict__ out) {
  BEOM_CELL(D)
  const int l = blockIdx.z;
  const double mk = m_n(f);
  const double hl = D.hlay[(size_t)l * D.plane + c];
  float r = (float)(-D.ocrp / (double)(D.nsal - 1) * D.hsal * mk * cube(D.hsal / (D.hmin * (1.0 - mk) + hl)));
  double s = 0.0;
  for (int i = 0; i < l; i++) r = r - (float)((D.rhon[l] - D.rhon[i]) * D.hlay[(size_t)i * D.plane + c] / D.rhon[l]);
  for (int i = 0; i < D.nlay; i++) s += D.hlay[(size_t)i * D.plane + c];
  r = r + (float)(s - D.h_th[c]);
  out[(size_t)l * D.plane + c] = r;
}
// 'pvor' record, pm:2951-2974
__global__ void k_rec_pvor(const __grid_constant__ Dev D, float *__restrict__ out) {
  BEOM_CELL(D)
  const size_t L = (size_t)blockIdx.z * D.plane;
  const double *u = D.u + L, *v = D.v + L, *h = D.hlay + L;
  const double zeta = ((v[c] - v[c - 1]) / D.dl - (u[c] - u[c - NX]) / D.dl) * m_pe(f);
  const double msum = m_n(f) + m_n(D.flags[c - 1]) + m_n(D.flags[c - NX]) + m_n(D.flags[c - NX - 1]);
  out[L + c] = (float)((D.fcor[c] + zeta * D.uadv) * m_pi(f) * msum / (h[c] + h[c - 1] + h[c - NX - 1] + h[c - NX]));
}

// ---- the state records of write_array (private_mod.f95:2848-2883) straight in the reference's record layout ----
// One thread per vector point p = p0 + k of this rank: eta_ = layer thickness minus the float32 rest thickness h_0.bin,
// cumulated upward from the bottom layer IN FLOAT32 exactly as the reference does (pm:2848-2870; under the rigid lid the
// top record is pi_s, pm:2871-2875), u___ and v___ rounded to float32.  rec = [3][nlay][n] floats (eta, u, v).
// The same pass takes the min / max thickness of every layer over the wet points (the report and the `hlay < hmin / 2'
// halt of write_outputs, pm:2772-2808): one partial pair per block and layer, finished by k_rec_minmax.
__global__ void k_rec_state(const __grid_constant__ Dev D, const float *__restrict__ h0r4, const int *__restrict__ cell, int p0, int n,
                            int ndeg, float *__restrict__ rec, double *__restrict__ mm_partial) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int nlay = D.nlay;
  const int cc = k < n ? cell[p0 + k] : -1;
  float *eta = rec, *ru = rec + (size_t)nlay * n, *rv = rec + (size_t)2 * nlay * n;
  const bool wet = cc >= 0 && (D.flags[cc] & F_N);
  __shared__ double sh[2][8];
  float acc = 0.0f;
  for (int l = nlay - 1; l >= 0; l--) {
    double h = 0.0;
    if (k < n) {
      float e = 0.0f, fu = 0.0f, fv = 0.0f;
      if (cc >= 0) {
        const size_t c = (size_t)l * D.plane + cc;
        h = D.hlay[c];
        const double dh = h - (double)h0r4[(size_t)l * ndeg + (p0 + k - 1)];
        acc = (l == nlay - 1) ? (float)dh : (float)(dh + (double)acc);
        e = acc;
        if (l == 0 && D.rgld > 0.5) e = (float)D.pi_s[cc];
        fu = (float)D.u[c];
        fv = (float)D.v[c];
      }
      eta[(size_t)l * n + k] = e;
      ru[(size_t)l * n + k] = fu;
      rv[(size_t)l * n + k] = fv;
    }
    double lo = wet ? h : INFINITY, hi = wet ? h : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { sh[0][w] = lo; sh[1][w] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < (int)(blockDim.x >> 5); i++) { lo = fmin(lo, sh[0][i]); hi = fmax(hi, sh[1][i]); }
      mm_partial[((size_t)l * gridDim.x + blockIdx.x) * 2 + 0] = lo;
      mm_partial[((size_t)l * gridDim.x + blockIdx.x) * 2 + 1] = hi;
    }
  }
}
__global__ void k_rec_minmax(const double *__restrict__ mm_partial, int per_layer, double *__restrict__ out) {  // one block (32 lanes) per layer
  const int l = blockIdx.x, lane = threadIdx.x;
  double lo = INFINITY, hi = -INFINITY;
  for (int b = lane; b < per_layer; b += 32) {
    lo = fmin(lo, mm_partial[((size_t)l * per_layer + b) * 2 + 0]);
    hi = fmax(hi, mm_partial[((size_t)l * per_layer + b) * 2 + 1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { out[2 * l] = lo; out[2 * l + 1] = hi; }
}

// ---- conservation integrals (testcases/conservation.m:116-211) over the vector points whose rows this rank owns ----
// per layer l: q[0] = sum over wet points of hlay, q[1] = 0.5 sum(0.5 (U + U(E))) + 0.5 sum(0.5 (V + V(N))) with
// U = u^2 * 0.5 (h(W) + h), V = v^2 * 0.5 (h(S) + h) (dry or missing thickness counts as 0); layer 0 also q[2] =
// sum over wet points of eta_1^2, eta_1 = sum over layers of (hlay - h_0); and, conservation.m:169-211,
// q[3] = sum of 0.5 pvor^2 hatp (potential enstrophy), q[4] = sum of zeta = pvor hatp - fcor (relative vorticity),
// q[5] = sum of zeta^2, with pvor as write_array computes it (private_mod.f95:2951-2974, here kept in double) and hatp the mean
// thickness of the cells around the psi point that hold a value (vector points and periodic images).
// One thread per vector point (frozen periodic duplicates have no cell and are skipped, as conservation.m:196-201
// discards them); neighbours are the dense cells around it, so periodic aliases are honoured.  Warp-shuffle tree + one
// partial per block; the partials are added in block order by k_sum_partials, so the result does not depend on scheduling.
constexpr int kNQ = 7;  // + q[6] = number of points counted
__device__ __forceinline__ double wet_h(const Dev &D, const double *h, size_t c) { return (D.flags[c] & F_N) ? h[c] : 0.0; }
__global__ void k_conservation(const __grid_constant__ Dev D, const double *__restrict__ h_0, const int *__restrict__ cell, int p0, int n,
                               double *__restrict__ partial) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y, NX = D.NX;
  double q[kNQ];
#pragma unroll
  for (int k = 0; k < kNQ; k++) q[k] = 0.0;
  const int cc = t < n ? cell[p0 + t] : -1;
  if (cc >= 0 && cc / NX >= D.y_lo && cc / NX <= D.y_hi) {
    const size_t c = (size_t)cc, L = (size_t)l * D.plane;
    const double *h = D.hlay + L, *u = D.u + L, *v = D.v + L;
    const uint8_t f = D.flags[c];
    const bool wet = f & F_N;
    if (wet) q[0] = h[c];
    q[6] = 1.0;
    const double hc = wet_h(D, h, c), hW = wet_h(D, h, c - 1), hE = wet_h(D, h, c + 1), hS = wet_h(D, h, c - NX), hN = wet_h(D, h, c + NX);
    const double U0 = u[c] * u[c] * (0.5 * (hW + hc)), U1 = u[c + 1] * u[c + 1] * (0.5 * (hc + hE));
    const double V0 = v[c] * v[c] * (0.5 * (hS + hc)), V1 = v[c + NX] * v[c + NX] * (0.5 * (hc + hN));
    q[1] = 0.5 * (0.5 * (U0 + U1)) + 0.5 * (0.5 * (V0 + V1));
    if (l == 0 && wet) {
      double eta = 0.0;
      for (int i = D.nlay - 1; i >= 0; i--) eta += D.hlay[(size_t)i * D.plane + c] - h_0[(size_t)i * D.plane + c];
      q[2] = eta * eta;
    }
    {
      const uint8_t fW = D.flags[c - 1], fS = D.flags[c - NX], fSW = D.flags[c - NX - 1];
      const double zr = ((v[c] - v[c - 1]) / D.dl - (u[c] - u[c - NX]) / D.dl) * m_pe(f);
      const double msum = m_n(f) + m_n(fW) + m_n(fS) + m_n(fSW);
      const double pv = (D.fcor[c] + zr * D.uadv) * m_pi(f) * msum / (h[c] + h[c - 1] + h[c - NX - 1] + h[c - NX]);
      const uint8_t val = F_ACT | F_GHOST;  // cells that hold a value (get_field leaves NaN elsewhere)
      double cnt = 1.0 + ((fW & val) ? 1.0 : 0.0) + ((fS & val) ? 1.0 : 0.0) + ((fSW & val) ? 1.0 : 0.0);
      const double hsum = h[c] + ((fW & val) ? h[c - 1] : 0.0) + ((fS & val) ? h[c - NX] : 0.0) + ((fSW & val) ? h[c - NX - 1] : 0.0);
      const double hatp = hsum / cnt;
      const double z = pv * hatp - D.fcor[c];
      if (pv == pv) {  // (a psi point with no thickness around it: 0/0, which nanmean skips)
        q[3] = pv * pv * hatp * 0.5;
        q[4] = z;
        q[5] = z * z;
      }
    }
  }
  __shared__ double sh[kNQ][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < kNQ; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    if (lane == 0) sh[k][w] = q[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < kNQ; k++) {
      double s = lane < nw ? sh[k][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) partial[kNQ * ((size_t)l * gridDim.x + blockIdx.x) + k] = s;
    }
  }
}
// one block per layer: adds the layer's partials in a fixed order (lane-strided, then a shuffle tree)
__global__ void k_sum_partials(const double *__restrict__ partial, int per_layer, double *__restrict__ out) {
  const int l = blockIdx.x, lane = threadIdx.x;
  double q[kNQ];
  for (int k = 0; k < kNQ; k++) q[k] = 0.0;
  for (int b = lane; b < per_layer; b += 32)
    for (int k = 0; k < kNQ; k++) q[k] += partial[kNQ * ((size_t)l * per_layer + b) + k];
  for (int k = 0; k < kNQ; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    if (lane == 0) out[kNQ * l + k] = q[k];
  }
}

}  // namespace beom
#endif
