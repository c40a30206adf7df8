// rigid.cuh -- surf_pressure's iteration (private_mod.f95:1756-1803) on every SM, with the reference's iterates.
//
// The reference sweeps the points in vector order (row by row, west to east) and updates pi_s in place: lexicographic
// Gauss-Seidel (PSOR with rp = 1), each point using the NEW west / south values and the OLD east / north values, repeated
// until the largest change of a sweep is <= pi_tol or 1000 sweeps are done.  A point (x, y) of sweep k therefore depends
// on (x-1, y, k), (x, y-1, k), (x+1, y, k-1) and (x, y+1, k-1) only, which leaves two kinds of parallelism that do not
// change a single bit of the iterates:
//   * inside a sweep, the anti-diagonals x + y = const (k_surf_pressure in split.cuh walks them with one thread block);
//   * ACROSS sweeps: sweep k+1 may start in the south-west corner as soon as sweep k has moved two tiles on.
// This file does both.  The domain is cut into tiles of kPiBx columns x 32 rows; a warp owns a tile for a sweep (lane =
// row, marching east: the south value arrives by shuffle from the lane below, the west value is the lane's own previous
// result), and tile (I, J) of sweep k waits -- on a per-tile counter of completed sweeps in global memory -- for
// (I-1, J, k), (I, J-1, k), (I+1, J, k-1), (I, J+1, k-1).  Sweep k writes into array X[k mod kPiNA] and reads its old
// values from X[(k-1) mod kPiNA], so up to kPiNA - 1 sweeps are in flight at once and the result of any of them is still
// intact when its convergence verdict (the max-norm of its change, an atomicMax over its tiles) comes in: the first
// sweep K with maxdiff <= pi_tol (or K = maxiters) is final, exactly as in the reference, and the sweeps begun beyond it
// are abandoned.  Workers (warps) are persistent and visit their tiles in the order (k, J, I), a linear extension of
// the dependency order, so the smallest unfinished tile is always runnable: no deadlock as long as all workers are
// resident (the launch is sized by the occupancy query).
#ifndef BEOM_RIGID_CUH
#define BEOM_RIGID_CUH
#include "dev.cuh"

namespace beom {

constexpr int kPiBx = 64;  // tile width
constexpr int kPiBy = 32;  // tile height = lanes
constexpr int kPiNA = 16;  // rotating pi_s arrays (sweeps in flight + 1)
constexpr int kPiDoneBorder = 1 << 30;

struct PiSolve {
  double *X[kPiNA];                      // X[0] = the model's pi_s plane
  double *c0;                            // rp * Osum_ * pi_rhs, rebuilt every step
  const double *cE, *cN, *cW, *cS;       // rp * Osum_ * Ow(E), Os(N), Ow, Os: static
  int *done;                             // [(TJ + 2) x (TI + 2)] sweeps completed per tile; border entries = kPiDoneBorder
  int *count;                            // [kPiNA] tiles that have finished the sweep using the slot
  unsigned long long *maxbits;           // [kPiNA] max |change| of that sweep (bits of a non-negative double)
  int *decided;                          // sweeps 1 .. *decided are complete and were not final
  int *final_sweep;                      // 0 until the final sweep K is known
  double *md;                            // [kPiNA] y-slab runs: this rank's max |change| per sweep slot (then reduced over the ranks)
  int TI, TJ, maxiters;
  double tol;
};

// static coefficient planes (once at init): cE = (rp * Osum_) * Ow(E) etc., the products the reference forms first in
// `rp * Osum_(ipnt) * Ow(c__1) * pi_s(c__1)' (pm:1772-1787); 0 outside the vector points
__global__ void k_pi_coeff(const __grid_constant__ Dev D, double *__restrict__ cE, double *__restrict__ cN, double *__restrict__ cW,
                           double *__restrict__ cS) {
  const int x = D.x_lo + blockIdx.x * blockDim.x + threadIdx.x, y = D.y_lo + blockIdx.y * blockDim.y + threadIdx.y;
  if (x > D.x_hi || y > D.y_hi) return;
  const int NX = D.NX;
  const size_t c = (size_t)y * NX + x;
  if (!(D.flags[c] & F_ACT)) return;
  const double rp = 1.0, os = D.Osum_[c];
  cE[c] = rp * os * D.Ow[c + 1];
  cN[c] = rp * os * D.Os[c + NX];
  cW[c] = rp * os * D.Ow[c];
  cS[c] = rp * os * D.Os[c];
}
// per step: c0 = rp * Osum_ * pi_rhs (after k_pi_rhs), and the bookkeeping of a new solve
__global__ void k_pi_begin(const __grid_constant__ Dev D, const __grid_constant__ PiSolve S) {
  const int x = D.x_lo + blockIdx.x * blockDim.x + threadIdx.x, y = D.y_lo + blockIdx.y * blockDim.y + threadIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    const int t = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    const int DW = S.TI + 2, n = DW * (S.TJ + 2);
    for (int i = t; i < n; i += nt) {
      const int I = i % DW, J = i / DW;
      S.done[i] = (I == 0 || J == 0 || I == DW - 1 || J == S.TJ + 1) ? kPiDoneBorder : 0;
    }
    for (int i = t; i < kPiNA; i += nt) { S.count[i] = 0; S.maxbits[i] = 0ull; S.md[i] = 0.0; }
    if (t == 0) { *S.decided = 0; *S.final_sweep = 0; }
  }
  if (x > D.x_hi || y > D.y_hi) return;
  const size_t c = (size_t)y * D.NX + x;
  if (!(D.flags[c] & F_ACT)) return;
  const double rp = 1.0;
  S.c0[c] = rp * D.Osum_[c] * D.pi_rhs[c];
}

__device__ __forceinline__ int ld_volatile(const int *p) { return *reinterpret_cast<const volatile int *>(p); }

// One launch runs sweeps k_first .. k_last.  One rank: 1 .. maxiters with the verdicts taken inside the kernel.  y-slabs
// (multi != 0): ONE sweep per launch -- the south row of this sweep and the north row of the previous one are the
// neighbouring ranks' and arrive between launches (beom_gpu.cu, pi_solve_slabs); the sweep's max |change| goes to S.md.
__global__ void __launch_bounds__(1024) k_pi_wave(const __grid_constant__ Dev D, const __grid_constant__ PiSolve S, int k_first, int k_last, int multi) {
  const int lane = threadIdx.x & 31;
  const int worker = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nworkers = (int)((gridDim.x * blockDim.x) >> 5);
  const int ntiles = S.TI * S.TJ, NX = D.NX, DW = S.TI + 2;
  const double rp = 1.0;
  for (int k = k_first; k <= k_last; k++) {
    double *__restrict__ Xn = S.X[k % kPiNA];
    const double *__restrict__ Xo = S.X[(k - 1) % kPiNA];
    for (int t = worker; t < ntiles; t += nworkers) {
      const int I = t % S.TI, J = t / S.TI;
      int *dn = S.done + (J + 1) * DW + (I + 1);
      int fin = 0;
      if (lane == 0) {
        for (;;) {
          if (!multi) fin = ld_volatile(S.final_sweep);
          if (fin) break;
          if (ld_volatile(dn - 1) >= k && ld_volatile(dn - DW) >= k && ld_volatile(dn + 1) >= k - 1 && ld_volatile(dn + DW) >= k - 1 &&
              (multi || ld_volatile(S.decided) >= k - kPiNA))
            break;
          __nanosleep(40);
        }
        __threadfence();  // acquire: the neighbours' values below are read after their counters
      }
      fin = __shfl_sync(0xffffffffu, fin, 0);
      if (fin) return;
      const int x0 = D.x_lo + I * kPiBx, y = D.y_lo + J * kPiBy + lane;
      const int sj = y - D.j_off;
      const bool row_ok = y <= D.y_hi;
      const size_t crow = (size_t)y * NX;
      double vlast = 0.0, lmax = 0.0;
      for (int s = 0; s < kPiBx + kPiBy - 1; s++) {
        const int rx = s - lane, x = x0 + rx;
        double south = __shfl_up_sync(0xffffffffu, vlast, 1);  // (x, y-1) of this sweep: the lane below, one step ago
        const bool on = row_ok && rx >= 0 && rx < kPiBx && x <= D.x_hi;
        double v = 0.0;
        if (on) {
          const size_t c = crow + x;
          if (D.flags[c] & F_ACT) {
            const int si = x - D.i_off;
            if (lane == 0) south = __ldcg(Xn + c - NX);              // the tile below, this sweep
            const double west = rx == 0 ? __ldcg(Xn + c - 1) : vlast;  // the tile to the west, this sweep / own previous result
            const double prev = __ldcg(Xo + c);
            v = (1 - rp) * prev - S.c0[c];
            if (si < D.lm) v = v + __ldg(S.cE + c) * __ldcg(Xo + c + 1);
            if (sj < D.mm) v = v + __ldg(S.cN + c) * __ldcg(Xo + c + NX);
            if (si > 1) v = v + __ldg(S.cW + c) * west;
            if (sj > 1) v = v + __ldg(S.cS + c) * south;
            Xn[c] = v;
            lmax = fmax(lmax, fabs(v - prev));
          }
        }
        if (rx >= 0) vlast = v;  // cells that are not vector points hold 0 in every array
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
      __threadfence();  // release: the tile's values before its counter
      __syncwarp();
      if (lane == 0) {
        const int slot = k % kPiNA;
        atomicMax(S.maxbits + slot, (unsigned long long)__double_as_longlong(lmax));
        __threadfence();
        *reinterpret_cast<volatile int *>(dn) = k;
        if (atomicAdd(S.count + slot, 1) == ntiles - 1) {  // this was the last tile of sweep k: its verdict (pm:1756, 1795-1802)
          const double m = __longlong_as_double((long long)atomicExch(S.maxbits + slot, 0ull));
          S.count[slot] = 0;
          __threadfence();
          if (multi) S.md[slot] = m;
          else if (!(m > S.tol) || k >= S.maxiters) *reinterpret_cast<volatile int *>(S.final_sweep) = k;
          else *reinterpret_cast<volatile int *>(S.decided) = k;
        }
      }
    }
  }
}

// y-slabs: md = max(md, what the two neighbours hold) -- nranks - 1 rounds of this after a neighbour exchange make every rank
// hold the maxima over all ranks (the ranks form a line; an allreduce that needs nothing but the halo exchange)
__global__ void k_pi_max_merge(double *__restrict__ md, const double *__restrict__ from_lo, const double *__restrict__ from_hi, int has_lo, int has_hi) {
  const int i = threadIdx.x;
  if (i >= kPiNA) return;
  double m = md[i];
  if (has_lo) m = fmax(m, from_lo[i]);
  if (has_hi) m = fmax(m, from_hi[i]);
  md[i] = m;
}
// y-slabs: the verdict on sweeps b0 .. b1 from their reduced max-norms (pm:1756, 1795-1802); the slots are cleared for the next batch
__global__ void k_pi_verdict(const __grid_constant__ PiSolve S, int b0, int b1) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int fin = 0;
  for (int k = b0; k <= b1 && !fin; k++)
    if (!(S.md[k % kPiNA] > S.tol) || k >= S.maxiters) fin = k;
  for (int i = 0; i < kPiNA; i++) S.md[i] = 0.0;
  *S.final_sweep = fin;
}

// the final sweep's array becomes pi_s (X[0]); iters_out (optional) = number of sweeps, as the reference counts them
__global__ void k_pi_select(const __grid_constant__ Dev D, const __grid_constant__ PiSolve S, int *iters_out) {
  const int x = D.x_lo + blockIdx.x * blockDim.x + threadIdx.x, y = D.y_lo + blockIdx.y * blockDim.y + threadIdx.y;
  const int K = *S.final_sweep;
  if (iters_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) *iters_out = K;
  if (x > D.x_hi || y > D.y_hi || K % kPiNA == 0) return;
  const size_t c = (size_t)y * D.NX + x;
  if (D.flags[c] & F_ACT) S.X[0][c] = S.X[K % kPiNA][c];  // (y-slabs: the halo rows of pi_s are refreshed by the caller)
}

}  // namespace beom
#endif
