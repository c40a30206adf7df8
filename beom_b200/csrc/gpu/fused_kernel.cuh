// fused_kernel.cuh -- the fused single-pass time step (gener_forward_backward, private_mod.f95:2225-2316, and, with
// gene = 0 after the caller's centred-flux rebuild, first_three_timesteps, :2151-2223).
//
// One kernel per step.  It reads every persistent field once (hlay,u,v,h_u,h_v, 2 rs_h, 3 dmdx,
// 3 dmdy: 13 doubles per cell-layer) and writes the 8 new ones (hlay,u,v,h_u,h_v + newest rs_h, dmdx,
// dmdy): the 168 algorithmic bytes of SURVEY.md 8(d).  Everything the reference keeps in 2-D scratch
// arrays between its loops (mont, rvor, pvor, dive, d2hx, d2hy, v_cc, v_ll; private_mod.f95:48-61)
// lives in registers here.
//
// Mapping.  A warp owns 32 consecutive columns of ONE layer and marches north (row by row) through
// its y-chunk; lane k is column x0-2+k, lanes 2..29 produce results, lanes 0,1,30,31 are the x halo
// (recomputed by the neighbouring warp).  Raw inputs are staged by TMA bulk copies into a shared-memory
// ring per LAYER, shared by the column-group warps of that layer (one copy per stream and row for the whole
// CTA width; 4 rows deep for h_u,h_v,u,v, which are read at three row lags, 2 rows for the rest).  The new
// thickness, the Montgomery potential, P and F (below) live in small shared-memory rings as well, so their
// east/west and south neighbours are plain shared-memory loads; the other computed values travel by warp
// shuffles (E/W) and in registers (S): a 3-row software pipeline
//     row R   : update_h, rvor, dive, d2hx, pvor, Montgomery/Bernoulli potential          (front)
//     row R-1 : d2hy, Leith v_cc / v_ll, and the first momentum component if it is v
//     row R-2 : u (and v when u goes first)
// The carried state is kept in 4-entry rings indexed by (phase - age) & 3.  The row loop is unrolled four times
// with the phase a compile-time constant, so the rings are plain registers that are never moved and (in the LEAN
// instantiation, whose stream slots and column-group count are compile-time too) every shared-memory address is an
// immediate offset.
//
// What is carried is chosen so that no product or squared difference is evaluated twice (each is the
// SAME IEEE operation on the SAME operands the reference performs, so results stay bit-identical):
//   A1,A3,B1,B3  the squared vorticity / divergence differences of the Leith stencils (pm:2477-2502);
//                their x-shifted copies are shuffled instead of recomputed
//   P = v_cc*dive, Q = v_ll*rvor   the viscous flux products of update_u / update_v (pm:1474-1477, 1560-1563)
//   q = 0.25*pvor,  T = q*(h_v + h_v(W))  /  W = q*(h_u + h_u(S))   the Coriolis terms (pm:1461-1462, 1546-1547)
//   F = 0.16667*d2hx, Gy = 0.16667*d2hy                             (pm:1493-1495, 1579-1581)
// The upstream flux 0.5(w+|w|)(hc-Ff) + 0.5(w-|w|)(hc-Fh) is evaluated as w*(hc - (w>0 ? Ff : Fh)),
// which is the same number (one of the two products is an exact zero).
//
// A CTA is (column groups) x (layers) warps; the layer-warps of a column group exchange the new layer
// thickness of the front row through shared memory for the column sums of the Montgomery potential
// (private_mod.f95:2357-2373) with a split-phase mbarrier.  The warps that share a scheduler (warp id mod 4) belong to
// different layers AND different column groups, so they never wait at the same barrier.  State is double buffered
// (in -> out), so halo lanes and neighbouring CTAs always read time level n.
//
// Masks.  F_ACT cells are evaluated and stored; F_GHOST cells (periodic images, dev.cuh) are evaluated like the
// cell they mirror -- the halo is recomputed, not re-read -- but never stored.  Rows whose three rows in flight are
// open water on all 32 columns (a bitmap built at init) take a copy of the row code without any select.
//
// Arithmetic: the same expressions in the same order as split.cuh / the reference (-fmad=false), so
// the two paths agree bit for bit; 0/1 masks are applied as selects (x*1 = x, x*0 = +-0).
#ifndef BEOM_FUSED_KERNEL_CUH
#define BEOM_FUSED_KERNEL_CUH
#include <type_traits>

#include "dev.cuh"

namespace beom {
namespace fusedk {

constexpr int kHalo = 2;              // halo lanes on each side of a warp
constexpr int kUse = 32 - 2 * kHalo;  // 28 result columns per warp
constexpr int kMaxLay = 8;            // layers per CTA (shared-memory exchange, warps per CTA)
#ifndef BEOM_TRYWAIT_HINT
#define BEOM_TRYWAIT_HINT 20000   // suspend-time hint (ns) of mbarrier.try_wait (the warp sleeps in hardware instead of spinning); 0 = none
#endif
#ifndef BEOM_FUSED_WARPS
#define BEOM_FUSED_WARPS 16
#endif
#ifndef BEOM_LEAN4_GROUPS
#define BEOM_LEAN4_GROUPS (BEOM_FUSED_WARPS / 4)  // column groups of the 4-layer bare step (experiment: 3 = 12 warps at 168 registers)
#endif
constexpr int kMaxWarps = BEOM_FUSED_WARPS;  // warps per CTA: 16 -> <= 128 registers per thread, 12 -> <= 168
constexpr int kPad = 4;               // staged columns on each side of a CTA's result columns (16-byte aligned rows)
__host__ __device__ constexpr int seg_doubles(int groups) { return groups * kUse + 2 * kPad; }  // one staged row segment
#ifndef BEOM_WROW
#define BEOM_WROW 34
#endif
#ifndef BEOM_OWORDS
#define BEOM_OWORDS 32
#endif
#ifndef BEOM_MIN_CTAS
#define BEOM_MIN_CTAS 1   // experiment: 2 = two CTAs of <= 256 threads per SM (needs <= 113 KB of shared memory per CTA)
#endif
constexpr int kWRow = BEOM_WROW;      // per-warp state ring: 32 lanes + 1 pad column on the west side (33; 34 = padded on both)
constexpr int kOWords = BEOM_OWORDS;  // words of a chunk's open-water bitmap kept in shared memory (one bit per 4-row group)
constexpr int kWRings = 3;            // mo, P, F
#ifndef BEOM_L2_AHEAD
#define BEOM_L2_AHEAD 0   // rows of L2 prefetch ahead of the staging copies (0 = none)
#endif

// ---- raw-input streams: one 36-double row segment per (field, row), staged by TMA bulk copies ----
enum {
  S_HU, S_HV, S_U, S_V,  // 4 rows deep
  S_HL, S_R1, S_R2, S_DX1, S_DX2, S_DX3, S_DY1, S_DY2, S_DY3, S_FCOR, S_HTH,  // 2 rows deep, always present
  kMandatory,
  S_HDOT = kMandatory, S_FNN, S_NUDN, S_FNU, S_NUDU, S_FNV, S_NUDV, S_TBX, S_TUX, S_TBY, S_TUY,  // general extras
  S_TTXU, S_TTYV, S_TTYVS, S_TTYU, S_TTXV, S_TTXVS,  // wind streams come last (..S = south neighbour row)
  S_COUNT
};
static_assert(S_COUNT <= 32, "one lane per stream");

struct StreamTab {
  int n;                       // enabled streams
  int n_nowind;                // enabled streams for a layer that receives no wind stress
  signed char slot[S_COUNT];   // stream -> compact slot (-1 = disabled)
  const double *base[S_COUNT];  // by slot
  short lag[S_COUNT];          // row = R - lag
  unsigned char lstride[S_COUNT];  // layer stride in planes (0 = 2-D field, 1, or 2 for [nlay][2] arrays)
};

__host__ __device__ constexpr int ring_segments(int nstreams) { return 16 + (nstreams - 4) * 2; }

// ---- option sets.  FEAT < 0: the GENERAL instantiation, every option a run-time (warp-uniform) switch, stream slots and
// column-group count read from the launch arguments.  FEAT >= 0: a SPECIALISED instantiation -- the option set is this
// compile-time mask, gene is exactly 1 (or exactly 0 in the G0 copy for the start-up steps), there is no hdot, top drag or
// body force; the stream slots below and the column-group count are compile-time constants, so every shared-memory
// address of the row code is an immediate.  FEAT == 0 is the bare step of the benchmark workload ("lean").
constexpr int FB_NUDG = 1;  // sponge: relaxation of hlay, u, v towards fnud (nudg.bin)
constexpr int FB_OCRP = 2;  // outcropping: Salmon's term in the Montgomery potential, curvature limiter
constexpr int FB_BDRG = 4;  // bottom drag (tb3d from distribute_stress)
// slot of a stream in make_streams' order (fused.cu) for a specialised option set
__host__ __device__ constexpr int spec_slot(int feat, int stream) {
  const int nudg = (feat & FB_NUDG) ? 6 : 0, bdrg = (feat & FB_BDRG) ? 2 : 0;
  if (stream < kMandatory) return stream;
  if (stream >= S_FNN && stream <= S_NUDV) return kMandatory + (stream - S_FNN);
  if (stream == S_TBX || stream == S_TBY) return kMandatory + nudg + (stream - S_TBX) / 2;  // (S_TBX, S_TUX, S_TBY, S_TUY are interleaved in the enum)
  if (stream >= S_TTXU && stream <= S_TTYVS) return kMandatory + nudg + bdrg + (stream - S_TTXU);
  if (stream >= S_TTYU) return kMandatory + nudg + bdrg + 3 + (stream - S_TTYU);
  return -1;
}

__device__ __forceinline__ double sel(bool p, double a) { return p ? a : 0.0; }
__device__ __forceinline__ double shup(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }     // value of lane-1 (west)
__device__ __forceinline__ double shdn(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }   // value of lane+1 (east)
__device__ __forceinline__ double sq(double a) { return a * a; }

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
#if BEOM_TRYWAIT_HINT > 0
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
      "r"(parity), "r"((unsigned)BEOM_TRYWAIT_HINT)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

// L2 prefetch of a row segment the CTA stages BEOM_L2_AHEAD rows later (no destination, no completion to wait for).  An
// experiment (build switch, off): measured SLOWER the further ahead it reaches (profiles/r2_l2_prefetch_ab.txt)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// One momentum update: update_u (private_mod.f95:1437-1500) when IS_U, update_v (private_mod.f95:1520-1586)
// otherwise.  c1, c2 are the two Coriolis products, Ph/Pf and Qf/Qh the viscous flux products (here / far),
// f_far / f_here the scaled thickness curvatures of the upstream flux.
struct MomX {  // optional inputs (general instantiation and wind layers)
  double tw_a = 0.0, tw_b = 0.0;  // wind stress component along the velocity at (W or S), (point)
  double te_a = 0.0, te_b = 0.0;  // cross component for the Ekman term of the sponge target
  double tb = 0.0, tu = 0.0;      // bottom / top drag at the point
  double fn = 0.0, nud = 0.0;     // sponge target and rate
  double bodf = 0.0;
};
template <bool IS_U, bool VISC, bool MASKED, int FEAT, bool G0>
__device__ __forceinline__ void momentum(const Dev &D, const bool wind, const double mask, const double hsum, const double m_far,
                                         const double m_here, const double c1, const double c2, const double old, const double h1,
                                         const double h2, const double h3, const double Ph, const double Pf, const double Qf,
                                         const double Qh, const double f_far, const double f_here, const MomX &q, double &vel,
                                         double &flux, double &dmd4) {
  constexpr bool GEN = FEAT < 0, SPEC = FEAT >= 0;
  const bool has_bdrg = GEN ? (D.has_bdrg != 0) : ((FEAT & FB_BDRG) != 0);
  const bool has_tdrg = GEN && D.has_tdrg;
  const bool has_nudg = GEN ? (D.has_nudg != 0) : ((FEAT & FB_NUDG) != 0);
  const double hcen = hsum * (MASKED ? (mask != 0.0 ? 0.5 : 1.0) : 0.5);
  dmd4 = (m_far - m_here) * D.i_dl * D.grav;
  if (MASKED) dmd4 = sel(mask != 0.0, dmd4);
  double rhsi;
  if (SPEC && !G0) {  // gene = 1 exactly: dmd4*(1-gene) is an exact zero
    rhsi = IS_U ? (c1 + c2) : ((-c1) - c2);
  } else if (SPEC) {  // gene = 0 exactly (start-up steps): dmd4*(1-gene) = dmd4
    rhsi = IS_U ? (dmd4 + c1 + c2) : (dmd4 - c1 - c2);
  } else {
    if (IS_U) rhsi = dmd4 * (1.0 - D.gene) + c1 + c2;
    else      rhsi = dmd4 * (1.0 - D.gene) - c1 - c2;
  }
  double i__h = 0.0;
  if (wind || has_bdrg || has_tdrg) i__h = 1.0 / (hcen + 1.0 - (MASKED ? mask : 1.0));
  if (wind) rhsi = rhsi + 0.5 * (q.tw_a + q.tw_b) * D.ramp * D.i_r0 * i__h;
  if (has_bdrg) rhsi = rhsi - q.tb * D.i_r0 * i__h;
  if (has_tdrg) rhsi = rhsi - q.tu * D.i_r0 * i__h;
  const double hist = D.del1 * dmd4 + D.del2 * h3 + D.gamm * h2 + D.epsi * h1;
  if (SPEC && !G0) rhsi = rhsi + hist;  // bodf = 0 and gene = 1 (checked by fused_configure)
  else if (GEN)    rhsi = rhsi + q.bodf + hist * D.gene;  // (specialised, gene = 0: + 0 + hist*0 changes nothing)
  if (VISC) {
    if (IS_U) rhsi = rhsi + (Ph - Pf) * D.i_dl - (Qf - Qh) * D.i_dl;
    else      rhsi = rhsi + (Ph - Pf) * D.i_dl + (Qf - Qh) * D.i_dl;
  }
  double w = old + (MASKED ? sel(mask != 0.0, rhsi) : rhsi) * D.dt;
  if (has_nudg) {
    double tgt = q.fn;
    if (wind) {
      if (IS_U) tgt = tgt + 0.5 * (q.te_b + q.te_a) * D.i_r1 * D.invf * i__h * D.ramp;
      else      tgt = tgt - 0.5 * (q.te_b + q.te_a) * D.i_r1 * D.invf * i__h * D.ramp;
    }
    w = tgt * q.nud + w * (1.0 - q.nud);
  }
  vel = w;
  flux = w * (hcen - (w > 0.0 ? f_far : f_here));
}

template <int V> using ic = std::integral_constant<int, V>;

// shared memory carve-up (host and device agree through these helpers)
struct SmemPlan {
  int tpad;        // columns of one thickness row: groups*32 + 2
  size_t off_bars, off_wring, off_ring, total;
  size_t seg_bytes;  // one staged row segment
  int nseg_all, nseg_nowind, wind_layers;
  // input ring of layer l (layers that receive wind stress stage three more streams)
  __host__ __device__ size_t ring_bytes(int l) const { return (size_t)(((wind_layers >> l) & 1) ? nseg_all : nseg_nowind) * seg_bytes; }
  __host__ __device__ size_t ring_off(int l) const {
    const int nw = __builtin_popcount((unsigned)wind_layers & ((1u << l) - 1u));
    return off_ring + ((size_t)nw * nseg_all + (size_t)(l - nw) * nseg_nowind) * seg_bytes;
  }
};
__host__ __device__ inline SmemPlan smem_plan(int nlay, int groups, int n_all, int n_nowind, int wind_layers) {
  SmemPlan p;
  p.tpad = groups * 32 + 2;
  size_t o = (size_t)4 * nlay * p.tpad * 8;          // thickness ring [4][nlay][tpad]
  p.off_bars = o;
  o += (size_t)(8 * nlay + 2 * groups) * 8;           // full[nlay][4], empty[nlay][4], gbar[groups][2]
  o = (o + 15) & ~(size_t)15;
  p.off_wring = o;
  o += (size_t)nlay * groups * (kWRings * 4 * kWRow * 8 + kOWords * 4);  // per-warp state rings + open-water bitmap
  o = (o + 127) & ~(size_t)127;
  p.off_ring = o;
  p.seg_bytes = (size_t)seg_doubles(groups) * 8;
  p.nseg_all = ring_segments(n_all);
  p.nseg_nowind = ring_segments(n_nowind);
  p.wind_layers = wind_layers & ((1 << nlay) - 1);
  p.total = p.ring_off(nlay);
  return p;
}

// FLAVOR only distinguishes the symbol of the copy compiled with FMA contraction (fused_inst_lean*_fma.cu, BEOM_FMA=1)
template <bool UFIRST, bool VISC, int NL, int FEAT, int GROUPS, int FLAVOR = 0, bool G0 = false>
// (a specialised instantiation knows its block size: fewer than 16 warps leave each thread more than 128 registers)
__global__ void __launch_bounds__((FEAT >= 0 && NL > 0 && GROUPS > 0) ? NL * GROUPS * 32 : kMaxWarps * 32,
                                  (FEAT >= 0 && NL > 0 && GROUPS > 0 && NL * GROUPS * 32 <= 256) ? BEOM_MIN_CTAS : 1)
k_fused_step(const __grid_constant__ Dev D, const __grid_constant__ Dev O, const __grid_constant__ StreamTab T,
             const uint8_t *__restrict__ open, const unsigned *__restrict__ open4, int open4_words, int groups_rt, int rows_per_chunk,
             int wind_layers) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool GEN = FEAT < 0, SPEC = FEAT >= 0;
  static_assert(GEN || (GROUPS > 0 && NL > 0), "a specialised option set has its layer and column-group counts compiled in");
  const bool has_nudg = GEN ? (D.has_nudg != 0) : ((FEAT & FB_NUDG) != 0);
  const bool has_bdrg = GEN ? (D.has_bdrg != 0) : ((FEAT & FB_BDRG) != 0);
  const bool has_tdrg = GEN && D.has_tdrg, has_hdot = GEN && D.has_hdot;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int groups = GROUPS > 0 ? GROUPS : groups_rt;
  const int l = wid / groups;  // layer of this warp
  // column group, rotated by the layer: the warps that share a scheduler (wid mod 4) then belong to different layers
  // AND different column groups, so no two of them wait for each other at the same barrier
  const int grp = (wid + l) % groups;
  const int nlay = NL > 0 ? NL : D.nlay;
  const int NX = D.NX, NY = D.NY;
  const int wseg = seg_doubles(groups);   // doubles per staged row segment (whole CTA width)
  const int tpad = groups * 32 + 2;
  const int tcol = grp * 32 + lane;
  const SmemPlan sp = smem_plan(nlay, groups, T.n, T.n_nowind, D.has_wind ? wind_layers : 0);
  double *sh_h = reinterpret_cast<double *>(smem_raw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + sp.off_bars);
  double *wring = reinterpret_cast<double *>(smem_raw + sp.off_wring) + (size_t)wid * (kWRings * 4 * kWRow + kOWords / 2);
  unsigned *obits = reinterpret_cast<unsigned *>(wring + kWRings * 4 * kWRow);  // [32]: open-water bit of every 4-row group of the chunk
  double *ring = reinterpret_cast<double *>(smem_raw + sp.ring_off(l));  // input ring of this layer

  // bit 16 of wind_layers: the grid is (chunks, strips) instead of (strips, chunks) -- which CTAs run side by side in a wave
  const int bx = ((wind_layers >> 16) & 1) ? blockIdx.y : blockIdx.x, by = ((wind_layers >> 16) & 1) ? blockIdx.x : blockIdx.y;
  const int tile = bx * groups + grp;
  const int xs = D.x_lo + bx * groups * kUse - kPad;  // first staged column of the CTA (even)
  const int x = xs + 2 + grp * kUse + lane;                    // lane 0 = first result column - 2
#ifdef BEOM_DBG_NOSTORE  // timing experiment (tools/ab.sh): every global store predicated off at run time; results are garbage
  const bool col_ok = lane >= kHalo && lane < 32 - kHalo && x <= D.x_hi && D.dt < -1.0;
#else
  const bool col_ok = lane >= kHalo && lane < 32 - kHalo && x <= D.x_hi;
#endif
  // rows of this CTA.  rows_per_chunk < 0: the two "edge" chunks of a y-slab (the -rows_per_chunk rows next to each
  // neighbouring rank), which are computed first so that their exchange overlaps the interior rows
  // (rows_per_chunk > 0: base | nbig << 16 -- chunks of base rows, the first nbig of base + 4, the last one up to y_hi; fused.cu)
  const int rpc = rows_per_chunk & 0xffff, nbig = rows_per_chunk >> 16;
  const int nchunks = (int)(((wind_layers >> 16) & 1) ? gridDim.x : gridDim.y);
  const int ya = rows_per_chunk > 0 ? D.y_lo + by * rpc + 4 * min(by, nbig) : (by == 0 ? D.y_lo : D.y_hi + rows_per_chunk + 1);
  const int yb = rows_per_chunk > 0 ? (by == nchunks - 1 ? D.y_hi : ya + rpc + (by < nbig ? 4 : 0) - 1) : ya - rows_per_chunk - 1;
  const size_t L = (size_t)l * D.plane;
  const bool wind = D.has_wind && ((wind_layers >> l) & 1);
  const int nstr = wind ? T.n : T.n_nowind;

  // output planes of this layer as byte pointers; off0 / off2 are the running byte offsets of (R, x) / (R-2, x)
  char *__restrict__ o_hlay = reinterpret_cast<char *>(O.hlay + L), *__restrict__ o_u = reinterpret_cast<char *>(O.u + L);
  char *__restrict__ o_v = reinterpret_cast<char *>(O.v + L), *__restrict__ o_hu = reinterpret_cast<char *>(O.h_u + L);
  char *__restrict__ o_hv = reinterpret_cast<char *>(O.h_v + L), *__restrict__ o_rs = reinterpret_cast<char *>(D.rs_new + L);
  char *__restrict__ o_dx = reinterpret_cast<char *>(D.dx_new + L), *__restrict__ o_dy = reinterpret_cast<char *>(D.dy_new + L);

  double cb[kMaxLay];  // (rhon(l) - rhon(i)) * i_rn(l), private_mod.f95:2359
#pragma unroll
  for (int i = 0; i < kMaxLay; i++) cb[i] = (i < l) ? (D.rhon[l] - D.rhon[i]) * D.i_rn[l] : 0.0;
  const double kin = 0.25 * D.uadv * D.i_gr;  // private_mod.f95:2381
  const bool ocrp = GEN ? (D.ocrp > 0.5) : ((FEAT & FB_OCRP) != 0);

  // ---- producer side: the warps of a layer share the staging; lane j of warp (l, grp) owns stream grp + j*groups ----
  const unsigned full0 = smem_u32(bars + 4 * l);               // [4]: inputs of front row R & 3 have landed (tx)
  const unsigned empty0 = smem_u32(bars + 4 * nlay + 4 * l);   // [4]: every column group has finished row R & 3
  const unsigned gbar0 = smem_u32(bars + 8 * nlay + 2 * grp);  // [2]: thickness exchange of a column group (split-phase)
  const unsigned ring0 = smem_u32(ring);
  const int my_s = grp + lane * groups;  // slots are compact and wind-only streams come last
  const bool my_on = my_s < nstr;
  const char *my_src = nullptr;
  int my_lag = 0;
  if (my_on) {
    my_src = reinterpret_cast<const char *>(T.base[my_s] + (size_t)T.lstride[my_s] * L + xs);
    my_lag = T.lag[my_s];
  }
  const unsigned segb = (unsigned)(wseg * 8);
  const unsigned my_dst = ring0 + (unsigned)(my_s < 4 ? my_s * 4 : 16 + (my_s - 4) * 2) * segb;
  const int my_mask = my_s < 4 ? 3 : 1;
  const unsigned my_bytes = (grp < nstr ? (unsigned)((nstr - grp + groups - 1) / groups) : 0u) * segb;  // this warp's share of a row
  {  // rows below the chunk read as 0 until staged; so do the state rings
    const int n = (int)(sp.ring_bytes(l) / 8);
    for (int i = grp * 32 + lane; i < n; i += groups * 32) ring[i] = 0.0;
    for (int i = lane; i < kWRings * 4 * kWRow; i += 32) wring[i] = 0.0;
    for (int i = threadIdx.x; i < 4 * nlay * tpad; i += blockDim.x) sh_h[i] = 0.0;
  }
  if (lane == 0) {
    if (grp == 0) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        mbar_init(full0 + 8 * k, (unsigned)groups);
        mbar_init(empty0 + 8 * k, (unsigned)groups);
      }
    }
    if (l == 0) {
      mbar_init(gbar0, (unsigned)nlay);
      mbar_init(gbar0 + 8, (unsigned)nlay);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  {
    // bits 17..28 of wind_layers (experiment): CTAs start up to that many x 64 ns apart (a hash of the block index), so that
    // the CTAs of a wave do not issue their row loads in the same instant
    const unsigned jit = ((unsigned)wind_layers >> 17) & 0xfffu;
    if (jit && threadIdx.x == 0) {
      unsigned hsh = (blockIdx.x * 2654435761u) ^ (blockIdx.y * 40503u + 0x9e3779b9u);
      hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
      const unsigned ns = (hsh % jit) * 64u;
      for (unsigned t = 0; t < ns; t += 1000u) __nanosleep(min(1000u, ns - t));
    }
  }
  __syncthreads();
  const int R0 = ya - 3, R1 = yb + 2;
  // The row loop works in groups of 4 rows from R0 on: ring slot = (R - R0) & 3 = unroll phase.  (It used to start on a
  // multiple of 4, which cost every chunk 3 rows of alignment on average; the host now picks chunks of 4k + 3 rows, whose
  // rows + 5 of pipeline lead-in and tail are whole groups.)
  const int Rend = R0 + ((R1 - R0) | 3);  // last row the loop visits
  const size_t row_bytes = (size_t)NX * 8;
  auto issue = [&](int Rt, int slot) {  // stage this warp's share of the inputs of front row Rt (slot = (Rt - R0) & 3)
    const unsigned bar = full0 + 8 * slot;
    if (lane == 0) mbar_expect_tx(bar, my_bytes);
    if (my_on) bulk_g2s(my_dst + (unsigned)(slot & my_mask) * segb, my_src + (size_t)min(max(Rt - my_lag, 0), NY - 1) * row_bytes, segb, bar);
    if (BEOM_L2_AHEAD > 0 && my_on && Rt + BEOM_L2_AHEAD <= Rend)
      bulk_prefetch_l2(my_src + (size_t)min(max(Rt + BEOM_L2_AHEAD - my_lag, 0), NY - 1) * row_bytes, segb);
  };
  issue(R0, 0);
  issue(R0 + 1, 1);
  for (int k = 2; k < BEOM_L2_AHEAD; k++)  // rows the first two issues do not reach
    if (my_on && R0 + k <= Rend) bulk_prefetch_l2(my_src + (size_t)min(max(R0 + k - my_lag, 0), NY - 1) * row_bytes, segb);
  size_t off0 = ((size_t)R0 * NX + x) * 8;                   // only dereferenced for rows this chunk owns
  size_t off2 = off0 - 2 * row_bytes;

  // ---- values carried from earlier rows: X[(phase - age) & 3] is X of row R - age ----
  double rv[4] = {0, 0, 0, 0}, dv[4] = {0, 0, 0, 0};
  double A1[4] = {0, 0, 0, 0}, A3[4] = {0, 0, 0, 0}, B1[4] = {0, 0, 0, 0}, B3[4] = {0, 0, 0, 0};
  double Qv[4] = {0, 0, 0, 0}, qp[4] = {0, 0, 0, 0};
  double Gy[4] = {0, 0, 0, 0}, Tc[4] = {0, 0, 0, 0}, fl[4] = {0, 0, 0, 0};
  double vold = 0.0;              // v(R-2) at time n when u goes first (its ring slot is being refilled)
  unsigned fw_m1 = 0, fw_m2 = 0;  // flags of (own | W<<8 | E<<16), masked rows only

  const double *sgp = ring + 2 + grp * kUse + lane;
  double *shp = sh_h + (size_t)l * tpad + 1 + tcol;  // + slot*nlay*tpad: this thread's thickness of row slot
  double *wrp = wring + 1 + lane;                     // + (ring*4 + slot)*kWRow
  const int tt_base = SPEC ? spec_slot(FEAT, S_TTXU) : (int)T.slot[S_TTXU];  // wind streams keep their order: TTXU, TTYV, TTYVS

#define AT(a, age) a[(PH - (age)) & 3]
#define SLOT(age) ((PH - (age)) & 3)
#define LD4(s, age, dx) sgp[((s) * 4 + SLOT(age)) * wseg + (dx)]
#define LD2(slot, dx) sgp[(16 + ((slot)-4) * 2 + (SLOT(0) & 1)) * wseg + (dx)]
#define LDX(stream, dx) LD2((SPEC ? spec_slot(FEAT, stream) : (int)T.slot[stream]), dx)
#define HN(age, dx) shp[SLOT(age) * nlay * tpad + (dx)]
#define WR(ring_id, age, dx) wrp[((ring_id)*4 + SLOT(age)) * kWRow + (dx)]
#define SELM(p, a) (MASKED ? sel((p), (a)) : (a))
#define MKN(f) (MASKED ? m_n(f) : 1.0)
  enum { W_MO = 0, W_PV = 1, W_FX = 2 };
  // One row of the pipeline.  PH = phase of the carried rings = ring slot of the row, (R - R0) & 3.
  // MASKED = false is the open-water fast path: every mask of the three rows in flight is 1 on all 32 lanes,
  // so no select is needed.
  auto row = [&](auto ph_tag, auto masked_tag, const int R, const unsigned f_own, const unsigned fw_0, const unsigned bpar) {
    constexpr int PH = decltype(ph_tag)::value;
    constexpr bool MASKED = decltype(masked_tag)::value;
    const int hslot = SLOT(0) & 1;
    constexpr unsigned hpar = (PH >> 1) & 1;
    mbar_wait(full0 + 8 * SLOT(0), bpar);  // staged inputs of front row R have landed
    const bool act = f_own & (F_ACT | F_GHOST);  // evaluated (ghost cells: like the cell they mirror) ...
    const bool own = f_own & F_ACT;              // ... and stored
    const bool row_own = (R >= ya && R <= yb);
    const bool row2_own = (R - 2 >= ya && R - 2 <= yb);

    // -------------------------------------------------------------------------------- update_h, row R (pm:1610-1643)
    const double hu_0 = LD4(S_HU, 0, 0), huE_0 = LD4(S_HU, 0, 1);
    const double hv_p1 = LD4(S_HV, 0, 0), hv_0 = LD4(S_HV, 1, 0);
    double rs_3 = (hu_0 - huE_0) * D.i_dl + (hv_0 - hv_p1) * D.i_dl;
    if (has_hdot) rs_3 = rs_3 + LDX(S_HDOT, 0);
    rs_3 = SELM(f_own & F_N, rs_3);
    const double r1 = LD2(S_R1, 0), r2 = LD2(S_R2, 0);
    double rhs_h;
    if (SPEC && G0) rhs_h = rs_3 * D.dt;  // gene = 0 (start-up steps): the extrapolated term is an exact zero
    else if (SPEC) rhs_h = (D.c_ab1 * rs_3 - D.c_ab2 * r2 + D.beta * r1) * D.dt;  // gene = 1: the plain term is an exact zero
    else      rhs_h = (D.c_ab1 * rs_3 - D.c_ab2 * r2 + D.beta * r1) * D.dt * D.gene + rs_3 * D.dt * (1.0 - D.gene);
    double hn_0 = LD2(S_HL, 0) + rhs_h;
    if (has_nudg) {
      const double fnn_0 = LDX(S_FNN, 0), nudn_0 = LDX(S_NUDN, 0);
      hn_0 = fnn_0 * nudn_0 + (1.0 - nudn_0) * hn_0;
    }
    hn_0 = SELM(act, hn_0);
    if (col_ok && row_own && (!MASKED || own)) {
      __stcs(reinterpret_cast<double *>(o_hlay + off0), hn_0);
      __stcs(reinterpret_cast<double *>(o_rs + off0), rs_3);
    }
    HN(0, 0) = hn_0;
    __syncwarp();
    if (lane == 0) mbar_arrive(gbar0 + 8 * hslot);  // split-phase: waited for at the end of the row

    // -------------------------------------------------------------------------------- rvor, dive, row R (pm:2388, 2435)
    const double u_0 = LD4(S_U, 0, 0), uE_0 = LD4(S_U, 0, 1), u_m1 = LD4(S_U, 1, 0);
    const double v_p1 = LD4(S_V, 0, 0), v_0 = LD4(S_V, 1, 0), vW_0 = LD4(S_V, 1, -1);
    const double rv_0 = SELM(act && (f_own & F_PE), (v_0 - vW_0 - u_0 + u_m1) * D.i_dl);
    const double dv_0 = SELM(act, (uE_0 - u_0 + v_p1 - v_0) * D.i_dl);
    const double ke = kin * (uE_0 * uE_0 + u_0 * u_0 + v_p1 * v_p1 + v_0 * v_0);  // kinetic part of the Bernoulli potential (pm:2381)

    // -------------------------------------------------------------------------------- d2hx, pvor, row R; d2hy, row R-1
    const double hnE_0 = HN(0, 1), hnW_0 = HN(0, -1), hn_m1 = HN(1, 0), hn_m2 = HN(2, 0);
    double d2x_0 = SELM((fw_0 & (F_N << 16)) && (fw_0 & (F_N << 8)) && (f_own & F_N), hnE_0 + hnW_0 - hn_0 * 2.0);
    if (ocrp && (hnE_0 < D.two_hs || hnW_0 < D.two_hs || hn_0 < D.two_hs)) d2x_0 = 0.0;
    d2x_0 = SELM(act, d2x_0);
    WR(W_FX, 0, 0) = 0.16667 * d2x_0;
    double d2y_m1 = SELM((f_own & F_N) && (fw_m2 & F_N) && (fw_m1 & F_N), hn_0 + hn_m2 - hn_m1 * 2.0);
    if (ocrp && (hn_0 < D.two_hs || hn_m2 < D.two_hs || hn_m1 < D.two_hs)) d2y_m1 = 0.0;
    d2y_m1 = SELM(fw_m1 & (F_ACT | F_GHOST), d2y_m1);
    AT(Gy, 1) = 0.16667 * d2y_m1;
    {
      const double have = hn_0 + hnW_0 + HN(1, -1) + hn_m1;
      const double msum = MKN((uint8_t)f_own) + MKN((uint8_t)(fw_0 >> 8)) + MKN((uint8_t)(fw_m1 >> 8)) + MKN((uint8_t)fw_m1);
      const double pv_0 = SELM(act, SELM(f_own & F_PI, LD2(S_FCOR, 0) + rv_0 * D.uadv) * msum / have);
      AT(qp, 0) = 0.25 * pv_0;
    }
    if (UFIRST) AT(Tc, 0) = AT(qp, 0) * (hv_0 + LD4(S_HV, 1, -1));  // Coriolis term of u with h_v at time n (pm:1461-1462)

    // -------------------------------------------------------------------------------- Leith viscosity, row R-1 (pm:2477-2502)
    double P_m1 = 0.0;
    if (VISC) {
      AT(A1, 0) = sq(shdn(rv_0) - rv_0);
      AT(A3, 0) = sq(rv_0 - AT(rv, 1));
      AT(B1, 0) = sq(dv_0 - shup(dv_0));
      AT(B3, 0) = sq(dv_0 - AT(dv, 1));
      const double tll = AT(A1, 1) + shup(AT(A1, 1)) + AT(A3, 0) + AT(A3, 1) + AT(B1, 1) + AT(B1, 2) + AT(B3, 1) + shup(AT(B3, 1));
      const double tcc = AT(A1, 1) + AT(A1, 0) + AT(A3, 0) + shdn(AT(A3, 0)) + shdn(AT(B1, 1)) + AT(B1, 1) + AT(B3, 0) + AT(B3, 1);
      const bool a = fw_m1 & (F_ACT | F_GHOST);
      const double vll_m1 = SELM(a, sqrt(tll) * D.dvis * D.dl * D.dl + D.bvis);
      const double vcc_m1 = SELM(a, sqrt(tcc) * D.dvis * D.dl * D.dl + D.bvis);
      P_m1 = vcc_m1 * AT(dv, 1);
      WR(W_PV, 1, 0) = P_m1;
      AT(Qv, 1) = vll_m1 * AT(rv, 1);
    }
    AT(rv, 0) = rv_0;
    AT(dv, 0) = dv_0;

    // -------------------------------------------------------------------------------- momentum
    const bool a2 = fw_m2 & (F_ACT | F_GHOST);
    const bool sto2 = col_ok && (!MASKED || (fw_m2 & F_ACT)) && row2_own;
    MomX xu, xv;
    if (wind) {
      xu.tw_b = LD2(tt_base, 0); xu.tw_a = LD2(tt_base, -1);
      xv.tw_b = LD2(tt_base + 1, 0); xv.tw_a = LD2(tt_base + 2, 0);
    }
    if (wind && has_nudg) {
      xu.te_b = LDX(S_TTYU, 0); xu.te_a = LDX(S_TTYU, -1);
      xv.te_b = LDX(S_TTXV, 0); xv.te_a = LDX(S_TTXVS, 0);
    }
    if (has_bdrg) { xu.tb = LDX(S_TBX, 0); xv.tb = LDX(S_TBY, 0); }
    if (has_tdrg) { xu.tu = LDX(S_TUX, 0); xv.tu = LDX(S_TUY, 0); }
    if (has_nudg) { xu.fn = LDX(S_FNU, 0); xu.nud = LDX(S_NUDU, 0); xv.fn = LDX(S_FNV, 0); xv.nud = LDX(S_NUDV, 0); }
    if (GEN) { xu.bodf = D.bodf[0][l]; xv.bodf = D.bodf[1][l]; }
    const double ux1 = LD2(S_DX1, 0), ux2 = LD2(S_DX2, 0), ux3 = LD2(S_DX3, 0);
    const double vy1 = LD2(S_DY1, 0), vy2 = LD2(S_DY2, 0), vy3 = LD2(S_DY3, 0);
    const double mo_m2 = WR(W_MO, 2, 0), P_m2 = VISC ? WR(W_PV, 2, 0) : 0.0, PW_m2 = VISC ? WR(W_PV, 2, -1) : 0.0;
    const double hn_m2b = HN(2, 0);
    if (UFIRST) {
      // ---- u at row R-2 (pm:1422-1503) ----
      double un, hun, dm;
      momentum<true, VISC, MASKED, FEAT, G0>(D, wind, MASKED ? m_u((uint8_t)fw_m2) : 1.0, HN(2, -1) + hn_m2b, WR(W_MO, 2, -1), mo_m2,
                                         AT(Tc, 2), AT(Tc, 1), LD4(S_U, 2, 0), ux1, ux2, ux3, P_m2, PW_m2, AT(Qv, 1),
                                         AT(Qv, 2), WR(W_FX, 2, -1), WR(W_FX, 2, 0), xu, un, hun, dm);
      hun = SELM(a2, hun);
      if (sto2) {
        __stcs(reinterpret_cast<double *>(o_u + off2), un);
        __stcs(reinterpret_cast<double *>(o_hu + off2), hun);
        __stcs(reinterpret_cast<double *>(o_dx + off2), dm);
      }
      AT(fl, 2) = hun;
      // ---- v at row R-2 (pm:1505-1591), using the new h_u of rows R-2 and R-3 ----
      const double wc = AT(qp, 2) * (hun + AT(fl, 3));
      double vn, hvn;
      momentum<false, VISC, MASKED, FEAT, G0>(D, wind, MASKED ? m_v((uint8_t)fw_m2) : 1.0, hn_m2b + HN(3, 0), WR(W_MO, 3, 0), mo_m2, wc,
                                          shdn(wc), vold, vy1, vy2, vy3, P_m2, VISC ? WR(W_PV, 3, 0) : 0.0, shdn(AT(Qv, 2)), AT(Qv, 2),
                                          AT(Gy, 3), AT(Gy, 2), xv, vn, hvn, dm);
      if (sto2) {
        __stcs(reinterpret_cast<double *>(o_v + off2), vn);
        __stcs(reinterpret_cast<double *>(o_hv + off2), SELM(a2, hvn));
        __stcs(reinterpret_cast<double *>(o_dy + off2), dm);
      }
      vold = LD4(S_V, 2, 0);  // v(R-1): the old v of the next row's update
    } else {
      // ---- v at row R-1 (pm:1505-1591), h_u at time n ----
      const bool a1 = fw_m1 & (F_ACT | F_GHOST);
      const double wc = AT(qp, 1) * (LD4(S_HU, 1, 0) + LD4(S_HU, 2, 0));
      double vn, hvn, dm;
      momentum<false, VISC, MASKED, FEAT, G0>(D, wind, MASKED ? m_v((uint8_t)fw_m1) : 1.0, hn_m1 + hn_m2b, mo_m2, WR(W_MO, 1, 0), wc,
                                          shdn(wc), LD4(S_V, 2, 0), vy1, vy2, vy3, P_m1, P_m2, shdn(AT(Qv, 1)), AT(Qv, 1),
                                          AT(Gy, 2), AT(Gy, 1), xv, vn, hvn, dm);
      hvn = SELM(a1, hvn);
      if (col_ok && (!MASKED || (fw_m1 & F_ACT)) && (R - 1 >= ya) && (R - 1 <= yb)) {
        __stcs(reinterpret_cast<double *>(o_v + (off2 + row_bytes)), vn);
        __stcs(reinterpret_cast<double *>(o_hv + (off2 + row_bytes)), hvn);
        __stcs(reinterpret_cast<double *>(o_dy + (off2 + row_bytes)), dm);
      }
      AT(Tc, 1) = AT(qp, 1) * (hvn + shup(hvn));  // Coriolis term of u with the new h_v (pm:1461-1462)
      // ---- u at row R-2 (pm:1422-1503), using the new h_v of rows R-2 and R-1 ----
      double un, hun;
      momentum<true, VISC, MASKED, FEAT, G0>(D, wind, MASKED ? m_u((uint8_t)fw_m2) : 1.0, HN(2, -1) + hn_m2b, WR(W_MO, 2, -1), mo_m2,
                                         AT(Tc, 2), AT(Tc, 1), LD4(S_U, 2, 0), ux1, ux2, ux3, P_m2, PW_m2, AT(Qv, 1),
                                         AT(Qv, 2), WR(W_FX, 2, -1), WR(W_FX, 2, 0), xu, un, hun, dm);
      if (sto2) {
        __stcs(reinterpret_cast<double *>(o_u + off2), un);
        __stcs(reinterpret_cast<double *>(o_hu + off2), SELM(a2, hun));
        __stcs(reinterpret_cast<double *>(o_dx + off2), dm);
      }
    }

    // -------------------------------------------------------------------------------- mont, row R (pm:2351-2383)
    double mpot;
    if (ocrp) {
      mpot = hn_0 + D.hmin * (1.0 - MKN((uint8_t)f_own));
      mpot = cube(D.hsal / mpot);
      mpot = mpot * (-D.ocrp * D.i_ns * D.hsal * MKN((uint8_t)f_own));
    } else {
      mpot = -0.0;
    }
    mpot = mpot - 0.0;
    const double hth = LD2(S_HTH, 0);
    mbar_wait(gbar0 + 8 * hslot, hpar);  // every layer of this column group has published hn(R)
    {
      const double *col = sh_h + (size_t)SLOT(0) * nlay * tpad + 1 + tcol;
      double hcol = 0.0;
#pragma unroll
      for (int i = 0; i < (NL > 0 ? NL : kMaxLay); i++) {
        if (i < nlay) {
          const double hi = col[(size_t)i * tpad];
          if (i < l) mpot = mpot - cb[i] * hi;
          hcol = hcol + hi;
        }
      }
      mpot = hcol - hth + mpot;
    }
    WR(W_MO, 0, 0) = SELM(act, mpot + ke);

    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * SLOT(0));  // this warp has finished reading the slots row R + 2 refills
    if (R + 2 <= Rend) {
      mbar_wait(empty0 + 8 * SLOT(0), bpar);           // ... and so has every other column group of the layer
      issue(R + 2, (PH + 2) & 3);
    }
    off0 += row_bytes;
    off2 += row_bytes;
  };
#undef LDX
#undef LD2
#undef LD4
#undef AT
#undef SLOT
#undef HN
#undef WR
#undef SELM
#undef MKN

  constexpr unsigned kAllMasks = 0x3f | (0x3f << 8) | (0x3f << 16);
  const uint8_t *fl_p = D.flags + x;
  using Tt = std::true_type;
  using Ft = std::false_type;
  auto flags_of = [&](int R) -> unsigned { return (unsigned)fl_p[(size_t)min(R, NY - 1) * NX]; };
  auto widen = [&](unsigned f) -> unsigned { return f | (__shfl_up_sync(0xffffffffu, f, 1) << 8) | (__shfl_down_sync(0xffffffffu, f, 1) << 16); };
  {
    unsigned bpar = 0;
    // open-water bit of every 4-row group of the chunk, kept in shared memory: word k covers groups 32k .. 32k+31
    // past the first (a global load per group would sit on the critical path of every group).  The groups of this chunk are
    // the rows congruent to R0 mod 4: open4 holds one bitmap per residue, bit (R >> 2) of it says rows R-2 .. R+3 are open
    const int w0 = R0 >> 7;
    const unsigned *o4 = open4 + ((size_t)tile * 4 + (R0 & 3)) * open4_words;  // [tile][residue][word]
    if (lane < kOWords) obits[lane] = (w0 + lane < open4_words) ? o4[w0 + lane] : 0u;
    __syncwarp();
#pragma unroll 1
    for (int R = R0; R <= R1; R += 4) {
      const unsigned o = obits[(R >> 7) - w0] >> ((R >> 2) & 31);
      if (o & 1) {  // rows R-2 .. R+3 are open water on all 32 columns
        row(ic<0>{}, Ft{}, R, kAllMasks, kAllMasks, bpar);
        row(ic<1>{}, Ft{}, R + 1, kAllMasks, kAllMasks, bpar);
        row(ic<2>{}, Ft{}, R + 2, kAllMasks, kAllMasks, bpar);
        row(ic<3>{}, Ft{}, R + 3, kAllMasks, kAllMasks, bpar);
        fw_m1 = fw_m2 = kAllMasks;
      } else {
        const unsigned f0 = flags_of(R), f1 = flags_of(R + 1), f2 = flags_of(R + 2), f3 = flags_of(R + 3);
        const unsigned e0 = widen(f0), e1 = widen(f1), e2 = widen(f2), e3 = widen(f3);
        row(ic<0>{}, Tt{}, R, f0, e0, bpar);
        fw_m2 = fw_m1; fw_m1 = e0;
        row(ic<1>{}, Tt{}, R + 1, f1, e1, bpar);
        fw_m2 = fw_m1; fw_m1 = e1;
        row(ic<2>{}, Tt{}, R + 2, f2, e2, bpar);
        fw_m2 = fw_m1; fw_m1 = e2;
        row(ic<3>{}, Tt{}, R + 3, f3, e3, bpar);
        fw_m2 = fw_m1; fw_m1 = e3;
      }
      bpar ^= 1;
    }
  }
}

}  // namespace fusedk
}  // namespace beom
#endif
