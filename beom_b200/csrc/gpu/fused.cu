// fused.cu -- the fused single-pass time step (generalized forward-backward regime, tstp >= 4).
//
// One kernel per step.  It reads every persistent field once (hlay,u,v,h_u,h_v, 2 rs_h, 3 dmdx,
// 3 dmdy: 13 doubles per cell-layer) and writes the 8 new ones (hlay,u,v,h_u,h_v + newest rs_h, dmdx,
// dmdy): the 168 algorithmic bytes of SURVEY.md 8(d).  Everything the reference keeps in 2-D scratch
// arrays between its loops (mont, rvor, pvor, dive, d2hx, d2hy, v_cc, v_ll; private_mod.f95:48-61)
// lives in registers here.
//
// Mapping.  A warp owns 32 consecutive columns of ONE layer and marches north (row by row) through
// its y-chunk; lane k is column x0-2+k, lanes 2..29 produce results, lanes 0,1,30,31 are the x halo
// (recomputed by the neighbouring warp).  East/west neighbours come from warp shuffles, south
// neighbours from values the thread kept from earlier rows: a 3-row software pipeline
//     row R   : update_h, rvor, dive, Montgomery/Bernoulli potential, d2hx, pvor       (front)
//     row R-1 : d2hy, Leith v_cc / v_ll, and the first momentum component if it is v
//     row R-2 : u (and v when u goes first)
// A CTA is (column groups) x (layers) warps; the only shared memory is the new layer thickness of
// the front row, exchanged between the layer-warps for the column sums of the Montgomery potential
// (private_mod.f95:2357-2373), one __syncthreads per row.  State is double buffered (in -> out), so
// halo lanes and neighbouring CTAs always read time level n.
//
// Arithmetic: the same expressions in the same order as split.cuh / the reference (-fmad=false), so
// the two paths agree bit for bit; 0/1 masks are applied as selects (x*1 = x, x*0 = +-0).
#include "fused.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <type_traits>

namespace beom {

namespace {

constexpr int kHalo = 2;             // halo lanes on each side of a warp
constexpr int kUse = 32 - 2 * kHalo;  // 28 result columns per warp
constexpr int kMaxLay = 8;           // layers per CTA (shared-memory exchange, warps per CTA)
constexpr int kMaxThreads = 384;     // 12 warps: 168 registers per thread, no spills

struct FusedCfg {
  bool ok = false;
  int groups = 1;    // column groups (warps per layer) per CTA
  int strips = 1;    // CTAs along x
  int chunks = 1;    // CTAs along y
  int rows_per_chunk = 1;
  bool visc = false;  // Leith viscosity recomputed every step
  int wind_layers = 0;  // bit l set: tt3d of layer l may be non-zero
} cfg;

__device__ __forceinline__ double sel(bool p, double a) { return p ? a : 0.0; }
__device__ __forceinline__ double shup(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }     // value of lane-1 (west)
__device__ __forceinline__ double shdn(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }   // value of lane+1 (east)

struct MomIn {  // everything one momentum update needs, gathered by the caller
  double mask;        // mk_u / mk_v of the point
  double h_a, h_b;    // thickness of the two cells: hcen = (h_a + h_b)/(1+mask)
  double m_far, m_here;  // mont(W or S), mont(point)
  double pv_a, f_a0, f_a1;  // first Coriolis pair:  0.25*pv_a*(f_a0+f_a1)
  double pv_b, f_b0, f_b1;  // second pair
  double old;         // u or v at time n
  double h1, h2, h3;  // dmdx/dmdy(1..3)
  double vcc_here, dv_here, vcc_far, dv_far;  // (v_cc*dive)(point) - (v_cc*dive)(W or S)
  double vll_far, rv_far, vll_here, rv_here;  // (v_ll*rvor)(N or E) - (v_ll*rvor)(point)
  double d2_far, d2_here;                     // d2hx/d2hy (W or S), (point)
  double tw_a, tw_b;   // wind stress component along the velocity at (W or S), (point)
  double te_a, te_b;   // cross component for the Ekman term of the sponge target
  double tb, tu;       // bottom / top drag at the point
  double fn, nud;      // sponge target and rate
  double bodf;
};

// update_u (private_mod.f95:1437-1500) when IS_U, update_v (private_mod.f95:1520-1586) otherwise.
template <bool IS_U, bool VISC, bool MASKED>
__device__ __forceinline__ void momentum(const Dev &D, const MomIn &q, double &vel, double &flux, double &dmd4) {
  const double hcen = (q.h_a + q.h_b) * (MASKED ? (q.mask != 0.0 ? 0.5 : 1.0) : 0.5);
  const double i__h = 1.0 / (hcen + 1.0 - (MASKED ? q.mask : 1.0));
  dmd4 = (q.m_far - q.m_here) * D.i_dl * D.grav;
  if (MASKED) dmd4 = sel(q.mask != 0.0, dmd4);
  double rhsi;
  if (IS_U) rhsi = dmd4 * (1.0 - D.gene) + 0.25 * q.pv_a * (q.f_a0 + q.f_a1) + 0.25 * q.pv_b * (q.f_b0 + q.f_b1);
  else      rhsi = dmd4 * (1.0 - D.gene) - 0.25 * q.pv_a * (q.f_a0 + q.f_a1) - 0.25 * q.pv_b * (q.f_b0 + q.f_b1);
  if (D.has_wind) rhsi = rhsi + 0.5 * (q.tw_a + q.tw_b) * D.ramp * D.i_r0 * i__h;
  if (D.has_bdrg) rhsi = rhsi - q.tb * D.i_r0 * i__h;
  if (D.has_tdrg) rhsi = rhsi - q.tu * D.i_r0 * i__h;
  rhsi = rhsi + q.bodf + (D.del1 * dmd4 + D.del2 * q.h3 + D.gamm * q.h2 + D.epsi * q.h1) * D.gene;
  if (VISC) {
    if (IS_U) rhsi = rhsi + (q.vcc_here * q.dv_here - q.vcc_far * q.dv_far) * D.i_dl - (q.vll_far * q.rv_far - q.vll_here * q.rv_here) * D.i_dl;
    else      rhsi = rhsi + (q.vcc_here * q.dv_here - q.vcc_far * q.dv_far) * D.i_dl + (q.vll_far * q.rv_far - q.vll_here * q.rv_here) * D.i_dl;
  }
  double w = q.old + (MASKED ? sel(q.mask != 0.0, rhsi) : rhsi) * D.dt;
  if (D.has_nudg) {
    double tgt = q.fn;
    if (D.has_wind) {
      if (IS_U) tgt = tgt + 0.5 * (q.te_b + q.te_a) * D.i_r1 * D.invf * i__h * D.ramp;
      else      tgt = tgt - 0.5 * (q.te_b + q.te_a) * D.i_r1 * D.invf * i__h * D.ramp;
    }
    w = tgt * q.nud + w * (1.0 - q.nud);
  }
  vel = w;
  flux = 0.5 * (w + fabs(w)) * (hcen - 0.16667 * q.d2_far) + 0.5 * (w - fabs(w)) * (hcen - 0.16667 * q.d2_here);
}

// ---- raw-input streams: one 36-double row segment per (field, row lag), staged by TMA bulk copies ----
enum {
  S_HU, S_HV, S_HL, S_R1, S_R2, S_U, S_V, S_DX1, S_DX2, S_DX3, S_DY1, S_DY2, S_DY3, S_FCOR, S_HTH,  // always
  S_HDOT, S_FNN, S_NUDN,                                       // update_h extras
  S_TTXU, S_TTYU, S_TBX, S_TUX, S_FNU, S_NUDU,                 // u stage
  S_TTYV, S_TTYVS, S_TTXV, S_TTXVS, S_TBY, S_TUY, S_FNV, S_NUDV,  // v stage (..S = south neighbour row)
  S_COUNT
};
constexpr int kSeg = 36;  // doubles per staged row segment: columns xw0-2 .. xw0+33 (16-byte aligned)

struct StreamTab {
  int n;                       // enabled streams
  int n_nowind;                // enabled streams for a layer that receives no wind stress
  signed char slot[S_COUNT];   // stream -> compact slot (-1 = disabled)
  const double *base[S_COUNT];  // by slot
  short lag[S_COUNT];          // row = R - lag
  unsigned char lstride[S_COUNT];  // layer stride in planes (0 = 2-D field, 1, or 2 for [nlay][2] arrays)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

template <bool UFIRST, bool VISC, int NL>
__global__ void __launch_bounds__(kMaxThreads, 1) k_fused_step(const __grid_constant__ Dev D, const __grid_constant__ Dev O,
                                                     const __grid_constant__ StreamTab T, int groups, int rows_per_chunk, int wind_layers) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int grp = wid % groups;
  const int l = wid / groups;  // layer of this warp
  const int nlay = NL > 0 ? NL : D.nlay;
  const int NX = D.NX;
  const int tcols = groups * 32;
  const int tcol = grp * 32 + lane;
  // shared memory: [2][nlay][tcols] new thickness | per-warp mbarriers | per-warp 2-stage raw-input ring
  double *sh_h = reinterpret_cast<double *>(smem_raw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sh_h + (size_t)2 * nlay * tcols);
  unsigned long long *gbars = bars + 2 * nwarps;  // [groups][2]: thickness exchange of a column group (split-phase)
  double *ring = reinterpret_cast<double *>(gbars + 2 * groups) + (size_t)wid * 2 * T.n * kSeg;

  const int xw0 = D.x_lo + (blockIdx.x * groups + grp) * kUse - kHalo;  // column of lane 0
  const int xs = min(xw0 - 2, NX - kSeg);                                // first staged column (even)
  const int x = xs + 2 + lane;
  const bool col_ok = (xs == xw0 - 2) && lane >= kHalo && lane < 32 - kHalo && x <= D.x_hi;
  const int ya = D.y_lo + blockIdx.y * rows_per_chunk;
  const int yb = min(ya + rows_per_chunk - 1, D.y_hi);
  const size_t L = (size_t)l * D.plane;
  const bool wind = D.has_wind && ((wind_layers >> l) & 1);
  const int nstr = wind ? T.n : T.n_nowind;

  double *__restrict__ o_hlay = O.hlay + L, *__restrict__ o_u = O.u + L, *__restrict__ o_v = O.v + L;
  double *__restrict__ o_hu = O.h_u + L, *__restrict__ o_hv = O.h_v + L;
  double *__restrict__ o_rs = D.rs_new + L, *__restrict__ o_dx = D.dx_new + L, *__restrict__ o_dy = D.dy_new + L;

  double cb[kMaxLay];  // (rhon(l) - rhon(i)) * i_rn(l), private_mod.f95:2359
#pragma unroll
  for (int i = 0; i < kMaxLay; i++) cb[i] = (i < l) ? (D.rhon[l] - D.rhon[i]) * D.i_rn[l] : 0.0;
  const double kin = 0.25 * D.uadv * D.i_gr;  // private_mod.f95:2381
  const double bodf_u = D.bodf[0][l], bodf_v = D.bodf[1][l];
  const bool ocrp = D.ocrp > 0.5;

  // ---- producer side: every lane owns (at most) one stream of this warp ----
  const unsigned bar0 = smem_u32(bars + 2 * wid);
  const unsigned ring0 = smem_u32(ring);
  const int my_slot = lane;  // slots are compact and wind-only streams come last
  const bool my_on = my_slot < nstr;
  const double *my_src = nullptr;
  int my_lag = 0;
  if (my_on) {
    my_src = T.base[my_slot] + (size_t)T.lstride[my_slot] * L + xs;
    my_lag = T.lag[my_slot];
  }
  const unsigned gbar0 = smem_u32(gbars + 2 * grp);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    if (l == 0) {
      mbar_init(gbar0, (unsigned)(nlay * 32));
      mbar_init(gbar0 + 8, (unsigned)(nlay * 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int R0 = ya - 3, R1 = yb + 2;
  auto issue = [&](int Rt) {  // stage the inputs of front row Rt
    const int st = (Rt - R0) & 1;
    const unsigned bar = bar0 + 8 * st;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(nstr * kSeg * 8));
    __syncwarp();
    if (my_on) bulk_g2s(ring0 + (unsigned)((st * T.n + my_slot) * kSeg * 8), my_src + (size_t)max(Rt - my_lag, 0) * NX, kSeg * 8, bar);
  };
  issue(R0);
  if (R0 + 1 <= R1) issue(R0 + 1);

  // ---- values carried from earlier rows (suffix _mK = row R-K, _p1 = row R+1) ----
  double u_m1 = 0, u_m2 = 0, v_0 = 0, v_m1 = 0, v_m2 = 0, vW_0 = 0;
  double hv_0 = 0, hv_m1 = 0, hv_m2 = 0, hvW_m1 = 0, hvW_m2 = 0;  // h_v at time n (and its west neighbour)
  double hu_m1 = 0, hu_m2 = 0, huE_m1 = 0, huE_m2 = 0;            // h_u at time n (and its east neighbour)
  double hn_m1 = 0, hn_m2 = 0, hn_m3 = 0, hnW_m1 = 0, hnW_m2 = 0;
  double rv_m1 = 0, rv_m2 = 0, rvE_m1 = 0, rvE_m2 = 0, rvW_m1 = 0;
  double dv_m1 = 0, dv_m2 = 0, dv_m3 = 0, dvE_m1 = 0, dvW_m1 = 0, dvW_m2 = 0;
  double mo_m1 = 0, mo_m2 = 0, mo_m3 = 0;
  double pv_m1 = 0, pv_m2 = 0;
  double vcc_m2 = 0, vcc_m3 = 0, vll_m2 = 0;
  double d2x_m1 = 0, d2x_m2 = 0, d2y_m2 = 0, d2y_m3 = 0;
  double fl2_m2 = 0, fl2E_m2 = 0, fl2_m3 = 0, fl2E_m3 = 0;  // new flux of the first component (and E/W neighbour)
  unsigned fw_m1 = 0, fw_m2 = 0;  // flags of (own | W<<8 | E<<16)

  {  // prime the values of rows R0-1 / R0 that the loop expects to inherit (plain loads, once per chunk)
    const size_t c = (size_t)R0 * NX + x;
    v_0 = __ldg(D.v + L + c);
    vW_0 = __ldg(D.v + L + c - 1);
    hv_0 = __ldg(D.h_v + L + c);
    u_m1 = __ldg(D.u + L + c - NX);
  }
  unsigned f_next = D.flags[(size_t)R0 * NX + x];

#define SG(stream, dx) sg[(int)T.slot[stream] * kSeg + (dx)]
#define SGF(stream, dx) sg[(stream) * kSeg + (dx)]
#define SELM(p, a) (MASKED ? sel((p), (a)) : (a))
#define MKN(f) (MASKED ? m_n(f) : 1.0)
#define MKU(f) (MASKED ? m_u(f) : 1.0)
#define MKV(f) (MASKED ? m_v(f) : 1.0)
  // One row of the pipeline.  MASKED = false is the open-water fast path: every mask of the three rows in
  // flight is 1 on all 32 lanes (and their E/W neighbours), so no select is needed.
  auto row = [&](auto masked_tag, const int R, const unsigned f_own, const unsigned fw_0) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    const size_t c = (size_t)R * NX + x;
    const int st = (R - R0) & 1;
    const double *sg = ring + (size_t)st * T.n * kSeg + 2 + lane;
    const bool act = f_own & F_ACT;
    const double hu_0 = SGF(S_HU, 0), huE_0 = SGF(S_HU, 1);
    const double hv_p1 = SGF(S_HV, 0);
    const double hold = SGF(S_HL, 0);
    const double r1 = SGF(S_R1, 0), r2 = SGF(S_R2, 0);
    const double u_0 = SGF(S_U, 0), uE_0 = SGF(S_U, 1);
    const double v_p1 = SGF(S_V, 0), vW_p1 = SGF(S_V, -1);
    const double fcor_0 = SGF(S_FCOR, 0);
    const double hdot_0 = D.has_hdot ? SG(S_HDOT, 0) : 0.0;
    double fnn_0 = 0.0, nudn_0 = 0.0;
    if (D.has_nudg) { fnn_0 = SG(S_FNN, 0); nudn_0 = SG(S_NUDN, 0); }
    MomIn qu, qv;


    // -------------------------------------------------------------------------------- update_h, row R (pm:1610-1643)
    double rs_3 = (hu_0 - huE_0) * D.i_dl + (hv_0 - hv_p1) * D.i_dl;
    if (D.has_hdot) rs_3 = rs_3 + hdot_0;
    rs_3 = SELM(f_own & F_N, rs_3);
    const double rhs_h = (D.c_ab1 * rs_3 - D.c_ab2 * r2 + D.beta * r1) * D.dt * D.gene + rs_3 * D.dt * (1.0 - D.gene);
    double hn_0 = hold + rhs_h;
    if (D.has_nudg) hn_0 = fnn_0 * nudn_0 + (1.0 - nudn_0) * hn_0;
    hn_0 = SELM(act, hn_0);
    const bool row_own = (R >= ya && R <= yb);
    if (col_ok && row_own && (!MASKED || act)) {
      __stcs(o_hlay + c, hn_0);
      __stcs(o_rs + c, rs_3);
    }
    sh_h[((size_t)(R & 1) * nlay + l) * tcols + tcol] = hn_0;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gbar0 + 8 * st) : "memory");  // split-phase: waited for at the end of the row

    // -------------------------------------------------------------------------------- rvor, dive, row R (pm:2388, 2435)
    const double rv_0 = SELM(act && (f_own & F_PE), (v_0 - vW_0 - u_0 + u_m1) * D.i_dl);
    const double dv_0 = SELM(act, (uE_0 - u_0 + v_p1 - v_0) * D.i_dl);


    // -------------------------------------------------------------------------------- d2hx, pvor, row R; d2hy, row R-1
    const double hnE_0 = shdn(hn_0), hnW_0 = shup(hn_0);
    double d2x_0 = SELM((fw_0 & (F_N << 16)) && (fw_0 & (F_N << 8)) && (f_own & F_N), hnE_0 + hnW_0 - hn_0 * 2.0);
    if (ocrp && (hnE_0 < D.two_hs || hnW_0 < D.two_hs || hn_0 < D.two_hs)) d2x_0 = 0.0;
    d2x_0 = SELM(act, d2x_0);
    double d2y_m1 = SELM((f_own & F_N) && (fw_m2 & F_N) && (fw_m1 & F_N), hn_0 + hn_m2 - hn_m1 * 2.0);
    if (ocrp && (hn_0 < D.two_hs || hn_m2 < D.two_hs || hn_m1 < D.two_hs)) d2y_m1 = 0.0;
    d2y_m1 = SELM(fw_m1 & F_ACT, d2y_m1);
    double pv_0;
    {
      const double have = hn_0 + hnW_0 + hnW_m1 + hn_m1;
      const double msum = MKN((uint8_t)f_own) + MKN((uint8_t)(fw_0 >> 8)) + MKN((uint8_t)(fw_m1 >> 8)) + MKN((uint8_t)fw_m1);
      pv_0 = SELM(act, SELM(f_own & F_PI, fcor_0 + rv_0 * D.uadv) * msum / have);
    }

    // -------------------------------------------------------------------------------- Leith viscosity, row R-1 (pm:2477-2502)
    const double rvE_0 = shdn(rv_0), rvW_0 = shup(rv_0), dvE_0 = shdn(dv_0), dvW_0 = shup(dv_0);
    double vcc_m1 = 0.0, vll_m1 = 0.0;
    if (VISC) {
      const double r_bl = rv_m1, r_br = rvE_m1, r_tr = rvE_0, r_tl = rv_0, rbll = rvW_m1, rbbl = rv_m2;
      const double d_cc = dv_m1, d_ri = dvE_m1, d_to = dv_0, d_le = dvW_m1, d_bl = dvW_m2, d_bo = dv_m2;
      const double a1 = (r_br - r_bl) * (r_br - r_bl), a3 = (r_tl - r_bl) * (r_tl - r_bl);
      const double b1 = (d_cc - d_le) * (d_cc - d_le), b3 = (d_cc - d_bo) * (d_cc - d_bo);
      const double tll = a1 + (r_bl - rbll) * (r_bl - rbll) + a3 + (r_bl - rbbl) * (r_bl - rbbl) + b1 + (d_bo - d_bl) * (d_bo - d_bl) + b3 +
                         (d_le - d_bl) * (d_le - d_bl);
      const double tcc = a1 + (r_tr - r_tl) * (r_tr - r_tl) + a3 + (r_tr - r_br) * (r_tr - r_br) + (d_ri - d_cc) * (d_ri - d_cc) + b1 +
                         (d_to - d_cc) * (d_to - d_cc) + b3;
      const bool a = fw_m1 & F_ACT;
      vll_m1 = SELM(a, sqrt(tll) * D.dvis * D.dl * D.dl + D.bvis);
      vcc_m1 = SELM(a, sqrt(tcc) * D.dvis * D.dl * D.dl + D.bvis);
    }

    // -------------------------------------------------------------------------------- momentum
    const double hvW_0 = shup(hv_0);  // west neighbour of the old h_v, row R (rows R-1, R-2 later)
    double fl2_m1 = 0.0, fl2E_m1 = 0.0;  // v-first: new h_v of row R-1 and its west neighbour
    if (UFIRST) {
      // ---- u at row R-2 (pm:1422-1503) ----
      const size_t c2 = (size_t)max(R - 2, 0) * NX + x;
      const bool a2 = fw_m2 & F_ACT;
      qu.h1 = SGF(S_DX1, 0); qu.h2 = SGF(S_DX2, 0); qu.h3 = SGF(S_DX3, 0);
      qu.tw_a = qu.tw_b = qu.te_a = qu.te_b = 0.0;
      if (wind) {
        qu.tw_b = SG(S_TTXU, 0); qu.tw_a = SG(S_TTXU, -1);
        if (D.has_nudg) { qu.te_b = SG(S_TTYU, 0); qu.te_a = SG(S_TTYU, -1); }
      }
      qu.tb = D.has_bdrg ? SG(S_TBX, 0) : 0.0;
      qu.tu = D.has_tdrg ? SG(S_TUX, 0) : 0.0;
      qu.fn = qu.nud = 0.0;
      if (D.has_nudg) { qu.fn = SG(S_FNU, 0); qu.nud = SG(S_NUDU, 0); }
      qu.bodf = bodf_u;
      qu.mask = MKU((uint8_t)fw_m2);
      qu.h_a = hnW_m2; qu.h_b = hn_m2;
      qu.m_far = shup(mo_m2); qu.m_here = mo_m2;
      qu.pv_a = pv_m2; qu.f_a0 = hv_m2; qu.f_a1 = hvW_m2;
      qu.pv_b = pv_m1; qu.f_b0 = hv_m1; qu.f_b1 = hvW_m1;
      qu.old = u_m2;
      qu.vcc_here = vcc_m2; qu.dv_here = dv_m2; qu.vcc_far = shup(vcc_m2); qu.dv_far = dvW_m2;
      qu.vll_far = vll_m1; qu.rv_far = rv_m1; qu.vll_here = vll_m2; qu.rv_here = rv_m2;
      qu.d2_far = shup(d2x_m2); qu.d2_here = d2x_m2;
      double un, hun, dm;
      momentum<true, VISC, MASKED>(D, qu, un, hun, dm);
      hun = SELM(a2, hun);
      const bool sto = col_ok && (!MASKED || a2) && (R - 2 >= ya) && (R - 2 <= yb);
      if (sto) {
        __stcs(o_u + c2, un);
        __stcs(o_hu + c2, hun);
        __stcs(o_dx + c2, dm);
      }
      fl2_m2 = hun;
      fl2E_m2 = shdn(hun);
      // ---- v at row R-2 (pm:1505-1591), using the new h_u of rows R-2 and R-3 ----
      qv.h1 = SGF(S_DY1, 0); qv.h2 = SGF(S_DY2, 0); qv.h3 = SGF(S_DY3, 0);
      qv.tw_a = qv.tw_b = qv.te_a = qv.te_b = 0.0;
      if (wind) {
        qv.tw_b = SG(S_TTYV, 0); qv.tw_a = SG(S_TTYVS, 0);
        if (D.has_nudg) { qv.te_b = SG(S_TTXV, 0); qv.te_a = SG(S_TTXVS, 0); }
      }
      qv.tb = D.has_bdrg ? SG(S_TBY, 0) : 0.0;
      qv.tu = D.has_tdrg ? SG(S_TUY, 0) : 0.0;
      qv.fn = qv.nud = 0.0;
      if (D.has_nudg) { qv.fn = SG(S_FNV, 0); qv.nud = SG(S_NUDV, 0); }
      qv.bodf = bodf_v;
      qv.mask = MKV((uint8_t)fw_m2);
      qv.h_a = hn_m2; qv.h_b = hn_m3;
      qv.m_far = mo_m3; qv.m_here = mo_m2;
      qv.pv_a = pv_m2; qv.f_a0 = fl2_m2; qv.f_a1 = fl2_m3;
      qv.pv_b = shdn(pv_m2); qv.f_b0 = fl2E_m2; qv.f_b1 = fl2E_m3;
      qv.old = v_m2;
      qv.vcc_here = vcc_m2; qv.dv_here = dv_m2; qv.vcc_far = vcc_m3; qv.dv_far = dv_m3;
      qv.vll_far = shdn(vll_m2); qv.rv_far = rvE_m2; qv.vll_here = vll_m2; qv.rv_here = rv_m2;
      qv.d2_far = d2y_m3; qv.d2_here = d2y_m2;
      double vn, hvn;
      momentum<false, VISC, MASKED>(D, qv, vn, hvn, dm);
      if (sto) {
        __stcs(o_v + c2, vn);
        __stcs(o_hv + c2, SELM(a2, hvn));
        __stcs(o_dy + c2, dm);
      }
      fl2_m3 = fl2_m2;
      fl2E_m3 = fl2E_m2;
    } else {
      // ---- v at row R-1 (pm:1505-1591), old h_u ----
      const size_t c1 = (size_t)max(R - 1, 0) * NX + x, c2 = (size_t)max(R - 2, 0) * NX + x;
      const bool a1 = fw_m1 & F_ACT;
      qv.h1 = SGF(S_DY1, 0); qv.h2 = SGF(S_DY2, 0); qv.h3 = SGF(S_DY3, 0);
      qv.tw_a = qv.tw_b = qv.te_a = qv.te_b = 0.0;
      if (wind) {
        qv.tw_b = SG(S_TTYV, 0); qv.tw_a = SG(S_TTYVS, 0);
        if (D.has_nudg) { qv.te_b = SG(S_TTXV, 0); qv.te_a = SG(S_TTXVS, 0); }
      }
      qv.tb = D.has_bdrg ? SG(S_TBY, 0) : 0.0;
      qv.tu = D.has_tdrg ? SG(S_TUY, 0) : 0.0;
      qv.fn = qv.nud = 0.0;
      if (D.has_nudg) { qv.fn = SG(S_FNV, 0); qv.nud = SG(S_NUDV, 0); }
      qv.bodf = bodf_v;
      qv.mask = MKV((uint8_t)fw_m1);
      qv.h_a = hn_m1; qv.h_b = hn_m2;
      qv.m_far = mo_m2; qv.m_here = mo_m1;
      qv.pv_a = pv_m1; qv.f_a0 = hu_m1; qv.f_a1 = hu_m2;
      qv.pv_b = shdn(pv_m1); qv.f_b0 = huE_m1; qv.f_b1 = huE_m2;
      qv.old = v_m1;
      qv.vcc_here = vcc_m1; qv.dv_here = dv_m1; qv.vcc_far = vcc_m2; qv.dv_far = dv_m2;
      qv.vll_far = shdn(vll_m1); qv.rv_far = rvE_m1; qv.vll_here = vll_m1; qv.rv_here = rv_m1;
      qv.d2_far = d2y_m2; qv.d2_here = d2y_m1;
      double vn, hvn, dm;
      momentum<false, VISC, MASKED>(D, qv, vn, hvn, dm);
      hvn = SELM(a1, hvn);
      if (col_ok && (!MASKED || a1) && (R - 1 >= ya) && (R - 1 <= yb)) {
        __stcs(o_v + c1, vn);
        __stcs(o_hv + c1, hvn);
        __stcs(o_dy + c1, dm);
      }
      fl2_m1 = hvn;
      fl2E_m1 = shup(hvn);  // WEST neighbour in this order
      // ---- u at row R-2 (pm:1422-1503), using the new h_v of rows R-2 and R-1 ----
      const bool a2 = fw_m2 & F_ACT;
      qu.h1 = SGF(S_DX1, 0); qu.h2 = SGF(S_DX2, 0); qu.h3 = SGF(S_DX3, 0);
      qu.tw_a = qu.tw_b = qu.te_a = qu.te_b = 0.0;
      if (wind) {
        qu.tw_b = SG(S_TTXU, 0); qu.tw_a = SG(S_TTXU, -1);
        if (D.has_nudg) { qu.te_b = SG(S_TTYU, 0); qu.te_a = SG(S_TTYU, -1); }
      }
      qu.tb = D.has_bdrg ? SG(S_TBX, 0) : 0.0;
      qu.tu = D.has_tdrg ? SG(S_TUX, 0) : 0.0;
      qu.fn = qu.nud = 0.0;
      if (D.has_nudg) { qu.fn = SG(S_FNU, 0); qu.nud = SG(S_NUDU, 0); }
      qu.bodf = bodf_u;
      qu.mask = MKU((uint8_t)fw_m2);
      qu.h_a = hnW_m2; qu.h_b = hn_m2;
      qu.m_far = shup(mo_m2); qu.m_here = mo_m2;
      qu.pv_a = pv_m2; qu.f_a0 = fl2_m2; qu.f_a1 = fl2E_m2;
      qu.pv_b = pv_m1; qu.f_b0 = fl2_m1; qu.f_b1 = fl2E_m1;
      qu.old = u_m2;
      qu.vcc_here = vcc_m2; qu.dv_here = dv_m2; qu.vcc_far = shup(vcc_m2); qu.dv_far = dvW_m2;
      qu.vll_far = vll_m1; qu.rv_far = rv_m1; qu.vll_here = vll_m2; qu.rv_here = rv_m2;
      qu.d2_far = shup(d2x_m2); qu.d2_here = d2x_m2;
      double un, hun;
      momentum<true, VISC, MASKED>(D, qu, un, hun, dm);
      if (col_ok && (!MASKED || a2) && (R - 2 >= ya) && (R - 2 <= yb)) {
        __stcs(o_u + c2, un);
        __stcs(o_hu + c2, SELM(a2, hun));
        __stcs(o_dx + c2, dm);
      }
      fl2_m2 = fl2_m1;
      fl2E_m2 = fl2E_m1;
    }

    // -------------------------------------------------------------------------------- mont, row R (pm:2351-2383)
    mbar_wait(gbar0 + 8 * st, (unsigned)(((R - R0) >> 1) & 1));  // every layer of this column group has published hn(R)
    double mpot;
    if (ocrp) {
      mpot = hn_0 + D.hmin * (1.0 - MKN((uint8_t)f_own));
      mpot = cube(D.hsal / mpot);
      mpot = mpot * (-D.ocrp * D.i_ns * D.hsal * MKN((uint8_t)f_own));
    } else {
      mpot = -0.0;
    }
    mpot = mpot - 0.0;
    {
      const double *col = sh_h + (size_t)(R & 1) * nlay * tcols + tcol;
      double hcol = 0.0;
#pragma unroll
      for (int i = 0; i < (NL > 0 ? NL : kMaxLay); i++) {
        if (i < nlay) {
          const double hi = col[(size_t)i * tcols];
          if (i < l) mpot = mpot - cb[i] * hi;
          hcol = hcol + hi;
        }
      }
      mpot = hcol - SGF(S_HTH, 0) + mpot;
    }
    const double mo_0 = SELM(act, mpot + kin * (uE_0 * uE_0 + u_0 * u_0 + v_p1 * v_p1 + v_0 * v_0));

    __syncwarp();
    if (R + 2 <= R1) issue(R + 2);  // refill this stage (all lanes have read it)

    // -------------------------------------------------------------------------------- roll the pipeline
    u_m2 = u_m1; u_m1 = u_0;
    v_m2 = v_m1; v_m1 = v_0; v_0 = v_p1; vW_0 = vW_p1;
    hv_m2 = hv_m1; hv_m1 = hv_0; hv_0 = hv_p1;
    hvW_m2 = hvW_m1; hvW_m1 = hvW_0;
    hu_m2 = hu_m1; hu_m1 = hu_0;
    huE_m2 = huE_m1; huE_m1 = huE_0;
    hn_m3 = hn_m2; hn_m2 = hn_m1; hn_m1 = hn_0;
    hnW_m2 = hnW_m1; hnW_m1 = hnW_0;
    rv_m2 = rv_m1; rv_m1 = rv_0;
    rvE_m2 = rvE_m1; rvE_m1 = rvE_0;
    rvW_m1 = rvW_0;
    dv_m3 = dv_m2; dv_m2 = dv_m1; dv_m1 = dv_0;
    dvE_m1 = dvE_0;
    dvW_m2 = dvW_m1; dvW_m1 = dvW_0;
    mo_m3 = mo_m2; mo_m2 = mo_m1; mo_m1 = mo_0;
    pv_m2 = pv_m1; pv_m1 = pv_0;
    vcc_m3 = vcc_m2; vcc_m2 = vcc_m1;
    vll_m2 = vll_m1;
    d2x_m2 = d2x_m1; d2x_m1 = d2x_0;
    d2y_m3 = d2y_m2; d2y_m2 = d2y_m1;
    fw_m2 = fw_m1; fw_m1 = fw_0;
  };
  constexpr unsigned kAllMasks = 0x3f | (0x3f << 8) | (0x3f << 16);
#pragma unroll 1
  for (int R = R0; R <= R1; R++) {
    mbar_wait(bar0 + 8 * ((R - R0) & 1), (unsigned)(((R - R0) >> 1) & 1));  // staged inputs of front row R have landed
    const unsigned f_own = f_next;
    f_next = D.flags[(size_t)(R + 1) * NX + x];
    const unsigned fw_0 = f_own | (__shfl_up_sync(0xffffffffu, f_own, 1) << 8) | (__shfl_down_sync(0xffffffffu, f_own, 1) << 16);
    const bool open_water = __all_sync(0xffffffffu, ((fw_0 & fw_m1 & fw_m2) & kAllMasks) == kAllMasks);
    if (open_water) row(std::false_type{}, R, f_own, fw_0);
    else row(std::true_type{}, R, f_own, fw_0);
  }
#undef SGF
#undef SELM
#undef MKN
#undef MKU
#undef MKV
#undef SG
}

}  // namespace

namespace {
// Stream table for one step: where each staged row segment comes from (pointers follow the state
// double buffering, so the table is rebuilt per launch; it is ~20 entries).
StreamTab make_streams(const Dev &D, bool ufirst) {
  StreamTab T;
  memset(&T, 0, sizeof T);
  for (auto &s : T.slot) s = -1;
  auto add = [&](int stream, const double *base, int lag, int lstride) {
    const int k = T.n++;
    T.slot[stream] = (signed char)k;
    T.base[k] = base;
    T.lag[k] = (short)lag;
    T.lstride[k] = (unsigned char)lstride;
  };
  const size_t pl = D.plane, nl = (size_t)D.nlay;
  const int LV = ufirst ? 2 : 1;  // row lag of the v stage
  add(S_HU, D.h_u, 0, 1); add(S_HV, D.h_v, -1, 1); add(S_HL, D.hlay, 0, 1);
  add(S_R1, D.rs1, 0, 1); add(S_R2, D.rs2, 0, 1);
  add(S_U, D.u, 0, 1); add(S_V, D.v, -1, 1);
  add(S_DX1, D.dx1, 2, 1); add(S_DX2, D.dx2, 2, 1); add(S_DX3, D.dx3, 2, 1);
  add(S_DY1, D.dy1, LV, 1); add(S_DY2, D.dy2, LV, 1); add(S_DY3, D.dy3, LV, 1);
  add(S_FCOR, D.fcor, 0, 0); add(S_HTH, D.h_th, 0, 0);
  if (D.has_hdot) add(S_HDOT, D.hdot, 0, 1);
  if (D.has_nudg) {
    add(S_FNN, D.fnud, 0, 1); add(S_NUDN, D.nudg, 0, 0);
    add(S_FNU, D.fnud + nl * pl, 2, 1); add(S_NUDU, D.nudg + pl, 2, 0);
    add(S_FNV, D.fnud + 2 * nl * pl, LV, 1); add(S_NUDV, D.nudg + 2 * pl, LV, 0);
  }
  if (D.has_bdrg) { add(S_TBX, D.tb3d, 2, 2); add(S_TBY, D.tb3d + pl, LV, 2); }
  if (D.has_tdrg) { add(S_TUX, D.tu3d, 2, 2); add(S_TUY, D.tu3d + pl, LV, 2); }
  T.n_nowind = T.n;
  if (D.has_wind) {  // wind-only streams last: a layer without wind stress stages only the first n_nowind
    add(S_TTXU, D.tt3d, 2, 2);
    add(S_TTYV, D.tt3d + pl, LV, 2); add(S_TTYVS, D.tt3d + pl, LV + 1, 2);
    if (D.has_nudg) { add(S_TTYU, D.tt3d + pl, 2, 2); add(S_TTXV, D.tt3d, LV, 2); add(S_TTXVS, D.tt3d, LV + 1, 2); }
  }
  return T;
}
size_t fused_smem_bytes(int nlay, int groups, int nstreams) {
  const size_t nwarps = (size_t)groups * nlay;
  return (size_t)2 * nlay * groups * 32 * 8 + nwarps * 16 + (size_t)groups * 16 + nwarps * 2 * (size_t)nstreams * kSeg * 8;
}
template <bool UF, bool VI, int NL>
int launch(const Dev &in, const Dev &out, const StreamTab &T, dim3 grid, dim3 block, size_t shmem, cudaStream_t s) {
  static size_t configured = 0;
  if (shmem > configured) {
    if (cudaFuncSetAttribute(k_fused_step<UF, VI, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem) != cudaSuccess) return -61;
    configured = shmem;
  }
  k_fused_step<UF, VI, NL><<<grid, block, shmem, s>>>(in, out, T, cfg.groups, cfg.rows_per_chunk, cfg.wind_layers);
  return 0;
}
}  // namespace

int fused_configure(const Dev &D, const beom_params &P, int nmir, int nranks, bool *enabled) {
  *enabled = false;
  cfg = FusedCfg();
  (void)nranks;
  if (nmir != 0) return 0;  // periodic aliases: split path (for now)
  if (P.rgld > 0.5 || P.svis > 0.0 || P.variant != BEOM_VARIANT_STANDARD) return 0;
  if (D.has_tide) return 0;
  if (D.nlay > kMaxLay) return 0;
  const double dtd8 = P.dt / 24.0 / 3600.0;
  const bool every_step = std::floor(P.dt3d / dtd8 + 0.5) < 2.0;  // n_3d = 1
  if (P.dvis > 1.e-3) {
    if (!every_step) return 0;  // v_cc/v_ll persist between updates: split path
    cfg.visc = true;
  } else if (P.dvis == 0.0 && P.bvis == 0.0) {
    cfg.visc = false;           // v_cc = v_ll = 0 for ever: the viscous terms vanish
  } else if (P.dvis == 0.0) {
    cfg.visc = true;            // constant bvis: sqrt(.)*0*dl*dl + bvis
  } else {
    return 0;                   // 0 < dvis <= 1e-3: viscosity frozen at its step-3 value (reference quirk)
  }
  if ((D.has_bdrg || D.has_tdrg) && !every_step) return 0;
  // wind stress reaches only the layers that can hold a share of hsbl
  cfg.wind_layers = 0;
  if (D.has_wind) cfg.wind_layers = (P.ocrp > 0.5) ? ((1 << D.nlay) - 1) : 1;

  int dev = 0, sms = 148, max_smem = 227 * 1024;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int nstreams = make_streams(D, true).n;
  if (nstreams > 32) return 0;
  const int width = D.x_hi - D.x_lo + 1, rows = D.y_hi - D.y_lo + 1;
  int groups = std::max(1, std::min(std::min((kMaxThreads / 32) / D.nlay, 15), (width + kUse - 1) / kUse));  // <= 12 warps per CTA, <= 15 named barriers
  while (groups > 1 && fused_smem_bytes(D.nlay, groups, nstreams) > (size_t)max_smem - 1024) groups--;
  if (fused_smem_bytes(D.nlay, groups, nstreams) > (size_t)max_smem - 1024) return 0;
  cfg.groups = groups;
  cfg.strips = (width + groups * kUse - 1) / (groups * kUse);
  int chunks = std::max(1, (2 * sms) / cfg.strips);  // about two CTAs' worth of work per SM
  chunks = std::min(chunks, std::max(1, rows / 16));
  cfg.rows_per_chunk = (rows + chunks - 1) / chunks;
  cfg.chunks = (rows + cfg.rows_per_chunk - 1) / cfg.rows_per_chunk;
  cfg.ok = true;
  *enabled = true;
  return 0;
}

bool fused_supports(bool first_three, bool upst) { return cfg.ok && !first_three && upst; }

int fused_step(const Dev &in, const Dev &out, int tstp, bool first_three, cudaStream_t s, int *nlaunch) {
  if (!cfg.ok || first_three) return -1;
  const dim3 grid((unsigned)cfg.strips, (unsigned)cfg.chunks, 1), block((unsigned)(cfg.groups * 32 * in.nlay), 1, 1);
  const bool ufirst = (tstp % 2 == 0);
  const StreamTab T = make_streams(in, ufirst);
  const size_t shmem = fused_smem_bytes(in.nlay, cfg.groups, T.n);
  int rc;
#define BEOM_LAUNCH(NL)                                                                                                          \
  (cfg.visc ? (ufirst ? launch<true, true, NL>(in, out, T, grid, block, shmem, s) : launch<false, true, NL>(in, out, T, grid, block, shmem, s)) \
            : (ufirst ? launch<true, false, NL>(in, out, T, grid, block, shmem, s) : launch<false, false, NL>(in, out, T, grid, block, shmem, s)))
  switch (in.nlay) {
    case 1: rc = BEOM_LAUNCH(1); break;
    case 2: rc = BEOM_LAUNCH(2); break;
    case 3: rc = BEOM_LAUNCH(3); break;
    case 4: rc = BEOM_LAUNCH(4); break;
    default: rc = BEOM_LAUNCH(0); break;
  }
#undef BEOM_LAUNCH
  if (rc) return rc;
  *nlaunch = 1;
  return cudaGetLastError() == cudaSuccess ? 0 : -60;
}

}  // namespace beom
