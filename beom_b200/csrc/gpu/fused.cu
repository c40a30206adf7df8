// fused.cu -- host side of the fused single-pass time step: decides whether a case can run on it,
// sizes the grid / shared memory, builds the per-launch stream table and dispatches the instantiation
// (kernel: fused_kernel.cuh; instantiations: fused_inst_*.cu).
#include "fused.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fused_kernel.cuh"

namespace beom {

using namespace fusedk;

namespace {

// The specialised instantiations that exist: option set (fused_kernel.cuh FB_*), layers, viscosity, column groups.  Whatever
// is not listed -- or does not fit the shared memory with its wind streams -- runs the general instantiation.
struct SpecEntry {
  int feat, nlay;
  bool visc;
  int groups;
  int (*launch)(const FusedLaunch &, bool ufirst, bool gene0);
};
const SpecEntry kSpec[] = {
    {0, 1, true, kMaxWarps / 1, fused_launch_lean1}, {0, 2, true, kMaxWarps / 2, fused_launch_lean2},
    {0, 3, true, kMaxWarps / 3, fused_launch_lean3}, {0, 4, true, BEOM_LEAN4_GROUPS, fused_launch_lean4},
    {FB_NUDG | FB_OCRP, 2, true, 7, fused_launch_spec_3_2_1},            // sill_exchange3D / 2D
    {FB_NUDG | FB_OCRP, 4, true, 3, fused_launch_spec_3_4_1},            // bench.py --workload sill_like
    {FB_NUDG, 2, true, 7, fused_launch_spec_1_2_1}, {FB_NUDG, 4, true, 3, fused_launch_spec_1_4_1},
    {FB_NUDG | FB_OCRP | FB_BDRG, 2, true, 6, fused_launch_spec_7_2_1},  // ... with bottom drag
    {FB_NUDG | FB_OCRP | FB_BDRG, 4, true, 3, fused_launch_spec_7_4_1},
    {FB_BDRG, 1, false, 15, fused_launch_spec_4_1_0},                    // stommel1948: wind, linear drag, no viscosity
};

struct FusedCfg {
  bool ok = false;
  const SpecEntry *spec = nullptr;  // compile-time specialised instantiation (see fused_configure), or the general one
  int groups = 1;     // column groups (warps per layer) per CTA
  int strips = 1;     // CTAs along x
  int chunks = 1;     // CTAs along y
  int rows_per_chunk = 1;  // packed: see chunk_rows()
  bool visc = false;    // Leith viscosity recomputed every step
  int wind_layers = 0;  // bit l set: tt3d of layer l may be non-zero
  uint8_t *open = nullptr;  // [tiles][NY]: bit 0 = rows R-2..R open water on the tile's 32 columns, bit 1 = the same for R..R+3
  unsigned *open4 = nullptr;  // [tiles][4][open4_words]: bit g of the row-group bitmap of residue r = bit 1 of open[4g + r]
  int open4_words = 0;
} cfg;

// open-water summary of the flag plane, one byte per (warp tile, row)
__global__ void k_open_rows(const uint8_t *__restrict__ flags, uint8_t *__restrict__ open, int NX, int NY, int x_lo) {
  const int tile = blockIdx.x, R = blockIdx.y * blockDim.y + threadIdx.y, lane = threadIdx.x;
  if (R >= NY) return;
  const int x = x_lo + tile * kUse - kHalo + lane;  // the warp's 32 columns (fused_kernel.cuh)
  bool ok = true;
  for (int d = 0; d < 3; d++) {
    const int r = R - d;
    ok = ok && r >= 0 && x >= 0 && x < NX && (flags[(size_t)r * NX + x] & 0x3f) == 0x3f;
  }
  const bool all = __all_sync(0xffffffffu, ok);
  if (lane == 0) open[(size_t)tile * NY + R] = all ? 1 : 0;
}
__global__ void k_open_groups(uint8_t *__restrict__ open, int NY, int ntiles) {
  const int R = blockIdx.x * blockDim.x + threadIdx.x, tile = blockIdx.y;
  if (R >= NY || tile >= ntiles) return;
  uint8_t *o = open + (size_t)tile * NY;
  bool all = R + 3 < NY;
  for (int k = 0; all && k < 4; k++) all = o[R + k] & 1;
  if (all) o[R] |= 2;  // bit 1 is read by nobody in this kernel
}

// one bitmap per residue of the group's first row mod 4 (a chunk's groups start at its own first row): bit g of
// open4[tile][res] says rows 4g + res - 2 .. 4g + res + 3 are open water on the tile's 32 columns
__global__ void k_open_bits(const uint8_t *__restrict__ open, unsigned *__restrict__ open4, int NY, int nwords) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x, tile = blockIdx.y, res = blockIdx.z;
  if (w >= nwords) return;
  unsigned bits = 0;
  for (int b = 0; b < 32; b++) {
    const int R = (w * 32 + b) * 4 + res;
    if (R < NY && (open[(size_t)tile * NY + R] & 2)) bits |= 1u << b;
  }
  open4[((size_t)tile * 4 + res) * nwords + w] = bits;
}

// Rows of the y-chunks of a strip.  A chunk runs its rows + 5 (3 of pipeline lead-in, 2 of tail) in groups of 4, so chunks of
// 4k + 3 rows waste nothing: `rows` are cut into `chunks` pieces of base or base + 4 rows, base = 3 mod 4 (the first `nbig` get
// the 4 more; the last one also takes the 0..3 rows that are left).  Returns the kernel's argument: base | nbig << 16.
int chunk_rows(int rows, int chunks) {
  const int q = rows / chunks;
  if (chunks <= 1 || q < 7) return rows;  // one piece (the last chunk takes everything that is left)
  const int base = ((q - 3) & ~3) + 3;
  const int nbig = (rows - base * chunks) / 4;
  return base | (nbig << 16);
}

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
  add(S_HU, D.h_u, 0, 1); add(S_HV, D.h_v, -1, 1); add(S_U, D.u, 0, 1); add(S_V, D.v, -1, 1);
  add(S_HL, D.hlay, 0, 1); add(S_R1, D.rs1, 0, 1); add(S_R2, D.rs2, 0, 1);
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
size_t fused_smem_bytes(int nlay, int groups, const StreamTab &T, int wind_layers) {
  return smem_plan(nlay, groups, T.n, T.n_nowind, wind_layers).total;
}
}  // namespace

const char *fused_variant() {
  static char b[96];
  if (!cfg.ok) return "none";
  if (cfg.spec) snprintf(b, sizeof b, "specialised (options %d, %d layers, %d column groups%s)", cfg.spec->feat, cfg.spec->nlay, cfg.spec->groups, cfg.spec->visc ? "" : ", no viscosity");
  else snprintf(b, sizeof b, "general (%d column groups)", cfg.groups);
  return b;
}

void fused_release() {
  if (cfg.open) cudaFree(cfg.open);
  if (cfg.open4) cudaFree(cfg.open4);
  cfg = FusedCfg();
}

int fused_configure(const Dev &D, const beom_params &P, int nmir, int nranks, bool *enabled) {
  *enabled = false;
  fused_release();
  (void)nranks;
  if (nmir != 0) return 0;  // periodic aliases: split path (for now)
  if (P.rgld > 0.5 || P.svis > 0.0 || P.variant != BEOM_VARIANT_STANDARD) return 0;
  // tidal targets: the caller hands the step fnud + tide term (k_tide_targets); exact only when the Ekman term of the
  // target, which the reference adds between the two, is absent or an exact zero; periodic images of the targets are not kept
  if (D.has_tide && ((D.has_wind && D.invf != 0.0) || P.xper > 0.5 || P.yper > 0.5)) return 0;
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
  // Specialised instantiations: generalized forward-backward (gene = 1 exactly; the start-up steps run their gene = 0 copy), no
  // hdot, top drag or body force, an option set and layer count that has been instantiated (kSpec), and its column-group count
  // must fit the shared memory with the wind streams of this case.  Everything else runs the general instantiation.
  bool bodf0 = true;
  for (int l = 0; l < D.nlay; l++) bodf0 = bodf0 && D.bodf[0][l] == 0.0 && D.bodf[1][l] == 0.0;
  const int feat = (D.has_nudg ? FB_NUDG : 0) | (P.ocrp > 0.5 ? FB_OCRP : 0) | (D.has_bdrg ? FB_BDRG : 0);
  const bool spec_ok = P.g_fb == 1.0 && !D.has_hdot && !D.has_tdrg && bodf0 && !(getenv("BEOM_FUSED_GENERAL") && atoi(getenv("BEOM_FUSED_GENERAL")) > 0);

  int dev = 0, sms = 148, max_smem = 227 * 1024;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const StreamTab T0 = make_streams(D, true);
  if (T0.n > 32) return 0;
  const int width = D.x_hi - D.x_lo + 1, rows = D.y_hi - D.y_lo + 1;
  const int wl = D.has_wind ? cfg.wind_layers : 0;
  cfg.spec = nullptr;
  if (spec_ok)
    for (const SpecEntry &e : kSpec)
      if (e.feat == feat && e.nlay == D.nlay && e.visc == cfg.visc && fused_smem_bytes(D.nlay, e.groups, T0, wl) <= (size_t)max_smem - 1024) cfg.spec = &e;
  // the specialised instantiations have their column-group count compiled in (idle groups on a narrow domain are harmless)
  int groups = cfg.spec ? cfg.spec->groups : std::max(1, std::min(kMaxWarps / D.nlay, (width + kUse - 1) / kUse));
  while (groups > 1 && fused_smem_bytes(D.nlay, groups, T0, wl) > (size_t)max_smem - 1024) groups--;
  if (fused_smem_bytes(D.nlay, groups, T0, wl) > (size_t)max_smem - 1024) return 0;
  cfg.groups = groups;
  cfg.strips = (width + groups * kUse - 1) / (groups * kUse);
  // At least two CTAs' worth of work per SM, and chunks of about 256 rows: many short waves instead of two long ones.  Measured
  // (8192 x 8192 x 4, profiles/r2_chunk_sweep.txt): 1 wave 11.6 ms, 2 waves 10.2, 4 waves 10.1, 8 waves 9.2, 16 waves 8.8, 32 waves
  // 8.9 -- CTAs that start together stay in step and hit HBM in bursts (all row loads of a wave within the same microsecond);
  // waves that end at different moments spread the requests out.
  // y-slabs of a multi-GPU run are short (1024 rows at 8 GPUs): there at least eight waves, of chunks no shorter than 64 rows
  // (a chunk runs 8 rows more than it owns: pipeline fill and the 4-row unrolling); 1024 rows: 4 chunks 1.307 ms, 16 chunks 1.243 ms.
  int chunks = std::max(1, (2 * sms + cfg.strips / 2) / cfg.strips);
  chunks = std::max(chunks, (rows + 128) / 256);
  chunks = std::max(chunks, std::min((8 * sms + cfg.strips - 1) / cfg.strips, rows / 64));
  if (const char *e = getenv("BEOM_FUSED_CHUNKS")) chunks = std::max(1, atoi(e));  // experiment: y-chunks per strip
  chunks = std::min(chunks, std::max(1, rows / 16));
  {  // the kernel keeps a chunk's open-water bitmap in kOWords words of 32 four-row groups (a chunk visits 8 rows more than it owns)
    const int max_rows = kOWords * 128 - 256;
    chunks = std::max(chunks, (rows + max_rows - 1) / max_rows);
  }
  {
    const int rpc = (rows + chunks - 1) / chunks;
    cfg.chunks = (rows + rpc - 1) / rpc;
    cfg.rows_per_chunk = chunk_rows(rows, cfg.chunks);
  }
  {
    const int ntiles = cfg.strips * groups;
    if (cudaMalloc(&cfg.open, (size_t)ntiles * D.NY) != cudaSuccess) return -62;
    k_open_rows<<<dim3((unsigned)ntiles, (unsigned)((D.NY + 7) / 8)), dim3(32, 8)>>>(D.flags, cfg.open, D.NX, D.NY, D.x_lo);
    k_open_groups<<<dim3((unsigned)((D.NY + 127) / 128), (unsigned)ntiles), 128>>>(cfg.open, D.NY, ntiles);
    cfg.open4_words = (D.NY + 127) / 128;
    if (cudaMalloc(&cfg.open4, (size_t)ntiles * 4 * cfg.open4_words * sizeof(unsigned)) != cudaSuccess) return -62;
    k_open_bits<<<dim3((unsigned)((cfg.open4_words + 63) / 64), (unsigned)ntiles, 4), 64>>>(cfg.open, cfg.open4, D.NY, cfg.open4_words);
    if (cudaDeviceSynchronize() != cudaSuccess) return -63;
  }
  cfg.ok = true;
  *enabled = true;
  return 0;
}

// The three start-up steps (plain forward-backward, gene = 0) run on the general instantiation after the caller has
// rebuilt the centred fluxes (private_mod.f95:2166-2177).
// Whatever fused_configure accepted either recomputes its upst-dependent terms every step (n_3d = 1: Leith viscosity,
// drag) or does not depend on upst at all (dvis = 0: v_cc = v_ll = bvis for ever, private_mod.f95:2268-2270), so steps
// between two 3-D updates run here as well -- the split path's stored v_cc / v_ll would never have been filled.
bool fused_supports(bool first_three, bool upst) {
  (void)first_three; (void)upst;
  return cfg.ok;
}

// part: 0 = every row of the slab, 1 = only the `edge` rows next to each neighbouring rank (two chunks), 2 = the rows
// in between (the part whose computation hides the halo exchange of part 1)
int fused_step(const Dev &in_, const Dev &out, int tstp, bool first_three, cudaStream_t s, int *nlaunch, int part, int edge) {
  (void)first_three;
  if (!cfg.ok) return -1;
  Dev in = in_;
  FusedLaunch a;
  int chunks = cfg.chunks, rpc = cfg.rows_per_chunk;
  if (part == 1) {
    chunks = 2;
    rpc = -edge;
  } else if (part == 2) {
    in.y_lo += edge;
    in.y_hi -= edge;
    const int rows = in.y_hi - in.y_lo + 1;
    if (rows <= 0) return 0;
    chunks = std::max(1, std::min(cfg.chunks, rows / 16));
    rpc = (rows + chunks - 1) / chunks;
    chunks = (rows + rpc - 1) / rpc;
    rpc = chunk_rows(rows, chunks);
  }
  static const bool swap = getenv("BEOM_FUSED_ORDER") && atoi(getenv("BEOM_FUSED_ORDER")) == 1;  // experiment: chunk index fastest
  a.grid = swap ? dim3((unsigned)chunks, (unsigned)cfg.strips, 1) : dim3((unsigned)cfg.strips, (unsigned)chunks, 1);
  a.block = dim3((unsigned)(cfg.groups * 32 * in.nlay), 1, 1);
  const bool ufirst = (tstp % 2 == 0);
  const StreamTab T = make_streams(in, ufirst);
  a.shmem = fused_smem_bytes(in.nlay, cfg.groups, T, in.has_wind ? cfg.wind_layers : 0);
  a.in = &in; a.out = &out; a.tab = &T; a.open = cfg.open; a.open4 = cfg.open4; a.open4_words = cfg.open4_words;
  a.groups = cfg.groups; a.rows_per_chunk = rpc; static const int jitter = getenv("BEOM_FUSED_JITTER") ? std::min(4095, std::max(0, atoi(getenv("BEOM_FUSED_JITTER")) / 64)) : 0;  // ns
  a.wind_layers = cfg.wind_layers | (swap ? 1 << 16 : 0) | (jitter << 17);
  a.stream = s;
  // the specialised instantiations assume gene = 1 (tstp >= 4 with g_fb = 1) or gene = 0 (the start-up steps) exactly
  const bool g0 = in.gene == 0.0;  // the start-up steps
  const bool spec = cfg.spec && (in.gene == 1.0 || g0);
  int rc;
  if (spec) {
    // BEOM_FMA=1: the copy compiled with FMA contraction (tolerance parity instead of bit-exact parity; off by default)
    static const bool fma = getenv("BEOM_FMA") && atoi(getenv("BEOM_FMA")) > 0;
    if (fma && !g0 && cfg.spec->feat == 0) {
      switch (in.nlay) {
        case 1: rc = fused_launch_lean1_fma(a, ufirst); break;
        case 2: rc = fused_launch_lean2_fma(a, ufirst); break;
        case 3: rc = fused_launch_lean3_fma(a, ufirst); break;
        default: rc = fused_launch_lean4_fma(a, ufirst); break;
      }
    } else {
      rc = cfg.spec->launch(a, ufirst, g0);
    }
  } else {
    rc = fused_launch_general(a, ufirst, cfg.visc, in.nlay);
  }
  if (rc) return rc;
  *nlaunch += 1;
  return cudaGetLastError() == cudaSuccess ? 0 : -60;
}

}  // namespace beom
