// beom_gpu.cu -- the C ABI of include/beom_gpu.h: context, HBM layout, upload/download, step
// sequencing (first_three_timesteps / gener_forward_backward, private_mod.f95:2151-2316).
// sm_100a only; no CPU path exists in this library.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <string>
#include <vector>

#include "comm.h"
#include "dev.cuh"
#include "fused.cuh"
#include "split.cuh"
#include "diag.cuh"
#include "rigid.cuh"
#include "gridinit.cuh"
#include "grid_index.h"
#include "layout.h"
#include "orphans.h"

using namespace beom;

namespace {

std::string g_err;
int fail(int code, const char *fmt, ...) {
  char b[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b, sizeof b, fmt, ap);
  va_end(ap);
  g_err = b;
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(-100 - (int)e_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct Ctx {
  bool ready = false;
  beom_params P;
  beom_gpu_options opt;
  int lm = 0, mm = 0, nlay = 0, ndeg = 0;
  int rank = 0, nranks = 1;
  int j0 = 1, j1 = 1;      // owned grid rows (inclusive)
  int p_lo = 1, p_hi = 0;  // vector points held on this device (owned + halo rows), inclusive
  int NX = 0, NY = 0;
  size_t plane = 0;
  Dev D;  // template for kernel arguments
  std::vector<void *> allocs;
  // host-side maps
  std::vector<int> cell_of_point;  // [ndeg+1], -1 = not on this device, -2 = orphan (periodic duplicate)
  int *d_cell = nullptr;
  std::vector<int> orphans;        // vector indices whose dense cell is a periodic mirror
  Orphans orph;                    // their hlay,u,v (host side): frozen, or following the sponge recurrence (orphans.h)
  int nmir = 0;
  int *d_mir_dst = nullptr, *d_mir_src = nullptr;
  SegDev *d_seg = nullptr;
  int nseg = 0;          // open-boundary segments whose points this rank owns
  bool obc_any = false;  // some rank has open-boundary segments: every rank takes part in the exchanges around k_obc
  // buffers
  uint8_t *flags = nullptr;
  double *st[5][2] = {{nullptr}};  // hlay,u,v,h_u,h_v x {A,B}
  int cur = 0;                     // which buffer holds the current state
  double *rs[3] = {nullptr}, *dx[4] = {nullptr}, *dy[4] = {nullptr};
  int rs_o = 0, dx_o = 0, dy_o = 0;  // index of the oldest slot
  double *stage = nullptr;           // (ndeg_local) * nlay * 3 doubles: vector-layout staging
  size_t stage_elems = 0;
  bool split_alloc = false;
  bool use_fused = false;
  bool any_taus = false;
  bool visc_valid = false;
  bool stress_const_done = false;  // tt3d = taus*layt does not change when ocrp = 0 and no drag
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;  // halo exchange of the fused step, overlapped with the interior rows
  cudaEvent_t ev_edge = nullptr, ev_comm = nullptr;
  double *halo_send[2] = {nullptr, nullptr}, *halo_recv[2] = {nullptr, nullptr};  // [lower, upper] neighbour
  size_t halo_cap = 0;
  size_t big_allocs = 0;
  bool torus = false;              // periodic domain whose aliases form a complete torus: deep ghost cells exist, the fused step may run
  bool ring = false;               // y-periodic domain split into y-slabs: the halo exchange is ring-closed (layout.h)
  HaloRows halo;                   // peers and rows of the packed halo exchange
  int own_first = 0, own_last = -1;  // vector points of the rows this rank owns (a contiguous range)
  double adv_ramp = 1.0, adv_gene = 0.0;  // beom_gpu_advance: ramp and gene of the previous step
  double *fnud_tide = nullptr;     // [3][nlay] planes: the step's relaxation targets with the tidal term (fused step only)
  float *rec_f32 = nullptr;        // [nlay] dense float32 planes: one diagnostic record
  float *rec_stage = nullptr;      // vector-layout staging of one layer of a record
  // asynchronous output records (beom_gpu_records_begin / _wait): two slots of [6][nlay][n] floats, device + page-locked host
  struct RecSlot {
    float *dev = nullptr, *host = nullptr;
    double *mm_dev = nullptr, *mm_host = nullptr;  // [nlay][2] min / max thickness
    cudaEvent_t ready = nullptr, done = nullptr;
    bool busy = false, diag = false, used = false;
    std::vector<double> orph;  // hlay,u,v of the frozen periodic duplicates when the set was begun
  } rec[2];
  int rec_head = 0, rec_pending = 0;  // oldest busy slot, number of busy slots
  float *h0r4 = nullptr;              // [nlay][ndeg] float32 rest thickness (h_0.bin)
  std::vector<float> h0r4_orph;       // [nlay][orphans]
  double *mm_partial = nullptr;
  cudaStream_t copy_stream = nullptr;
  double *diag_h0 = nullptr;       // [nlay] dense h_0 (beom_gpu_diagnostics)
  double *diag_partial = nullptr, *diag_out = nullptr;
  size_t diag_blocks = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  long long launches = 0;
  // CUDA graphs of whole steps for latency-bound grids (step_graphed): one executable graph per phase of the rotating buffers
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;            // kernel nodes
    int cur = 0, rs_o = 0, dx_o = 0, dy_o = 0;  // the rotating indices after the step
    bool stress_done = false;
  };
  std::map<unsigned, StepGraph> graphs;
  bool graphs_on = false;      // decided at init (one rank, no tides, no rigid lid, a latency-bound grid; BEOM_GRAPH overrides)
  double graph_gene = -1.0;    // the gene the graphs were captured with
  int direct_steps = 0;        // steady steps launched directly since the last upload (the first ones are never captured)
  long long graph_launches = 0;
  size_t win_first = 0, win_stride = 0;  // host state arrays cover points win_first .. win_first+win_stride-1 per layer
  int n_3d = 1;
  PiSolve pis;           // rigid lid: the wavefront solver's arrays (rigid.cuh)
  bool pis_ready = false;
  int pis_blocks = 0, pis_threads = 0;
  int *pi_iters = nullptr;  // device: sweeps of the last solve
};
Ctx g;
void drop_graphs_fwd();  // (defined next to step_graphed)

// Device allocation.  Large planes are staggered: allocation k starts (k mod 8) x 2 MiB + (k mod 5) x 4 KiB into its
// block, so that the ~90 row streams the fused step reads and writes at the same time do not all sit at the same
// offset modulo the page / TLB-set / DRAM-bank interleave (every field has the same size, so without the stagger
// their bases are congruent modulo any power of two up to 2 MiB x 8).
template <class T>
int dalloc(T **p, size_t n, bool zero = true) {
  void *q = nullptr;
  n += 1024;  // slack: the fused step stages whole-CTA row segments that may run past the last row of a plane
  size_t bytes = n * sizeof(T), shift = 0;
  static int stagger = -1;
  if (stagger < 0) {
    const char *e = getenv("BEOM_STAGGER");
    stagger = e ? atoi(e) : 1;
  }
  if (stagger && bytes >= ((size_t)64 << 20)) {
    const size_t k = g.big_allocs++;
    shift = (k % 8) * ((size_t)2 << 20) + (k % 5) * 4096;
  }
  CK(cudaMalloc(&q, bytes + shift));
  g.allocs.push_back(q);
  q = (char *)q + shift;
  if (zero) CK(cudaMemsetAsync(q, 0, bytes, g.stream));
  *p = (T *)q;
  return 0;
}

dim3 cell_grid(const Dev &D, int layers, dim3 block, int xpad = 0, int ypad = 0) {
  const int w = D.x_hi - D.x_lo + 1 + xpad, h = D.y_hi - D.y_lo + 1 + ypad;
  return dim3((unsigned)((w + block.x - 1) / block.x), (unsigned)((h + block.y - 1) / block.y), (unsigned)layers);
}
const dim3 kBlock(128, 2, 1);

// Upload one reference-layout array (vector of ndeg+1 doubles per plane) into dense planes.
int upload_planes(double *dense, const double *vec, int nplanes) {
  const int n = g.p_hi - g.p_lo + 1;
  if (n <= 0) return 0;
  for (int p = 0; p < nplanes; p++) {
    CK(cudaMemcpyAsync(g.stage, vec + (size_t)p * (g.ndeg + 1) + g.p_lo, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, g.stream));
    k_scatter<double><<<(n + 255) / 256, 256, 0, g.stream>>>(dense + (size_t)p * g.plane, g.stage, g.d_cell, g.p_lo, n);
    g.launches++;
    CK(cudaStreamSynchronize(g.stream));  // staging buffer is reused
  }
  return 0;
}

struct Item {
  double *p;
  int planes;
};

// Make the cells that shadow other cells consistent after `items' were written: periodic aliases on
// this device (k_mirror) and, with y-slabs, the G halo rows owned by the neighbouring ranks (one packed
// NCCL send/recv per neighbour).
int sync_fields(std::initializer_list<Item> items, bool remote = true, cudaStream_t st = nullptr) {
  if (!st) st = g.stream;
  if (g.nmir) {
    for (const Item &it : items) {
      if (!it.p) continue;
      k_mirror<<<(g.nmir + 127) / 128, 128, 0, st>>>(it.p, g.plane, it.planes, g.d_mir_dst, g.d_mir_src, g.nmir);
      g.launches++;
    }
  }
  if (g.nranks > 1 && remote) {
    HaloTab t;
    memset(&t, 0, sizeof t);
    for (const Item &it : items) {
      if (!it.p) continue;
      if (t.n >= kHaloMaxItems) return fail(-70, "sync_fields: too many fields");
      t.p[t.n] = it.p;
      t.planes[t.n] = it.planes;
      t.off[t.n] = t.total;
      t.total += it.planes;
      t.n++;
    }
    if (t.n == 0) return 0;
    const size_t per_plane = (size_t)G * g.NX, count = per_plane * t.total;
    if (count > g.halo_cap) return fail(-71, "sync_fields: halo buffer too small");
    const HaloRows &h = g.halo;  // plain slab neighbours, or the ring of a y-periodic domain (layout.h)
    const int lo = h.peer_lo, hi = h.peer_hi;
    const unsigned blocks = (unsigned)((per_plane + 255) / 256);
    k_halo_pack<<<dim3(blocks, (unsigned)t.total, 2), 256, 0, st>>>(t, g.plane, g.NX, h.send_lo, h.send_hi, g.halo_send[0], g.halo_send[1]);
    g.launches++;
    int rc = comm_exchange(g.halo_send[0], g.halo_recv[0], lo, g.halo_send[1], g.halo_recv[1], hi, count, st, &g_err);
    if (rc) return rc;
    k_halo_unpack<<<dim3(blocks, (unsigned)t.total, 2), 256, 0, st>>>(t, g.plane, g.NX, h.recv_lo, h.recv_hi, g.halo_recv[0], g.halo_recv[1], lo >= 0, hi >= 0);
    g.launches++;
  }
  return 0;
}
int mirror(double *field, int nplanes) { return sync_fields({{field, nplanes}}); }

void set_state_pointers(Dev &D) {
  D.hlay = g.st[0][g.cur]; D.u = g.st[1][g.cur]; D.v = g.st[2][g.cur]; D.h_u = g.st[3][g.cur]; D.h_v = g.st[4][g.cur];
  D.rs1 = g.rs[g.rs_o]; D.rs2 = g.rs[(g.rs_o + 1) % 3]; D.rs_new = g.rs[(g.rs_o + 2) % 3];
  D.dx1 = g.dx[g.dx_o]; D.dx2 = g.dx[(g.dx_o + 1) % 4]; D.dx3 = g.dx[(g.dx_o + 2) % 4]; D.dx_new = g.dx[(g.dx_o + 3) % 4];
  D.dy1 = g.dy[g.dy_o]; D.dy2 = g.dy[(g.dy_o + 1) % 4]; D.dy3 = g.dy[(g.dy_o + 2) % 4]; D.dy_new = g.dy[(g.dy_o + 3) % 4];
}

int alloc_split_buffers() {
  if (g.split_alloc) return 0;
  const size_t pl = g.plane, nl = (size_t)g.nlay;
  Dev &D = g.D;
  int rc;
  if ((rc = dalloc(&D.mont, pl * nl)) || (rc = dalloc(&D.rvor, pl * nl)) || (rc = dalloc(&D.pvor, pl * nl)) || (rc = dalloc(&D.dive, pl * nl)) ||
      (rc = dalloc(&D.d2hx, pl * nl)) || (rc = dalloc(&D.d2hy, pl * nl)) || (rc = dalloc(&D.v_cc, pl * nl)) || (rc = dalloc(&D.v_ll, pl * nl)))
    return rc;
  if (g.P.svis > 0.0)
    if ((rc = dalloc(&D.delu, pl * nl)) || (rc = dalloc(&D.delv, pl * nl)) || (rc = dalloc(&D.UU4, pl * nl)) || (rc = dalloc(&D.VV4, pl * nl))) return rc;
  g.split_alloc = true;
  return 0;
}

// distribute_stress, private_mod.f95:1921-2149
int run_stress() {
  Dev D = g.D;
  set_state_pointers(D);
  if (!(D.has_wind || D.has_bdrg || D.has_tdrg)) return 0;
  // wind only, no outcropping: layt = (1,0,...,0) and taus are constants, so tt3d computed once stays
  // bit-identical to what the reference recomputes every n_3d steps (private_mod.f95:1960-1967, 2136-2146)
  const bool constant = !D.has_bdrg && !D.has_tdrg && !(g.P.ocrp > 0.5);
  if (constant && g.stress_const_done) return 0;
  g.stress_const_done = constant;
  k_stress_fractions<<<cell_grid(D, 1, kBlock, 1, 1), kBlock, 0, g.stream>>>(D);
  g.launches++;
  if (D.has_bdrg) { k_stress_drag<<<cell_grid(D, 1, kBlock), kBlock, 0, g.stream>>>(D, 0); g.launches++; }
  if (D.has_tdrg) { k_stress_drag<<<cell_grid(D, 1, kBlock), kBlock, 0, g.stream>>>(D, 1); g.launches++; }
  k_stress_apply<<<cell_grid(D, g.nlay, kBlock), kBlock, 0, g.stream>>>(D);
  g.launches++;
  int rc;
  if (D.has_wind && (rc = mirror(D.tt3d, 2 * g.nlay))) return rc;
  if (D.has_bdrg && (rc = mirror(D.tb3d, 2 * g.nlay))) return rc;
  if (D.has_tdrg && (rc = mirror(D.tu3d, 2 * g.nlay))) return rc;
  CK(cudaGetLastError());
  return 0;
}

// surf_pressure's iteration across y-slabs (rigid.cuh).  A sweep of rank r needs the south row of the SAME sweep from rank r-1
// and the north row of the PREVIOUS sweep from rank r+1, so the ranks work in rounds: rank r runs sweep k in round 2k + r (its
// neighbours rest in that round), and every round ends with the exchange of one row per neighbour.  Sweeps are done in batches of
// kPiNA - 1 into rotating arrays; after a batch the per-sweep max-norms are reduced over the ranks and the first sweep that
// satisfies the reference's stopping rule is final (the later ones of the batch wrote into other arrays).
int pi_solve_slabs(const Dev &D) {
  PiSolve &S = g.pis;
  const int R = g.nranks, r = g.rank, NX = g.NX;
  const int lo = r > 0 ? r - 1 : -1, hi = r < R - 1 ? r + 1 : -1;
  const int B = kPiNA - 1;
  int rc;
  double *scratch = g.halo_recv[0];                 // (two rows + the reduction buffers: far below the halo capacity)
  double *red_lo = g.halo_recv[1], *red_hi = g.halo_recv[1] + kPiNA;
  int fin = 0;
  for (int b0 = 1; b0 <= S.maxiters && !fin; b0 += B) {
    const int b1 = std::min(b0 + B - 1, S.maxiters);
    auto sweep_of = [&](int rank, int t) {  // the sweep `rank' runs in round t (0 = it rests)
      const int q = t - rank;
      return (rank >= 0 && rank < R && q % 2 == 0 && q / 2 >= b0 && q / 2 <= b1) ? q / 2 : 0;
    };
    for (int t = 2 * b0; t <= 2 * b1 + R - 1; t++) {
      const int k = sweep_of(r, t), k_lo = sweep_of(lo, t), k_hi = sweep_of(hi, t);
      if (k) {
        k_pi_wave<<<(unsigned)g.pis_blocks, (unsigned)g.pis_threads, 0, g.stream>>>(D, S, k, k, 1);
        g.launches++;
      }
      // what this rank wrote in this round goes to both neighbours (a resting rank sends rows nobody looks at)
      const double *mine = S.X[(k ? k : 0) % kPiNA];
      double *from_lo = k_lo ? S.X[k_lo % kPiNA] + (size_t)(D.y_lo - 1) * NX : scratch;
      double *from_hi = k_hi ? S.X[k_hi % kPiNA] + (size_t)(D.y_hi + 1) * NX : scratch + NX;
      if ((rc = comm_exchange(mine + (size_t)D.y_lo * NX, from_lo, lo, mine + (size_t)D.y_hi * NX, from_hi, hi, (size_t)NX, g.stream, &g_err))) return rc;
    }
    for (int it = 0; it < R - 1; it++) {  // max over the ranks of every sweep's max-norm
      if ((rc = comm_exchange(S.md, red_lo, lo, S.md, red_hi, hi, (size_t)kPiNA, g.stream, &g_err))) return rc;
      k_pi_max_merge<<<1, 32, 0, g.stream>>>(S.md, red_lo, red_hi, lo >= 0, hi >= 0);
      g.launches++;
    }
    k_pi_verdict<<<1, 32, 0, g.stream>>>(S, b0, b1);
    g.launches++;
    CK(cudaMemcpyAsync(&fin, S.final_sweep, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
  }
  return 0;
}

// One step of the split path: the reference's call sequence, one kernel per loop.
int step_split(int tstp, bool upst, bool first_three) {
  int rc = alloc_split_buffers();
  if (rc) return rc;
  Dev D = g.D;
  set_state_pointers(D);
  const int nl = g.nlay;
  const dim3 grid1 = cell_grid(D, 1, kBlock), gridL = cell_grid(D, nl, kBlock);
  const bool rgld = g.P.rgld > 0.5;
  if (first_three) {  // pm:2166-2177
    k_centred_flux<<<gridL, kBlock, 0, g.stream>>>(D);
    g.launches++;
    if ((rc = sync_fields({{D.h_u, nl}, {D.h_v, nl}}))) return rc;
  } else if (rgld) {  // pm:2237-2257
    k_upstream_flux<<<gridL, kBlock, 0, g.stream>>>(D);
    g.launches++;
    if ((rc = sync_fields({{D.h_u, nl}, {D.h_v, nl}}))) return rc;
  }
  k_update_h<<<grid1, kBlock, 0, g.stream>>>(D);  // pm:2181 / 2259
  g.launches++;
  if (rgld) { k_rgld_correct<<<grid1, kBlock, 0, g.stream>>>(D); g.launches++; }
  if ((rc = sync_fields({{D.hlay, nl}, {(g.nranks > 1 || g.torus) ? D.rs_new : nullptr, nl}}))) return rc;
  g.rs_o = (g.rs_o + 1) % 3;

  k_diag<<<gridL, kBlock, 0, g.stream>>>(D);  // pm:2187 / 2266
  g.launches++;
  if ((g.nmir || g.nranks > 1) && (rc = sync_fields({{D.mont, nl}, {D.rvor, nl}, {D.pvor, nl}, {D.dive, nl}, {D.d2hx, nl}, {D.d2hy, nl}}))) return rc;
  if (first_three || (g.P.dvis > 1.e-3 && upst) || g.P.svis > 0) {  // pm:2188 / 2268-2270
    k_visc<<<gridL, kBlock, 0, g.stream>>>(D);
    g.launches++;
    if ((g.nmir || g.nranks > 1) && (rc = sync_fields({{D.v_cc, nl}, {D.v_ll, nl}}))) return rc;
    if (g.P.svis > 0.0) {
      if ((g.nmir || g.nranks > 1) && (rc = sync_fields({{D.delu, nl}, {D.delv, nl}}))) return rc;
      k_biharm<<<gridL, kBlock, 0, g.stream>>>(D);
      g.launches++;
      if ((g.nmir || g.nranks > 1) && (rc = sync_fields({{D.UU4, nl}, {D.VV4, nl}}))) return rc;
    }
  }
  for (int pass = 0; pass < 2; pass++) {  // pm:2193-2199 / 2276-2282
    const bool do_u = ((tstp % 2 == 0) == (pass == 0));
    if (do_u) {
      k_update_u<<<gridL, kBlock, 0, g.stream>>>(D);
      g.launches++;
      if ((rc = sync_fields({{D.u, nl}, {D.h_u, nl}, {(g.nranks > 1 || g.torus) ? D.dx_new : nullptr, nl}}))) return rc;
    } else {
      k_update_v<<<gridL, kBlock, 0, g.stream>>>(D);
      g.launches++;
      if ((rc = sync_fields({{D.v, nl}, {D.h_v, nl}, {(g.nranks > 1 || g.torus) ? D.dy_new : nullptr, nl}}))) return rc;
    }
  }
  g.dx_o = (g.dx_o + 1) % 4;
  g.dy_o = (g.dy_o + 1) % 4;
  if (g.D.has_nudg && g.P.mcbc < 0.5 && g.obc_any) {  // pm:2201-2204 / 2285-2288
    for (int pass = 0; pass < 2 && g.nseg > 0; pass++) {
      k_obc<<<(g.nseg + 63) / 64, 64, 0, g.stream>>>(D, g.d_seg, g.nseg, pass);
      g.launches++;
    }
    if ((g.nmir || g.nranks > 1) && (rc = sync_fields({{D.u, nl}, {D.v, nl}, {D.h_u, nl}, {D.h_v, nl}}))) return rc;
  }
  if (rgld) {  // pm:2207-2221 / 2292-2314
    if (first_three) k_centred_flux<<<gridL, kBlock, 0, g.stream>>>(D);
    else k_upstream_flux<<<gridL, kBlock, 0, g.stream>>>(D);
    if (g.nranks > 1 && (rc = sync_fields({{D.h_u, nl}, {D.h_v, nl}}))) return rc;  // the right-hand side reads h_v of the row above
    k_pi_rhs<<<grid1, kBlock, 0, g.stream>>>(D);
    static const bool one_cta = getenv("BEOM_PI_ONE_CTA") && atoi(getenv("BEOM_PI_ONE_CTA")) > 0;  // the single-block wavefront (cross-check)
    if (g.nranks > 1) {
      k_pi_begin<<<grid1, kBlock, 0, g.stream>>>(D, g.pis);
      if ((rc = pi_solve_slabs(D))) return rc;
      k_pi_select<<<grid1, kBlock, 0, g.stream>>>(D, g.pis, g.pi_iters);
      g.launches += 2;
      if ((rc = sync_fields({{D.pi_s, 1}}))) return rc;  // the projection reads pi_s of the row below
    } else if (one_cta || !g.pis_ready) {
      k_surf_pressure<<<1, 1024, 0, g.stream>>>(D, 1000, 1.e-5, g.pi_iters);
      g.launches += 1;
    } else {
      // every SM: tiles of anti-diagonal wavefronts, several sweeps in flight, the reference's iterates (rigid.cuh)
      k_pi_begin<<<grid1, kBlock, 0, g.stream>>>(D, g.pis);
      k_pi_wave<<<(unsigned)g.pis_blocks, (unsigned)g.pis_threads, 0, g.stream>>>(D, g.pis, 1, g.pis.maxiters, 0);
      k_pi_select<<<grid1, kBlock, 0, g.stream>>>(D, g.pis, g.pi_iters);
      g.launches += 3;
    }
    k_pi_correct<<<gridL, kBlock, 0, g.stream>>>(D);
    g.launches += 3;
    if (g.nranks > 1 && (rc = sync_fields({{D.u, nl}, {D.v, nl}}))) return rc;  // the projected velocities of the neighbours' rows
  }
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

const char *beom_gpu_version(void) { return "beom_b200 0.1 (sm_100a)"; }
int beom_gpu_abi_version(void) { return BEOM_ABI_VERSION; }
int beom_gpu_last_error(char *buf, int len) {
  if (buf && len > 0) snprintf(buf, (size_t)len, "%s", g_err.c_str());
  return (int)g_err.size();
}
void beom_gpu_default_options(beom_gpu_options *opt) {
  memset(opt, 0, sizeof *opt);
  opt->device = -1;
  opt->fused = 2;  // by size
  opt->rank = 0;
  opt->nranks = 1;
}
const char *beom_gpu_path(void) { return g.use_fused ? "fused" : "split"; }
const char *beom_gpu_fused_variant(void) { return g.use_fused ? fused_variant() : "none"; }
long long beom_gpu_launch_count(void) { return g.launches; }

int beom_gpu_finalize(void) {
  if (g.stream) cudaStreamSynchronize(g.stream);
  drop_graphs_fwd();
  for (void *p : g.allocs) cudaFree(p);
  g.allocs.clear();
  for (auto &e : g.ev)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  if (g.stream) { cudaStreamDestroy(g.stream); g.stream = nullptr; }
  if (g.comm_stream) { cudaStreamDestroy(g.comm_stream); g.comm_stream = nullptr; }
  if (g.copy_stream) { cudaStreamSynchronize(g.copy_stream); cudaStreamDestroy(g.copy_stream); g.copy_stream = nullptr; }
  for (auto &r : g.rec) {
    if (r.host) cudaFreeHost(r.host);
    if (r.mm_host) cudaFreeHost(r.mm_host);
    if (r.ready) cudaEventDestroy(r.ready);
    if (r.done) cudaEventDestroy(r.done);
  }
  if (g.ev_edge) { cudaEventDestroy(g.ev_edge); g.ev_edge = nullptr; }
  if (g.ev_comm) { cudaEventDestroy(g.ev_comm); g.ev_comm = nullptr; }
  fused_release();
  g = Ctx();
  return 0;
}

static int init_impl(const beom_params *par, const beom_fields *fld, const beom_gpu_options *opt_in);
int beom_gpu_init(const beom_params *par, const beom_fields *fld, const beom_gpu_options *opt_in) {
  if (g.ready) beom_gpu_finalize();
  const int rc = init_impl(par, fld, opt_in);
  if (rc) {  // release the streams, events and allocations made before the failure; the message survives
    const std::string keep = g_err;
    beom_gpu_finalize();
    g_err = keep;
  }
  return rc;
}
// device, streams and events: common to beom_gpu_init and beom_gpu_init_grids
static int init_prologue(const beom_params *par, const beom_gpu_options *opt_in, cudaDeviceProp *prop_out) {
  beom_gpu_options opt;
  if (opt_in) opt = *opt_in;
  else beom_gpu_default_options(&opt);
  if (par->nlay < 1 || par->nlay > BEOM_MAXLAY) return fail(-3, "beom_gpu_init: nlay out of range");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(-10, "beom_gpu_init: no CUDA device (%s); this library has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
  int dev = opt.device;
  if (dev < 0) {
    const char *lr = getenv("LOCAL_RANK");
    dev = lr ? atoi(lr) % ndev : 0;
  }
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(-11, "beom_gpu_init: device %d is sm_%d%d; kernels are built for sm_100a only", dev, prop.major, prop.minor);
  if (prop_out) *prop_out = prop;

  g.P = *par;
  g.opt = opt;
  g.lm = par->lm; g.mm = par->mm; g.nlay = par->nlay; g.ndeg = par->ndeg;
  g.rank = opt.nranks > 1 ? opt.rank : 0;
  g.nranks = opt.nranks > 1 ? opt.nranks : 1;
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  {
    int lo_pri = 0, hi_pri = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    CK(cudaStreamCreateWithPriority(&g.comm_stream, cudaStreamNonBlocking, hi_pri));
    CK(cudaEventCreateWithFlags(&g.ev_edge, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_comm, cudaEventDisableTiming));
  }
  CK(cudaEventCreate(&g.ev[0]));
  CK(cudaEventCreate(&g.ev[1]));
  return 0;
}
static int init_tail(double invf, double w_ti, const double *bodf);

// open-boundary segments as dense cells (private_mod.f95:1060-1240; segm is [18][nseg]: columns 1-3 the face point and its (i, j),
// 4 / 5 zonal / meridional, 10-12 the wet cell, 13-15 the interior normal-velocity point, 16-18 the interior cell); cell_of(p, i, j) = the
// dense cell of vector point p at grid point (i, j), or < 0.  A rank keeps the segments whose wet or face point lies in its rows.
static int build_segments(const int32_t *segm, int nseg, bool flag_nudging, const std::function<int(int, int, int)> &cell_of) {
  g.nseg = 0;
  g.obc_any = segm && nseg > 0 && flag_nudging;
  if (!g.obc_any) return 0;
  std::vector<SegDev> segs;
  for (int s = 0; s < nseg; s++) {
    auto col = [&](int cidx) { return segm[(size_t)(cidx - 1) * nseg + s]; };
    if (col(10) < 1 || col(12) < g.j0 || col(12) > g.j1)
      if (col(1) < 1 || col(3) < g.j0 || col(3) > g.j1) continue;
    SegDev sd;
    sd.zonal = col(4); sd.merid = col(5);
    sd.c_face = cell_of(col(1), col(2), col(3)); sd.c_wet = cell_of(col(10), col(11), col(12));
    sd.c_norm = cell_of(col(13), col(14), col(15)); sd.c_int = cell_of(col(16), col(17), col(18));
    if (sd.c_face < 0 || sd.c_wet < 0 || sd.c_norm < 0 || sd.c_int < 0) {
      if (g.P.mcbc < 0.5) return fail(-12, "beom_gpu_init: open-boundary segment %d touches a cell outside the grid", s + 1);
      continue;
    }
    segs.push_back(sd);
  }
  g.nseg = (int)segs.size();
  if (g.nseg) {
    int rc;
    if ((rc = dalloc(&g.d_seg, (size_t)g.nseg, false))) return rc;
    CK(cudaMemcpy(g.d_seg, segs.data(), sizeof(SegDev) * g.nseg, cudaMemcpyHostToDevice));
  }
  return 0;
}

// what analyse_layout (layout.h) found becomes this rank's context: slab rows, plane geometry, cell map, flags, mirror lists, duplicates
static int adopt_layout(const Layout &lay, const int32_t *sj) {
  const int nlay = g.nlay;
  const size_t nd1 = (size_t)g.ndeg + 1;
  g.j0 = lay.j0; g.j1 = lay.j1;
  g.NX = lay.NX; g.NY = lay.NY; g.plane = lay.plane;
  g.p_lo = lay.p_lo; g.p_hi = lay.p_hi;
  g.cell_of_point = lay.cell_of_point;
  g.orphans = lay.orphans;
  g.torus = lay.torus;
  g.ring = lay.ring;
  g.halo = halo_rows(g.rank, g.nranks, g.ring, G, G + (g.j1 - g.j0));
  g.own_first = 0; g.own_last = -1;
  for (int p = g.p_lo; p <= g.p_hi; p++)
    if (sj[p] >= g.j0 && sj[p] <= g.j1) { if (!g.own_first) g.own_first = p; g.own_last = p; }
  const std::vector<uint8_t> &hflags = lay.flags;
  int rc;
  g.nmir = (int)lay.mdst.size();
  if (g.nmir) {
    if ((rc = dalloc(&g.d_mir_dst, (size_t)g.nmir, false)) || (rc = dalloc(&g.d_mir_src, (size_t)g.nmir, false))) return rc;
    CK(cudaMemcpy(g.d_mir_dst, lay.mdst.data(), sizeof(int) * g.nmir, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.d_mir_src, lay.msrc.data(), sizeof(int) * g.nmir, cudaMemcpyHostToDevice));
  }
  g.orph = Orphans();
  g.orph.nlay = nlay;
  g.orph.n = g.orphans.size();
  g.orph.val.assign((size_t)3 * nlay * g.orphans.size(), 0.0);

  if ((rc = dalloc(&g.flags, g.plane, false))) return rc;
  CK(cudaMemcpy(g.flags, hflags.data(), g.plane, cudaMemcpyHostToDevice));
  if ((rc = dalloc(&g.d_cell, nd1, false))) return rc;
  CK(cudaMemcpy(g.d_cell, g.cell_of_point.data(), sizeof(int) * nd1, cudaMemcpyHostToDevice));
  const int nloc = g.p_hi - g.p_lo + 1;
  g.stage_elems = (size_t)nloc * nlay * 3;
  if ((rc = dalloc(&g.stage, g.stage_elems, false))) return rc;
  return 0;
}

static int init_impl(const beom_params *par, const beom_fields *fld, const beom_gpu_options *opt_in) {
  if (!par || !fld) return fail(-1, "beom_gpu_init: null argument");
  if (!fld->neig || !fld->subc || !fld->mk_u || !fld->mk_v || !fld->mk_n || !fld->mkpe || !fld->mkpi || !fld->fcor || !fld->h_th)
    return fail(-2, "beom_gpu_init: a required static field is NULL");
  if (fld->nudg && !fld->fnud) return fail(-4, "beom_gpu_init: nudg given without fnud");
  cudaDeviceProp prop = cudaDeviceProp();
  {
    const int rc0 = init_prologue(par, opt_in, &prop);
    if (rc0) return rc0;
  }

  const int lm = g.lm, mm = g.mm, nlay = g.nlay, ndeg = g.ndeg;
  const size_t nd1 = (size_t)ndeg + 1;
  const int32_t *sj = fld->subc + nd1;

  // dense layout of this rank: slab rows, cell map, flags, periodic images, duplicates (layout.h)
  Layout lay;
  if (analyse_layout(lay, lm, mm, ndeg, par->xper > 0.5, par->yper > 0.5, g.rank, g.nranks, fld->subc, fld->neig, fld->mk_n, fld->mk_u,
                     fld->mk_v, fld->mkpe, fld->mkpi))
    return fail(lay.rc, "%s", lay.error.c_str());
  int rc;
  if ((rc = adopt_layout(lay, sj))) return rc;
  const int j_off = lay.j_off;  // Y = j + j_off

  // ---- Dev template ----
  Dev &D = g.D;
  memset(&D, 0, sizeof D);
  D.NX = g.NX; D.NY = g.NY; D.plane = g.plane;
  D.x_lo = 1 + GX0; D.x_hi = lm + 1 + GX0;
  D.y_lo = G; D.y_hi = G + (g.j1 - g.j0);
  D.i_off = GX0; D.j_off = j_off;
  D.lm = lm; D.mm = mm; D.nlay = nlay;
  D.flags = g.flags;
  const size_t pl = g.plane, nl = (size_t)nlay;

  double *tmp = nullptr;
  if ((rc = dalloc(&tmp, pl))) return rc;
  if ((rc = upload_planes(tmp, fld->fcor, 1))) return rc;
  D.fcor = tmp;
  if ((rc = dalloc(&tmp, pl))) return rc;
  if ((rc = upload_planes(tmp, fld->h_th, 1))) return rc;
  D.h_th = tmp;
  if (g.torus) {  // the fused step reads the 2-D statics at its (recomputed) halo cells
    sync_fields({{const_cast<double *>(D.fcor), 1}, {const_cast<double *>(D.h_th), 1}}, false);
  }
  D.has_nudg = 0;
  if (fld->nudg) {
    bool any = false;
    for (size_t k = 0; k < 3 * nd1 && !any; k++) any = fld->nudg[k] != 0.0;
    if (any) {
      if ((rc = dalloc(&tmp, pl * 3))) return rc;
      if ((rc = upload_planes(tmp, fld->nudg, 3))) return rc;
      D.nudg = tmp;
      D.has_nudg = 1;
      if (g.torus) sync_fields({{tmp, 3}}, false);
    }
  }
  D.has_tide = 0;
  if (fld->tide && fld->w_ti != 0.0 && D.has_nudg) {
    // tide(2,1,0:ndeg,3): de-interleave amplitude / phase into [3][2] planes
    std::vector<double> v(nd1);
    if ((rc = dalloc(&tmp, pl * 6))) return rc;
    for (int c3 = 0; c3 < 3; c3++)
      for (int a = 0; a < 2; a++) {
        for (size_t p = 0; p < nd1; p++) v[p] = fld->tide[((size_t)c3 * nd1 + p) * 2 + a];
        if ((rc = upload_planes(tmp + ((size_t)c3 * 2 + a) * pl, v.data(), 1))) return rc;
      }
    D.tide = tmp;
    D.has_tide = 1;
  }
  if (fld->fnud && (D.has_nudg || g.P.variant != BEOM_VARIANT_STANDARD)) {
    if ((rc = dalloc(&tmp, pl * nl * 3))) return rc;
    if ((rc = upload_planes(tmp, fld->fnud, 3 * nlay))) return rc;
    D.fnud = tmp;
    if (g.torus) sync_fields({{tmp, 3 * nlay}}, false);
  }
  D.has_hdot = 0;
  if (fld->hdot) {
    bool any = false;
    for (size_t k = 0; k < nl * nd1 && !any; k++) any = fld->hdot[k] != 0.0;
    if (any) {
      if ((rc = dalloc(&tmp, pl * nl))) return rc;
      if ((rc = upload_planes(tmp, fld->hdot, nlay))) return rc;
      D.hdot = tmp;
      D.has_hdot = 1;
      if (g.torus) sync_fields({{tmp, nlay}}, false);
    }
  }
  g.any_taus = false;  // any(abs(taus) > 1.e-7), private_mod.f95:1945
  if (fld->taus)
    for (size_t k = 0; k < 2 * nd1 && !g.any_taus; k++) g.any_taus = std::fabs(fld->taus[k]) > 1.e-7;
  D.has_wind = g.any_taus;
  D.has_bdrg = par->bdrg > 1.e-7;
  D.has_tdrg = par->tdrg > 1.e-7;
  // periodic duplicates inside a sponge are not frozen: the relaxation moves them (orphans.h)
  if (D.has_nudg && !g.orphans.empty()) {
    Orphans &O = g.orph;
    const size_t no = O.n;
    O.nud.assign(3 * no, 0.0);
    bool uv = false;
    for (int f = 0; f < 3; f++)
      for (size_t k = 0; k < no; k++) {
        const double c = fld->nudg[(size_t)f * nd1 + g.orphans[k]];
        O.nud[(size_t)f * no + k] = c;
        O.live = O.live || c != 0.0;
        uv = uv || (f > 0 && c != 0.0);
      }
    if (O.live) {
      if (!fld->fnud) return fail(-9, "beom_gpu_init: nudged periodic duplicates need fnud");
      if (par->variant != BEOM_VARIANT_STANDARD || par->rgld > 0.5)
        return fail(-9, "beom_gpu_init: unsupported periodic connectivity (a sponge over the duplicate row/column with the 1d/3d/plume variants or the rigid lid)");
      if (uv && D.has_wind && fld->invf != 0.0)
        return fail(-9, "beom_gpu_init: unsupported periodic connectivity (a velocity sponge over the duplicate row/column under wind stress: the Ekman term of its target, private_mod.f95:1449-1452)");
      O.fnud.resize(3 * nl * no);
      for (int f = 0; f < 3; f++)
        for (int l = 0; l < nlay; l++)
          for (size_t k = 0; k < no; k++) O.fnud[((size_t)f * nl + l) * no + k] = fld->fnud[((size_t)f * nl + l) * nd1 + g.orphans[k]];
      O.has_tide = D.has_tide != 0;
      O.w_ti = fld->w_ti;
      if (O.has_tide) {
        O.tide.resize(6 * no);
        for (int f = 0; f < 3; f++)
          for (size_t k = 0; k < no; k++)
            for (int a = 0; a < 2; a++) O.tide[((size_t)f * no + k) * 2 + a] = fld->tide[((size_t)f * nd1 + g.orphans[k]) * 2 + a];
      }
    }
  }
  if (D.has_wind) {
    if ((rc = dalloc(&tmp, pl * 2))) return rc;
    if ((rc = upload_planes(tmp, fld->taus, 2))) return rc;
    D.taus = tmp;
    if ((rc = dalloc(&D.tt3d, pl * nl * 2)) || (rc = dalloc(&D.layt, pl * nl))) return rc;
  }
  if (D.has_bdrg)
    if ((rc = dalloc(&D.tb3d, pl * nl * 2)) || (rc = dalloc(&D.layb, pl * nl)) || (rc = dalloc(&D.taub, pl * 2))) return rc;
  if (D.has_tdrg)
    if ((rc = dalloc(&D.tu3d, pl * nl * 2)) || (rc = dalloc(&D.layu, pl * nl)) || (rc = dalloc(&D.taum, pl * 2))) return rc;

  if (par->rgld > 0.5) {  // rigid lid: Poisson operators and the start pressure (private_mod.f95:505-563)
    if (g.nmir || par->xper > 0.5 || par->yper > 0.5) return fail(-30, "beom_gpu_init: rgld = 1 is supported on non-periodic domains only");
    if (!fld->Ow || !fld->Os || !fld->Osum_) return fail(-14, "beom_gpu_init: rgld = 1 needs Ow, Os, Osum_");
    if (nlay != 2) return fail(-14, "beom_gpu_init: the reference's rigid lid is written for two layers (private_mod.f95:1653-1654)");
    double *o1 = nullptr, *o2 = nullptr, *o3 = nullptr;
    if ((rc = dalloc(&o1, pl)) || (rc = dalloc(&o2, pl)) || (rc = dalloc(&o3, pl)) || (rc = dalloc(&D.pi_s, pl)) || (rc = dalloc(&D.pi_rhs, pl))) return rc;
    if ((rc = upload_planes(o1, fld->Ow, 1)) || (rc = upload_planes(o2, fld->Os, 1)) || (rc = upload_planes(o3, fld->Osum_, 1))) return rc;
    if (fld->pi_s && (rc = upload_planes(D.pi_s, fld->pi_s, 1))) return rc;
    D.Ow = o1; D.Os = o2; D.Osum_ = o3;
    if ((rc = dalloc(&g.pi_iters, (size_t)16))) return rc;
    {  // the all-SM solver (rigid.cuh): kPiNA - 1 further pi_s arrays, five coefficient planes, per-tile sweep counters
      PiSolve &S = g.pis;
      memset(&S, 0, sizeof S);
      S.X[0] = D.pi_s;
      for (int k = 1; k < kPiNA; k++)
        if ((rc = dalloc(&S.X[k], pl))) return rc;
      double *cE = nullptr, *cN = nullptr, *cW = nullptr, *cS = nullptr;
      if ((rc = dalloc(&S.c0, pl)) || (rc = dalloc(&cE, pl)) || (rc = dalloc(&cN, pl)) || (rc = dalloc(&cW, pl)) || (rc = dalloc(&cS, pl))) return rc;
      const int w = D.x_hi - D.x_lo + 1, h = D.y_hi - D.y_lo + 1;
      S.TI = (w + kPiBx - 1) / kPiBx;
      S.TJ = (h + kPiBy - 1) / kPiBy;
      S.maxiters = 1000;   // pm:1717
      S.tol = 1.e-5;       // pi_tol, pm:1716
      unsigned long long *mb = nullptr;
      if ((rc = dalloc(&S.done, (size_t)(S.TI + 2) * (S.TJ + 2))) || (rc = dalloc(&S.count, (size_t)kPiNA)) || (rc = dalloc(&mb, (size_t)kPiNA)) ||
          (rc = dalloc(&S.decided, (size_t)4)) || (rc = dalloc(&S.md, (size_t)kPiNA)))
        return rc;
      S.maxbits = mb;
      S.final_sweep = S.decided + 1;
      k_pi_coeff<<<cell_grid(D, 1, kBlock), kBlock, 0, g.stream>>>(D, cE, cN, cW, cS);
      g.launches++;
      S.cE = cE; S.cN = cN; S.cW = cW; S.cS = cS;
      // persistent workers (one warp per tile and sweep): all of them must be resident at once
      const int ntiles = S.TI * S.TJ;
      g.pis_threads = 256;
      int per_sm = 0;
      CK((cudaError_t)cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pi_wave, g.pis_threads, 0));
      const int warps_per_block = g.pis_threads / 32;
      g.pis_blocks = std::max(1, std::min((ntiles + warps_per_block - 1) / warps_per_block, per_sm * prop.multiProcessorCount));
      g.pis_ready = true;
    }
  }

  rc = build_segments(fld->segm, fld->nseg, fld->flag_nudging != 0, [&](int p, int, int) { return (p >= 1 && p <= ndeg) ? g.cell_of_point[p] : -1; });
  if (rc) return rc;

  return init_tail(fld->invf, fld->w_ti, fld->bodf);
}

// state buffers, exchange buffers, the scalars of the Dev template, the fused step: common to both ways of initialising
static int init_tail(double invf, double w_ti, const double *bodf) {
  const beom_params *par = &g.P;
  const beom_gpu_options &opt = g.opt;
  Dev &D = g.D;
  const int nlay = g.nlay;
  const size_t pl = g.plane, nl = (size_t)nlay, nd1 = (size_t)g.ndeg + 1;
  int rc;
  // state
  for (int f = 0; f < 5; f++)
    if (!g.st[f][0] && (rc = dalloc(&g.st[f][0], pl * nl))) return rc;
  for (auto &p : g.rs)
    if ((rc = dalloc(&p, pl * nl))) return rc;
  for (auto &p : g.dx)
    if ((rc = dalloc(&p, pl * nl))) return rc;
  for (auto &p : g.dy)
    if ((rc = dalloc(&p, pl * nl))) return rc;

  if (g.nranks > 1) {
    if (!comm_ready() || comm_size() != g.nranks || comm_rank() != g.rank)
      return fail(-13, "beom_gpu_init: nranks = %d but beom_gpu_comm_init was not called with the same rank/size", g.nranks);
    g.halo_cap = (size_t)G * g.NX * (size_t)nlay * 12;  // up to 12 layered fields per exchange
    for (int k = 0; k < 2; k++)
      if ((rc = dalloc(&g.halo_send[k], g.halo_cap)) || (rc = dalloc(&g.halo_recv[k], g.halo_cap))) return rc;
    if (g.ring && g.torus) {
      // the statics of the deep rows across the seam of a y-periodic slab chain are another rank's rows: one ring
      // exchange each (every rank initialises, so this is a collective like the rest of a multi-rank init)
      if ((rc = sync_fields({{const_cast<double *>(D.fcor), 1}, {const_cast<double *>(D.h_th), 1}}))) return rc;
      if (D.has_nudg && (rc = sync_fields({{const_cast<double *>(D.nudg), 3}}))) return rc;
      if (D.fnud && (rc = sync_fields({{const_cast<double *>(D.fnud), 3 * nlay}}))) return rc;
      if (D.has_hdot && (rc = sync_fields({{const_cast<double *>(D.hdot), nlay}}))) return rc;
    }
  }

  // scalars and constants, evaluated like the reference does
  D.invf = invf; D.w_ti = w_ti;
  D.dl = par->dl; D.dt = par->dt; D.grav = par->grav;
  D.i_dl = 1.0 / par->dl; D.i_gr = 1.0 / par->grav; D.i_r0 = 1.0 / par->rho0; D.i_r1 = 1.0 / par->rhon[0];
  D.uadv = par->uadv; D.ocrp = par->ocrp; D.qdrg = par->qdrg; D.rgld = par->rgld;
  D.hsal = par->hsal; D.hmin = par->hmin; D.hsbl = par->hsbl; D.hbbl = par->hbbl;
  D.i_ns = 1.0 / (double)(par->nsal - 1);
  D.two_hs = 2.0 * par->hsal;
  D.bvis = par->bvis; D.dvis = par->dvis; D.svis = par->svis; D.bdrg = par->bdrg; D.tdrg = par->tdrg;
  D.beta = par->beta; D.epsi = par->epsi; D.gamm = par->gamm; D.del1 = par->del1; D.del2 = par->del2;
  D.plum = par->plum; D.pi = par->pi;
  D.c_ab1 = 1.5 + par->beta;
  D.c_ab2 = 0.5 + 2.0 * par->beta;
  for (int l = 0; l < nlay; l++) {
    D.rhon[l] = par->rhon[l];
    D.i_rn[l] = 1.0 / par->rhon[l];
    D.bodf[0][l] = bodf ? bodf[l] : 0.0;
    D.bodf[1][l] = bodf ? bodf[(size_t)nlay + l] : 0.0;
  }
  D.bstress_thr = (par->variant == BEOM_VARIANT_1D ? 0.0 : 2.0) * par->hsal;
  D.variant = par->variant; D.nsal = par->nsal;
  D.gene = 0.0; D.ramp = 1.0; D.ctim = 0.0;
  {
    const double dtd8 = par->dt / 24.0 / 3600.0;
    const double q = par->dt3d / dtd8;
    g.n_3d = std::max((int)std::floor(q + 0.5), 1);
  }

  g.use_fused = false;
  // fused = 2 (the default): small grids are latency bound and one kernel per loop over every SM beats one CTA marching through
  // a strip (sill_exchange3D 0.028 against 0.093 ms per step, stommel1948 0.032 against 0.063)
  const double cell_layers = (double)(D.x_hi - D.x_lo + 1) * (double)(D.y_hi - D.y_lo + 1) * (double)nlay;
  if (opt.fused == 1 || (opt.fused == 2 && cell_layers >= 250000.0)) {
    // periodic images that are not a complete single-rank torus (incl. the ring of a y-periodic slab chain): split path
    rc = fused_configure(g.D, g.P, g.torus ? 0 : g.nmir + (g.ring ? 1 : 0), g.nranks, &g.use_fused);
    if (rc) return rc;
    if (g.use_fused)
      for (int f = 0; f < 5; f++)
        if ((rc = dalloc(&g.st[f][1], pl * nl))) return rc;
  }
  {
    // whole steps as CUDA graphs where the step is launch bound (step_graphed): the same size rule as the kernel path
    const char *e = getenv("BEOM_GRAPH");
    const bool want = e ? atoi(e) > 0 : cell_layers < 250000.0;
    g.graphs_on = want && g.nranks == 1 && !D.has_tide && !(par->rgld > 0.5);
    g.graph_gene = -1.0;
    g.direct_steps = 0;
    g.graph_launches = 0;
  }
  CK(cudaStreamSynchronize(g.stream));
  g.win_first = 0;
  g.win_stride = nd1;
  g.ready = true;
  g_err.clear();
  return 0;
}

}  // extern "C"
namespace {
__global__ void k_flags_as_double(const uint8_t *f, double *out, size_t n) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = (double)f[k];
}
int download_planes_fwd(double *dst, const double *dense, int nplanes);
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// beom_gpu_init_grids: read_input_data's grid-shaped work on the device (gridinit.cuh).  The caller hands over the raw
// contents of the input files; masks, vector numbering, rest thickness, relaxation targets, forcing planes and the initial
// state are built straight in the dense layout of this rank, with the reference's arithmetic.  Returns
// BEOM_GRIDS_UNSUPPORTED (and says why in the error string) for what still needs the host path: the rigid lid, the 1d/3d/plume
// variants, h_to, restarts.  Periodic aliases, the layout analysis and the open-boundary segment table are host work here too
// (grid_index.h, layout.h), from the depth grid; everything plane-shaped is device work.
// ---------------------------------------------------------------------------------------------------------------------
static int init_grids_impl(const beom_params *par, const beom_grids *gr, const beom_gpu_options *opt_in) {
  if (!par || !gr) return fail(-1, "beom_gpu_init_grids: null argument");
  const char *why = nullptr;
  if (par->rgld > 0.5) why = "rigid lid";
  else if (par->variant != BEOM_VARIANT_STANDARD) why = "1d / 3d / plume variant";
  else if (par->topt > 0.5 || gr->has_h_to) why = "h_to.bin";
  else if (gr->tide && !gr->nudg) why = "tide.bin without nudg.bin";
  else if (par->rsta > 0.5) why = "restart";
  if (why) {
    fail(BEOM_GRIDS_UNSUPPORTED, "beom_gpu_init_grids: %s is initialised on the host (read_input_data + beom_gpu_init)", why);
    return BEOM_GRIDS_UNSUPPORTED;
  }
  cudaDeviceProp prop = cudaDeviceProp();
  int rc;
  if ((rc = init_prologue(par, opt_in, &prop))) return rc;
  const int lm = g.lm, mm = g.mm, nlay = g.nlay, ndeg = g.ndeg;
  const size_t nd1 = (size_t)ndeg + 1, nl = (size_t)nlay;

  const bool periodic = par->xper > 0.5 || par->yper > 0.5;
  const double flat_depth = (par->cext * par->cext) / par->grav;
  auto host_depth = [&](int i, int j) -> double {  // the depth of cell (i, j) as read_input_data sees it (pm:119-121, 827-839)
    if (i < 1 || i > lm || j < 1 || j > mm) return 0.0;
    if (!gr->h_bo) return flat_depth;
    const double d = (double)gr->h_bo[(size_t)j * (lm + 2) + i];
    return d < par->hdry ? 0.0 : d;
  };
  std::vector<int32_t> h_subc;  // periodic domains: the grid coordinates of the vector points (for the duplicates' host-side state)
  int j_off;
  if (!periodic) {
    // y-slab and dense layout of this rank: the same rules as analyse_layout (layout.h) on a non-periodic domain
    {
      const int rows = mm + 1, base = rows / g.nranks, rem = rows % g.nranks;
      g.j0 = 1 + g.rank * base + std::min(g.rank, rem);
      g.j1 = g.j0 + base + (g.rank < rem ? 1 : 0) - 1;
      if (g.j1 < g.j0) return fail(-5, "beom_gpu_init: more ranks than grid rows");
    }
    g.NX = ((lm + GX0 + 34) + 15) / 16 * 16;
    g.NY = (g.j1 - g.j0 + 1) + 2 * G;
    g.plane = (size_t)g.NX * g.NY;
    if (g.plane > 0x7fffffffull) return fail(-6, "beom_gpu_init: plane too large for 32-bit cell offsets");
    j_off = G - g.j0;
    g.torus = false; g.ring = false; g.nmir = 0;
    g.halo = halo_rows(g.rank, g.nranks, false, G, G + (g.j1 - g.j0));
    g.orph = Orphans();
    g.orph.nlay = nlay;
    g.orphans.clear();
    g.cell_of_point.clear();
  } else {
    // Periodic domains: the aliases of index_grid_points (pm:614-685) decide which cells are images of which and which points are
    // displaced duplicates -- the analysis of the neighbour table beom_gpu_init runs (layout.h).  The table is built here on the
    // host from the depth grid (grid_index.h, the host driver's own source) and analysed the same way; the planes are then filled
    // on the device like on any other domain.
    std::vector<int32_t> neig(nd1 * 8, 0), posc(nd1, 0);
    h_subc.assign(nd1 * 2, 0);
    std::vector<double> mk[5];
    for (auto &m : mk) m.assign(nd1, 0.0);
    const int count = index_grid_points_core(lm, mm, ndeg, par->hdry, par->xper > 0.5, par->yper > 0.5, host_depth, neig.data(), h_subc.data(),
                                             posc.data(), mk[0].data(), mk[1].data(), mk[2].data(), mk[3].data(), mk[4].data());
    if (count != ndeg) return fail(-15, " wrong input parameter! Please set ndeg = %d inside file shared_mod.f95.", count);
    Layout lay;
    if (analyse_layout(lay, lm, mm, ndeg, par->xper > 0.5, par->yper > 0.5, g.rank, g.nranks, h_subc.data(), neig.data(), mk[2].data(), mk[0].data(),
                       mk[1].data(), mk[3].data(), mk[4].data()))
      return fail(lay.rc, "%s", lay.error.c_str());
    if ((rc = adopt_layout(lay, h_subc.data() + nd1))) return rc;
    j_off = lay.j_off;
  }
  const size_t pl = g.plane;

  // the raw files on the device (released again at the end)
  std::vector<void *> raw;
  auto put = [&](const float *src, size_t count, const float **dst) -> int {
    *dst = nullptr;
    if (!src) return 0;
    float *d = nullptr;
    CK(cudaMalloc(&d, count * sizeof(float)));
    raw.push_back(d);
    CK(cudaMemcpyAsync(d, src, count * sizeof(float), cudaMemcpyHostToDevice, g.stream));
    *dst = d;
    return 0;
  };
  struct Release {
    std::vector<void *> &v;
    ~Release() { for (void *q : v) cudaFree(q); }
  } release{raw};
  GridIn A;
  memset(&A, 0, sizeof A);
  A.lm = lm; A.mm = mm; A.nlay = nlay; A.NX = g.NX; A.NY = g.NY; A.j_off = j_off;
  A.jlo = std::max(g.j0 - G, 0); A.jhi = std::min(g.j1 + G, mm + 1);
  A.hdry = par->hdry;
  A.flat = flat_depth;
  A.tauw[0] = par->tauw[0]; A.tauw[1] = par->tauw[1]; A.f0 = par->f0;
  for (int l = 0; l < nlay; l++) A.topl[l] = par->topl[l];
  const size_t gp = (size_t)(lm + 2) * (mm + 2);
  if ((rc = put(gr->h_bo, gp, &A.h_bo)) || (rc = put(gr->init, gp * nl * 3, &A.init)) || (rc = put(gr->nudg, gp * 3, &A.nudg)) ||
      (rc = put(gr->taus, gp * 2, &A.taus)) || (rc = put(gr->fcor, gp, &A.fcor)) || (rc = put(gr->hdot, gp * nl, &A.hdot)) ||
      (rc = put(gr->tide, gp * 6, &A.tide)))
    return rc;

  // ---- index_grid_points (pm:567-764): points per row, their prefix sum, the depth extremes (pm:134-135)
  int *d_rowcnt = nullptr, *d_rowoff = nullptr;
  double *d_rowmm = nullptr;
  if ((rc = dalloc(&d_rowcnt, (size_t)mm + 3, false)) || (rc = dalloc(&d_rowoff, (size_t)mm + 3, false)) || (rc = dalloc(&d_rowmm, 2 * ((size_t)mm + 2), false))) return rc;
  k_gi_row_counts<<<(unsigned)(mm + 2), 256, 0, g.stream>>>(A, d_rowcnt, d_rowmm, d_rowmm + (mm + 2));
  g.launches++;
  std::vector<int> rowcnt(mm + 2), rowoff(mm + 3, 0);
  std::vector<double> rowmm(2 * ((size_t)mm + 2));
  CK(cudaMemcpyAsync(rowcnt.data(), d_rowcnt, sizeof(int) * (mm + 2), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaMemcpyAsync(rowmm.data(), d_rowmm, sizeof(double) * rowmm.size(), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  double dmin = INFINITY, dmax = 0.0;
  for (int j = 0; j <= mm + 1; j++) {
    rowoff[j + 1] = rowoff[j] + rowcnt[j];
    dmin = std::min(dmin, rowmm[j]);
    dmax = std::max(dmax, rowmm[(size_t)(mm + 2) + j]);
  }
  if (rowoff[mm + 2] != ndeg) return fail(-15, " wrong input parameter! Please set ndeg = %d inside file shared_mod.f95.", rowoff[mm + 2]);
  if (par->ocrp < 0.5 && nlay > 1) {  // pm:137-152
    if (par->topl[nlay - 1] * dmax + 10.0 * par->hmin >= dmin) return fail(-16, " Please modify topl so that bathymetry is contained within lower layer.");
  } else if (par->ocrp < 0.5 && nlay == 1) {
    if (dmin <= 10.0 * par->hmin) return fail(-16, " Please adjust h_bo or hmin so that min(h_bo) > 10. * hmin.");
  }
  A.dmax = dmax;
  if (!periodic) {
    CK(cudaMemcpyAsync(d_rowoff, rowoff.data(), sizeof(int) * (mm + 3), cudaMemcpyHostToDevice, g.stream));
    g.p_lo = 1 + rowoff[A.jlo]; g.p_hi = rowoff[A.jhi + 1];
    if (g.p_hi < g.p_lo) return fail(-8, "beom_gpu_init: no grid points on rank %d", g.rank);
    g.own_first = 1 + rowoff[g.j0]; g.own_last = rowoff[g.j1 + 1];
    if (g.own_last < g.own_first) { g.own_first = 0; g.own_last = -1; }
  }

  // ---- Dev template, flags, cell map, h_th
  Dev &D = g.D;
  memset(&D, 0, sizeof D);
  D.NX = g.NX; D.NY = g.NY; D.plane = pl;
  D.x_lo = 1 + GX0; D.x_hi = lm + 1 + GX0;
  D.y_lo = G; D.y_hi = G + (g.j1 - g.j0);
  D.i_off = GX0; D.j_off = j_off;
  D.lm = lm; D.mm = mm; D.nlay = nlay;
  double *h_th = nullptr, *fcor = nullptr;
  if ((rc = dalloc(&h_th, pl)) || (rc = dalloc(&fcor, pl))) return rc;
  if (!periodic) {
    if ((rc = dalloc(&g.flags, pl)) || (rc = dalloc(&g.d_cell, nd1, false))) return rc;
    CK(cudaMemsetAsync(g.d_cell, 0xff, sizeof(int) * nd1, g.stream));  // -1: not held on this rank
    k_gi_index<<<(unsigned)(A.jhi - A.jlo + 1), 256, 0, g.stream>>>(A, d_rowoff, g.d_cell, g.flags, h_th);
    const int nloc = g.p_hi - g.p_lo + 1;
    g.stage_elems = (size_t)nloc * nlay * 3;
    if ((rc = dalloc(&g.stage, g.stage_elems, false))) return rc;
  } else {  // (flags, cell map and staging buffer: adopt_layout)
    k_gi_hth<<<dim3((unsigned)((g.NX + 127) / 128), (unsigned)g.NY), 128, 0, g.stream>>>(A, g.flags, h_th);
  }
  g.launches++;
  D.flags = g.flags;
  D.h_th = h_th; D.fcor = fcor;

  // ---- rest thickness (pm:154-183)
  if ((rc = dalloc(&g.diag_h0, pl * nl))) return rc;
  const unsigned cb = (unsigned)((pl + 255) / 256);
  if (par->ocrp < 0.5) {
    k_gi_h0_stack<<<cb, 256, 0, g.stream>>>(A, g.flags, h_th, g.diag_h0, pl);
  } else {
    RestSolver S;
    memset(&S, 0, sizeof S);
    S.nlay = nlay; S.nsal = par->nsal; S.itmx = par->itmx;
    S.hsal = par->hsal; S.thre = par->tole; S.sor = par->sor; S.dmax = dmax;
    for (int l = 0; l < nlay; l++) { S.rho[l] = par->rhon[l]; S.topl[l] = par->topl[l]; }
    S.prepare();
    int *d_bad = nullptr, bad = -1;
    if ((rc = dalloc(&d_bad, (size_t)4, false))) return rc;
    CK(cudaMemcpyAsync(d_bad, &bad, sizeof(int), cudaMemcpyHostToDevice, g.stream));
    k_gi_h0_newton<<<(unsigned)((pl + 127) / 128), 128, 0, g.stream>>>(S, g.flags, h_th, g.diag_h0, pl, d_bad);
    CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    if (bad >= 0) return fail(-17, " calculation of h_layers did not converge, tolerance (meters) was %g, at cell %d", S.thre, bad);
  }
  g.launches++;

  // ---- forcing files and the initial state (pm:198-216, 840-964)
  for (int f = 0; f < 3; f++)
    if ((rc = dalloc(&g.st[f][0], pl * nl))) return rc;
  GridOut O;
  memset(&O, 0, sizeof O);
  O.plane = pl; O.flags = g.flags; O.h_0 = g.diag_h0;
  O.hlay = g.st[0][0]; O.u = g.st[1][0]; O.v = g.st[2][0];
  O.fcor = fcor;
  O.set_state = 1;
  double *tmp = nullptr;
  if (A.nudg) {
    if ((rc = dalloc(&tmp, pl * 3))) return rc;
    O.nudg = tmp;
    if ((rc = dalloc(&tmp, pl * nl * 3))) return rc;
    O.fnud = tmp;
  }
  const bool tauw_live = std::fabs(par->tauw[0]) > 1.e-7 || std::fabs(par->tauw[1]) > 1.e-7;
  if (A.taus || tauw_live) {
    if ((rc = dalloc(&tmp, pl * 2))) return rc;
    O.taus = tmp;
  }
  if (A.hdot) {
    if ((rc = dalloc(&tmp, pl * nl))) return rc;
    O.hdot = tmp;
  }
  const double w_ti = gr->tide ? (double)gr->tide[0] : 0.0;  // tide(1,1,0,0,1) is the frequency (pm:951-964)
  if (A.tide && w_ti != 0.0) {
    if ((rc = dalloc(&tmp, pl * 6))) return rc;
    O.tide = tmp;
  }
  unsigned *d_any = nullptr, any = 0;
  if ((rc = dalloc(&d_any, (size_t)4))) return rc;
  O.any = d_any;
  k_gi_forcing<<<dim3((unsigned)((g.NX + 127) / 128), (unsigned)g.NY), 128, 0, g.stream>>>(A, O);
  g.launches++;
  CK(cudaMemcpyAsync(&any, d_any, sizeof(unsigned), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  D.has_nudg = (any & 2u) ? 1 : 0;
  if (D.has_nudg) { D.nudg = O.nudg; D.fnud = O.fnud; }
  D.has_tide = (O.tide && D.has_nudg) ? 1 : 0;
  if (D.has_tide) D.tide = O.tide;
  D.has_hdot = (any & 8u) ? 1 : 0;
  if (D.has_hdot) D.hdot = O.hdot;
  g.any_taus = (any & 4u) != 0;
  D.has_wind = g.any_taus;
  D.has_bdrg = par->bdrg > 1.e-7;
  D.has_tdrg = par->tdrg > 1.e-7;
  if (D.has_wind) {
    D.taus = O.taus;
    if ((rc = dalloc(&D.tt3d, pl * nl * 2)) || (rc = dalloc(&D.layt, pl * nl))) return rc;
  }
  if (D.has_bdrg)
    if ((rc = dalloc(&D.tb3d, pl * nl * 2)) || (rc = dalloc(&D.layb, pl * nl)) || (rc = dalloc(&D.taub, pl * 2))) return rc;
  if (D.has_tdrg)
    if ((rc = dalloc(&D.tu3d, pl * nl * 2)) || (rc = dalloc(&D.layu, pl * nl)) || (rc = dalloc(&D.taum, pl * 2))) return rc;
  g.nseg = 0;
  g.obc_any = false;
  if (gr->nudg && (any & 1u)) {  // flag_nudging: the nudged open-boundary faces (index_boundary_points, pm:1060-1240; O(N) pass on the host)
    std::vector<int32_t> segm;
    const int nseg = index_boundary_points_core(lm, mm, par->hdry, par->xper > 0.5, par->yper > 0.5, host_depth, gr->nudg, segm);
    if (nseg == 0) return fail(-18, " the nudged open boundary segments could not be identified.");
    const int NX = g.NX;
    if (periodic) rc = build_segments(segm.data(), nseg, true, [&](int p, int, int) { return (p >= 1 && p <= ndeg) ? g.cell_of_point[p] : -1; });
    else rc = build_segments(segm.data(), nseg, true, [&](int p, int i, int j) {
      return (p >= 1 && j >= g.j0 - G && j <= g.j1 + G) ? (j + j_off) * NX + (i + GX0) : -1;
    });
    if (rc) return rc;
  }
  if (g.torus) {  // the fused step reads the statics at its (recomputed) halo cells: periodic images on this rank
    if ((rc = sync_fields({{fcor, 1}, {h_th, 1}}, false))) return rc;
    if (D.has_nudg && ((rc = sync_fields({{const_cast<double *>(D.nudg), 3}}, false)) || (rc = sync_fields({{const_cast<double *>(D.fnud), 3 * nlay}}, false)))) return rc;
    if (D.has_hdot && (rc = sync_fields({{const_cast<double *>(D.hdot), nlay}}, false))) return rc;
  }

  // invf = 1 / mean(fcor(0:ndeg)) (pm:223-229): a sum in vector order, so it is taken on the host from what the vector holds
  double invf;
  {
    double acc = 0.0;
    if (!gr->fcor) {
      for (size_t q = 0; q < nd1; q++) acc += par->f0;
    } else {
      const float *f = gr->fcor, *hb = gr->h_bo;
      float mean = 0.0f;
      for (size_t k = 0; k < gp; k++) mean = mean + f[k];
      acc += (double)(mean / (float)gp);  // fcor(0)
      auto depth = [&](int i, int j) -> double {
        if (i < 1 || i > lm || j < 1 || j > mm) return 0.0;
        if (!hb) return A.flat;
        const double d = (double)hb[(size_t)j * (lm + 2) + i];
        return d < par->hdry ? 0.0 : d;
      };
      auto wet = [&](int i, int j) { return depth(i, j) > par->hdry; };
      for (int j = 0; j <= mm + 1; j++)
        for (int i = 0; i <= lm + 1; i++) {
          if (!(wet(i, j) || wet(i - 1, j) || wet(i, j - 1) || wet(i - 1, j - 1))) continue;
          const size_t k0 = (size_t)j * (lm + 2) + i;
          if (i > 0 && j > 0) {
            float t = f[k0] * 0.25f;
            t = t + f[k0 - 1] * 0.25f;
            t = t + f[k0 - (lm + 2)] * 0.25f;
            t = t + f[k0 - (lm + 2) - 1] * 0.25f;
            acc += (double)t;
          } else {
            acc += (double)f[k0];
          }
        }
    }
    invf = acc / (double)nd1;
    invf = std::fabs(invf) > 1.25e-5 ? 1.0 / invf : 0.0;
  }
  std::vector<double> bodf((size_t)nlay * 2, 0.0);
  if (gr->bodf)
    for (size_t k = 0; k < (size_t)nlay * 2; k++) bodf[k] = (double)gr->bodf[k];

  // Displaced periodic duplicates (orphans.h) have no cell: their state and, where a sponge covers them, the data of their
  // relaxation recurrence are kept on the host -- the same statements as k_gi_forcing, for those few points (their masks are 0)
  if (!g.orphans.empty()) {
    Orphans &O = g.orph;
    const size_t no = O.n;
    const int lm2 = lm + 2;
    const int32_t *si = h_subc.data(), *sj = h_subc.data() + nd1;
    O.nud.assign(3 * no, 0.0);
    std::vector<double> ofn(3 * nl * no, 0.0), otide(gr->tide ? 6 * no : 0, 0.0);
    bool uv = false;
    for (size_t k = 0; k < no; k++) {
      const int i = si[g.orphans[k]], j = sj[g.orphans[k]];
      const size_t k0 = (size_t)j * lm2 + i;
      if (gr->nudg) {
        const float *n0 = gr->nudg, *n1 = gr->nudg + gp, *n2 = gr->nudg + 2 * gp;
        double nu = 0.0, nv = 0.0;
        if (i >= 1 && n1[k0 - 1] > 1.e-9f && n1[k0] > 1.e-9f) nu = (double)n1[k0] * 0.5 + (double)n1[k0 - 1] * 0.5;
        if (j >= 1 && n2[k0 - lm2] > 1.e-9f && n2[k0] > 1.e-9f) nv = (double)n2[k0] * 0.5 + (double)n2[k0 - lm2] * 0.5;
        O.nud[k] = (double)n0[k0]; O.nud[no + k] = nu; O.nud[2 * no + k] = nv;
        for (int f = 0; f < 3; f++) {
          O.live = O.live || O.nud[(size_t)f * no + k] != 0.0;
          uv = uv || (f > 0 && O.nud[(size_t)f * no + k] != 0.0);
        }
      }
      for (int l = 0; l < nlay; l++) {
        double hl = 0.0 * 0.0, fn = hl, fu = 0.0, fv = 0.0;  // h_0 * mk_n with mk_n = 0 (pm:198-200)
        if (gr->init) {
          double t = hl + (double)gr->init[((size_t)0 * nl + l) * gp + k0];
          if (l < nlay - 1) t = t - (double)gr->init[((size_t)0 * nl + l + 1) * gp + k0];
          fn = t * 0.0;
          fu = (double)gr->init[((size_t)1 * nl + l) * gp + k0];
          fv = (double)gr->init[((size_t)2 * nl + l) * gp + k0];
          hl = fn * 0.0;
        }
        O.val[((size_t)0 * nl + l) * no + k] = hl;
        O.val[((size_t)1 * nl + l) * no + k] = fu;
        O.val[((size_t)2 * nl + l) * no + k] = fv;
        ofn[((size_t)0 * nl + l) * no + k] = fn; ofn[((size_t)1 * nl + l) * no + k] = fu; ofn[((size_t)2 * nl + l) * no + k] = fv;
      }
      if (gr->tide)
        for (int c3 = 0; c3 < 3; c3++)
          for (int a = 0; a < 2; a++) otide[((size_t)c3 * no + k) * 2 + a] = (double)gr->tide[((size_t)c3 * gp + k0) * 2 + a];
    }
    if (O.live) {
      if (uv && D.has_wind && invf != 0.0)
        return fail(-9, "beom_gpu_init: unsupported periodic connectivity (a velocity sponge over the duplicate row/column under wind stress: the Ekman term of its target, private_mod.f95:1449-1452)");
      O.fnud = ofn;
      O.has_tide = D.has_tide != 0;
      O.w_ti = w_ti;
      if (O.has_tide) O.tide = otide;
    } else {
      O.nud.clear();
    }
    O.forget();
  }

  if ((rc = init_tail(invf, w_ti, gr->bodf ? bodf.data() : nullptr))) return rc;
  if (periodic) {  // images of the initial state (across the seam of a y-periodic slab chain they are another rank's rows)
    for (int f = 0; f < 3; f++)
      if ((rc = sync_fields({{g.st[f][0], nlay}}, g.ring))) return rc;
  }

  // h_0.bin's content for the output records (pm:185-194)
  if ((rc = dalloc(&g.h0r4, (size_t)ndeg * nl))) return rc;
  {
    const int p0 = std::max(g.p_lo, 1), n = g.p_hi - p0 + 1;
    k_gi_h0r4<<<(unsigned)((n + 255) / 256), 256, 0, g.stream>>>(g.diag_h0, pl, g.d_cell, p0, n, ndeg, nlay, g.h0r4);
    g.launches++;
  }
  g.h0r4_orph.clear();
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

extern "C" int beom_gpu_init_grids(const beom_params *par, const beom_grids *gr, const beom_gpu_options *opt_in) {
  if (g.ready) beom_gpu_finalize();
  const int rc = init_grids_impl(par, gr, opt_in);
  if (rc) {
    const std::string keep = g_err;
    beom_gpu_finalize();
    g_err = keep;
  }
  return rc;
}

// grid coordinates (i, j) of the vector points first .. first + count - 1 this rank holds (beom_gpu_point_range)
extern "C" int beom_gpu_download_subc(int32_t *si, int32_t *sj) {
  if (!g.ready) return fail(-20, "beom_gpu_download_subc: not initialised");
  const int n = g.p_hi - g.p_lo + 1;
  int *d = reinterpret_cast<int *>(g.stage);  // (3 nlay n doubles: room for 2 n ints)
  k_gi_subc<<<(unsigned)((n + 255) / 256), 256, 0, g.stream>>>(g.d_cell, g.p_lo, n, g.NX, g.D.j_off, d, d + n);
  g.launches++;
  CK(cudaMemcpyAsync(si, d, sizeof(int) * n, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaMemcpyAsync(sj, d + n, sizeof(int) * n, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

// grid.bin's records (posc, mk_n, mk_u, mk_v, mkpi: 5 x ndeg int32, pm:732-749) and h_0.bin's (float32 [nlay][ndeg], pm:185-194)
// for a host driver that initialised on the device (one rank)
extern "C" int beom_gpu_download_grid_files(int32_t *grid5, float *h_0_r4) {
  if (!g.ready) return fail(-20, "beom_gpu_download_grid_files: not initialised");
  if (g.nranks > 1 || g.p_lo != 1 || g.p_hi != g.ndeg || !g.orphans.empty())
    return fail(-27, "beom_gpu_download_grid_files: one rank holding every point in a cell only (no displaced periodic duplicates)");
  const int n = g.ndeg;
  if (grid5) {
    int *d = nullptr, rc;
    if ((rc = dalloc(&d, (size_t)5 * n, false))) return rc;
    k_gi_grid_record<<<(unsigned)((n + 255) / 256), 256, 0, g.stream>>>(g.d_cell, g.flags, 1, n, g.NX, g.D.j_off, g.lm, d);
    g.launches++;
    CK(cudaMemcpyAsync(grid5, d, sizeof(int) * 5 * (size_t)n, cudaMemcpyDeviceToHost, g.stream));
  }
  if (h_0_r4) {
    if (!g.h0r4) return fail(-22, "beom_gpu_download_grid_files: no rest thickness on the device");
    CK(cudaMemcpyAsync(h_0_r4, g.h0r4, sizeof(float) * (size_t)n * g.nlay, cudaMemcpyDeviceToHost, g.stream));
  }
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

// A static plane in the reference's vector layout (tests: the device-side initialisation against read_input_data):
// name = fcor | h_th | nudg | fnud | taus | hdot | h_0 | flags (the flag byte as a double), index = plane within the array
extern "C" int beom_gpu_debug_static(const char *name, int index, double *out) {
  if (!g.ready) return fail(-20, "beom_gpu_debug_static: not initialised");
  const std::string nm = name ? name : "";
  const Dev &D = g.D;
  const double *base = nm == "fcor" ? D.fcor : nm == "h_th" ? D.h_th : nm == "nudg" ? D.nudg : nm == "fnud" ? D.fnud : nm == "taus" ? D.taus
                       : nm == "hdot" ? D.hdot : nm == "h_0" ? g.diag_h0 : nullptr;
  const int n = g.p_hi - g.p_lo + 1;
  const size_t keep_first = g.win_first, keep_stride = g.win_stride;
  g.win_first = 0; g.win_stride = (size_t)g.ndeg + 1;
  int rc = 0;
  if (nm == "flags") {
    double *tmp = nullptr;
    if ((rc = dalloc(&tmp, g.plane, false))) return rc;
    k_flags_as_double<<<(unsigned)((g.plane + 255) / 256), 256, 0, g.stream>>>(g.flags, tmp, g.plane);
    rc = download_planes_fwd(out, tmp, 1);
  } else if (!base) {
    rc = fail(-26, "beom_gpu_debug_static: no such plane (%s)", nm.c_str());
  } else {
    rc = download_planes_fwd(out, base + (size_t)index * g.plane, 1);
  }
  g.win_first = keep_first; g.win_stride = keep_stride;
  (void)n;
  return rc;
}

extern "C" {

int beom_gpu_upload_state(const double *hlay, const double *u, const double *v) {
  if (!g.ready) return fail(-20, "beom_gpu_upload_state: not initialised");
  const size_t pl = g.plane, nl = (size_t)g.nlay;
  const double *src[3] = {hlay, u, v};
  g.cur = 0;
  for (int f = 0; f < 5; f++) CK(cudaMemsetAsync(g.st[f][0], 0, pl * nl * sizeof(double), g.stream));
  for (auto p : g.rs) CK(cudaMemsetAsync(p, 0, pl * nl * sizeof(double), g.stream));
  for (auto p : g.dx) CK(cudaMemsetAsync(p, 0, pl * nl * sizeof(double), g.stream));
  for (auto p : g.dy) CK(cudaMemsetAsync(p, 0, pl * nl * sizeof(double), g.stream));
  g.rs_o = g.dx_o = g.dy_o = 0;
  g.stress_const_done = false;
  drop_graphs_fwd();  // (captured with the stress state and buffer phases of the run so far)
  g.direct_steps = 0;
  // one H2D copy per field (all layers), then scatter into the dense planes
  const int n = g.p_hi - g.p_lo + 1;
  for (int f = 0; f < 3; f++) {
    for (int l = 0; l < g.nlay; l++)
      CK(cudaMemcpyAsync(g.stage + ((size_t)f * nl + l) * n, src[f] + (size_t)l * g.win_stride + (g.p_lo - g.win_first), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  }
  for (int f = 0; f < 3; f++)
    for (int l = 0; l < g.nlay; l++) {
      k_scatter<double><<<(n + 255) / 256, 256, 0, g.stream>>>(g.st[f][0] + (size_t)l * pl, g.stage + ((size_t)f * nl + l) * n, g.d_cell, g.p_lo, n);
      g.launches++;
    }
  // halo rows come straight from the caller's arrays: no exchange here (upload is not a collective) -- except across the
  // seam of a y-periodic slab chain, whose images are another rank's rows (every rank uploads, so the ring closes)
  for (int f = 0; f < 3; f++)
    if (int rc = sync_fields({{g.st[f][0], g.nlay}}, g.ring)) return rc;
  const size_t no = g.orphans.size();
  for (int f = 0; f < 3; f++)
    for (int l = 0; l < g.nlay; l++)
      for (size_t k = 0; k < no; k++) g.orph.val[((size_t)f * nl + l) * no + k] = src[f][(size_t)l * g.win_stride + (g.orphans[k] - g.win_first)];
  g.orph.forget();
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

int beom_gpu_stress(void) {
  if (!g.ready) return fail(-20, "beom_gpu_stress: not initialised");
  return run_stress();
}

}  // extern "C"
namespace {

// One step, launched kernel by kernel on g.stream (g.D.ctim / ramp / gene already set).
int step_direct(int tstp, int upst, int first_three) {
  if (g.use_fused && fused_supports(first_three != 0, upst != 0)) {
    Dev D = g.D;
    set_state_pointers(D);
    Dev Dout = D;
    const int nxt = g.cur ^ 1;
    Dout.hlay = g.st[0][nxt]; Dout.u = g.st[1][nxt]; Dout.v = g.st[2][nxt]; Dout.h_u = g.st[3][nxt]; Dout.h_v = g.st[4][nxt];
    int nlaunch = 0, rc;
    if (first_three) {  // pm:2166-2177: the start-up steps first rebuild centred fluxes from u, v, hlay
      k_centred_flux<<<cell_grid(D, g.nlay, kBlock), kBlock, 0, g.stream>>>(D);
      g.launches++;
      if ((rc = sync_fields({{D.h_u, g.nlay}, {D.h_v, g.nlay}}))) return rc;
    }
    const double *fnud_plain = D.fnud;
    if (D.has_tide) {  // this step's targets with the tidal term (the fused step's stream table has no room for the tidal planes)
      if (!g.fnud_tide && (rc = dalloc(&g.fnud_tide, g.plane * (size_t)g.nlay * 3))) return rc;
      k_tide_targets<<<dim3((unsigned)((g.plane + 255) / 256), (unsigned)g.nlay), 256, 0, g.stream>>>(D, g.fnud_tide);
      g.launches++;
      D.fnud = g.fnud_tide;
    }
    const bool obc = g.D.has_nudg && g.P.mcbc < 0.5 && g.obc_any;  // (decided alike on every rank: the exchanges are collective)
    // The edge rows first, their exchange on the communication stream while the interior rows are computed.  Measured on
    // 8192 x 8192 x 4 (profiles/r2_bench_n8*.json, r2_bench_n2*.json): 8 GPUs 1.308 ms per step against 1.378 without, 2 GPUs
    // 4.70 against 4.75.  (Round 1 had it slower: the grid was sized for exactly two waves then, and NCCL's copy CTAs displaced
    // some of them; with chunks of ~64-256 rows there are eight or more short waves.)  BEOM_OVERLAP=0 switches it off.
    static const bool want_overlap = !(getenv("BEOM_OVERLAP") && atoi(getenv("BEOM_OVERLAP")) == 0);
    // (not with periodic images: they mirror every owned row, so they can only be refreshed once the interior rows are done)
    const bool overlap = want_overlap && g.nranks > 1 && !obc && g.nmir == 0 && (D.y_hi - D.y_lo + 1) >= 4 * G;
    if (overlap) {
      // y-slabs: the G rows next to each neighbour first, then their exchange on the communication stream while the
      // rows in between are computed (the interior reads time level n only, the exchange touches halo rows of n+1)
      if ((rc = fused_step(D, Dout, tstp, false, g.stream, &nlaunch, 1, G))) return fail(rc, "fused_step (edge rows) failed: %s", cudaGetErrorString(cudaGetLastError()));
      CK(cudaEventRecord(g.ev_edge, g.stream));
      CK(cudaStreamWaitEvent(g.comm_stream, g.ev_edge, 0));
      rc = sync_fields({{Dout.hlay, g.nlay}, {Dout.u, g.nlay}, {Dout.v, g.nlay}, {Dout.h_u, g.nlay}, {Dout.h_v, g.nlay},
                        {D.rs_new, g.nlay}, {D.dx_new, g.nlay}, {D.dy_new, g.nlay}}, true, g.comm_stream);
      if (rc) return rc;
      CK(cudaEventRecord(g.ev_comm, g.comm_stream));
      if ((rc = fused_step(D, Dout, tstp, false, g.stream, &nlaunch, 2, G))) return fail(rc, "fused_step (interior rows) failed: %s", cudaGetErrorString(cudaGetLastError()));
      CK(cudaStreamWaitEvent(g.stream, g.ev_comm, 0));
      g.launches += nlaunch;
    } else {
      rc = fused_step(D, Dout, tstp, first_three != 0, g.stream, &nlaunch);
      if (rc) return fail(rc, "fused_step failed: %s", cudaGetErrorString(cudaGetLastError()));
      g.launches += nlaunch;
      const bool shadows = g.nranks > 1 || g.nmir;  // halo rows of the neighbouring ranks / periodic images (deep torus ghosts)
      if (obc) {  // no_gradient_obc after both components (pm:2285-2288)
        // the open-boundary copy reads hlay, u, v at neighbours of the segment points (neig 5 / 7, pm:2635-2676), which
        // may be periodic images or a neighbouring rank's rows: the step wrote only the rows it owns, so refresh them first
        if (shadows) {
          rc = sync_fields({{Dout.hlay, g.nlay}, {Dout.u, g.nlay}, {Dout.v, g.nlay}, {Dout.h_u, g.nlay}, {Dout.h_v, g.nlay},
                            {D.rs_new, g.nlay}, {D.dx_new, g.nlay}, {D.dy_new, g.nlay}});
          if (rc) return rc;
        }
        Dev Dobc = D;
        Dobc.fnud = fnud_plain;  // (the open-boundary copy uses the plain targets, pm:2613-2679)
        Dobc.hlay = Dout.hlay; Dobc.u = Dout.u; Dobc.v = Dout.v; Dobc.h_u = Dout.h_u; Dobc.h_v = Dout.h_v;
        for (int pass = 0; pass < 2 && g.nseg > 0; pass++) {
          k_obc<<<(g.nseg + 63) / 64, 64, 0, g.stream>>>(Dobc, g.d_seg, g.nseg, pass);
          g.launches++;
        }
      }
      if (shadows) {
        if (obc) rc = sync_fields({{Dout.u, g.nlay}, {Dout.v, g.nlay}, {Dout.h_u, g.nlay}, {Dout.h_v, g.nlay}});  // what k_obc rewrote
        else rc = sync_fields({{Dout.hlay, g.nlay}, {Dout.u, g.nlay}, {Dout.v, g.nlay}, {Dout.h_u, g.nlay}, {Dout.h_v, g.nlay},
                               {D.rs_new, g.nlay}, {D.dx_new, g.nlay}, {D.dy_new, g.nlay}});
        if (rc) return rc;
      }
    }
    g.cur = nxt;
    g.rs_o = (g.rs_o + 1) % 3;
    g.dx_o = (g.dx_o + 1) % 4;
    g.dy_o = (g.dy_o + 1) % 4;
    return 0;
  }
  return step_split(tstp, upst != 0, first_three != 0);
}

void drop_graphs() {
  for (auto &kv : g.graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  g.graphs.clear();
}
void drop_graphs_fwd() { drop_graphs(); }

// One step (with distribute_stress in front of it when `with_stress'), as ONE graph launch where the step is launch bound.
// The reference's small configurations (322 .. 63 252 points) take 5-9 kernels of a few microseconds per step; what varies
// from step to step in their arguments is only which of the rotating buffers is which (state double buffer, rs_h / dmdx /
// dmdy rings, u-first or v-first): a period of 12 steps.  So each phase is captured once -- the very launch sequence of
// step_direct, recorded instead of run -- and replayed from then on.  Not captured: the start-up steps, the first two steady
// steps after an upload (allocations and function attributes happen there), steps with ramp != 1 or tides (ctim and ramp are
// kernel arguments), the rigid lid, more than one rank.  A capture that fails switches the graphs off; the step then runs
// directly, as before.  BEOM_GRAPH=0 / 1 overrides the size rule.
int step_graphed(int tstp, double ctim, double ramp, double gene, int upst, int first_three, bool with_stress) {
  g.D.ctim = ctim; g.D.ramp = ramp; g.D.gene = gene;
  g.orph.record(ctim, ramp);
  int rc;
  const bool steady = !first_three && ramp == 1.0;
  if (!(g.graphs_on && steady && g.direct_steps >= 2)) {
    if (with_stress && (rc = run_stress())) return rc;
    rc = step_direct(tstp, upst, first_three);
    if (!rc && steady) g.direct_steps++;
    return rc;
  }
  if (gene != g.graph_gene) {
    drop_graphs();
    g.graph_gene = gene;
  }
  const unsigned key = (with_stress ? 1u : 0u) | (upst ? 2u : 0u) | ((unsigned)(tstp & 1) << 2) | ((unsigned)g.cur << 3) |
                       ((unsigned)g.rs_o << 4) | ((unsigned)g.dx_o << 6) | ((unsigned)g.dy_o << 8);
  auto it = g.graphs.find(key);
  if (it == g.graphs.end()) {
    const int cur = g.cur, rs = g.rs_o, dx = g.dx_o, dy = g.dy_o;
    const bool sd = g.stress_const_done;
    const long long l0 = g.launches;
    cudaGraph_t graph = nullptr;
    Ctx::StepGraph sg;
    bool bad = cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeRelaxed) != cudaSuccess;
    if (!bad) {
      rc = with_stress ? run_stress() : 0;
      if (!rc) rc = step_direct(tstp, upst, 0);
      const cudaError_t e = cudaStreamEndCapture(g.stream, &graph);
      bad = rc != 0 || e != cudaSuccess || !graph;
    }
    if (!bad) bad = cudaGraphInstantiate(&sg.exec, graph, 0) != cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (bad) {  // nothing has run: put the indices back and do this step (and all later ones) directly
      cudaGetLastError();
      g.cur = cur; g.rs_o = rs; g.dx_o = dx; g.dy_o = dy;
      g.stress_const_done = sd;
      g.launches = l0;
      g.graphs_on = false;
      drop_graphs();
      if (with_stress && (rc = run_stress())) return rc;
      return step_direct(tstp, upst, 0);
    }
    sg.launches = g.launches - l0;
    sg.cur = g.cur; sg.rs_o = g.rs_o; sg.dx_o = g.dx_o; sg.dy_o = g.dy_o;
    sg.stress_done = g.stress_const_done;
    g.launches = l0;
    it = g.graphs.emplace(key, sg).first;
  }
  const Ctx::StepGraph &sg = it->second;
  CK(cudaGraphLaunch(sg.exec, g.stream));
  g.cur = sg.cur; g.rs_o = sg.rs_o; g.dx_o = sg.dx_o; g.dy_o = sg.dy_o;
  g.stress_const_done = sg.stress_done;
  g.launches += sg.launches;
  g.graph_launches++;
  return 0;
}

}  // namespace
extern "C" {

int beom_gpu_step(int tstp, double ctim, double ramp, double gene, int upst, int first_three) {
  if (!g.ready) return fail(-20, "beom_gpu_step: not initialised");
  return step_graphed(tstp, ctim, ramp, gene, upst, first_three, false);
}

long long beom_gpu_graph_launch_count(void) { return g.graph_launches; }

int beom_gpu_advance(int tstp0, int tstp1, double tres) {
  if (!g.ready) return fail(-20, "beom_gpu_advance: not initialised");
  const beom_params &P = g.P;
  const double dtd8 = P.dt / 24.0 / 3600.0;
  double &ramp = g.adv_ramp, &gene = g.adv_gene;  // carried from call to call (steps 1-3 keep the ramp of step 1, pm:1862-1866)
  for (int tstp = tstp0; tstp <= tstp1; tstp++) {
    const double ctim = tres + dtd8 * (double)tstp;
    int rc;
    if (tstp <= 3) {
      if (tstp == 1) {
        ramp = 1.0; gene = 0.0;
        if ((rc = run_stress())) return rc;
        if (P.rsta < 0.5 && ctim < P.dt_r) ramp = ctim / P.dt_r;
      }
      if ((rc = beom_gpu_step(tstp, ctim, ramp, gene, 1, 1))) return rc;
      if (tstp == 3) {
        gene = P.g_fb;
        if (gene > 0.5 && P.rgld > 0.5) gene = 0.0;
      }
    } else {
      const bool upst = (tstp % g.n_3d) == 0;
      if (tstp == tstp0) {
        gene = P.g_fb;
        if (gene > 0.5 && P.rgld > 0.5) gene = 0.0;
      }
      ramp = 1.0;
      if (P.rsta < 0.5 && ctim < P.dt_r) ramp = ctim / P.dt_r;
      // distribute_stress (pm:1895, before the ramp is set: it does not use it) + the step, one graph launch where that pays
      if ((rc = step_graphed(tstp, ctim, ramp, gene, upst ? 1 : 0, 0, upst))) return rc;
    }
  }
  return 0;
}

static int download_planes(double *dst, const double *dense, int nplanes) {
  // dst: reference layout [nplanes][ndeg+1]; only this rank's owned+halo points are written
  const int n = g.p_hi - g.p_lo + 1;
  for (int p0 = 0; p0 < nplanes; p0 += 3 * g.nlay) {
    const int np = std::min(3 * g.nlay, nplanes - p0);
    for (int p = 0; p < np; p++) {
      k_gather<double><<<(n + 255) / 256, 256, 0, g.stream>>>(g.stage + (size_t)p * n, dense + (size_t)(p0 + p) * g.plane, g.d_cell, g.p_lo, n);
      g.launches++;
    }
    for (int p = 0; p < np; p++)
      CK(cudaMemcpyAsync(dst + (size_t)(p0 + p) * g.win_stride + (g.p_lo - g.win_first), g.stage + (size_t)p * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
  }
  return 0;
}

}  // extern "C"
namespace {
int download_planes_fwd(double *dst, const double *dense, int nplanes) { return download_planes(dst, dense, nplanes); }
}  // namespace
extern "C" {

int beom_gpu_download_state(double *hlay, double *u, double *v) {
  if (!g.ready) return fail(-20, "beom_gpu_download_state: not initialised");
  double *dst[3] = {hlay, u, v};
  const size_t nl = (size_t)g.nlay, no = g.orphans.size();
  g.orph.replay();  // the steps since the last download, for the duplicates a sponge moves
  for (int f = 0; f < 3; f++) {
    if (!dst[f]) continue;
    int rc = download_planes(dst[f], g.st[f][g.cur], g.nlay);
    if (rc) return rc;
    for (int l = 0; l < g.nlay; l++) {
      if (g.win_first == 0) dst[f][(size_t)l * g.win_stride] = 0.0;  // the discarded cell
      for (size_t k = 0; k < no; k++) dst[f][(size_t)l * g.win_stride + (g.orphans[k] - g.win_first)] = g.orph.val[((size_t)f * nl + l) * no + k];
    }
  }
  CK(cudaGetLastError());
  return 0;
}

int beom_gpu_download_aux(double *h_u, double *h_v, double *rs_h, double *dmdx, double *dmdy) {
  if (!g.ready) return fail(-20, "beom_gpu_download_aux: not initialised");
  int rc;
  if (h_u && (rc = download_planes(h_u, g.st[3][g.cur], g.nlay))) return rc;
  if (h_v && (rc = download_planes(h_v, g.st[4][g.cur], g.nlay))) return rc;
  const int n = g.p_hi - g.p_lo + 1;
  auto hist = [&](double *dst, int nh, double *a, double *b, double *c) -> int {
    for (int l = 0; l < g.nlay; l++) {
      const size_t L = (size_t)l * g.plane;
      k_gather_hist<<<(n + 255) / 256, 256, 0, g.stream>>>(g.stage, a + L, b + L, c ? c + L : nullptr, nh, g.d_cell, g.p_lo, n);
      g.launches++;
      CK(cudaMemcpyAsync(dst + ((size_t)l * g.win_stride + (g.p_lo - g.win_first)) * nh, g.stage, (size_t)n * nh * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
      CK(cudaStreamSynchronize(g.stream));
    }
    return 0;
  };
  if (rs_h && (rc = hist(rs_h, 2, g.rs[g.rs_o], g.rs[(g.rs_o + 1) % 3], nullptr))) return rc;
  if (dmdx && (rc = hist(dmdx, 3, g.dx[g.dx_o], g.dx[(g.dx_o + 1) % 4], g.dx[(g.dx_o + 2) % 4]))) return rc;
  if (dmdy && (rc = hist(dmdy, 3, g.dy[g.dy_o], g.dy[(g.dy_o + 1) % 4], g.dy[(g.dy_o + 2) % 4]))) return rc;
  CK(cudaGetLastError());
  return 0;
}

// One float32 record: out[l*ndeg + (p-1)] for the vector points p this rank holds (whole-array layout).
static int download_record(float *out) {
  const int n = g.p_hi - std::max(g.p_lo, 1) + 1, p0 = std::max(g.p_lo, 1);
  if (n <= 0) return 0;
  for (int l = 0; l < g.nlay; l++) {
    k_gather<float><<<(n + 255) / 256, 256, 0, g.stream>>>(g.rec_stage, g.rec_f32 + (size_t)l * g.plane, g.d_cell, p0, n);
    g.launches++;
    CK(cudaMemcpyAsync(out + (size_t)l * g.ndeg + (p0 - 1), g.rec_stage, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));  // the staging buffer is reused
  }
  return 0;
}
int beom_gpu_download_diag(float *pvor, float *mont, float *v_cc) {
  if (!g.ready) return fail(-20, "beom_gpu_download_diag: not initialised");
  int rc;
  const size_t pl = g.plane, nl = (size_t)g.nlay;
  if (!g.rec_f32) {
    if ((rc = dalloc(&g.rec_f32, pl * nl)) || (rc = dalloc(&g.rec_stage, (size_t)(g.p_hi - g.p_lo + 1), false))) return rc;
  }
  Dev D = g.D;
  set_state_pointers(D);
  const dim3 grid = cell_grid(D, g.nlay, kBlock);
  if (pvor) {
    k_rec_pvor<<<grid, kBlock, 0, g.stream>>>(D, g.rec_f32);
    g.launches++;
    if ((rc = download_record(pvor))) return rc;
  }
  if (mont) {
    k_rec_mont<<<grid, kBlock, 0, g.stream>>>(D, g.rec_f32);
    g.launches++;
    if ((rc = download_record(mont))) return rc;
  }
  if (v_cc) {
    if ((rc = alloc_split_buffers())) return rc;  // wrk1 / wrk2 live in the split path's rvor / dive planes
    D = g.D;
    set_state_pointers(D);
    k_rec_vort_dive<<<cell_grid(D, g.nlay, kBlock, 2, 2), kBlock, 0, g.stream>>>(D, D.rvor, D.dive);
    g.launches++;
    if ((rc = sync_fields({{D.rvor, g.nlay}, {D.dive, g.nlay}}))) return rc;
    k_rec_vcc<<<grid, kBlock, 0, g.stream>>>(D, D.rvor, D.dive, g.rec_f32);
    g.launches++;
    if ((rc = download_record(v_cc))) return rc;
  }
  CK(cudaGetLastError());
  return 0;
}
int beom_gpu_set_rest_thickness(const float *h_0_r4) {
  if (!g.ready) return fail(-20, "beom_gpu_set_rest_thickness: not initialised");
  if (!h_0_r4) return fail(-1, "beom_gpu_set_rest_thickness: null argument");
  int rc;
  const size_t cnt = (size_t)g.ndeg * g.nlay;
  if (!g.h0r4 && (rc = dalloc(&g.h0r4, cnt, false))) return rc;
  CK(cudaMemcpyAsync(g.h0r4, h_0_r4, cnt * sizeof(float), cudaMemcpyHostToDevice, g.stream));
  CK(cudaStreamSynchronize(g.stream));  // (the caller's array may be pageable and short-lived)
  const size_t no = g.orphans.size();
  g.h0r4_orph.resize((size_t)g.nlay * no);
  for (int l = 0; l < g.nlay; l++)
    for (size_t k = 0; k < no; k++) g.h0r4_orph[(size_t)l * no + k] = h_0_r4[(size_t)l * g.ndeg + (g.orphans[k] - 1)];
  return 0;
}

int beom_gpu_records_begin(int with_diag) {
  if (!g.ready) return fail(-20, "beom_gpu_records_begin: not initialised");
  if (!g.h0r4) return fail(-22, "beom_gpu_records_begin: beom_gpu_set_rest_thickness has not been called");
  if (g.rec_pending >= 2) return fail(-23, "beom_gpu_records_begin: two record sets are in flight already (call beom_gpu_records_wait)");
  int rc;
  const int p0 = std::max(g.p_lo, 1), n = g.p_hi - p0 + 1;
  if (n <= 0) return fail(-24, "beom_gpu_records_begin: this rank holds no vector point");
  const size_t nl = (size_t)g.nlay, per = nl * (size_t)n, pl = g.plane;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (!g.copy_stream) CK(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
  if (!g.mm_partial && (rc = dalloc(&g.mm_partial, (size_t)blocks * nl * 2, false))) return rc;
  Ctx::RecSlot &r = g.rec[(g.rec_head + g.rec_pending) % 2];
  if (!r.dev) {
    if ((rc = dalloc(&r.dev, 6 * per, false)) || (rc = dalloc(&r.mm_dev, 2 * nl, false))) return rc;
    CK(cudaHostAlloc(&r.host, 6 * per * sizeof(float), cudaHostAllocDefault));
    CK(cudaHostAlloc(&r.mm_host, 2 * nl * sizeof(double), cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&r.ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming));
  }
  if (r.used) CK(cudaStreamWaitEvent(g.stream, r.done, 0));  // the previous copy out of this slot's device buffer
  Dev D = g.D;
  set_state_pointers(D);
  k_rec_state<<<blocks, 256, 0, g.stream>>>(D, g.h0r4, g.d_cell, p0, n, g.ndeg, r.dev, g.mm_partial);
  k_rec_minmax<<<(unsigned)g.nlay, 32, 0, g.stream>>>(g.mm_partial, (int)blocks, r.mm_dev);
  g.launches += 2;
  if (with_diag) {
    if (!g.rec_f32 && ((rc = dalloc(&g.rec_f32, pl * nl)) || (rc = dalloc(&g.rec_stage, (size_t)(g.p_hi - g.p_lo + 1), false)))) return rc;
    if ((rc = alloc_split_buffers())) return rc;  // wrk1 / wrk2 of the v_cc record live in the split path's rvor / dive planes
    D = g.D;
    set_state_pointers(D);
    const dim3 grid = cell_grid(D, g.nlay, kBlock);
    auto gather = [&](int which) {
      for (int l = 0; l < g.nlay; l++) {
        k_gather<float><<<blocks, 256, 0, g.stream>>>(r.dev + (size_t)which * per + (size_t)l * n, g.rec_f32 + (size_t)l * pl, g.d_cell, p0, n);
        g.launches++;
      }
    };
    k_rec_pvor<<<grid, kBlock, 0, g.stream>>>(D, g.rec_f32);
    gather(3);
    k_rec_mont<<<grid, kBlock, 0, g.stream>>>(D, g.rec_f32);
    gather(4);
    k_rec_vort_dive<<<cell_grid(D, g.nlay, kBlock, 2, 2), kBlock, 0, g.stream>>>(D, D.rvor, D.dive);
    g.launches += 3;
    if ((rc = sync_fields({{D.rvor, g.nlay}, {D.dive, g.nlay}}))) return rc;
    k_rec_vcc<<<grid, kBlock, 0, g.stream>>>(D, D.rvor, D.dive, g.rec_f32);
    g.launches++;
    gather(5);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(r.ready, g.stream));
  CK(cudaStreamWaitEvent(g.copy_stream, r.ready, 0));
  CK(cudaMemcpyAsync(r.host, r.dev, (with_diag ? 6 : 3) * per * sizeof(float), cudaMemcpyDeviceToHost, g.copy_stream));
  CK(cudaMemcpyAsync(r.mm_host, r.mm_dev, 2 * nl * sizeof(double), cudaMemcpyDeviceToHost, g.copy_stream));
  CK(cudaEventRecord(r.done, g.copy_stream));
  g.orph.replay();  // the duplicates a sponge moves, as of this step (host side, orphans.h)
  r.orph = g.orph.val;
  r.busy = true; r.used = true; r.diag = with_diag != 0;
  g.rec_pending++;
  return 0;
}

int beom_gpu_records_wait(beom_records *out) {
  if (!g.ready) return fail(-20, "beom_gpu_records_wait: not initialised");
  if (!out) return fail(-1, "beom_gpu_records_wait: null argument");
  if (g.rec_pending == 0) return fail(-25, "beom_gpu_records_wait: no record set has been begun");
  Ctx::RecSlot &r = g.rec[g.rec_head];
  CK(cudaEventSynchronize(r.done));
  const int p0 = std::max(g.p_lo, 1), n = g.p_hi - p0 + 1;
  const size_t nl = (size_t)g.nlay, per = nl * (size_t)n, no = g.orphans.size();
  // frozen periodic duplicates have no cell: their state lives on the host (same float32 arithmetic as k_rec_state)
  for (size_t k = 0; k < no; k++) {
    const int o = g.orphans[k];
    if (o < p0 || o > g.p_hi) continue;
    float acc = 0.0f;
    for (int l = g.nlay - 1; l >= 0; l--) {
      const double dh = r.orph[((size_t)0 * nl + l) * no + k] - (double)g.h0r4_orph[(size_t)l * no + k];
      acc = (l == g.nlay - 1) ? (float)dh : (float)(dh + (double)acc);
      if (!(l == 0 && g.P.rgld > 0.5)) r.host[(size_t)l * n + (o - p0)] = acc;
      r.host[per + (size_t)l * n + (o - p0)] = (float)r.orph[((size_t)1 * nl + l) * no + k];
      r.host[2 * per + (size_t)l * n + (o - p0)] = (float)r.orph[((size_t)2 * nl + l) * no + k];
    }
  }
  memset(out, 0, sizeof *out);
  out->eta = r.host; out->u = r.host + per; out->v = r.host + 2 * per;
  if (r.diag) { out->pvor = r.host + 3 * per; out->mont = r.host + 4 * per; out->v_cc = r.host + 5 * per; }
  out->first_point = p0; out->count = n;
  for (int l = 0; l < g.nlay; l++) {
    out->hmin[l] = r.mm_host[2 * l]; out->hmax[l] = r.mm_host[2 * l + 1];
    if (!out->thin_layer && out->hmin[l] < 0.5 * g.P.hmin) out->thin_layer = l + 1;
  }
  r.busy = false;
  g.rec_head = (g.rec_head + 1) % 2;
  g.rec_pending--;
  return 0;
}

int beom_gpu_pi_iterations(int *iters) {
  if (!g.ready) return fail(-20, "beom_gpu_pi_iterations: not initialised");
  if (!g.pi_iters) return fail(-30, "beom_gpu_pi_iterations: rgld = 0, there is no surface-pressure solve");
  CK(cudaMemcpyAsync(iters, g.pi_iters, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}
int beom_gpu_download_pi_s(double *pi_s) {
  if (!g.ready) return fail(-20, "beom_gpu_download_pi_s: not initialised");
  if (!g.D.pi_s) return fail(-30, "beom_gpu_download_pi_s: rgld = 0, there is no surface pressure");
  const size_t keep_first = g.win_first, keep_stride = g.win_stride;
  g.win_first = 0; g.win_stride = (size_t)g.ndeg + 1;  // pi_s(0:ndeg) is always a whole array
  int rc = download_planes(pi_s, g.D.pi_s, 1);
  g.win_first = keep_first; g.win_stride = keep_stride;
  if (!rc) pi_s[0] = 0.0;
  return rc;
}
int beom_gpu_diagnostics_all(const double *h_0, double *vol, double *ke, double *pe, double *enst, double *zeta, double *zeta2, double *npts) {
  if (!g.ready) return fail(-20, "beom_gpu_diagnostics: not initialised");
  int rc;
  const size_t pl = g.plane, nl = (size_t)g.nlay;
  Dev D = g.D;
  set_state_pointers(D);
  const int p0 = std::max(g.p_lo, 1), np = g.p_hi - p0 + 1;  // this rank's vector points (the kernel keeps the rows it owns)
  const dim3 block(256, 1, 1);
  const dim3 grid((unsigned)((std::max(np, 1) + 255) / 256), (unsigned)g.nlay, 1);
  const size_t per_layer = (size_t)grid.x;
  if (!g.diag_h0 && (rc = dalloc(&g.diag_h0, pl * nl))) return rc;
  if (!g.diag_partial) {
    if ((rc = dalloc(&g.diag_partial, kNQ * per_layer * nl)) || (rc = dalloc(&g.diag_out, kNQ * nl + 1))) return rc;
    g.diag_blocks = per_layer;
  }
  if (h_0) {  // rest thickness h_0(0:ndeg, nlay), reference layout; static, but cheap enough to refresh per call
    const size_t keep_first = g.win_first, keep_stride = g.win_stride;
    g.win_first = 0; g.win_stride = (size_t)g.ndeg + 1;
    rc = upload_planes(g.diag_h0, h_0, g.nlay);
    g.win_first = keep_first; g.win_stride = keep_stride;
    if (rc) return rc;
  }
  k_conservation<<<grid, block, 0, g.stream>>>(D, g.diag_h0, g.d_cell, p0, np, g.diag_partial);
  k_sum_partials<<<(unsigned)g.nlay, 32, 0, g.stream>>>(g.diag_partial, (int)per_layer, g.diag_out);
  g.launches += 2;
  if (g.nranks > 1 && (rc = comm_allreduce_sum(g.diag_out, kNQ * nl, g.stream, &g_err))) return fail(rc, "beom_gpu_diagnostics: %s", g_err.c_str());
  std::vector<double> h(kNQ * nl);
  CK(cudaMemcpyAsync(h.data(), g.diag_out, kNQ * nl * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  for (int l = 0; l < g.nlay; l++) {
    if (vol) vol[l] = h[kNQ * l + 0];
    if (ke) ke[l] = h[kNQ * l + 1];
    if (enst) enst[l] = h[kNQ * l + 3];
    if (zeta) zeta[l] = h[kNQ * l + 4];
    if (zeta2) zeta2[l] = h[kNQ * l + 5];
  }
  if (pe) pe[0] = h[2];
  if (npts) *npts = h[6];  // vector points that are not frozen periodic duplicates
  return 0;
}
int beom_gpu_diagnostics(const double *h_0, double *vol, double *ke, double *pe) {
  return beom_gpu_diagnostics_all(h_0, vol, ke, pe, nullptr, nullptr, nullptr, nullptr);
}

int beom_gpu_point_range(int *first, int *count, int *own_first, int *own_count) {
  if (!g.ready) return fail(-20, "beom_gpu_point_range: not initialised");
  if (first) *first = g.p_lo;
  if (count) *count = g.p_hi - g.p_lo + 1;
  // owned = the vector points of rows j0..j1, periodic duplicates included (their state is patched in by the downloads)
  if (own_first) *own_first = g.own_first;
  if (own_count) *own_count = g.own_last - g.own_first + 1;
  return 0;
}
int beom_gpu_set_window(int first, int count) {
  if (!g.ready) return fail(-20, "beom_gpu_set_window: not initialised");
  if (count <= 0) { g.win_first = 0; g.win_stride = (size_t)g.ndeg + 1; return 0; }
  if (first > g.p_lo || first + count - 1 < g.p_hi) return fail(-21, "beom_gpu_set_window: window [%d,%d] does not cover this rank's points [%d,%d]", first, first + count - 1, g.p_lo, g.p_hi);
  g.win_first = (size_t)first;
  g.win_stride = (size_t)count;
  return 0;
}

}  // extern "C"
namespace {
// CPUs of the NUMA node the current device hangs off (sysfs: the PCI function's numa_node, the node's cpulist); false if the
// machine does not say (one node, a VM, no sysfs)
bool device_node_cpus(cpu_set_t *set) {
  int dev = 0;
  char bus[32] = {0};
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, (int)sizeof bus, dev) != cudaSuccess) return false;
  for (char *c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  if (!f) return false;
  int node = -1;
  const int got = fscanf(f, "%d", &node);
  fclose(f);
  if (got != 1 || node < 0) return false;
  snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
  if (!(f = fopen(path, "r"))) return false;
  CPU_ZERO(set);
  int a, b, n = 0;
  while (fscanf(f, "%d", &a) == 1) {  // "0-31,64-95"
    b = a;
    int ch = fgetc(f);
    if (ch == '-') {
      if (fscanf(f, "%d", &b) != 1) break;
      ch = fgetc(f);
    }
    for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, set); n++; }
    if (ch != ',') break;
  }
  fclose(f);
  return n > 0;
}
}  // namespace
extern "C" {

// The pages are allocated (and pinned) by the calling thread inside cudaHostAlloc, on the node of the CPU it runs on: run it on
// the device's own node for the duration of the call, so that with one process per GPU on a two-socket host every rank's
// boundary copies stay on its socket.  (The B200 boxes of this pool are single-node VMs whose sysfs reports numa_node = -1, so
// there it does nothing -- profiles/r2_topology_4gpu.txt, same end-to-end time with and without, r2_bench_n4*.json; the
// aggregate host<->device rate of 4-8 ranks, 85-120 GB/s against 50 GB/s for one, is the VM's.)  BEOM_HOST_NUMA=0 switches it off.
void *beom_gpu_host_alloc(size_t bytes) {
  void *p = nullptr;
  cpu_set_t old_set, node_set;
  static const bool want = !(getenv("BEOM_HOST_NUMA") && atoi(getenv("BEOM_HOST_NUMA")) == 0);
  bool moved = false;
  if (want && sched_getaffinity(0, sizeof old_set, &old_set) == 0 && device_node_cpus(&node_set)) {
    cpu_set_t both;
    CPU_AND(&both, &old_set, &node_set);  // never widen what the caller was given
    if (CPU_COUNT(&both) > 0) moved = sched_setaffinity(0, sizeof both, &both) == 0;
  }
  const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
  if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
  if (e != cudaSuccess) { fail(-50, "cudaHostAlloc(%zu) failed", bytes); return nullptr; }
  return p;
}
void beom_gpu_host_free(void *p) { if (p) cudaFreeHost(p); }

int beom_gpu_sync(void) {
  if (!g.ready) return fail(-20, "beom_gpu_sync: not initialised");
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}
int beom_gpu_mark(int which) {
  if (!g.ready || which < 0 || which > 1) return fail(-20, "beom_gpu_mark: bad call");
  CK(cudaEventRecord(g.ev[which], g.stream));
  return 0;
}
int beom_gpu_elapsed_ms(double *ms) {
  if (!g.ready) return fail(-20, "beom_gpu_elapsed_ms: not initialised");
  CK(cudaEventSynchronize(g.ev[1]));
  float t = 0;
  CK(cudaEventElapsedTime(&t, g.ev[0], g.ev[1]));
  *ms = (double)t;
  return 0;
}

int beom_gpu_comm_unique_id(char id[128]) { return comm_unique_id(id, &g_err); }
int beom_gpu_comm_init(const char id[128], int rank, int nranks, int device) { return comm_init(id, rank, nranks, device, &g_err); }
int beom_gpu_comm_finalize(void) { return comm_finalize(); }

}  // extern "C"
