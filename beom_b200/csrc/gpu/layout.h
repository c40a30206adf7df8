// layout.h -- the dense-plane layout of a rank and the analysis of the reference's neighbour table that produces it
// (plain C++, no CUDA: beom_gpu_init calls it, tools/layout_host.cc exposes it to CPU-only tests).
//
// Every vector point of the reference is a point (i, j) of the padded (lm+2) x (mm+2) grid and neig(k, p) is the
// vector entry of grid point (i+di, j+dj) -- or 0, or, on periodic domains, the entry of the periodic image
// (private_mod.f95:614-726).  analyse_layout() turns the table into
//   * cell_of_point : vector index -> dense cell of this rank's planes (-1: not held here, -2: periodic duplicate)
//   * flags         : one byte per cell (the five 0/1 masks + "is a vector point")
//   * mirror lists  : cells that must show another cell's values (periodic images), refreshed after each kernel;
//                     "deep" images for the fused step when the aliases form a complete torus (one rank only)
//   * orphans       : vector points displaced by a mirror (the duplicate column/row): state kept on the host (orphans.h)
//   * ring          : a y-periodic domain split into y-slabs: the images in row 0 (of row mm) and in row mm+1 (of
//                     row 1) live on the other end of the slab chain, so the halo exchange is ring-closed:
//                     rank 0 <-> rank n-1, with the last rank sending rows mm-G+1..mm and receiving into mm+1..mm+G
//                     (one row lower than a plain slab boundary: row mm+1 is the duplicate of row 1, not a row of its own)
#pragma once
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace beom {

constexpr int G = 4;     // halo width (cells); the fused step needs 3
constexpr int GX0 = 15;  // X = i + GX0  -> i = 1 sits at X = 16: the fused step's row segments start on 128-byte lines

// per-cell flag bits (the reference's 0./1. masks, private_mod.f95:54-58, plus "is a vector point")
// F_GHOST: the cell shows a periodic image of an active cell (a mirror cell): never updated in place, but the fused
// step, which recomputes its halo instead of re-reading it, evaluates it like the cell it mirrors
enum : uint8_t { F_N = 1, F_U = 2, F_V = 4, F_PE = 8, F_PI = 16, F_ACT = 32, F_GHOST = 64 };

struct Layout {
  int lm = 0, mm = 0, ndeg = 0;
  int rank = 0, nranks = 1;
  int j0 = 1, j1 = 1;      // owned grid rows (inclusive)
  int p_lo = 1, p_hi = 0;  // vector points held on this rank (owned + halo rows), inclusive
  int NX = 0, NY = 0, j_off = 0;
  size_t plane = 0;
  std::vector<int> cell_of_point;  // [ndeg+1]
  std::vector<uint8_t> flags;      // [plane]
  std::vector<int> point_of_cell;  // [plane], 0 = none
  std::vector<int> mdst, msrc;     // local mirrors: dst cell <- src cell
  std::vector<int> orphans;
  bool torus = false;
  bool ring = false;
  int rc = 0;
  std::string error;
};

inline int layout_fail(Layout &L, int rc, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  L.rc = rc;
  L.error = buf;
  return rc;
}

// subc: [2][ndeg+1] (i, then j); neig: [ndeg+1][8] (E, NE, N, NW, W, SW, S, SE); masks: [ndeg+1] of 0./1.
inline int analyse_layout(Layout &L, int lm, int mm, int ndeg, bool xper, bool yper, int rank, int nranks, const int32_t *subc,
                          const int32_t *neig, const double *mk_n, const double *mk_u, const double *mk_v, const double *mkpe,
                          const double *mkpi) {
  L = Layout();
  L.lm = lm; L.mm = mm; L.ndeg = ndeg; L.rank = rank; L.nranks = nranks;
  const size_t nd1 = (size_t)ndeg + 1;
  const int32_t *si = subc, *sj = subc + nd1;

  // y-slab owned by this rank: rows 1..mm+1 split evenly
  {
    const int rows = mm + 1, base = rows / nranks, rem = rows % nranks;
    L.j0 = 1 + rank * base + std::min(rank, rem);
    L.j1 = L.j0 + base + (rank < rem ? 1 : 0) - 1;
    if (L.j1 < L.j0) return layout_fail(L, -5, "beom_gpu_init: more ranks than grid rows");
  }
  L.NX = ((lm + GX0 + 34) + 15) / 16 * 16;  // room for the fused kernel's last 28-column tile (+ 4 staged columns) past x_hi
  L.NY = (L.j1 - L.j0 + 1) + 2 * G;
  L.plane = (size_t)L.NX * L.NY;
  if (L.plane > 0x7fffffffull) return layout_fail(L, -6, "beom_gpu_init: plane too large for 32-bit cell offsets");
  const int j_off = L.j_off = G - L.j0;  // Y = j + j_off
  const int NX = L.NX;

  // vector points on this device: rows j0-G .. j1+G (vector order is j outer, i inner: contiguous)
  L.cell_of_point.assign(nd1, -1);
  L.p_lo = ndeg + 1; L.p_hi = 0;
  for (int p = 1; p <= ndeg; p++) {
    const int j = sj[p], i = si[p];
    if (j < L.j0 - G || j > L.j1 + G) continue;
    if (i < 1 - GX0 || i + GX0 >= NX) return layout_fail(L, -7, "beom_gpu_init: subc out of range at point %d", p);
    L.cell_of_point[p] = (j + j_off) * NX + (i + GX0);
    L.p_lo = std::min(L.p_lo, p);
    L.p_hi = std::max(L.p_hi, p);
  }
  if (L.p_hi < L.p_lo) return layout_fail(L, -8, "beom_gpu_init: no grid points on rank %d", rank);

  // dense flags + point-of-cell map
  auto flags_of_point = [&](int p) -> uint8_t {
    uint8_t f = F_ACT;
    if (mk_n[p] > 0.5) f |= F_N;
    if (mk_u[p] > 0.5) f |= F_U;
    if (mk_v[p] > 0.5) f |= F_V;
    if (mkpe[p] > 0.5) f |= F_PE;
    if (mkpi[p] > 0.5) f |= F_PI;
    return f;
  };
  std::vector<uint8_t> &hflags = L.flags;
  std::vector<int> &point_of_cell = L.point_of_cell;
  hflags.assign(L.plane, 0);
  point_of_cell.assign(L.plane, 0);
  for (int p = L.p_lo; p <= L.p_hi; p++) {
    const int c = L.cell_of_point[p];
    if (c < 0) continue;
    hflags[c] = flags_of_point(p);
    point_of_cell[c] = p;
  }
  // mirror cells: wherever neig(k,p) is not the point sitting at (i+di, j+dj) (periodic aliases,
  // private_mod.f95:614-685).  A vector point whose own cell must show another point's values is an
  // "orphan": nothing ever reads it and its masks are zero, so its state is frozen (or follows the sponge, orphans.h).
  static const int di[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dj[8] = {0, 1, 1, 1, 0, -1, -1, -1};
  std::vector<int> &mdst = L.mdst, &msrc = L.msrc;
  std::vector<int> alias_of_cell(L.plane, -1);
  std::vector<int> rdst, rsrc;  // remote images (ring): dst cell on this rank, source POINT on another rank
  auto wrap = [](int k, int n) { return ((k - 1) % n + n) % n + 1; };
  for (int p = L.p_lo; p <= L.p_hi; p++) {
    if (L.cell_of_point[p] < 0) continue;
    const int j = sj[p], i = si[p];
    // only cells whose neighbours are read -- plus, on the y-slabs of a domain periodic in x, the east/west images of
    // the deeper halo rows: the fused step recomputes those rows and must see them exactly as their owner does
    const bool near = j >= L.j0 - 1 && j <= L.j1 + 1;
    if (!near && !(xper && nranks > 1)) continue;
    for (int k = 0; k < 8; k++) {
      if (!near && dj[k] != 0) continue;
      const int q = neig[(size_t)p * 8 + k];
      const int c = (j + dj[k] + j_off) * NX + (i + di[k] + GX0);
      if (q == point_of_cell[c] && alias_of_cell[c] < 0) continue;
      if (q == 0) {
        if (alias_of_cell[c] == 0 || point_of_cell[c] == 0) continue;
        return layout_fail(L, -9, "beom_gpu_init: neig(%d,%d) = 0 but a grid point exists there", k + 1, p);
      }
      if (alias_of_cell[c] >= 0) {
        if (alias_of_cell[c] != q) return layout_fail(L, -9, "beom_gpu_init: inconsistent connectivity at point %d", p);
        continue;
      }
      const bool remote = q < L.p_lo || q > L.p_hi || L.cell_of_point[q] < 0;
      if (remote) {
        // The image lives on another rank.  Only the seam of a y-periodic domain split into y-slabs is understood:
        // row 0 on the first rank shows row mm, row mm+1 on the last rank shows row 1 (same column, modulo xper).
        const int jc = j + dj[k], ic = i + di[k];
        const bool seam_lo = yper && nranks > 1 && rank == 0 && jc == 0 && sj[q] == mm;
        const bool seam_hi = yper && nranks > 1 && rank == nranks - 1 && jc == mm + 1 && sj[q] == 1;
        const int iq = si[q];
        if (!((seam_lo || seam_hi) && (iq == ic || (xper && iq == wrap(ic, lm)))))
          return layout_fail(L, -9, "beom_gpu_init: neig(%d,%d) = %d is out of range", k + 1, p, q);
      }
      alias_of_cell[c] = q;
      if (point_of_cell[c] != 0) {  // orphan
        const int o = point_of_cell[c];
        if (mk_n[o] > 0.5 || mk_u[o] > 0.5 || mk_v[o] > 0.5)
          return layout_fail(L, -9, "beom_gpu_init: unsupported periodic connectivity (aliased point %d is not masked)", o);
        L.orphans.push_back(o);
        L.cell_of_point[o] = -2;
        point_of_cell[c] = 0;
      }
      if (remote) {
        rdst.push_back(c);
        rsrc.push_back(q);
      } else {
        // an image of another ROW must come from a row this rank owns: halo rows are only refreshed after the mirrors
        if (nranks > 1 && sj[q] != j + dj[k] && (sj[q] < L.j0 || sj[q] > L.j1))
          return layout_fail(L, -9, "beom_gpu_init: the y-periodic seam falls into the halo rows of rank %d (too few rows per rank)", rank);
        mdst.push_back(c);
        msrc.push_back(L.cell_of_point[q]);
      }
    }
  }
  // a mirror's source may itself have been turned into a mirror/orphan: forbid chains
  for (size_t k = 0; k < msrc.size(); k++)
    if (msrc[k] < 0) return layout_fail(L, -9, "beom_gpu_init: chained periodic aliases are not supported");
  // The ring is a property of the parameters, so that every rank takes part in the same exchange.  It copies whole
  // rows (row mm -> row 0 of the first rank, row 1 -> row mm+1 of the last), which equals the reference's aliasing when
  // every vector point of the seam rows has been displaced by an image (a seam with dry gaps keeps points of its own
  // there) and every point of the source rows is imaged; the ranks at the seam need G rows to send.
  L.ring = yper && nranks > 1;
  if (L.ring) {
    if (L.j1 - L.j0 + 1 < G + 1) return layout_fail(L, -9, "beom_gpu_init: a y-periodic domain needs at least %d rows per rank", G + 1);
    for (int p = L.p_lo; p <= L.p_hi; p++) {
      const bool seam = (rank == 0 && sj[p] == 0) || (rank == nranks - 1 && sj[p] == mm + 1);
      if (seam && L.cell_of_point[p] >= 0)
        return layout_fail(L, -9, "beom_gpu_init: a y-periodic seam with dry gaps cannot be split into y-slabs (point %d keeps a cell of its own)", p);
    }
    for (int q = 1; q <= ndeg; q++) {  // every rank knows the whole grid
      int jc = -1;
      if (rank == 0 && sj[q] == mm) jc = 0;
      if (rank == nranks - 1 && sj[q] == 1) jc = mm + 1;
      if (jc < 0 || si[q] < 1 || si[q] > lm) continue;  // (the margin / duplicate columns show their own images)
      if (alias_of_cell[(jc + j_off) * NX + (si[q] + GX0)] != q)
        return layout_fail(L, -9, "beom_gpu_init: unsupported y-periodic connectivity (point %d of the seam is not imaged on rank %d)", q, rank);
    }
  }
  // Deep torus ghosts.  The fused step recomputes its halo (2 columns, 3-4 rows) instead of re-reading it, so on a
  // periodic domain the cells up to 3 columns / 4 rows outside the core must show periodic images too -- cells the
  // reference never indexes.  Only when the reference's own aliases (above) form a complete torus: every row
  // 1..mm aliased in x (xper), every column 1..lm aliased in y (yper).
  // On y-slabs the x images stay inside a row, hence inside a slab; the y images of the seam rows are another rank's rows
  // (the ring above, already verified), so only the x part is looked at here and the deep rows across the seam get their
  // flags from the points they show (below).
  L.torus = false;
  if ((!mdst.empty() || L.ring) && (xper || yper)) {
    const bool xp = xper, yp = yper && nranks == 1;
    auto cell = [&](int i, int j) { return (j + j_off) * NX + (i + GX0); };
    std::vector<int> img(L.plane, -1);
    for (size_t k = 0; k < mdst.size(); k++) img[mdst[k]] = msrc[k];
    bool complete = true;
    // (on a ring, the rows next to the seam have already lost their corner images to the remote list: rows 1..mm only)
    if (xp) for (int j = std::max(1, L.j0 - G); j <= std::min(mm, L.j1 + G) && complete; j++) complete = img[cell(0, j)] == cell(lm, j) && img[cell(lm + 1, j)] == cell(1, j);
    if (yp) for (int i = 1; i <= lm && complete; i++) complete = img[cell(i, 0)] == cell(i, mm) && img[cell(i, mm + 1)] == cell(i, 1);
    for (size_t k = 0; k < mdst.size() && complete; k++) {  // and nothing else: every alias is the torus image
      const int X = mdst[k] % NX - GX0, Y = mdst[k] / NX - j_off;
      complete = msrc[k] == cell(xp ? wrap(X, lm) : X, yp ? wrap(Y, mm) : Y);
    }
    if (complete) {
      for (int j = 1 - G; j <= mm + 1 + G; j++)
        for (int i = -3; i <= lm + 4; i++) {
          if (j + j_off < 0 || j + j_off >= L.NY || i + GX0 < 0 || i + GX0 >= NX) continue;
          const int c = cell(i, j);
          if (img[c] >= 0 || point_of_cell[c] != 0) continue;  // an alias of the reference, or a vector point of its own
          const int is = xp ? wrap(i, lm) : i, js = yp ? wrap(j, mm) : j;
          if ((is == i && js == j) || is < 0 || is > lm + 1 || js < 0 || js > mm + 1) continue;
          const int src = cell(is, js);
          if (point_of_cell[src] == 0 || img[src] >= 0) continue;  // nothing there, or itself a mirror
          mdst.push_back(c);
          msrc.push_back(src);
        }
      L.torus = true;
    }
  }
  for (size_t k = 0; k < mdst.size(); k++)
    hflags[mdst[k]] = (uint8_t)((hflags[msrc[k]] & ~F_ACT) | ((hflags[msrc[k]] & F_ACT) ? F_GHOST : 0));
  // remote images: the masks of the point they show (every rank holds the static fields of the whole domain), never active
  for (size_t k = 0; k < rdst.size(); k++) hflags[rdst[k]] = (uint8_t)((flags_of_point(rsrc[k]) & ~F_ACT) | F_GHOST);
  if (L.ring && L.torus) {
    // the deep rows across the seam (the fused step recomputes them): G rows below row 1 on the first rank show rows
    // mm-G+1..mm, G rows from mm+1 on the last rank show rows 1..G -- flagged with the masks of the points they show
    std::vector<int> grid((size_t)(lm + 2) * (mm + 2), 0);
    for (int p = 1; p <= ndeg; p++) grid[(size_t)sj[p] * (lm + 2) + si[p]] = p;
    for (int k = 0; k < G && (rank == 0 || rank == nranks - 1); k++) {
      const int js = rank == 0 ? mm - k : 1 + k;        // the row shown ...
      const int jh = rank == 0 ? -k : mm + 1 + k;       // ... and the row showing it
      if (jh + j_off < 0 || jh + j_off >= L.NY || js < 1 || js > mm) continue;
      for (int i = -3; i <= lm + 4; i++) {
        if (i + GX0 < 0 || i + GX0 >= NX) continue;
        const int is = xper ? wrap(i, lm) : i;
        if (is < 0 || is > lm + 1) continue;
        const int q = grid[(size_t)js * (lm + 2) + is];
        if (q) hflags[(jh + j_off) * NX + (i + GX0)] = (uint8_t)((flags_of_point(q) & ~F_ACT) | F_GHOST);
      }
    }
  }
  return 0;
}

// rows (dense Y) of the packed halo exchange of a rank: what it sends to / receives from the rank below (lo) and above (hi)
struct HaloRows {
  int peer_lo = -1, peer_hi = -1;
  int send_lo = 0, send_hi = 0, recv_lo = 0, recv_hi = 0;  // first of G consecutive rows
};
inline HaloRows halo_rows(int rank, int nranks, bool ring, int y_lo, int y_hi) {
  HaloRows h;
  h.peer_lo = rank > 0 ? rank - 1 : (ring ? nranks - 1 : -1);
  h.peer_hi = rank < nranks - 1 ? rank + 1 : (ring ? 0 : -1);
  h.send_lo = y_lo;           h.recv_lo = y_lo - G;
  h.send_hi = y_hi - G + 1;   h.recv_hi = y_hi + 1;
  if (ring && rank == nranks - 1) {  // y_hi is the duplicate row mm+1: the seam sits one row lower
    h.send_hi = y_hi - G;
    h.recv_hi = y_hi;
  }
  return h;
}

}  // namespace beom
