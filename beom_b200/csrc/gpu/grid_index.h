// grid_index.h -- index_grid_points (private_mod.f95:567-764): which points of the padded (lm+2) x (mm+2) grid enter the vector,
// in which order, the neighbour table with its periodic aliases (:614-685) and the five masks (:700-714).  ONE source for the host
// restatement of read_input_data (csrc/host/init.cc) and for beom_gpu_init_grids, which needs the neighbour table of a PERIODIC
// domain to run the same layout analysis as beom_gpu_init (layout.h); on non-periodic domains it numbers the points on the
// device instead (gridinit.cuh).  Plain C++.
#ifndef BEOM_GRID_INDEX_H
#define BEOM_GRID_INDEX_H
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace beom {

// H(i, j): depth of cell (i, j) for i in -1 .. lm+2, j in -1 .. mm+2 (0 outside the basin).  neig is [ndeg+1][8], subc [2][ndeg+1], the
// masks and posc [ndeg+1], all zero-filled by the caller.  Returns the number of vector points; if that is not ndeg nothing is written.
template <class DepthFn>
inline int index_grid_points_core(int lm, int mm, int ndeg, double hdry, bool xper, bool yper, DepthFn H, int32_t *neig, int32_t *subc,
                                  int32_t *posc, double *mk_u, double *mk_v, double *mk_n, double *mkpe, double *mkpi) {
  const size_t nd1 = (size_t)ndeg + 1;
  const int ww = lm + 4;
  std::vector<unsigned char> wet_((size_t)ww * (mm + 4), 0);
  auto wet = [&](int i, int j) -> unsigned char & { return wet_[(size_t)(j + 1) * ww + (size_t)(i + 1)]; };
  for (int j = -1; j <= mm + 2; j++)
    for (int i = -1; i <= lm + 2; i++) wet(i, j) = H(i, j) > hdry;
  // a grid point enters the vector if it carries an eta, u, v or psi point that touches water
  auto carries = [&](int i, int j) { return wet(i, j) || wet(i - 1, j) || wet(i, j - 1) || wet(i - 1, j - 1); };

  std::vector<int32_t> alias_((size_t)ww * (mm + 4), 0);  // "indc": which vector entry a neighbour reference resolves to
  auto alias = [&](int i, int j) -> int32_t & { return alias_[(size_t)(j + 1) * ww + (size_t)(i + 1)]; };
  int count = 0;
  for (int j = 0; j <= mm + 1; j++)
    for (int i = 0; i <= lm + 1; i++)
      if (carries(i, j)) alias(i, j) = ++count;
  if (count != ndeg) return count;  // the caller reports it (pm:604-610); nothing has been written

  if (xper) {  // pm:614-640
    for (int j = 1; j <= mm; j++) {
      const bool both = wet(1, j) && wet(lm, j);
      if (both) {
        alias(0, j) = alias(lm, j);
        alias(lm + 1, j) = alias(1, j);
        mk_u[alias(1, j)] = 1.0;
      }
      if (j > 1 && wet(1, j - 1) && wet(1, j) && wet(lm, j - 1) && wet(lm, j)) mkpe[alias(1, j)] = 1.0;
      if (j == mm && both) {
        alias(0, mm + 1) = alias(lm, mm + 1);
        alias(lm + 1, mm + 1) = alias(1, mm + 1);
      }
    }
  }
  if (yper) {  // pm:642-668
    for (int i = 1; i <= lm; i++) {
      const bool both = wet(i, 1) && wet(i, mm);
      if (both) {
        alias(i, 0) = alias(i, mm);
        alias(i, mm + 1) = alias(i, 1);
        mk_v[alias(i, 1)] = 1.0;
      }
      if (i > 1 && wet(i - 1, 1) && wet(i, 1) && wet(i - 1, mm) && wet(i, mm)) mkpe[alias(i, 1)] = 1.0;
      if (i == lm && both) {
        alias(lm + 1, 0) = alias(lm + 1, mm);
        alias(lm + 1, mm + 1) = alias(lm + 1, 1);
      }
    }
  }
  if (xper && yper) {  // pm:672-685
    if (wet(1, 1) && wet(lm, 1) && wet(1, mm)) {
      alias(0, 0) = alias(lm, mm);
      mkpe[alias(1, 1)] = 1.0;
      alias(0, mm + 1) = alias(lm, 1);
    }
    if (wet(lm, mm) && wet(1, mm) && wet(lm, 1)) {
      alias(lm + 1, 0) = alias(1, mm);
      alias(lm + 1, mm + 1) = alias(1, 1);
    }
  }

  static const int di[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dj[8] = {0, 1, 1, 1, 0, -1, -1, -1};
  int p = 0;
  for (int j = 0; j <= mm + 1; j++)  // pm:692-730
    for (int i = 0; i <= lm + 1; i++) {
      if (!carries(i, j)) continue;
      ++p;
      if (wet(i, j)) mk_n[p] = 1.0;
      if (wet(i - 1, j) && wet(i, j)) mk_u[p] = 1.0;
      if (wet(i, j - 1) && wet(i, j)) mk_v[p] = 1.0;
      if (wet(i - 1, j - 1) && wet(i, j - 1) && wet(i - 1, j) && wet(i, j)) mkpe[p] = 1.0;
      mkpi[p] = 1.0;  // carries() already says one of the four cells is wet (pm:712-714)
      posc[p] = i + 1 + j * (lm + 2);
      subc[p] = i;
      subc[nd1 + p] = j;
      for (int k = 0; k < 8; k++) neig[(size_t)p * 8 + k] = alias(i + di[k], j + dj[k]);
    }
  return count;
}

// index_boundary_points (private_mod.f95:1060-1240): one entry per nudged open-boundary face; nf = nudg.bin's content (0:lm+1, 0:mm+1, 3).
// segm comes back as [18][nseg] (the reference's segm(nseg, 18)); returns nseg (0: "the nudged open boundary segments could not be identified")
template <class DepthFn>
inline int index_boundary_points_core(int lm, int mm, double hdry, bool xper, bool yper, DepthFn H, const float *nf, std::vector<int32_t> &segm) {
  const float tiny4 = std::numeric_limits<float>::min();
  const int ww = lm + 4;
  std::vector<int32_t> own_((size_t)ww * (mm + 4), 0);
  auto own = [&](int i, int j) -> int32_t & { return own_[(size_t)(j + 1) * ww + (size_t)(i + 1)]; };
  {
    int c = 0;
    for (int j = 0; j <= mm + 1; j++)
      for (int i = 0; i <= lm + 1; i++)
        if (H(i, j) > hdry || H(i - 1, j) > hdry || H(i, j - 1) > hdry || H(i - 1, j - 1) > hdry) own(i, j) = ++c;
  }
  auto coef = [&](int i, int j, int comp) -> float {  // comp: 1 eta, 2 u, 3 v
    if (i < 0 || i > lm + 1 || j < 0 || j > mm + 1) return 0.0f;
    return nf[((size_t)(comp - 1) * (mm + 2) + j) * (lm + 2) + i];
  };
  struct Seg { int32_t c[18]; };
  std::vector<Seg> segs;
  auto push = [&](int i, int j, bool zonal, int sign, int di, int dj, int wi, int wj, int ni, int nj, int ci, int cj) {
    Seg s;
    std::memset(&s, 0, sizeof s);
    s.c[0] = own(i, j); s.c[1] = i; s.c[2] = j;
    s.c[zonal ? 3 : 4] = 1;
    s.c[5] = sign;
    s.c[6] = own(di, dj); s.c[7] = di; s.c[8] = dj;      // the dry cell
    s.c[9] = own(wi, wj); s.c[10] = wi; s.c[11] = wj;    // the wet cell
    s.c[12] = own(ni, nj); s.c[13] = ni; s.c[14] = nj;   // interior normal-velocity point
    s.c[15] = own(ci, cj); s.c[16] = ci; s.c[17] = cj;   // interior cell
    segs.push_back(s);
  };
  for (int j = 0; j <= mm + 1; j++)
    for (int i = 0; i <= lm + 1; i++) {
      const bool here = H(i, j) > hdry, west = H(i - 1, j) > hdry, south = H(i, j - 1) > hdry;
      if (here && !west && coef(i, j, 2) > tiny4 && coef(i - 1, j, 2) > tiny4 && !xper)
        push(i, j, true, 1, i - 1, j, i, j, i + 1, j, i + 1, j);
      if (!here && west && coef(i - 1, j, 2) > tiny4 && coef(i, j, 2) > tiny4 && !xper)
        push(i, j, true, -1, i, j, i - 1, j, i - 1, j, i - 2, j);
      if (here && !south && coef(i, j, 3) > tiny4 && coef(i, j - 1, 3) > tiny4 && !yper)
        push(i, j, false, 1, i, j - 1, i, j, i, j + 1, i, j + 1);
      if (!here && south && coef(i, j - 1, 3) > tiny4 && coef(i, j, 3) > tiny4 && !yper)
        push(i, j, false, -1, i, j, i, j - 1, i, j - 1, i, j - 2);
    }
  const int nseg = (int)segs.size();
  segm.assign((size_t)nseg * 18, 0);
  for (int s = 0; s < nseg; s++)
    for (int c = 0; c < 18; c++) segm[(size_t)c * nseg + s] = segs[(size_t)s].c[c];
  return nseg;
}

}  // namespace beom
#endif
