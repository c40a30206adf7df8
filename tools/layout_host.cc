// Host-only test shim for beom_b200/csrc/gpu/layout.h (no CUDA): the analysis of the reference's neighbour table that
// beom_gpu_init runs (slab rows, dense cell map, flags, periodic images, duplicates, ring-closed halo rows), exposed with
// a C ABI so that tests/test_layout.py can check it on a machine without a GPU.
// Build: g++ -O2 -std=c++17 -shared -fPIC tools/layout_host.cc -o liblayout_host.so
#include "../beom_b200/csrc/gpu/layout.h"

#include <cstring>

static beom::Layout L;

extern "C" {
int layout_run(int lm, int mm, int ndeg, int xper, int yper, int rank, int nranks, const int32_t *subc, const int32_t *neig,
               const double *mk_n, const double *mk_u, const double *mk_v, const double *mkpe, const double *mkpi) {
  return beom::analyse_layout(L, lm, mm, ndeg, xper != 0, yper != 0, rank, nranks, subc, neig, mk_n, mk_u, mk_v, mkpe, mkpi);
}
const char *layout_error() { return L.error.c_str(); }
// NX, NY, j0, j1, j_off, p_lo, p_hi, nmir, norph, torus, ring, G, GX0
void layout_dims(int *out) {
  const int v[13] = {L.NX, L.NY, L.j0, L.j1, L.j_off, L.p_lo, L.p_hi, (int)L.mdst.size(), (int)L.orphans.size(), L.torus, L.ring, beom::G, beom::GX0};
  std::memcpy(out, v, sizeof v);
}
void layout_copy(int32_t *cell_of_point, uint8_t *flags, int32_t *point_of_cell, int32_t *mdst, int32_t *msrc, int32_t *orphans) {
  std::memcpy(cell_of_point, L.cell_of_point.data(), sizeof(int) * L.cell_of_point.size());
  std::memcpy(flags, L.flags.data(), L.flags.size());
  std::memcpy(point_of_cell, L.point_of_cell.data(), sizeof(int) * L.point_of_cell.size());
  if (!L.mdst.empty()) {
    std::memcpy(mdst, L.mdst.data(), sizeof(int) * L.mdst.size());
    std::memcpy(msrc, L.msrc.data(), sizeof(int) * L.msrc.size());
  }
  if (!L.orphans.empty()) std::memcpy(orphans, L.orphans.data(), sizeof(int) * L.orphans.size());
}
// peer_lo, peer_hi, send_lo, send_hi, recv_lo, recv_hi (dense rows) of this rank's packed halo exchange
void layout_halo(int *out) {
  const beom::HaloRows h = beom::halo_rows(L.rank, L.nranks, L.ring, beom::G, beom::G + (L.j1 - L.j0));
  const int v[6] = {h.peer_lo, h.peer_hi, h.send_lo, h.send_hi, h.recv_lo, h.recv_hi};
  std::memcpy(out, v, sizeof v);
}
}
