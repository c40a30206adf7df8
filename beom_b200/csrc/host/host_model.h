// host_model.h -- internal state of the host driver (not part of the C ABI).
#ifndef BEOM_HOST_MODEL_H
#define BEOM_HOST_MODEL_H
#include <cstdint>
#include <string>
#include <vector>

#include "beom_host.h"

// A 2-D array over the reference's padded grid indices (lo..hi inclusive in both directions),
// i fastest like the Fortran arrays it mirrors.
template <class T>
struct Grid2 {
  int ilo = 0, ihi = -1, jlo = 0, jhi = -1, w = 0;
  std::vector<T> d;
  void reset(int ilo_, int ihi_, int jlo_, int jhi_, T fill = T()) {
    ilo = ilo_; ihi = ihi_; jlo = jlo_; jhi = jhi_;
    w = ihi - ilo + 1;
    d.assign((size_t)w * (size_t)(jhi - jlo + 1), fill);
  }
  T &operator()(int i, int j) { return d[(size_t)(j - jlo) * w + (size_t)(i - ilo)]; }
  const T &operator()(int i, int j) const { return d[(size_t)(j - jlo) * w + (size_t)(i - ilo)]; }
  bool inside(int i, int j) const { return i >= ilo && i <= ihi && j >= jlo && j <= jhi; }
};

struct beom_host {
  beom_params p;
  std::string idir, odir, desc;
  int lm = 0, mm = 0, nlay = 0, ndeg = 0;
  size_t nd1 = 0;

  Grid2<double> h2d;  // h_2d(-1:lm+2,-1:mm+2), private_mod.f95:108
  std::vector<int32_t> neig, subc, posc, segm;
  int nseg = 0;
  bool flag_nudging = false;
  std::vector<double> mk_u, mk_v, mk_n, mkpe, mkpi, fcor, h_th;
  std::vector<double> nudg, fnud, hdot, taus, tide, bodf;
  std::vector<double> Ow, Os, Osum_, pi_s;
  std::vector<double> h_0, hlay, u, v;
  bool has_nudg = false, has_init = false, has_hdot = false, has_taus = false, has_tide = false, has_bodf = false,
       has_fcor = false;
  double invf = 0, w_ti = 0, tres = 0, ctim = 0;
  int nstp = 0, notp = 1, n_3d = 1;
  // write_outputs' saved state (private_mod.f95:2684-2685)
  bool out_init = false;
  int irec = 0;
  std::vector<float> h_0_r4;  // h_0.bin as written (private_mod.f95:185-194), read back by write_array
  // read_input_data was done on the device (beom_gpu_init_grids): the big vectors above stay empty, the GPU library is
  // initialised and holds the initial state already
  bool device_init = false;
  std::vector<int32_t> grid5;  // grid.bin's five records as the device made them
};

void beom_host_set_error(const std::string &s);

// io.cc
bool beom_host_write_grid_files(beom_host *h);
bool beom_host_save_metadata(beom_host *h);

#endif
