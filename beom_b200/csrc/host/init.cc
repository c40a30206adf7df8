// init.cc -- read_input_data (private_mod.f95:105-250): builds every static field the GPU library is
// given, in the reference's vector layout.  This part of the reference stays on the host (SURVEY
// section 2.1 row 4); it is restated in C++ because no Fortran compiler exists here.
//
// Layout reminders (private_mod.f95:27-93): point index 0 is the discarded cell; x(0:ndeg,nlay) is
// stored [nlay][ndeg+1]; neig(8,0:ndeg) is [ndeg+1][8]; fnud(0:ndeg,nlay,3) is [3][nlay][ndeg+1].
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>

#include "host_model.h"
#include "../gpu/grid_index.h"
#include "../gpu/rest_solver.h"

namespace {
// BEOM_HOST_TIMING=1: wall-clock of the stages of read_input_data on stderr (diagnostics only)
struct StageTimer {
  const char *what;
  double t0;
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  explicit StageTimer(const char *w) : what(w), t0(now()) {}
  ~StageTimer() {
    static const bool on = getenv("BEOM_HOST_TIMING") != nullptr;
    if (on) std::fprintf(stderr, "[host init] %-28s %.2f s\n", what, now() - t0);
  }
};

struct Fail {
  std::string msg;
};

// raw little-endian float32 file of exactly n values; absent file -> empty vector (private_mod.f95:775-776)
std::vector<float> slurp(const std::string &dir, const char *keyw, size_t n, bool *present) {
  *present = false;
  std::vector<float> buf;
  if (dir.empty()) return buf;
  std::string path = dir + keyw + ".bin";
  std::FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return buf;
  buf.resize(n);
  size_t got = std::fread(buf.data(), sizeof(float), n, f);
  std::fclose(f);
  if (got != n) throw Fail{std::string(" could not open/read file ") + keyw + ".bin from directory " + dir};
  *present = true;
  return buf;
}

inline long nint(double x) { return (long)(x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5)); }

// ------------------------------------------------------------------------------------------------
// index_grid_points (private_mod.f95:567-764)
// ------------------------------------------------------------------------------------------------
void index_grid_points(beom_host *h) {
  const Grid2<double> &H = h->h2d;
  const int count = beom::index_grid_points_core(h->lm, h->mm, h->ndeg, h->p.hdry, h->p.xper > 0.5, h->p.yper > 0.5,
                                                 [&](int i, int j) { return H(i, j); }, h->neig.data(), h->subc.data(), h->posc.data(),
                                                 h->mk_u.data(), h->mk_v.data(), h->mk_n.data(), h->mkpe.data(), h->mkpi.data());
  if (count != h->ndeg) {
    char b[160];
    std::snprintf(b, sizeof b, " wrong input parameter! Please set ndeg = %d inside file shared_mod.f95.", count);
    throw Fail{b};
  }
  for (size_t q = 0; q < h->nd1; q++) h->h_th[q] = H(h->subc[q], h->subc[h->nd1 + q]);  // pm:753-757
}

// ------------------------------------------------------------------------------------------------
// get_equilibrium_thickness_h_0 (private_mod.f95:309-502): the Newton iteration per water column lives in
// ../gpu/rest_solver.h (shared with the device-side initialisation).
// ------------------------------------------------------------------------------------------------
using beom::RestSolver;

void rest_thickness(beom_host *h) {
  const int nlay = h->nlay;
  RestSolver s;
  s.nlay = nlay; s.nsal = h->p.nsal; s.itmx = h->p.itmx;
  s.hsal = h->p.hsal; s.thre = h->p.tole; s.sor = h->p.sor;
  s.dmax = *std::max_element(h->h2d.d.begin(), h->h2d.d.end());  // pm:334
  for (int l = 0; l < nlay; l++) { s.rho[l] = h->p.rhon[l]; s.topl[l] = h->p.topl[l]; }
  s.prepare();
  int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 512)
#endif
  for (int p = 1; p <= h->ndeg; p++) {
    if (h->mk_n[p] < 0.5 || bad) continue;
    double col[BEOM_MAXLAY];
    const double hbot = h->h2d(h->subc[p], h->subc[h->nd1 + p]);
    if (!s.column(hbot, col)) { bad = p; continue; }
    for (int l = 0; l < nlay; l++) h->h_0[(size_t)l * h->nd1 + p] = col[l];
  }
  if (bad) {
    char b[200];
    std::snprintf(b, sizeof b, " calculation of h_layers did not converge, tolerance (meters) was %g, at point %d", s.thre, bad);
    throw Fail{b};
  }

  if (h->p.rgld > 0.5) {  // pm:505-563: start pressure and the Poisson operators
    const int lm = h->lm, mm = h->mm;
    const double dl = h->p.dl;
    const size_t n = h->nd1;
    std::vector<double> osum(n, 0.0);
    for (size_t q = 0; q < n; q++) {
      double tot = 0.0;
      for (int l = 0; l < nlay; l++) tot += h->h_0[(size_t)l * n + q];
      h->pi_s[q] = (tot - h->h_th[q]) * h->p.grav;
      h->Ow[q] = h->Os[q] = h->Osum_[q] = 0.0;
    }
    auto W = [&](int q) { return h->neig[(size_t)q * 8 + 4]; };
    auto S = [&](int q) { return h->neig[(size_t)q * 8 + 6]; };
    auto E = [&](int q) { return h->neig[(size_t)q * 8 + 0]; };
    auto N = [&](int q) { return h->neig[(size_t)q * 8 + 2]; };
    for (int q = 1; q <= h->ndeg; q++) {
      const int i = h->subc[q], j = h->subc[n + q];
      const bool xin = 1 < i && i < lm + 1, yin = 1 < j && j < mm + 1;
      if (xin && yin) {
        h->Ow[q] = 0.5 * (h->h_th[q] + h->h_th[W(q)]) / (dl * dl);
        h->Os[q] = 0.5 * (h->h_th[q] + h->h_th[S(q)]) / (dl * dl);
      } else if (i == 1 && yin) {
        h->Os[q] = 0.5 * (h->h_th[q] + h->h_th[S(q)]) / (dl * dl);
      } else if (xin && j == 1) {
        h->Ow[q] = 0.5 * (h->h_th[q] + h->h_th[W(q)]) / (dl * dl);
      }
    }
    for (int q = 1; q <= h->ndeg; q++) {
      const int i = h->subc[q], j = h->subc[n + q];
      if (i < lm && j < mm) osum[q] = h->Ow[q] + h->Ow[E(q)] + h->Os[q] + h->Os[N(q)];
      else if (i == lm && j < mm) osum[q] = h->Ow[q] + h->Os[q] + h->Os[N(q)];
      else if (j == mm && i < lm) osum[q] = h->Ow[q] + h->Os[q] + h->Ow[E(q)];
      else osum[q] = h->Ow[q] + h->Os[q];
      if (i > 0 && i < lm + 1 && j > 0 && j < mm + 1) h->Osum_[q] = 1 / osum[q];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// index_boundary_points (private_mod.f95:1060-1240): one entry per nudged open-boundary face
// ------------------------------------------------------------------------------------------------
void index_boundary_points(beom_host *h, const std::vector<float> &nf) {
  const Grid2<double> &H = h->h2d;
  h->nseg = beom::index_boundary_points_core(h->lm, h->mm, h->p.hdry, h->p.xper > 0.5, h->p.yper > 0.5, [&](int i, int j) { return H(i, j); },
                                             nf.data(), h->segm);
  if (h->nseg == 0) throw Fail{" the nudged open boundary segments could not be identified."};
}

// ------------------------------------------------------------------------------------------------
// read_input_file for everything after h_0 (private_mod.f95:840-964)
// ------------------------------------------------------------------------------------------------
void read_forcing_files(beom_host *h) {
  const int lm = h->lm, mm = h->mm, nlay = h->nlay, ndeg = h->ndeg;
  const size_t n = h->nd1, plane = (size_t)(lm + 2) * (mm + 2);
  const int32_t *si = h->subc.data(), *sj = h->subc.data() + n;
  bool present;

  std::vector<float> f = slurp(h->idir, "nudg", plane * 3, &present);  // pm:843-881
  h->has_nudg = present;
  if (present) {
    auto at = [&](int i, int j, int c) { return f[((size_t)c * (mm + 2) + j) * (lm + 2) + i]; };
    h->nudg.assign(n * 3, 0.0);
    bool live = false;
#pragma omp parallel for schedule(static) reduction(|| : live)
    for (int p = 1; p <= ndeg; p++) {
      const int i = si[p], j = sj[p];
      const double ne = (double)at(i, j, 0);
      double nu = 0.0, nv = 0.0;
      if (i >= 1 && at(i - 1, j, 1) > 1.e-9f && at(i, j, 1) > 1.e-9f) nu = (double)at(i, j, 1) * 0.5 + (double)at(i - 1, j, 1) * 0.5;
      if (j >= 1 && at(i, j - 1, 2) > 1.e-9f && at(i, j, 2) > 1.e-9f) nv = (double)at(i, j, 2) * 0.5 + (double)at(i, j - 1, 2) * 0.5;
      h->nudg[p] = ne;
      h->nudg[n + p] = nu;
      h->nudg[2 * n + p] = nv;
      live = live || ne > 1.e-9 || nu > 1.e-9 || nv > 1.e-9;
    }
    if (live) {
      h->flag_nudging = true;
      index_boundary_points(h, f);
    }
    h->fnud.assign(n * nlay * 3, 0.0);
#pragma omp parallel for collapse(2) schedule(static)
    for (int l = 0; l < nlay; l++)
      for (int p = 1; p <= ndeg; p++) h->fnud[(size_t)l * n + p] = h->hlay[(size_t)l * n + p];
  }

  StageTimer *t_init = new StageTimer("  init.bin: read");
  f = slurp(h->idir, "init", plane * nlay * 3, &present);  // pm:882-910
  delete t_init;
  h->has_init = present;
  if (present) {
    StageTimer t_g("  init.bin: gather");
    if (h->fnud.empty()) h->fnud.assign(n * nlay * 3, 0.0);
    auto at = [&](int i, int j, int l, int c) { return (double)f[(((size_t)c * nlay + l) * (mm + 2) + j) * (lm + 2) + i]; };
    double *fn = h->fnud.data(), *fu = fn + n * nlay, *fv = fu + n * nlay;
#pragma omp parallel for collapse(2) schedule(static)
    for (int l = 0; l < nlay; l++)
      for (int p = 1; p <= ndeg; p++) {
        const int i = si[p], j = sj[p];
        const size_t k = (size_t)l * n + p;
        double t = h->hlay[k] + at(i, j, l, 0);
        if (l < nlay - 1) t = t - at(i, j, l + 1, 0);
        fn[k] = t * h->mk_n[p];
        fu[k] = at(i, j, l, 1);
        fv[k] = at(i, j, l, 2);
        if (h->p.rsta < 0.5) {
          h->hlay[k] = fn[k] * h->mk_n[p];
          h->u[k] = fu[k];
          h->v[k] = fv[k];
        }
      }
  }

  f = slurp(h->idir, "bodf", (size_t)nlay * 2, &present);  // pm:840-842
  h->has_bodf = present;
  h->bodf.assign((size_t)nlay * 2, 0.0);
  if (present)
    for (size_t k = 0; k < (size_t)nlay * 2; k++) h->bodf[k] = (double)f[k];

  f = slurp(h->idir, "hdot", plane * nlay, &present);  // pm:911-919
  h->has_hdot = present;
  if (present) {
    h->hdot.assign(n * nlay, 0.0);
#pragma omp parallel for collapse(2) schedule(static)
    for (int l = 0; l < nlay; l++)
      for (int p = 1; p <= ndeg; p++) h->hdot[(size_t)l * n + p] = (double)f[((size_t)l * (mm + 2) + sj[p]) * (lm + 2) + si[p]];
  }

  // taus defaults to the uniform tauw (pm:302-303); a file replaces it and zeroes the sentinel (pm:921)
  h->taus.assign(n * 2, 0.0);
  for (size_t q = 0; q < n; q++) { h->taus[q] = h->p.tauw[0]; h->taus[n + q] = h->p.tauw[1]; }
  f = slurp(h->idir, "taus", plane * 2, &present);  // pm:920-931
  h->has_taus = present;
  if (present) {
    std::fill(h->taus.begin(), h->taus.end(), 0.0);
#pragma omp parallel for schedule(static)
    for (int p = 1; p <= ndeg; p++) {
      const size_t k = (size_t)sj[p] * (lm + 2) + si[p];
      h->taus[p] = (double)f[k];
      h->taus[n + p] = (double)f[plane + k];
    }
  }

  f = slurp(h->idir, "tide", plane * 2 * 3, &present);  // pm:951-964; (2,1,0:lm+1,0:mm+1,3)
  h->has_tide = present;
  if (present) {
    if (h->fnud.empty()) h->fnud.assign(n * nlay * 3, 0.0);
    h->tide.assign(n * 6, 0.0);
    h->w_ti = (double)f[0];
    for (int p = 1; p <= ndeg; p++)
      for (int c = 0; c < 3; c++)
        for (int a = 0; a < 2; a++)
          h->tide[((size_t)c * n + p) * 2 + a] = (double)f[(((size_t)c * (mm + 2) + sj[p]) * (lm + 2) + si[p]) * 2 + a];
  }

  f = slurp(h->idir, "fcor", plane, &present);  // pm:932-950: psi-point average taken in float32
  h->has_fcor = present;
  if (present) {
    float acc = 0.0f;
    for (size_t k = 0; k < plane; k++) acc = acc + f[k];
    h->fcor[0] = (double)(acc / (float)plane);
    auto at = [&](int i, int j) { return f[(size_t)j * (lm + 2) + i]; };
#pragma omp parallel for schedule(static)
    for (int p = 1; p <= ndeg; p++) {
      const int i = si[p], j = sj[p];
      if (i > 0 && j > 0) {
        float t = at(i, j) * 0.25f;
        t = t + at(i - 1, j) * 0.25f;
        t = t + at(i, j - 1) * 0.25f;
        t = t + at(i - 1, j - 1) * 0.25f;
        h->fcor[p] = (double)t;
      } else
        h->fcor[p] = (double)at(i, j);
    }
  }
}

void read_input_data(beom_host *h) {
  const int lm = h->lm, mm = h->mm, nlay = h->nlay;
  const beom_params &P = h->p;
  const size_t n = h->nd1;

  // initialize_variables (pm:252-307): only what the host keeps
  StageTimer t_all("read_input_data");
  {
  StageTimer t("initialize_variables");
  // first touch of ~250 bytes per point: one vector per thread at a time (4 GB at 4096 x 4096 x 4)
  std::vector<double> *dv[] = {&h->mk_u, &h->mk_v, &h->mk_n, &h->mkpe, &h->mkpi, &h->h_th, &h->Ow, &h->Os, &h->Osum_, &h->pi_s,
                               &h->fcor, &h->h_0, &h->hlay, &h->u, &h->v};
  const int ndv = (int)(sizeof dv / sizeof dv[0]);
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < ndv + 3; k++) {
    if (k == ndv) h->neig.assign(n * 8, 0);
    else if (k == ndv + 1) h->subc.assign(n * 2, 0);
    else if (k == ndv + 2) h->posc.assign(n, 0);
    else if (k == 10) dv[k]->assign(n, P.f0);
    else dv[k]->assign(k > 10 ? n * nlay : n, 0.0);
  }
  }

  // default flat bottom of depth cext**2/grav (pm:119-121), replaced by h_bo.bin if present
  h->h2d.reset(-1, lm + 2, -1, mm + 2, 0.0);
  for (int j = 1; j <= mm; j++)
    for (int i = 1; i <= lm; i++) h->h2d(i, j) = (P.cext * P.cext) / P.grav;

  bool present;
  std::vector<float> f = slurp(h->idir, "h_bo", (size_t)(lm + 2) * (mm + 2), &present);  // pm:827-839
  if (present) {
    std::vector<float> ft;
    if (P.topt > 0.5) {
      bool pt;
      ft = slurp(h->idir, "h_to", (size_t)(lm + 2) * (mm + 2), &pt);
      if (!pt) throw Fail{" could not open/read file h_to.bin from directory " + h->idir};
    }
    std::fill(h->h2d.d.begin(), h->h2d.d.end(), 0.0);
    for (int j = 0; j <= mm + 1; j++)
      for (int i = 0; i <= lm + 1; i++) {
        const size_t k = (size_t)j * (lm + 2) + i;
        const double d = ft.empty() ? (double)f[k] : (double)(f[k] - ft[k]);
        h->h2d(i, j) = (d < P.hdry || i == 0 || j == 0 || i == lm + 1 || j == mm + 1) ? 0.0 : d;
      }
  }

  {
    StageTimer t("index_grid_points");
    index_grid_points(h);
  }

  double dmin = std::numeric_limits<double>::infinity(), dmax = -dmin;  // pm:134-135
  for (double d : h->h2d.d) {
    if (d > P.hdry) dmin = std::min(dmin, d);
    dmax = std::max(dmax, d);
  }
  if (P.ocrp < 0.5 && nlay > 1) {  // pm:137-152
    if (P.topl[nlay - 1] * dmax + 10.0 * P.hmin >= dmin)
      throw Fail{" Please modify topl so that bathymetry is contained within lower layer."};
  } else if (P.ocrp < 0.5 && nlay == 1) {
    if (dmin <= 10.0 * P.hmin) throw Fail{" Please adjust h_bo or hmin so that min(h_bo) > 10. * hmin."};
  }

  if (P.ocrp < 0.5) {  // pm:154-175: layers stacked from the bottom, no outcrop
#pragma omp parallel for schedule(static)
    for (int p = 1; p <= h->ndeg; p++) {
      if (!(h->mk_n[p] > 0.5)) continue;
      const double depth = h->h2d(h->subc[p], h->subc[n + p]);
      for (int l = nlay - 1; l >= 0; l--) {
        const double above = l > 0 ? dmax * P.topl[l] : 0.0;
        double below = 0.0;
        for (int k = l + 1; k < nlay; k++) below += h->h_0[(size_t)k * n + p];
        h->h_0[(size_t)l * n + p] = depth - above - below;
      }
    }
  } else {
    rest_thickness(h);  // pm:177-183
  }

  // h_0.bin holds float32 (pm:185-194); write_array reads it back (pm:2839-2846)
  h->h_0_r4.resize((size_t)h->ndeg * nlay);
#pragma omp parallel for collapse(2) schedule(static)
  for (int l = 0; l < nlay; l++)
    for (int p = 1; p <= h->ndeg; p++) h->h_0_r4[(size_t)l * h->ndeg + (p - 1)] = (float)h->h_0[(size_t)l * n + p];

#pragma omp parallel for collapse(2) schedule(static)
  for (int l = 0; l < nlay; l++)  // pm:198-200
    for (size_t q = 0; q < n; q++) h->hlay[(size_t)l * n + q] = h->h_0[(size_t)l * n + q] * h->mk_n[q];

  {
    StageTimer t("read_forcing_files");
    read_forcing_files(h);  // pm:204-216
  }

  double acc = 0.0;  // pm:223-229
  for (size_t q = 0; q < n; q++) acc += h->fcor[q];
  h->invf = acc / (double)n;
  h->invf = std::fabs(h->invf) > 1.25e-5 ? 1.0 / h->invf : 0.0;

  const double dtd8 = P.dt / 24.0 / 3600.0;  // pm:1853-1856
  h->nstp = (int)nint(P.dt_s / dtd8);
  h->notp = std::max((int)nint(P.dt_o / dtd8), 1);
  h->n_3d = std::max((int)nint(P.dt3d / dtd8), 1);
}

void check_consistency_options(const beom_params &P) {  // pm:969-1058 (numerical range checks)
  std::string m;
  if (P.lm < 1 || P.mm < 1) m += " grid dimensions (lm,mm) should be >= 1.";
  if (P.dl < 1.e1) m += " mesh size (dl) should be >= 10 meters.";
  if (std::fabs(P.f0) > 2.e-4) m += " Coriolis parameter (f0, in s**(-1)) should be within: -2x10**(-4) < f0 < 2x10**(-4).";
  if (P.dvis < 0.0 || P.dvis > 5.0) m += " Viscosity coefficient should be within: 0 <= dvis < 5.0.";
  if ((P.bdrg < 0.0 || P.bdrg > 15.e-3) && P.qdrg > 0.5) m += " quadratic bottom drag coefficient bdrg should be within: 0 <= bdrg < 5x10**(-3).";
  else if (P.bdrg < 0.0 || P.bdrg > 5.e-2) m += " linear bottom drag coefficient bdrg should be within: 0 <= bdrg < 5x10**(-3) x u_max.";
  if ((P.tdrg < 0.0 || P.tdrg > 15.e-3) && P.qdrg > 0.5) m += " quadratic top drag coefficient tdrg should be within: 0 <= tdrg < 5x10**(-3).";
  else if (P.tdrg < 0.0 || P.tdrg > 5.e-2) m += " linear top drag coefficient tdrg should be within: 0 <= tdrg < 5x10**(-3) x u_max.";
  if (!m.empty()) throw Fail{m};
}

}  // namespace

extern "C" {

beom_host *beom_host_create(const beom_params *par, const char *idir, const char *odir, const char *desc) {
  beom_host *h = new beom_host();
  try {
    h->p = *par;
    h->lm = par->lm; h->mm = par->mm; h->nlay = par->nlay; h->ndeg = par->ndeg;
    h->nd1 = (size_t)par->ndeg + 1;
    auto slash = [](const char *s) {  // pm:1002-1009
      std::string r = s ? s : "";
      if (!r.empty() && r.back() != '/') r.push_back('/');
      return r;
    };
    h->idir = slash(idir);
    h->odir = slash(odir);
    h->desc = desc ? desc : "";
    if (par->nlay < 1 || par->nlay > BEOM_MAXLAY) throw Fail{" nlay must be within 1..16."};
    check_consistency_options(h->p);
    read_input_data(h);
    if (!h->odir.empty()) {
      if (!beom_host_write_grid_files(h)) throw Fail{" could not write grid.bin / h_0.bin into " + h->odir};
      if (h->p.rsta < 0.5 && !beom_host_save_metadata(h)) throw Fail{" could not write param_basin.txt into " + h->odir};
    }
    return h;
  } catch (const Fail &e) {
    beom_host_set_error("In main, in subroutine read_input_data," + e.msg);
  } catch (const std::exception &e) {
    beom_host_set_error(std::string("In main, in subroutine read_input_data, ") + e.what());
  }
  delete h;
  return nullptr;
}

namespace {
// a read-only mapping of <idir><keyw>.bin, or {nullptr, 0} if the file does not exist (private_mod.f95:775-776)
struct Mapped {
  const float *p = nullptr;
  size_t bytes = 0;
  void *base = nullptr;
};
}  // namespace
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
namespace {
Mapped map_file(const std::string &dir, const char *keyw, size_t n) {
  Mapped m;
  const std::string path = dir + keyw + ".bin";
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) return m;
  struct stat st;
  if (::fstat(fd, &st) != 0 || (size_t)st.st_size < n * sizeof(float)) {
    ::close(fd);
    throw Fail{" could not open/read file " + std::string(keyw) + ".bin from directory " + dir};
  }
  void *b = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (b == MAP_FAILED) throw Fail{" could not map file " + std::string(keyw) + ".bin from directory " + dir};
  m.base = b; m.bytes = (size_t)st.st_size; m.p = static_cast<const float *>(b);
  return m;
}
}  // namespace

beom_host *beom_host_create_on_device(const beom_params *par, const char *idir, const char *odir, const char *desc, const beom_gpu_options *opt) {
  // periodic domains: grid.bin lists the displaced duplicate points too, which have no cell on the device -- those (small) cases keep
  // the host's read_input_data
  if (par->xper > 0.5 || par->yper > 0.5) return beom_host_create(par, idir, odir, desc);
  beom_host *h = new beom_host();
  std::vector<Mapped> maps;
  auto unmap = [&]() { for (auto &m : maps) if (m.base) ::munmap(m.base, m.bytes); maps.clear(); };
  try {
    h->p = *par;
    h->lm = par->lm; h->mm = par->mm; h->nlay = par->nlay; h->ndeg = par->ndeg;
    h->nd1 = (size_t)par->ndeg + 1;
    auto slash = [](const char *s) {
      std::string r = s ? s : "";
      if (!r.empty() && r.back() != '/') r.push_back('/');
      return r;
    };
    h->idir = slash(idir);
    h->odir = slash(odir);
    h->desc = desc ? desc : "";
    if (par->nlay < 1 || par->nlay > BEOM_MAXLAY) throw Fail{" nlay must be within 1..16."};
    check_consistency_options(h->p);
    const size_t gp = (size_t)(h->lm + 2) * (h->mm + 2), nl = (size_t)h->nlay;
    beom_grids gr;
    std::memset(&gr, 0, sizeof gr);
    auto get = [&](const char *k, size_t n) { maps.push_back(map_file(h->idir, k, n)); return maps.back().p; };
    gr.h_bo = get("h_bo", gp); gr.init = get("init", gp * nl * 3); gr.nudg = get("nudg", gp * 3); gr.taus = get("taus", gp * 2);
    gr.fcor = get("fcor", gp); gr.hdot = get("hdot", gp * nl); gr.bodf = get("bodf", nl * 2); gr.tide = get("tide", gp * 6);
    {
      struct stat st;
      gr.has_h_to = ::stat((h->idir + "h_to.bin").c_str(), &st) == 0;
    }
    const int rc = beom_gpu_init_grids(&h->p, &gr, opt);
    unmap();
    if (rc == BEOM_GRIDS_UNSUPPORTED) {  // the host path serves it
      delete h;
      return beom_host_create(par, idir, odir, desc);
    }
    if (rc) {
      char b[1024];
      beom_gpu_last_error(b, sizeof b);
      throw Fail{std::string(" ") + b};
    }
    h->device_init = true;
    const double dtd8 = h->p.dt / 24.0 / 3600.0;  // pm:1853-1856
    h->nstp = (int)nint(h->p.dt_s / dtd8);
    h->notp = std::max((int)nint(h->p.dt_o / dtd8), 1);
    h->n_3d = std::max((int)nint(h->p.dt3d / dtd8), 1);
    if (!h->odir.empty()) {
      h->grid5.resize((size_t)5 * h->ndeg);
      h->h_0_r4.resize((size_t)h->ndeg * nl);
      if (beom_gpu_download_grid_files(h->grid5.data(), h->h_0_r4.data())) {
        char b[1024];
        beom_gpu_last_error(b, sizeof b);
        throw Fail{std::string(" ") + b};
      }
      if (!beom_host_write_grid_files(h)) throw Fail{" could not write grid.bin / h_0.bin into " + h->odir};
      if (!beom_host_save_metadata(h)) throw Fail{" could not write param_basin.txt into " + h->odir};
    }
    return h;
  } catch (const Fail &e) {
    beom_host_set_error("In main, in subroutine read_input_data," + e.msg);
  } catch (const std::exception &e) {
    beom_host_set_error(std::string("In main, in subroutine read_input_data, ") + e.what());
  }
  unmap();
  delete h;
  return nullptr;
}

void beom_host_destroy(beom_host *h) { delete h; }

double *beom_host_array(beom_host *h, const char *name) {
  std::string s = name;
#define F(x) if (s == #x) return h->x.empty() ? nullptr : h->x.data();
  F(mk_u) F(mk_v) F(mk_n) F(mkpe) F(mkpi) F(fcor) F(h_th) F(nudg) F(fnud) F(hdot) F(taus) F(tide) F(bodf)
  F(Ow) F(Os) F(Osum_) F(pi_s) F(h_0) F(hlay) F(u) F(v)
#undef F
  if (s == "h_2d") return h->h2d.d.data();
  return nullptr;
}
int32_t *beom_host_iarray(beom_host *h, const char *name) {
  std::string s = name;
  if (s == "neig") return h->neig.data();
  if (s == "subc") return h->subc.data();
  if (s == "posc") return h->posc.data();
  if (s == "segm") return h->segm.empty() ? nullptr : h->segm.data();
  return nullptr;
}
double beom_host_scalar(const beom_host *h, const char *name) {
  std::string s = name;
  if (s == "invf") return h->invf;
  if (s == "w_ti") return h->w_ti;
  if (s == "tres") return h->tres;
  if (s == "ctim") return h->ctim;
  if (s == "nseg") return (double)h->nseg;
  if (s == "flag_nudging") return h->flag_nudging ? 1.0 : 0.0;
  if (s == "irec") return (double)h->irec;
  return std::nan("");
}
const beom_params *beom_host_params(const beom_host *h) { return &h->p; }
void beom_host_counts(const beom_host *h, int *nstp, int *notp, int *n_3d) {
  if (nstp) *nstp = h->nstp;
  if (notp) *notp = h->notp;
  if (n_3d) *n_3d = h->n_3d;
}

void beom_host_fields(beom_host *h, beom_fields *f) {
  std::memset(f, 0, sizeof *f);
  auto opt = [](std::vector<double> &v) -> const double * { return v.empty() ? nullptr : v.data(); };
  f->neig = h->neig.data(); f->subc = h->subc.data();
  f->mk_u = h->mk_u.data(); f->mk_v = h->mk_v.data(); f->mk_n = h->mk_n.data();
  f->mkpe = h->mkpe.data(); f->mkpi = h->mkpi.data();
  f->fcor = h->fcor.data(); f->h_th = h->h_th.data();
  f->nudg = opt(h->nudg); f->fnud = opt(h->fnud); f->hdot = opt(h->hdot);
  f->taus = opt(h->taus); f->tide = opt(h->tide); f->bodf = opt(h->bodf);
  f->segm = h->segm.empty() ? nullptr : h->segm.data();
  f->nseg = h->nseg;
  if (h->p.rgld > 0.5) { f->Ow = h->Ow.data(); f->Os = h->Os.data(); f->Osum_ = h->Osum_.data(); f->pi_s = h->pi_s.data(); }
  f->flag_nudging = h->flag_nudging ? 1 : 0;
  f->invf = h->invf;
  f->w_ti = h->w_ti;
}

}  // extern "C"
