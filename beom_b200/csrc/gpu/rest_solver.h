// rest_solver.h -- get_equilibrium_thickness_h_0 (private_mod.f95:309-502): the Newton iteration per water column for the rest
// thickness with Salmon's outcrop term.  ONE source for the host restatement of read_input_data (csrc/host/init.cc) and for the
// device-side initialisation (gridinit.cuh, one thread per wet cell): strict IEEE on both sides (g++ -ffp-contract=off, nvcc
// -fmad=false), so the two produce the same bits.
#ifndef BEOM_REST_SOLVER_H
#define BEOM_REST_SOLVER_H
#include <cmath>

#include "../../../include/beom_gpu.h"

#ifdef __CUDACC__
#define BEOM_HD __host__ __device__
#else
#define BEOM_HD
#endif

namespace beom {

struct RestSolver {
  int nlay, nsal, itmx;
  double hsal, thre, sor, dmax;
  double rho[BEOM_MAXLAY], topl[BEOM_MAXLAY], cons[BEOM_MAXLAY];

  BEOM_HD static double cube(double x) { return (x * x) * x; }
  BEOM_HD static double quad(double x) { double y = x * x; return y * y; }
  BEOM_HD static double vmax(double a, double b) { return a < b ? b : a; }  // std::max / std::min, usable on the device too
  BEOM_HD static double vmin(double a, double b) { return b < a ? b : a; }

  void prepare() {  // pm:339-355
    double g[BEOM_MAXLAY];
    for (int l = 0; l < nlay; l++) {
      g[l] = dmax * (1.0 - topl[l]);
      if (l < nlay - 1) g[l] = g[l] - dmax * (1.0 - topl[l + 1]);
    }
    double total = 0.0;
    for (int l = 0; l < nlay; l++) total += g[l];
    for (int l = 0; l < nlay; l++) {
      cons[l] = dmax * (-1.0) + total;
      for (int k = 0; k < l; k++) cons[l] = cons[l] - (rho[l] - rho[k]) * g[k] / rho[l];
    }
  }

  // returns false if itmx iterations were not enough (pm:395-401)
  BEOM_HD bool column(double hbot, double *out) const {
    double g[BEOM_MAXLAY], f[BEOM_MAXLAY], A[BEOM_MAXLAY][BEOM_MAXLAY + 1];
    for (int l = nlay - 1; l >= 0; l--) {  // pm:370-380
      double below = 0.0;
      for (int k = l + 1; k < nlay; k++) below += g[k];
      g[l] = vmax(hbot - dmax * topl[l] - below, hsal);
    }
    for (int iter = 1; iter <= itmx; iter++) {
      for (int a = 0; a < nlay; a++) {  // pm:383-393
        double tot = 0.0;
        for (int k = 0; k < nlay; k++) tot += g[k];
        f[a] = (hbot - tot) + 1.0 / (double)(nsal - 1) * hsal * cube(hsal / g[a]) + cons[a];
        f[a] = f[a] * (-1.0);
        for (int k = 0; k < a; k++) f[a] = f[a] - (rho[a] - rho[k]) * g[k] / rho[a];
      }
      if (iter == itmx) return false;
      bool done = true;
      for (int a = 0; a < nlay; a++) done = done && (fabs(f[a]) < thre);
      if (done) {
        for (int a = 0; a < nlay; a++) out[a] = g[a];
        return true;
      }
      for (int a = 0; a < nlay; a++) {  // Jacobian, pm:410-421
        for (int b = 0; b < nlay; b++) {
          A[a][b] = vmin(rho[a], rho[b]) / rho[a];
          if (a == b) A[a][b] = A[a][b] + quad(hsal / g[b]);
        }
        A[a][nlay] = f[a] * (-1.0);
      }
      for (int k = 0; k < nlay; k++) {  // elimination with partial pivoting, pm:426-455
        int piv = -1;
        double best = 0.0;
        for (int r = k; r < nlay; r++)
          if (fabs(A[r][k]) > best) { best = fabs(A[r][k]); piv = r; }
        if (piv >= 0 && piv != k)
          for (int c = 0; c <= nlay; c++) { const double t = A[k][c]; A[k][c] = A[piv][c]; A[piv][c] = t; }
        for (int r = k + 1; r < nlay; r++) {
          for (int c = k; c <= nlay; c++) A[r][c] = A[r][c] - A[k][c] * (A[r][k] / A[k][k]);
          A[r][k] = 0.0;
        }
      }
      for (int r = nlay - 1; r >= 0; r--) {  // pm:459-466
        double acc = 0.0;
        for (int c = r + 1; c < nlay; c++) acc = acc + A[r][c] * A[c][nlay];
        A[r][nlay] = (A[r][nlay] - acc) / A[r][r];
      }
      bool tiny = false;
      for (int a = 0; a < nlay; a++) {  // pm:468-472
        g[a] = (1.0 - sor) * g[a] + sor * (A[a][nlay] + g[a]);
        tiny = tiny || (g[a] <= thre);
      }
      if (tiny)
        for (int a = 0; a < nlay; a++) g[a] = vmax(g[a], thre);
    }
    return false;
  }
};

// note on the elimination loop: the reference updates maug(ilay,l) for l = k..nlay+1 using
// maug(ilay,k) *before* it is zeroed; since l = k is processed first and overwrites maug(ilay,k),
// the factor must be taken per element exactly as written (pm:448-451).  The loop above re-reads
// A[r][k] each time, like the reference.

}  // namespace beom
#endif
