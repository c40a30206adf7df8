// Periodic duplicates ("orphans") that sit inside a sponge.
//
// With xper/yper the reference keeps the duplicate column i = lm+1 / row j = mm+1 as vector points of their own
// (private_mod.f95:614-685): every neighbour table entry that would point at them points at their periodic twin
// instead, their masks are 0, nothing ever reads them -- but they are written to every output record.  On the device
// their dense cell shows the twin (a mirror cell), so their own state lives on the host (Orphans::val) and is patched
// into the downloads.  With all masks 0 the only thing the reference's step does to such a point is the sponge
// relaxation at the end of update_h / update_u / update_v (private_mod.f95:1633-1638, 1449-1454 + 1482, 1534-1539 +
// 1567), e.g. baines_ridge.m, whose east/west sponges cover the duplicate row of its y-periodic channel:
//
//   hlay <- hfor nudg_n + (1 - nudg_n) hlay,   hfor = fnud_n + ramp A_n [layer 1] cos(phi_n - w_ti ctim)
//   u    <- ufor nudg_u + u (1 - nudg_u),      ufor = fnud_u + (Ekman term) + ramp A_u cos(phi_u - w_ti ctim)
//   v    <- vfor nudg_v + v (1 - nudg_v)       likewise
//
// (the masked tendency is an exact zero).  The recurrence needs nothing from the device, so it is not run per step: the
// library logs (ctim, ramp) of every step and replays the log on the host when a download asks for the state.  Same
// expressions, same order, same libm as a strict-IEEE CPU build of the reference: bit-identical.  The Ekman term reads
// the wind-stress share of the duplicate's own water column; a nudged duplicate under wind stress with invf != 0 is
// refused at init (beom_gpu_init) rather than approximated.
//
// Plain C++ (no CUDA) so that tests can compile it on a CPU-only machine (tools/orphans_host.cc).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

struct Orphans {
  int nlay = 0;
  size_t n = 0;              // number of duplicates on this rank
  bool live = false;         // at least one of them is nudged
  bool has_tide = false;
  double w_ti = 0.0;
  std::vector<double> nud;   // [3][n]        eta, u, v coefficients
  std::vector<double> fnud;  // [3][nlay][n]  relaxation targets
  std::vector<double> tide;  // [3][n][2]     amplitude, phase
  std::vector<double> val;   // [3][nlay][n]  hlay, u, v of the duplicates
  std::vector<double> log;   // (ctim, ramp) of the steps not yet applied to val

  void record(double ctim, double ramp) {
    if (!live) return;
    log.push_back(ctim);
    log.push_back(ramp);
  }
  void forget() { log.clear(); }

  // apply the logged steps to val
  void replay() {
    const size_t nstep = log.size() / 2;
    if (!live || nstep == 0) { log.clear(); return; }
    const size_t nl = (size_t)nlay;
    std::vector<double> tterm(has_tide ? nstep : 0);
    for (int f = 0; f < 3; f++)
      for (size_t k = 0; k < n; k++) {
        const double c = nud[(size_t)f * n + k];
        if (c == 0.0) continue;  // x*0 + x*(1-0) = x
        if (has_tide) {
          const double amp = tide[((size_t)f * n + k) * 2], pha = tide[((size_t)f * n + k) * 2 + 1];
          for (size_t s = 0; s < nstep; s++) tterm[s] = std::cos(pha - w_ti * log[2 * s]);
          for (size_t l = 0; l < nl; l++) {
            double x = val[((size_t)f * nl + l) * n + k];
            const double target = fnud[((size_t)f * nl + l) * n + k];
            for (size_t s = 0; s < nstep; s++) {
              const double ramp = log[2 * s + 1];
              double xfor;
              if (f == 0) {
                const double vecl = (l == 0) ? 1.0 : 0.0;
                xfor = target + ramp * amp * vecl * tterm[s];
                x = xfor * c + (1.0 - c) * x;
              } else {
                xfor = target + 0.0 + ramp * amp * tterm[s];
                x = xfor * c + x * (1.0 - c);
              }
            }
            val[((size_t)f * nl + l) * n + k] = x;
          }
        } else {
          for (size_t l = 0; l < nl; l++) {
            double x = val[((size_t)f * nl + l) * n + k];
            // target + 0 (Ekman term) + ramp*0*cos(): the reference's sum with no wind share and no tide
            const double xfor = fnud[((size_t)f * nl + l) * n + k] + 0.0 + 0.0;
            if (f == 0) for (size_t s = 0; s < nstep; s++) x = xfor * c + (1.0 - c) * x;
            else        for (size_t s = 0; s < nstep; s++) x = xfor * c + x * (1.0 - c);
            val[((size_t)f * nl + l) * n + k] = x;
          }
        }
      }
    log.clear();
  }
};
