// Host-only test shim for beom_b200/csrc/gpu/orphans.h (no CUDA): the sponge recurrence of nudged periodic
// duplicates, exposed with a C ABI so that tests/test_abi_and_io.py can check it against the CPU oracle on a machine
// without a GPU.  Build: g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC tools/orphans_host.cc -o liborphans_host.so
#include "../beom_b200/csrc/gpu/orphans.h"

extern "C" int orphans_replay(int nlay, int n, const double *nud, const double *fnud, const double *tide, double w_ti,
                              int nstep, const double *log, double *val) {
  Orphans O;
  O.nlay = nlay;
  O.n = (size_t)n;
  O.nud.assign(nud, nud + 3 * (size_t)n);
  O.fnud.assign(fnud, fnud + 3 * (size_t)nlay * n);
  O.has_tide = tide != nullptr;
  if (tide) O.tide.assign(tide, tide + 6 * (size_t)n);
  O.w_ti = w_ti;
  O.val.assign(val, val + 3 * (size_t)nlay * n);
  for (size_t k = 0; k < 3 * (size_t)n && !O.live; k++) O.live = O.nud[k] != 0.0;
  // in two halves: a download in the middle of a run must not change the result
  for (int s = 0; s < nstep / 2; s++) O.record(log[2 * s], log[2 * s + 1]);
  O.replay();
  for (int s = nstep / 2; s < nstep; s++) O.record(log[2 * s], log[2 * s + 1]);
  O.replay();
  for (size_t k = 0; k < O.val.size(); k++) val[k] = O.val[k];
  return O.live ? 1 : 0;
}
