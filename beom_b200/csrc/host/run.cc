// run.cc -- run() = read_input_data + integrate_time (private_mod.f95:99-103, 1840-1919) with the
// per-step routines executed by the GPU library (include/beom_gpu.h).  The time loop is textually
// the reference's: the same tstp/ctim/ramp/gene/upst bookkeeping, the same output cadence.
#include <cstdio>
#include <cstring>

#include "host_model.h"

extern "C" int beom_host_write_records(beom_host *h, double ctim, const beom_records *r);

namespace {
int gpu_fail(const char *where, int rc) {
  char b[512];
  beom_gpu_last_error(b, sizeof b);
  beom_host_set_error(std::string("In main, in subroutine integrate_time, ") + where + ": " + b);
  return rc;
}

// Output is pipelined: at an output step the records are made on the device behind the steps issued so far
// (beom_gpu_records_begin) and copied out on the library's copy stream while the time loop goes on; they are written --
// and the thickness report / halt of write_outputs evaluated -- when the next output step (or the end of the run) comes.
// The files are the reference's, record for record; only the moment the halt is noticed moves by one output interval.
struct Pending {
  bool any = false;
  double ctim = 0.0;
};
int finish_outputs(beom_host *h, Pending &pend) {
  if (!pend.any) return 0;
  pend.any = false;
  beom_records r;
  int rc = beom_gpu_records_wait(&r);
  if (rc) return gpu_fail("beom_gpu_records_wait", rc);
  if ((rc = beom_host_write_records(h, pend.ctim, &r))) return rc;
  std::printf(" ctim = %.15g days; dt_s = %.15g days; record = %d\n", pend.ctim, h->p.dt_s, h->irec - 1);
  return 0;
}
int begin_outputs(beom_host *h, double ctim, Pending &pend) {
  int rc = finish_outputs(h, pend);
  if (rc) return rc;
  if ((rc = beom_gpu_records_begin(h->p.diag > 0.5 ? 1 : 0))) return gpu_fail("beom_gpu_records_begin", rc);
  pend.any = true;
  pend.ctim = ctim;
  return 0;
}
}  // namespace

extern "C" int beom_host_run(beom_host *h, const beom_gpu_options *opt, int max_steps) {
  const beom_params &P = h->p;
  int rc = 0;
  if (!h->device_init) {
    beom_fields fld;
    beom_host_fields(h, &fld);
    rc = beom_gpu_init(&P, &fld, opt);
    if (rc) return gpu_fail("beom_gpu_init", rc);
  } else {
    std::printf(" read_input_data: on the device (beom_gpu_init_grids)\n");
  }

  std::printf(" lm = %d\n mm = %d\n", h->lm, h->mm);  // pm:233-234
  if (P.rsta > 0.5) {                                  // pm:236-243
    if ((rc = beom_host_read_restart(h))) return rc;
    std::printf(" *** Restarting from record number %d at time = %.15g\n", h->irec - 1, h->tres);
  }
  if (!h->device_init && (rc = beom_gpu_upload_state(h->hlay.data(), h->u.data(), h->v.data()))) return gpu_fail("beom_gpu_upload_state", rc);
  Pending pend;
  if (!h->odir.empty() && (rc = beom_gpu_set_rest_thickness(h->h_0_r4.data()))) return gpu_fail("beom_gpu_set_rest_thickness", rc);
  if (P.rsta < 0.5 && !h->odir.empty() && (rc = begin_outputs(h, 0.0, pend))) return rc;

  // integrate_time, pm:1840-1919
  std::printf(" dl = %.15g meters.\n dt = %.15g seconds.\n", P.dl, P.dt);
  const double dtd8 = P.dt / 24.0 / 3600.0;
  int nstp = h->nstp;
  if (max_steps > 0 && max_steps < nstp) nstp = max_steps;
  double ramp = 1.0, gene = 0.0;
  for (int tstp = 1; tstp <= nstp; tstp++) {
    const double ctim = h->tres + dtd8 * (double)tstp;
    if (tstp <= 3) {
      if (tstp == 1) {  // pm:1861-1866
        if ((rc = beom_gpu_stress())) return gpu_fail("beom_gpu_stress", rc);
        if (P.rsta < 0.5 && ctim < P.dt_r) ramp = ctim / P.dt_r;
      }
      if ((rc = beom_gpu_step(tstp, ctim, ramp, gene, 1, 1))) return gpu_fail("beom_gpu_step", rc);
      if (tstp == 3) {  // pm:1877-1884
        gene = P.g_fb;
        if (gene > 0.5 && P.rgld > 0.5) gene = 0.0;
      }
      continue;  // the reference writes no output during the first three steps
    }
    const bool upst = (tstp % h->n_3d) == 0;  // pm:1889-1896
    if (upst && (rc = beom_gpu_stress())) return gpu_fail("beom_gpu_stress", rc);
    ramp = 1.0;  // pm:1898-1901
    if (P.rsta < 0.5 && ctim < P.dt_r) ramp = ctim / P.dt_r;
    if ((rc = beom_gpu_step(tstp, ctim, ramp, gene, upst ? 1 : 0, 0))) return gpu_fail("beom_gpu_step", rc);
    if (tstp % h->notp == 0 && !h->odir.empty() && (rc = begin_outputs(h, ctim, pend))) return rc;  // pm:1908-1910
  }
  if ((rc = finish_outputs(h, pend))) return rc;
  if ((rc = beom_gpu_sync())) return gpu_fail("beom_gpu_sync", rc);
  if (h->device_init) {  // the final state for callers that look at the host arrays
    h->hlay.resize(h->nd1 * (size_t)h->nlay); h->u.resize(h->hlay.size()); h->v.resize(h->hlay.size());
  }
  if ((rc = beom_gpu_download_state(h->hlay.data(), h->u.data(), h->v.data()))) return gpu_fail("beom_gpu_download_state", rc);
  return 0;
}
