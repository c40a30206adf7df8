// main.cc -- the executable form of main.f95:26-37: errc/errm initialised, run(), quit().
// usage: beom_run <shared_mod.f95 | parameter block> [--steps N] [--split | --fused] [--variant 0..3] [--host-init]
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "beom_host.h"

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <shared_mod.f95> [--steps N] [--split | --fused] [--variant k] [--host-init]\n", argv[0]);
    return 2;
  }
  beom_params par;
  char idir[1024], odir[1024], desc[1024], err[2048];
  if (beom_params_parse_file(argv[1], &par, idir, odir, desc, sizeof idir)) {
    beom_host_last_error(err, sizeof err);
    std::fprintf(stderr, "  *** ERROR CODE = -1 ***\n In main.f95, %s\n", err);
    return 1;
  }
  int steps = 0;
  bool host_init = false;
  beom_gpu_options opt;
  beom_gpu_default_options(&opt);
  for (int a = 2; a < argc; a++) {
    if (!std::strcmp(argv[a], "--steps") && a + 1 < argc) steps = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--split")) opt.fused = 0;
    else if (!std::strcmp(argv[a], "--fused")) opt.fused = 1;  // (default: by size, include/beom_gpu.h)
    else if (!std::strcmp(argv[a], "--variant") && a + 1 < argc) par.variant = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--host-init")) host_init = true;
  }
  // read_input_data on the device where the device path covers the case (restart runs read their state on the host)
  beom_host *h = (host_init || par.rsta > 0.5) ? beom_host_create(&par, idir, odir, desc) : beom_host_create_on_device(&par, idir, odir, desc, &opt);
  if (!h) {
    beom_host_last_error(err, sizeof err);
    std::fprintf(stderr, "  *** ERROR CODE = -1 ***\n %s\n", err);
    return 1;
  }
  int rc = beom_host_run(h, &opt, steps);
  if (rc) {
    beom_host_last_error(err, sizeof err);
    std::fprintf(stderr, "  *** ERROR CODE = %d ***\n %s\n", rc, err);
  }
  beom_gpu_finalize();
  beom_host_destroy(h);
  return rc ? 1 : 0;
}
