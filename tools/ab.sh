#!/bin/bash
# A/B timing of experiment builds of libbeom_gpu.so (xlib/<name>, see beom_b200/build.py BEOM_LIBDIR / BEOM_NVCC_DEFS)
# on one GPU box: tools/ab.sh D X D X ...   (interleave a reference build; boxes differ by ~10 %)
for v in "$@"; do
  BEOM_LIBDIR=/root/repo/xlib/$v python bench.py --steps ${STEPS:-100} --warmup ${WARM:-20} --no-cpu --no-e2e 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), d['clocks'])"
done
