#!/bin/bash
# tools/ab_env.sh "VAR=a" "VAR=b" ... : time the default build under different environment settings (same box)
for v in "$@"; do
  env $v python bench.py --steps ${STEPS:-50} --warmup ${WARM:-10} --no-cpu --no-e2e 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), d['clocks']['reasons'])"
done
