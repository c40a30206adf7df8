#!/bin/bash
# round 2, call 9: the 4-layer bare step with 3 column groups (12 warps at 168 registers) against 4 (16 warps at 128)
out=gpurun_out; tag=r2c9; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
STEPS=40 WARM=10 bash tools/ab.sh ../beom_b200/lib lean4g3 ../beom_b200/lib lean4g3 > $out/${tag}_ab.log 2>&1
for v in "BEOM_FUSED_CHUNKS=24" "BEOM_FUSED_CHUNKS=48"; do
  env $v BEOM_LIBDIR=/root/repo/xlib/lean4g3 python bench.py --steps 40 --warmup 10 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lean4g3 $v', round(d['ms_per_step'],3), d['clocks']['reasons'])"
done >> $out/${tag}_ab.log 2>&1
for lib in /root/repo/beom_b200/lib /root/repo/xlib/lean4g3; do
  BEOM_LIBDIR=$lib python bench.py --rows 1024 --steps 100 --warmup 10 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rows=1024 $lib', round(d['ms_per_step'],4), d['clocks']['reasons'])"
done >> $out/${tag}_ab.log 2>&1
cat $out/${tag}_ab.log
