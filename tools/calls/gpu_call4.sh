#!/bin/bash
# round 2, call 4: full GPU suite on the new chunking + wavefront rigid lid, bench line, chunk/jitter sweep, slab-shaped runs, rigid-lid timing, ncu
out=gpurun_out; tag=r2c4; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" >> $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log; cut -c1-900 $out/${tag}_bench_n1.json
STEPS=40 WARM=10 bash tools/ab_env.sh "BEOM_FUSED_CHUNKS=32" "BEOM_FUSED_CHUNKS=24" "BEOM_FUSED_CHUNKS=40" "BEOM_FUSED_CHUNKS=48" "BEOM_FUSED_CHUNKS=32 BEOM_FUSED_JITTER=3000" \
   "BEOM_FUSED_CHUNKS=32 BEOM_FUSED_JITTER=20000" "BEOM_FUSED_CHUNKS=4 BEOM_FUSED_JITTER=3000" "BEOM_FUSED_CHUNKS=4 BEOM_FUSED_JITTER=50000" "BEOM_FUSED_CHUNKS=4 BEOM_FUSED_JITTER=250000" > $out/${tag}_ab.log 2>&1
cat $out/${tag}_ab.log
for v in "BEOM_FUSED_CHUNKS=4" "BEOM_FUSED_CHUNKS=8" "BEOM_FUSED_CHUNKS=16" "BEOM_FUSED_CHUNKS=4 BEOM_FUSED_JITTER=50000" "BEOM_FUSED_CHUNKS=8 BEOM_FUSED_JITTER=50000"; do
  env $v python bench.py --rows 1024 --steps 100 --warmup 10 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rows=1024 $v', round(d['ms_per_step'],4), d['clocks']['reasons'])"
done > $out/${tag}_slab.log 2>&1; cat $out/${tag}_slab.log
python tools/time_rigid.py 1024 6 > $out/${tag}_rigid_wave.json 2> $out/${tag}_rigid_wave.log; cut -c1-700 $out/${tag}_rigid_wave.json
BEOM_PI_ONE_CTA=1 timeout 600 python tools/time_rigid.py 1024 2 > $out/${tag}_rigid_onecta.json 2> $out/${tag}_rigid_onecta.log; cut -c1-500 $out/${tag}_rigid_onecta.json
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 5 -c 2 -o $out/${tag}_fused_full \
    python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e > $out/${tag}_ncu_full.log 2>&1
ls -la $out | tail -5
