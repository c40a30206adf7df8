#!/bin/bash
# round 2, call 2: the named configs at native size, A/B of the refill-placement / bulk-copy-form variants, one full ncu capture
out=gpurun_out; tag=r2c2; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_native_configs.py -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_native.log 2>&1
echo "pytest native: exit $?" >> $out/${tag}_pytest_native.log; tail -8 $out/${tag}_pytest_native.log
STEPS=40 WARM=10 bash tools/ab.sh ../beom_b200/lib late1 late2 bulkcta late1cta ../beom_b200/lib late1 > $out/${tag}_ab.log 2>&1
STEPS=40 WARM=10 bash tools/ab_env.sh "BEOM_FUSED_CHUNKS=2" "BEOM_FUSED_CHUNKS=8" "BEOM_FUSED_CHUNKS=6" >> $out/${tag}_ab.log 2>&1
cat $out/${tag}_ab.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 9 -c 2 -o $out/${tag}_fused_full \
    python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_ncu_full.log; ls -la $out | tail -8
