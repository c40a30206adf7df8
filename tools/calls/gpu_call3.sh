#!/bin/bash
# round 2, call 3: native-size configs + output records on hardware, y-chunk / CTA-order sweep of the fused step, one full ncu capture
out=gpurun_out; tag=r2c3; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_native_configs.py tests/test_gpu_outputs.py -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest.log 2>&1
echo "pytest: exit $?" >> $out/${tag}_pytest.log; tail -8 $out/${tag}_pytest.log
STEPS=40 WARM=10 bash tools/ab_env.sh "BEOM_FUSED_CHUNKS=4" "BEOM_FUSED_CHUNKS=8" "BEOM_FUSED_CHUNKS=16" "BEOM_FUSED_CHUNKS=32" "BEOM_FUSED_CHUNKS=64" \
   "BEOM_FUSED_CHUNKS=4" "BEOM_FUSED_CHUNKS=12" "BEOM_FUSED_CHUNKS=4 BEOM_FUSED_ORDER=1" "BEOM_FUSED_CHUNKS=8 BEOM_FUSED_ORDER=1" "BEOM_FUSED_CHUNKS=16 BEOM_FUSED_ORDER=1" > $out/${tag}_ab.log 2>&1
cat $out/${tag}_ab.log
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -s 5 -c 2 -o $out/${tag}_fused_full \
    python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_ncu_full.log | cut -c1-300; ls -la $out | tail -8
