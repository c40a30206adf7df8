#!/bin/bash
# round 2, call 7: full GPU suite (option-set instantiations, device-side init, ...), bench with device-side init, sill_like workload
out=gpurun_out; tag=r2c7; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" >> $out/${tag}_pytest_gpu.log; tail -8 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -3 $out/${tag}_smoke.log
( time python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log ) 2> $out/${tag}_bench_n1.time; cut -c1-300 $out/${tag}_bench_n1.json; grep -E "init|inputs" $out/${tag}_bench_n1.log; cat $out/${tag}_bench_n1.time
python bench.py --workload sill_like --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_sill_like.json 2> $out/${tag}_bench_sill_like.log; cut -c1-200 $out/${tag}_bench_sill_like.json; grep -o '"fused_variant": "[^"]*"' $out/${tag}_bench_sill_like.json; grep -E "init" $out/${tag}_bench_sill_like.log
BEOM_FUSED_GENERAL=1 python bench.py --workload sill_like --steps 20 --warmup 5 --no-cpu --no-e2e > $out/${tag}_bench_sill_like_general.json 2> $out/${tag}_bench_sill_like_general.log; cut -c1-200 $out/${tag}_bench_sill_like_general.json; grep -o '"fused_variant": "[^"]*"' $out/${tag}_bench_sill_like_general.json
BEOM_HOST_INIT=1 python bench.py --steps 5 --warmup 3 --no-cpu 2> $out/${tag}_bench_hostinit.log | grep -o '"state_sha256": "[0-9a-f]*", "state_sha256_nsteps": [0-9]*'; grep -E "init" $out/${tag}_bench_hostinit.log
python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | grep -o '"state_sha256": "[0-9a-f]*", "state_sha256_nsteps": [0-9]*'
