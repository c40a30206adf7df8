#!/bin/bash
# round 2, call 13: the device-side init on hardware for all 16 scripts (+ periodic, segments), the whole GPU suite at the final state, bench
out=gpurun_out; tag=r2c13; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" >> $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log; cut -c1-260 $out/${tag}_bench_n1.json; grep -E "init" $out/${tag}_bench_n1.log
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log; cut -c1-200 $out/${tag}_bench_ref.json
