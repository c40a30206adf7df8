#!/bin/bash
# round 2, call 14: the GPU suite with the reference-made golden vectors of all 16 scripts, smoke() beside the prebuilt translated
# reference (oracle/_ref/smoke_40x25x2), both bench arms (the reference arm with its translated-reference leg)
out=gpurun_out; tag=r2c14; mkdir -p $out
ls -la oracle/_ref oracle/_ref/* > $out/${tag}_ref_listing.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" >> $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log; cut -c1-260 $out/${tag}_bench_n1.json; grep -E "init" $out/${tag}_bench_n1.log
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log; cut -c1-200 $out/${tag}_bench_ref.json; tail -3 $out/${tag}_bench_ref.log
python -c "
import json; d=json.load(open('$out/${tag}_bench_ref.json')); print('translated leg:', d.get('reference_translated'))"
