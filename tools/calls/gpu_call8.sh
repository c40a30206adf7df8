#!/bin/bash
# round 2, call 8: sill_like with the 168-register specialised instantiation, bench with the split init timing, grid-init tests incl. tides
out=gpurun_out; tag=r2c8; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 900 python -m pytest tests/test_grid_init.py tests/test_gpu_parity.py -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest.log 2>&1
echo "pytest: exit $?" >> $out/${tag}_pytest.log; tail -5 $out/${tag}_pytest.log
python bench.py --workload sill_like --size 4096 --steps 40 --warmup 5 --no-cpu --no-e2e 2> $out/${tag}_sill_4096.log | cut -c1-200; grep init $out/${tag}_sill_4096.log
python bench.py --workload sill_like --steps 20 --warmup 5 --no-cpu --no-e2e > $out/${tag}_bench_sill_like.json 2> $out/${tag}_bench_sill_like.log; cut -c1-200 $out/${tag}_bench_sill_like.json; grep init $out/${tag}_bench_sill_like.log
python bench.py --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log; cut -c1-200 $out/${tag}_bench_n1.json; grep -E "init|inputs" $out/${tag}_bench_n1.log
