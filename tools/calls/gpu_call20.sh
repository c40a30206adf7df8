#!/bin/bash
# round 2, call 20 (the round's last): the whole GPU suite at the final state, smoke(), both bench arms, the ncu launch list of the
# bench command
out=gpurun_out; tag=r2c20; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" >> $out/${tag}_pytest_gpu.log; tail -6 $out/${tag}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log; cut -c1-260 $out/${tag}_bench_n1.json; grep -E "init" $out/${tag}_bench_n1.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log; cut -c1-200 $out/${tag}_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > $out/${tag}_ncu_list.log 2>&1
tail -1 $out/${tag}_ncu_list.log | cut -c1-200; wc -l $out/${tag}_launches.csv
