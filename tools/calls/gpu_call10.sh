#!/bin/bash
# round 2, call 10: ncu launch list of the bench command (per-launch durations) for profiles/
out=gpurun_out; tag=r2c10; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
python bench.py --steps 4 --warmup 3 --no-cpu > $out/${tag}_plain.json 2> $out/${tag}_plain.log; cut -c1-200 $out/${tag}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > $out/${tag}_ncu_list.log 2>&1
tail -2 $out/${tag}_ncu_list.log | cut -c1-200; wc -l $out/${tag}_launches.csv
