#!/bin/bash
# round 2, call 6 (8 GPUs): the bench at N = 8 (value, e2e, state checksum) and the edge-rows-first overlap variant
out=gpurun_out; tag=r2c6; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 ${@:2}; }
run 29521 --no-cpu > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.log; cut -c1-260 $out/${tag}_bench_n8.json; grep -o '"state_sha256": "[0-9a-f]*"' $out/${tag}_bench_n8.json; grep -o '"e2e": {"value": [0-9.e+]*' $out/${tag}_bench_n8.json
BEOM_OVERLAP=1 run 29522 --no-cpu --no-e2e --steps 100 > $out/${tag}_bench_n8_overlap.json 2> $out/${tag}_bench_n8_overlap.log; cut -c1-260 $out/${tag}_bench_n8_overlap.json
run 29523 --no-cpu --no-e2e --steps 100 > $out/${tag}_bench_n8_100.json 2> $out/${tag}_bench_n8_100.log; cut -c1-260 $out/${tag}_bench_n8_100.json
