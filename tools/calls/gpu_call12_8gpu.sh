#!/bin/bash
# round 2, call 12 (8 GPUs): the bench at N = 8 and N = 4 at the final state (overlap default, device-side init)
out=gpurun_out; tag=r2c12; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu ${@:3}; }
run 8 29541 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.log; cut -c1-240 $out/${tag}_bench_n8.json; grep -o '"state_sha256": "[0-9a-f]*"' $out/${tag}_bench_n8.json; grep -o '"e2e": {"value": [0-9.e+]*' $out/${tag}_bench_n8.json; grep -E "rank 0.*init" $out/${tag}_bench_n8.log
run 4 29542 > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.log; cut -c1-240 $out/${tag}_bench_n4.json; grep -o '"state_sha256": "[0-9a-f]*"' $out/${tag}_bench_n4.json; grep -o '"e2e": {"value": [0-9.e+]*' $out/${tag}_bench_n4.json
