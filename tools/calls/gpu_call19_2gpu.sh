#!/bin/bash
# round 2, call 19 (2 GPUs): the multi-GPU tests at the final state (chunk-own row groups in the fused step, edge / interior launches,
# y-periodic rings, x-periodic slabs, rigid lid across slabs, device-side init on every rank), bench N = 2
out=gpurun_out; tag=r2c19; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_mgpu.log 2>&1
echo "pytest multigpu: exit $?" >> $out/${tag}_pytest_mgpu.log; tail -6 $out/${tag}_pytest_mgpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.log
cut -c1-240 $out/${tag}_bench_n2.json; grep -o '"state_sha256": "[0-9a-f]*"' $out/${tag}_bench_n2.json; grep -o '"e2e": {"value": [0-9.e+]*' $out/${tag}_bench_n2.json
