#!/bin/bash
# round 2, call 11 (2 GPUs): the multi-GPU tests at the final state (overlap default, device-side init on every rank, rigid lid across slabs), bench N = 2
out=gpurun_out; tag=r2c11; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_mgpu.log 2>&1
echo "pytest multigpu: exit $?" >> $out/${tag}_pytest_mgpu.log; tail -6 $out/${tag}_pytest_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.log
cut -c1-240 $out/${tag}_bench_n2.json; grep -o '"state_sha256": "[0-9a-f]*"' $out/${tag}_bench_n2.json; grep -E "init" $out/${tag}_bench_n2.log
