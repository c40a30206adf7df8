#!/bin/bash
# round 2, call 5 (2 GPUs): the multi-GPU parity tests (ring-closed y-periodic, x-periodic slabs, rigid lid across slabs), the bench at N = 2
# with and without the edge-rows-first overlap; the state checksum must equal the N = 1 line's
out=gpurun_out; tag=r2c5; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_mgpu.log 2>&1
echo "pytest multigpu: exit $?" >> $out/${tag}_pytest_mgpu.log; tail -12 $out/${tag}_pytest_mgpu.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 ${@:2}; }
run 29511 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.log; cut -c1-260 $out/${tag}_bench_n2.json; grep -o '"state_sha256": "[0-9a-f]*"' $out/${tag}_bench_n2.json
BEOM_OVERLAP=1 run 29512 --no-cpu > $out/${tag}_bench_n2_overlap.json 2> $out/${tag}_bench_n2_overlap.log; cut -c1-260 $out/${tag}_bench_n2_overlap.json; grep -o '"state_sha256": "[0-9a-f]*"' $out/${tag}_bench_n2_overlap.json
run 29513 --no-cpu --no-e2e --steps 60 | cut -c1-220
BEOM_OVERLAP=1 run 29514 --no-cpu --no-e2e --steps 60 | cut -c1-220
