#!/bin/bash
# round 2, call 18 (4 GPUs): the bench at N = 4 at the final state (chunk-own row groups, page-locked host arrays on the device's
# NUMA node) and the same with BEOM_HOST_NUMA=0 -- what the placement does to the end-to-end leg; the box's GPU / NUMA topology
out=gpurun_out; tag=r2c18; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
{ nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do [ "$(cat $d/class 2>/dev/null)" = "0x030200" ] && echo "$d numa_node $(cat $d/numa_node)"; done; lscpu | grep -i -E "numa|socket|^CPU\(s\)"; } > $out/${tag}_topology.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 20 --warmup 5 --no-cpu ${@:3}; }
show() { cut -c1-200 $1; grep -o '"state_sha256": "[0-9a-f]*"' $1; grep -o '"e2e": {"value": [0-9.e+]*' $1; grep -o '"seconds": [0-9.e+-]*' $1; }
run 4 29551 > $out/${tag}_bench_n4.json 2> $out/${tag}_bench_n4.log; show $out/${tag}_bench_n4.json
BEOM_HOST_NUMA=0 run 4 29552 > $out/${tag}_bench_n4_nonuma.json 2> $out/${tag}_bench_n4_nonuma.log; show $out/${tag}_bench_n4_nonuma.json
run 4 29553 > $out/${tag}_bench_n4_b.json 2> $out/${tag}_bench_n4_b.log; show $out/${tag}_bench_n4_b.json
tail -3 $out/${tag}_topology.txt
