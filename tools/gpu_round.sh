#!/bin/bash
# One gpurun call that brings back everything a round needs from a B200 box (run from the repository root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r2'
# 1. the GPU test suite with the verdict of every xfail-marked test (-rxX): which of the cases that had never run on
#    hardware XPASS (promote them to plain tests) and which XFAIL (bugs to fix; the worker prints what differs)
# 2. the bench line (1 GPU) and the reference arm
# 3. the ncu launch list of the same bench command and one --set full capture of the fused step
# Everything lands under gpurun_out/<tag>_*; copy what is to be judged into profiles/.
# Two GPUs (the ring of y-periodic slabs, x-periodic slabs on the fused step -- so far only seen on the CPU emulation):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'python -m pytest tests/test_multigpu.py -m gpu -q -rxXs > gpurun_out/r2_mgpu.log 2>&1; tail -20 gpurun_out/r2_mgpu.log'
# Before any of it costs GPU minutes, the same kernels can be run on the CPU: python -m pytest tests/test_emulation.py
tag=${1:-round}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" >> $out/${tag}_pytest_gpu.log
tail -45 $out/${tag}_pytest_gpu.log
# the same workers, verbosely, for the cases that did not pass (what differs, which path ran)
for spec in "soliton 40" "baines_ridge 60" "carrier_beach 80" "upwelling_seaward_wind 40" "mixed_open_bc 40" "morel_upwelling 60" \
            "outcrop_seamount 40" "sill_exchange2D 40" "sill_exchange2Dtides 40" "tide_ridge 40" "wave_sponge 40"; do
  for fused in 0 1; do
    timeout 120 python tests/case_worker.py $spec $fused 2>&1 | tail -1
  done
done > $out/${tag}_case_workers.log 2>&1
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log && cat $out/${tag}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_step -c 2 -o $out/${tag}_fused_full \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/${tag}_ncu_full.log 2>&1
ls -la $out | tail -15
