#!/bin/bash
# round 2, call 1: the GPU suite without xfail marks, the bench line with the state checksum, store-path diagnostic, reference arm
out=gpurun_out; tag=r2c1; mkdir -p $out
(which gfortran flang nvfortran f95 2>&1; echo "gfortran probe exit $?"; nproc; free -g | head -2; nvidia-smi -L) > $out/${tag}_probe.log 2>&1
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest -m gpu: exit $?" >> $out/${tag}_pytest_gpu.log
tail -25 $out/${tag}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.log; cat $out/${tag}_bench_n1.json
STEPS=40 WARM=10 bash tools/ab.sh nostore ../beom_b200/lib nostore ../beom_b200/lib > $out/${tag}_ab.log 2>&1
cat $out/${tag}_ab.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/streambench tools/streambench.cu && ./tools/streambench > $out/${tag}_streambench.txt 2>&1; cat $out/${tag}_streambench.txt
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log; cat $out/${tag}_bench_ref.json; tail -3 $out/${tag}_bench_ref.log
