#!/bin/bash
# round 2, call 15: the drop-in tests on hardware (the reference's own program on libbeom_gpu.so), then an A/B of the bare 4-layer
# step with 2 column groups per CTA = 8 warps, two CTAs per SM (xlib/g2: -DBEOM_LEAN4_GROUPS=2 -DBEOM_MIN_CTAS=2 -DBEOM_WROW=33
# -DBEOM_OWORDS=16) against the default (4 groups = 16 warps, one CTA per SM), full grid and the 1024-row slab of 8 GPUs
out=gpurun_out; tag=r2c15; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -rxXs -p no:cacheprovider > $out/${tag}_pytest_dropin.log 2>&1
echo "pytest dropin: exit $?" >> $out/${tag}_pytest_dropin.log; tail -4 $out/${tag}_pytest_dropin.log
ab() { # name libdir extra-args
  BEOM_FUSED_VERBOSE=1 BEOM_LIBDIR=$2 python bench.py --steps 40 --warmup 10 --no-cpu --no-e2e $3 2> $out/${tag}_ab_err.log |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 $3', round(d['ms_per_step'],4), d['config']['fused_variant'], d['clocks']['reasons'], d['state_sha256'][:12] if d.get('state_sha256') else '')"
  grep "\[fused\]" $out/${tag}_ab_err.log | tail -1
}
{
ab default "" ""; ab g2 /root/repo/xlib/g2 ""; ab default "" ""; ab g2 /root/repo/xlib/g2 ""
ab default "" "--rows 1024"; ab g2 /root/repo/xlib/g2 "--rows 1024"; ab default "" "--rows 1024"; ab g2 /root/repo/xlib/g2 "--rows 1024"
} 2>&1 | tee $out/${tag}_ab.txt
