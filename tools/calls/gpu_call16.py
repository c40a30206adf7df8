#!/usr/bin/env python
"""round 2, call 16: A/B of the L2 prefetch of the fused step's staged rows (xlib/l2a<K>: -DBEOM_L2_AHEAD=K, the row segment a CTA
stages K rows later is prefetched into L2 by cp.async.bulk.prefetch.L2 when row R's staging copy is issued) against the default
build, on the full 8192 x 8192 x 4 basin and on the 1024-row slab of one of 8 ranks, with the y-chunk count varied: if the
prefetch hides the HBM latency that the "platooning" of long-running CTAs exposes, longer chunks (less pipeline
fill per owned row) should stop losing.  Adaptive: the best K of the first pass is the one the chunk sweeps use."""
import json
import os
import subprocess
import sys
import time

OUT = "gpurun_out"
TAG = "r2c16"
os.makedirs(OUT, exist_ok=True)
log = open(os.path.join(OUT, TAG + "_ab.txt"), "w")
T0 = time.time()
BUDGET = float(os.environ.get("AB_BUDGET", "400"))  # seconds of runs; what does not fit is skipped


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    log.write(line + "\n")
    log.flush()


def run(name, libdir=None, env=None, args=(), steps=40, warm=10):
    if time.time() - T0 > BUDGET:
        say(name, "skipped (time budget)")
        return None
    e = dict(os.environ)
    if libdir:
        e["BEOM_LIBDIR"] = "/root/repo/xlib/" + libdir
    e.update(env or {})
    t = time.time()
    p = subprocess.run([sys.executable, "bench.py", "--steps", str(steps), "--warmup", str(warm), "--no-cpu", "--no-e2e", *args],
                       env=e, capture_output=True, text=True, timeout=300)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
    except Exception:
        say(name, "FAILED rc", p.returncode, p.stderr[-400:])
        return None
    ms = d["ms_per_step"]
    say("%-34s %8.4f ms  frac %.4f  %s  sha %s  clocks %s %s  (%.0f s)" % (
        name, ms, d["roofline"]["frac"], d["config"].get("fused_variant", ""), (d.get("state_sha256") or "")[:12],
        d["clocks"]["sm_mhz"], d["clocks"]["reasons"], time.time() - t))
    return ms


say("# full grid 8192 x 8192 x 4, default chunking (32 chunks of 256 rows)")
res = {"default": run("default")}
for k in (2, 3, 4):
    res["l2a%d" % k] = run("l2a%d" % k, "l2a%d" % k)
res["default_b"] = run("default (again)")
best = min((k for k in ("l2a2", "l2a3", "l2a4") if res.get(k)), key=lambda k: res[k], default=None)
say("# best prefetch distance:", best)
if best:
    say("# full grid, chunk count varied (default build and %s)" % best)
    for ch in (16, 8):
        run("default chunks=%d" % ch, None, {"BEOM_FUSED_CHUNKS": str(ch)})
        run("%s chunks=%d" % (best, ch), best, {"BEOM_FUSED_CHUNKS": str(ch)})
    run("%s chunks=64" % best, best, {"BEOM_FUSED_CHUNKS": "64"})
    say("# the 1024-row slab of one of 8 ranks (default: 16 chunks of 64 rows)")
    run("default rows=1024", None, None, ("--rows", "1024"), 100, 20)
    for ch in (16, 8, 4, 2):
        run("%s rows=1024 chunks=%d" % (best, ch), best, {"BEOM_FUSED_CHUNKS": str(ch)}, ("--rows", "1024"), 100, 20)
    run("default rows=1024 chunks=8", None, {"BEOM_FUSED_CHUNKS": "8"}, ("--rows", "1024"), 100, 20)
    run("default rows=1024 (again)", None, None, ("--rows", "1024"), 100, 20)
    say("# the 2048-row slab of one of 4 ranks")
    run("default rows=2048", None, None, ("--rows", "2048"), 60, 10)
    run("%s rows=2048" % best, best, None, ("--rows", "2048"), 60, 10)
    say("# full grid once more")
    run("%s (again)" % best, best)
    run("default (third)")
say("# total %.0f s" % (time.time() - T0))
