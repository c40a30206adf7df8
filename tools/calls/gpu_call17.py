#!/usr/bin/env python
"""round 2, call 17: same-box A/B of the row groups that start at the chunk's own first row (commit 85ff445: chunks of 4k + 3 rows,
no alignment rows) against the build before it (xlib/prev = commit c01efe1: groups start on multiples of 4), interleaved, on the
full 8192 x 8192 x 4 basin and on the slabs of 8 and 4 ranks; the state hash of both builds after 20 steps (must be the one every
earlier bench line carries); a chunk-count sweep of the new build on the 1024-row slab."""
import json
import os
import subprocess
import sys
import time

OUT = "gpurun_out"
TAG = "r2c17"
os.makedirs(OUT, exist_ok=True)
log = open(os.path.join(OUT, TAG + "_ab.txt"), "w")
T0 = time.time()
BUDGET = float(os.environ.get("AB_BUDGET", "330"))


def say(*a):
    line = " ".join(str(x) for x in a)
    print(line, flush=True)
    log.write(line + "\n")
    log.flush()


def run(name, libdir=None, env=None, args=(), steps=40, warm=10, e2e=False):
    if time.time() - T0 > BUDGET:
        say(name, "skipped (time budget)")
        return None
    e = dict(os.environ)
    if libdir:
        e["BEOM_LIBDIR"] = "/root/repo/xlib/" + libdir
    e.update(env or {})
    t = time.time()
    cmd = [sys.executable, "bench.py", "--steps", str(steps), "--warmup", str(warm), "--no-cpu", *args]
    if not e2e:
        cmd.append("--no-e2e")
    p = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=300)
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


say("# state after 20 steps (8192 x 8192 x 4; earlier lines: 685074ca375c)")
run("new  (20 steps, e2e)", None, None, (), 20, 5, True)
run("prev (20 steps, e2e)", "prev", None, (), 20, 5, True)
say("# the 1024-row slab of one of 8 ranks")
for k in range(3):
    run("new  rows=1024", None, None, ("--rows", "1024"), 100, 20)
    run("prev rows=1024", "prev", None, ("--rows", "1024"), 100, 20)
say("# full grid")
for k in range(2):
    run("new", None)
    run("prev", "prev")
say("# the 2048-row slab of one of 4 ranks")
run("new  rows=2048", None, None, ("--rows", "2048"), 60, 10)
run("prev rows=2048", "prev", None, ("--rows", "2048"), 60, 10)
say("# new build, 1024-row slab, chunk count (default 16)")
for ch in (12, 20, 24, 32):
    run("new  rows=1024 chunks=%d" % ch, None, {"BEOM_FUSED_CHUNKS": str(ch)}, ("--rows", "1024"), 100, 20)
say("# new build, full grid, chunk count (default 32)")
for ch in (40, 48):
    run("new  chunks=%d" % ch, None, {"BEOM_FUSED_CHUNKS": str(ch)})
say("# total %.0f s" % (time.time() - T0))
