"""Runs in its own process, so that a wedged kernel is killed by the caller's timeout instead of hanging the suite:
one reference test case (small version) through the C ABI on the GPU against the CPU oracle.

usage: case_worker.py <case> <nsteps> <fused 0|1> ['{"param": "value", ...}' appended to the parameter block
                      ['{"kwarg": value, ...}' for the case generator instead of the small defaults [variant]]]
Prints one JSON line {"path": ..., "exact": ..., "worst": ..., "bad": [...]}; exit status 0 = parity holds."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from beom_b200 import cases, model  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from tests.conftest import SMALL  # noqa: E402

name, nsteps, fused = sys.argv[1], int(sys.argv[2]), bool(int(sys.argv[3]))
extra = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
kwargs = json.loads(sys.argv[5]) if len(sys.argv) > 5 else SMALL.get(name, {})
variant = int(sys.argv[6]) if len(sys.argv) > 6 else 0
c = cases.CASES[name](**kwargs)
c.params_text += "".join("%-10s = %s\n" % kv for kv in extra.items())
with tempfile.TemporaryDirectory() as d:
    hm = model.HostModel.from_block(c.write(d), variant=variant)
    orc = Oracle(hm.params, d)
    orc.advance(1, nsteps)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=fused))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, nsteps)
    state = gm.download_state()
    aux = gm.download_aux()
    path = gm.path
    variant = gm.fused_variant
    graph_steps = gm.graph_launch_count()  # steps that ran as one CUDA-graph launch (latency-bound grids)
    gm.close()
    # cos() of the tidal targets is the one libm-dependent operation: 1e-11 of the field's
    # magnitude there, bit for bit everywhere else
    tol = 1.0e-11 if "tide" in c.files else 0.0
    own = np.ones(c.ndeg + 1, dtype=bool)  # frozen periodic duplicates carry no meaningful fluxes
    if hm.params.xper > 0.5 or hm.params.yper > 0.5:
        sub = hm.iarray("subc")
        own &= ~((sub[0] == c.lm + 1) | (sub[1] == c.mm + 1))
    bad, worst = [], 0.0
    pairs = [(n, a, orc.array(n), None) for n, a in zip(("hlay", "u", "v"), state)]
    pairs += [("h_u", aux[0], orc.array("h_u"), own), ("h_v", aux[1], orc.array("h_v"), own)]
    pairs += [(n, a.transpose(0, 2, 1), orc.array(n).transpose(0, 2, 1), own) for n, a in zip(("rs_h", "dmdx", "dmdy"), aux[2:])]
    for n, got, want, mask in pairs:
        want = want.reshape(got.shape)
        if mask is not None:
            got, want = got[..., mask], want[..., mask]
        if not np.all(np.isfinite(want)):
            bad.append(n + ": oracle not finite")
            continue
        scale = max(float(np.abs(want).max()), 1e-300)
        err = float(np.abs(got - want).max() / scale)
        worst = max(worst, err)
        if (tol == 0.0 and not np.array_equal(got, want)) or err > tol:
            bad.append("%s: %d entries differ, max %.3e of the field's magnitude" % (n, int(np.count_nonzero(got != want)), err))
    moved = float(np.abs(orc.array("u")).max() + np.abs(orc.array("v")).max())
    print(json.dumps({"case": name, "path": path, "variant": variant, "exact": tol == 0.0, "worst": worst, "bad": bad, "moved": moved,
                      "cell_layers": c.ndeg * c.nlay, "graph_steps": graph_steps}))
    sys.exit(1 if bad else 0)
