"""Times the rigid-lid pressure solve (surf_pressure, private_mod.f95:1705-1838) on a GPU: the all-SM wavefront solver
(rigid.cuh) against the single-block wavefront it replaces (BEOM_PI_ONE_CTA=1), same iterates.
usage: python tools/time_rigid.py [n] [steps]      (run once per solver: the switch is read when the library loads)"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from beom_b200 import cases, model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
c = cases.synthetic_basin(n=n, nlay=2)
c.params_text += "rgld       = 1.\nocrp       = 1.\n"
with tempfile.TemporaryDirectory() as d:
    blk = c.write(d)
    hm = model.HostModel.from_block(blk)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    out = []
    for t in range(1, steps + 1):
        gm.sync()
        t0 = time.perf_counter()
        gm.mark(0)
        gm.advance(t, t)
        gm.mark(1)
        gm.sync()
        out.append({"step": t, "ms": gm.elapsed_ms(), "sweeps": gm.pi_iterations(), "wall_ms": 1e3 * (time.perf_counter() - t0)})
    pi_s = gm.download_pi_s()
    hl, u, v = gm.download_state()
    gm.close()
    import hashlib
    sha = hashlib.sha256(pi_s.tobytes() + hl.tobytes() + u.tobytes() + v.tobytes()).hexdigest()
    print(json.dumps({"grid": [n, n, 2], "solver": "one block" if os.environ.get("BEOM_PI_ONE_CTA") else "all SMs (rigid.cuh)",
                      "steps": out, "state_sha256": sha, "max_abs_pi_s": float(np.abs(pi_s).max())}))
