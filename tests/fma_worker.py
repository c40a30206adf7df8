"""Runs in its own process (BEOM_FMA is read once per process): the FMA-contracted fused step against the oracle."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BEOM_FMA"] = "1"

from beom_b200 import cases, model  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

nlay, nsteps = int(sys.argv[1]), int(sys.argv[2])
c = cases.synthetic_basin(n=300, mm=170, nlay=nlay)
with tempfile.TemporaryDirectory() as d:
    blk = c.write(d)
    hm = model.HostModel.from_block(blk)
    orc = Oracle(hm.params, d)
    orc.advance(1, nsteps)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, nsteps)
    got = gm.download_state()
    gm.close()
    worst, exact = 0.0, True
    for name, a in zip(("hlay", "u", "v"), got):
        w = orc.array(name).reshape(a.shape)
        scale = np.abs(w).max()
        worst = max(worst, float(np.abs(a - w).max() / scale))  # relative to the field's magnitude
        exact = exact and np.array_equal(a, w)
    print("fma max relative field difference %.3e bit_identical=%s" % (worst, exact))
    sys.exit(0 if worst <= 1.0e-10 else 1)
