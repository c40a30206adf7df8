"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/beom_oracle.c).

The reference ships no expected arrays, and it cannot be compiled here (no
Fortran compiler), so these vectors are NOT outputs of the reference: they freeze the oracle's own
results after it was pinned against the reference's analytical checks (tests/test_oracle_pins.py), and
give the GPU tests a fixture that does not depend on running the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from beom_b200 import cases, model  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

# name -> (generator kwargs, steps)
GOLDEN = {
    "stommel1948": (dict(dl=250.0e3), 120),
    "lock_exchange": (dict(), 200),
    "unstable_jet": (dict(dl=60.0e3), 80),
    "sill_exchange3D": (dict(lx=6.0e3, ly=100.0e3), 80),
    "conservation": (dict(dl=30.0e3), 100),
}


def run(name):
    kw, nsteps = GOLDEN[name]
    c = cases.CASES[name](**kw)
    with tempfile.TemporaryDirectory() as d:
        blk = c.write(d)
        hm = model.HostModel.from_block(blk)
        orc = Oracle(hm.params, d)
        orc.advance(1, nsteps)
        out = {k: orc.array(k).copy() for k in ("hlay", "u", "v", "h_u", "h_v")}
        out["nsteps"] = np.array(nsteps)
        out["eta_record"] = orc.record("eta_")
        return out


if __name__ == "__main__":
    for name in GOLDEN:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run(name))
        print("wrote", name)
