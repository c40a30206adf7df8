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
    # the further reference scripts (small versions as in tests/conftest.py)
    "soliton": (dict(dl=80.0e3), 60),
    "baines_ridge": (dict(scale=0.3), 60),
    "carrier_beach": (dict(mesh=20.0), 80),
    "upwelling_seaward_wind": (dict(lm=80), 60),
    "mixed_open_bc": (dict(lm=60, mm=40), 60),
    "morel_upwelling": (dict(dl=4.0e3), 60),
    "outcrop_seamount": (dict(), 60),
    "sill_exchange2D": (dict(lx=140.0e3), 60),
    "sill_exchange2Dtides": (dict(lx=60.0e3), 60),
    "tide_ridge": (dict(lm=200), 60),
    "wave_sponge": (dict(dl=20.0e3), 60),
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
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path) and "--all" not in sys.argv:
            continue  # frozen: only missing vectors are made (--all regenerates everything)
        np.savez_compressed(path, **run(name))
        print("wrote", name)
