"""Regenerates tests/golden/*.npz -- outputs of THE REFERENCE ITSELF, run in the development container.

The reference ships no expected arrays and there is no Fortran compiler here, so the reference's
own sources (/root/reference/{shared_mod,private_mod,main}.f95, read where they lie) are translated to C++ by
oracle/f95c (a language translator without any knowledge of the model), compiled with g++ in strict IEEE mode and run
on the inputs of each test-case script (oracle/refbuild.py).  Every vector below is what that program left in its module
arrays after ``nsteps`` time steps (its raw double-precision state, dumped at STOP) and the last record of the
``eta_.bin`` it wrote.  The hand-written oracle reproduces all of them bit for bit
(tests/test_oracle_pins.py::test_oracle_reproduces_the_reference_vectors runs wherever the fixtures are, also without
/root/reference; tests/test_reference_pin.py re-runs the translated reference and compares every module array).

    python tests/golden/make_golden.py [--all]          (needs /root/reference)

``run(name)`` is the oracle's side of the same computation (used by the tests)."""
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
KEYS = ("hlay", "u", "v", "h_u", "h_v")


def run(name):
    """The oracle's vectors for one case."""
    kw, nsteps = GOLDEN[name]
    c = cases.CASES[name](**kw)
    with tempfile.TemporaryDirectory() as d:
        blk = c.write(d)
        hm = model.HostModel.from_block(blk)
        orc = Oracle(hm.params, d)
        orc.advance(1, nsteps)
        out = {k: orc.array(k).copy() for k in KEYS}
        out["nsteps"] = np.array(nsteps)
        out["eta_record"] = orc.record("eta_")
        return out


def run_reference(name):
    """The same vectors from the translated reference (development container only)."""
    from oracle import refbuild, refcheck

    kw, nsteps = GOLDEN[name]
    c = cases.CASES[name](**kw)
    with tempfile.TemporaryDirectory(prefix="g_") as d:
        blk = c.write(d)
        text = open(blk).read()
        p, _, _, _ = model.parse_params(text)
        text = refcheck.block_for_steps(text, nsteps, p.dt)
        p, _, odir, _ = model.parse_params(text)
        dump, _ = refbuild.run_case(refbuild.build_case(text), odir)
        out = {k: np.ascontiguousarray(dump[k].T) for k in KEYS}  # (0:ndeg, nlay) -> [nlay][0:ndeg]
        out["nsteps"] = np.array(nsteps)
        out["eta_record"] = refcheck.read_records(os.path.join(odir, "eta_.bin"), p.ndeg, p.nlay)[-1].copy()
        out["source"] = np.array("translated reference (oracle/f95c), g++ -O2 -ffp-contract=off")
        return out


if __name__ == "__main__":
    for name in GOLDEN:
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path) and "--all" not in sys.argv:
            continue  # frozen: only missing vectors are made (--all regenerates everything)
        ref = run_reference(name)
        mine = run(name)
        same = all(np.array_equal(ref[k], mine[k]) for k in KEYS + ("eta_record",))
        np.savez_compressed(path, **ref)
        print("wrote %-24s (reference output; the oracle %s)" % (name, "agrees bit for bit" if same else "DIFFERS"))
