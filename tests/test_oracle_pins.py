"""What pins the CPU oracle.  The reference has no golden vectors; its correctness references are the
analytical solutions inside its test scripts and the conservation property its documentation states
(SURVEY.md section 8c).  The oracle must reproduce those, and must keep reproducing its own frozen
outputs (tests/golden, made by tests/golden/make_golden.py)."""
import os
import tempfile

import numpy as np
import pytest

from beom_b200 import cases, model, readers
from oracle.pyoracle import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make(case):
    d = tempfile.mkdtemp(prefix="beom_pin_")
    blk = case.write(d)
    hm = model.HostModel.from_block(blk)
    return hm, Oracle(hm.params, d)


def test_stommel_gyre_matches_the_analytical_solution():
    """testcases/stommel1948.m:98-104: sea-surface height of Stommel (1948) after the script's 40 days."""
    c = cases.stommel1948()
    hm, orc = make(c)
    nstp, _, _ = orc.counts()
    assert nstp == 3062
    orc.advance(1, nstp)
    eta = readers.vector_to_grid(orc.array("hlay")[0] - orc.array("h_0")[0], orc.iarray("subc"), c.lm, c.mm)[1:-1, 1:-1]
    ana = c.info["eta_analytic"]
    em, ea = eta - eta[0, 0], ana - ana[0, 0]  # the script references both to the south-west corner
    assert np.corrcoef(em.ravel(), ea.ravel())[0, 1] > 0.998
    assert np.sqrt(np.mean((em - ea) ** 2)) / np.sqrt(np.mean(ea ** 2)) < 0.05
    assert abs(em.min() / ea.min() - 1.0) < 0.05  # depth of the western-intensified low


def test_lock_exchange_front_speed():
    """testcases/lock_exchange.m:42,116-120: the gravity current advances at about
    cint = 0.5 sqrt(g' H) (Shchepetkin et al. 2015)."""
    c = cases.lock_exchange()
    hm, orc = make(c)
    nsteps = int(round(6 * 3600.0 / hm.params.dt))
    orc.advance(1, nsteps)
    h2 = readers.vector_to_grid(orc.array("hlay")[1], orc.iarray("subc"), c.lm, c.mm)[1:-1, 1]
    x = (np.arange(c.lm) + 0.5) * 400.0 - 0.5 * c.lm * 400.0
    front = x[h2 > 1.0].max()
    assert 0.9 < front / (c.info["cint"] * nsteps * hm.params.dt) < 1.15
    assert h2.min() > 0.5 * hm.params.hmin  # the guard of private_mod.f95:2798-2808


def test_equatorial_soliton_travels_west_at_boyds_speed():
    """testcases/soliton.m (Lavelle & Thacker 2008, after Boyd 1980): on the equatorial beta plane (fcor.bin, periodic
    in x, advection on, no dissipation) the Rossby soliton keeps its shape and moves west at about
    -(1/3 + 0.395 B^2) sqrt(g H); the script only animates it, the speed is the analytical reference."""
    c = cases.soliton(dl=40.0e3)
    hm, orc = make(c)
    assert hm.params.xper > 0.5 and c.ndeg == hm.params.ndeg

    def peak():
        eta = readers.vector_to_grid(orc.array("hlay")[0] - orc.array("h_0")[0], orc.iarray("subc"), c.lm, c.mm)[1:-1, 1:-1]
        prof = np.nan_to_num(eta).max(axis=1)
        i = int(np.argmax(prof))
        a, b, d = prof[(i - 1) % c.lm], prof[i], prof[(i + 1) % c.lm]
        return (i + 0.5 * (a - d) / (a - 2.0 * b + d)) * c.info["dl"], float(np.nanmax(eta))

    x0, a0 = peak()
    assert abs(a0 / c.info["amplitude"] - 1.0) < 1.0e-3
    nsteps = int(round(20.0 * 86400.0 / hm.params.dt))
    orc.advance(1, nsteps)
    x1, a1 = peak()
    L = c.lm * c.info["dl"]
    speed = (((x1 - x0 + 0.5 * L) % L) - 0.5 * L) / (nsteps * hm.params.dt)
    assert speed < 0.0 and abs(speed / c.info["speed"] - 1.0) < 0.04  # Boyd's first-order speed, -1.2355 m/s
    assert a1 / a0 > 0.90                                               # the soliton holds together


def test_volume_is_conserved_to_roundoff():
    """doc p.4-6 / testcases/conservation.m:116-141: the area-mean layer thickness stays within ~1e-10 m."""
    c = cases.conservation(dl=30.0e3)
    hm, orc = make(c)
    wet = orc.array("mk_n")[0] > 0.5
    v0 = orc.array("hlay")[:, wet].mean(axis=1).copy()
    orc.advance(1, 3000)
    v1 = orc.array("hlay")[:, wet].mean(axis=1)
    assert np.all(np.abs(v1 - v0) < 1.0e-10)
    ke = (orc.array("u") ** 2 + orc.array("v") ** 2).sum()
    assert np.isfinite(ke) and ke > 0


@pytest.mark.parametrize("name", ["stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D", "conservation"])
def test_oracle_reproduces_its_golden_vectors(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    got = mg.run(name)
    want = np.load(os.path.join(GOLD, name + ".npz"))
    for k in ("hlay", "u", "v", "h_u", "h_v", "eta_record"):
        assert np.array_equal(got[k], want[k]), k


def test_openmp_build_of_the_oracle_agrees_with_the_strict_build():
    """The -O3 -fopenmp build (CPU baseline timing) may contract FMAs: agreement to 1e-12 relative."""
    c = cases.synthetic_basin(n=48, mm=40, nlay=3)
    d = tempfile.mkdtemp(prefix="beom_pin_")
    blk = c.write(d)
    hm = model.HostModel.from_block(blk)
    a, b = Oracle(hm.params, d), Oracle(hm.params, d, omp=True)
    a.advance(1, 25)
    b.advance(1, 25)
    for k in ("hlay", "u", "v"):
        x, y = a.array(k), b.array(k)
        assert np.max(np.abs(x - y)) <= 1e-12 * max(1.0, np.max(np.abs(x))), k
