"""What pins the CPU oracle.  The reference holds no golden vectors of its own; its correctness references are the
analytical solutions inside its test scripts and the conservation property its documentation states
(SURVEY.md section 8c): the oracle must reproduce those.  Since round 2 it must also reproduce, bit for bit, outputs of
the reference itself -- tests/golden/*.npz, made by running the reference's own sources (translated to C++ by
oracle/f95c, no Fortran compiler exists here) in the development container; tests/test_reference_pin.py re-runs that
program where /root/reference is present and compares every module array."""
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


def _grid(orc, c, a):
    return np.nan_to_num(readers.vector_to_grid(a, orc.iarray("subc"), c.lm, c.mm))


def test_flow_over_a_ridge_matches_baines_and_leonard():
    """testcases/baines_ridge.m:100-138, 190-206: steady lower-layer thickness of a rotating supercritical flow over a
    cosine ridge (Baines & Leonard 1989, Eq. 5.1-5.5); the script overlays model and theory after its 10 days.  Exercises
    y-periodicity of a one-row channel, bodf.bin, east/west sponges and a moving initial state."""
    c = cases.baines_ridge()
    hm, orc = make(c)
    nstp, _, _ = orc.counts()
    assert nstp == 41087 and c.info["F_0"] > 1.0
    orc.advance(1, nstp)
    h2 = _grid(orc, c, orc.array("hlay")[1])[:, 1]
    u2 = _grid(orc, c, orc.array("u")[1])[:, 1]
    sel = np.abs(c.info["X"]) < 20.0  # the ridge (|X| < 5 Rossby radii) and the lee waves behind it
    d_an, u_an = c.info["d_an"], c.info["u_an"]
    assert np.corrcoef(h2[sel], d_an[sel])[0, 1] > 0.95
    assert np.sqrt(np.mean((h2[sel] - d_an[sel]) ** 2)) < 0.3 * np.sqrt(np.mean((d_an[sel] - c.info["d_0"]) ** 2))
    assert abs(h2[sel].min() - d_an[sel].min()) < 0.15 and abs(h2[sel].max() - d_an[sel].max()) < 0.15  # of +-0.5 m
    assert np.max(np.abs(u2[sel] - u_an[sel])) < 0.03  # m/s, of 1.2


def test_shoreline_on_a_beach_follows_carrier_and_greenspan():
    """testcases/carrier_beach.m:93-114, 195-255: wetting and drying -- the shoreline (where the layer thins to Salmon's
    thickness) against the analytical run-up and run-down of Carrier & Greenspan (1958), Eq. 3.23-3.29."""
    c = cases.carrier_beach()
    hm, orc = make(c)
    nstp, notp, _ = orc.counts()
    assert hm.params.ocrp > 0.5 and abs(hm.params.hsal / c.info["hsal"] - 1.0) < 1e-6
    track = []
    for k in range(0, nstp - notp + 1, notp):
        orc.advance(k + 1, k + notp)
        h = _grid(orc, c, orc.array("hlay")[0])[:, 1]
        i = int(np.nonzero(h[1:] < c.info["hsal"])[0][0]) + 1  # first dry cell east of the western margin
        xm = np.interp(c.info["hsal"], [h[i], h[i - 1]], [c.info["xref"][i], c.info["xref"][i - 1]])
        t = (k + notp) * hm.params.dt
        track.append((t, xm, np.interp(t, c.info["t_sl"], c.info["x_sl"])))
    t, xm, xa = np.array(track).T
    early = t < 3600.0  # run-up, run-down and first rebound; later the theory's tail is a few metres, below the mesh size
    assert xa[early].max() > 420.0 and xa[early].min() < -180.0
    assert abs(xm[early].max() / xa[early].max() - 1.0) < 0.12   # highest run-up, about 435 m
    assert abs(xm[early].min() / xa[early].min() - 1.0) < 0.12   # deepest run-down, about -190 m
    assert np.sqrt(np.mean((xm[early] - xa[early]) ** 2)) < 0.06 * (xa.max() - xa.min())
    assert np.all(np.isfinite(orc.array("hlay"))) and orc.array("hlay")[0, 1:].min() >= 0.0


def _inertial_mean(orc, c, hm, f0, fields):
    """Mean over the last inertial period (the steady solutions below carry superposed inertial oscillations)."""
    nstp, _, _ = orc.counts()
    per = int(round(2.0 * np.pi / f0 / hm.params.dt))
    orc.advance(1, nstp - per)
    acc, cnt = [0.0] * len(fields), 0
    for k in range(nstp - per, nstp - 19, 20):
        orc.advance(k + 1, k + 20)
        for q, f in enumerate(fields):
            acc[q] = acc[q] + f()
        cnt += 1
    return [a / cnt for a in acc]


def test_seaward_wind_upwelling_matches_millot_and_crepon():
    """testcases/upwelling_seaward_wind.m:57-119: interface displacement and along-shore current of the lower layer next
    to the coast against the steady two-layer solution of Millot & Crepon (1981).  No input file: uniform ``tauw`` with a
    4-day ramp, wind stress spread over the top hsbl = 10 m, channel periodic in y."""
    c = cases.upwelling_seaward_wind()
    hm, orc = make(c)
    assert orc.counts()[0] == 20529 and not c.files
    eta2, v2 = _inertial_mean(orc, c, hm, 1.0e-4, [lambda: _grid(orc, c, orc.array("hlay")[1] - orc.array("h_0")[1])[:, 1],
                                                    lambda: _grid(orc, c, orc.array("v")[1])[:, 1]])
    te, tv = c.info["t_eta"][1], c.info["t_v"][1]
    near = slice(1, 9)  # the first 8 km: about three internal Rossby radii (r_2 = 3.2 km)
    assert te[1] > 1.3 and np.max(np.abs(eta2[near] - te[near])) < 0.06       # metres, of 1.31 at the coast
    assert np.max(np.abs(v2[1:20] - tv[1:20])) < 0.0015                       # m/s, of 0.024 offshore
    assert abs(v2[20] / tv[20] - 1.0) < 0.06


def test_mixed_open_boundaries_keep_the_upwelling_solution():
    """testcases/mixed_open_bc.m:134-196: the same upwelling in a basin with a coast, a wave sponge (east) and weakly
    relaxed open boundaries (north/south, no_gradient_obc with mcbc = 0); the mid-basin section must still show the
    Millot & Crepon solution."""
    c = cases.mixed_open_bc(lm=80, mm=40)
    d = tempfile.mkdtemp(prefix="beom_pin_")
    hm = model.HostModel.from_block(c.write(d))
    orc = Oracle(hm.params, d, omp=True)
    assert orc.nseg() > 0 and hm.scalar("flag_nudging") and hm.params.mcbc < 0.5
    ix_0 = int(np.floor(0.5 * (c.mm + 2) + 0.5)) - 1
    eta2, v2 = _inertial_mean(orc, c, hm, 1.0e-4, [lambda: _grid(orc, c, orc.array("hlay")[1] - orc.array("h_0")[1])[:, ix_0],
                                                    lambda: _grid(orc, c, orc.array("v")[1])[:, ix_0]])
    te, tv = c.info["t_eta"][1], c.info["t_v"][1]
    # a 40-row basin of which 30 are sponge: within 10 % of the channel solution (1.31 m, 0.024 m/s)
    assert np.max(np.abs(eta2[1:9] - te[1:9])) < 0.13 and eta2[1] > 1.15
    assert np.max(np.abs(v2[1:20] - tv[1:20])) < 0.004


def test_alongshore_wind_upwelling_before_outcrop_matches_morel():
    """testcases/morel_upwelling.m:104-112: before the interface surfaces (t < t_o) the interface displacement and the
    two layers' along-shore currents follow Morel, Darr & Talandier (2006).  Periodic in x with lm = 1, wind as a body
    force on the top layer (bodf.bin), outcropping switched on."""
    c = cases.morel_upwelling()
    hm, orc = make(c)
    I = c.info
    n1 = int(0.6 * I["t_o"] / hm.params.dt)
    orc.advance(1, n1)
    tt = n1 * hm.params.dt
    U_c = I["T_w"] * tt / (1.0 + I["delt"])
    U_b = U_c * I["delt"]
    y = -((np.arange(c.mm + 2, 0, -1) - 1.5) * I["dl"])  # yy_r of the script: distance from the coast, negative
    e = np.exp(y / I["R_d"])
    eta2 = _grid(orc, c, orc.array("hlay")[1] - orc.array("h_0")[1])[1, :]
    u1 = _grid(orc, c, orc.array("u")[0])[1, :]
    u2 = _grid(orc, c, orc.array("u")[1])[1, :]
    J = slice(c.mm - 40, c.mm + 1)  # the 40 km next to the coast (R_d = 13 km)
    assert np.max(np.abs(eta2[J] - I["H_1"] * U_c * e[J] / I["f0"] / I["R_d"])) < 0.07 * I["H_1"] * U_c / I["f0"] / I["R_d"]
    assert np.max(np.abs(u1[J] - (U_c * e[J] + U_b))) < 0.1 * (U_c + U_b)
    assert np.max(np.abs(u2[J] - (-U_c * e[J] * I["delt"] + U_b))) < 0.1 * (U_c + U_b)


def test_outcropped_state_of_rest_stays_at_rest():
    """testcases/outcrop_seamount.m:1-3: 'verifies that the initial state corresponds to the state of rest' -- five
    layers over a seamount and sloping coasts, most isopycnals grounded (Salmon 2002); h_0 comes from the Newton solve
    of private_mod.f95:309-502."""
    c = cases.outcrop_seamount()
    hm, orc = make(c)
    nstp, _, _ = orc.counts()
    h0 = orc.array("hlay").copy()
    assert np.count_nonzero(h0[:, 1:] < 2.0 * c.info["hsal"]) > 0.5 * h0[:, 1:].size  # most layer cells are outcropped
    orc.advance(1, nstp)  # the script's 15 days
    assert np.max(np.abs(orc.array("u"))) < 1.0e-4 and np.max(np.abs(orc.array("v"))) < 1.0e-4  # m/s
    assert np.max(np.abs(orc.array("hlay") - h0)) < 1.0e-3                                        # metres, of 300


def test_wave_sponge_lets_the_gravity_waves_out():
    """testcases/wave_sponge.m:139-205: potential + kinetic energy of a collapsing mound in a basin with flow-relaxation
    sponges on four sides: once the barotropic waves reach the sponges the energy drops to the few per cent held by the
    geostrophically adjusted remainder, and does not come back (no reflection)."""
    c = cases.wave_sponge()
    hm, orc = make(c)
    nstp, _, _ = orc.counts()
    rhon = c.info["rhon"]

    def energy():
        h, h0, u, v = orc.array("hlay"), orc.array("h_0"), orc.array("u"), orc.array("v")
        pe = 0.5 * rhon[0] * 9.8 * np.sum((h - h0).sum(axis=0) ** 2)
        return pe + 0.5 * sum(rhon[k] * np.sum((u[k] ** 2 + v[k] ** 2) * h[k]) for k in range(2))

    e0 = energy()
    hist = []
    for k in range(0, nstp - 19, 20):
        orc.advance(k + 1, k + 20)
        hist.append(energy())
    assert 0.95 < hist[1] / e0 < 1.05      # nothing is lost while the waves cross the interior
    assert hist[-1] < 0.05 * e0            # and almost everything once they have left
    assert max(hist[6:]) < 0.06 * e0       # no return


def test_tidal_target_drives_the_sponge_current():
    """testcases/tide_ridge.m:73-92 / private_mod.f95:951-964, 1453-1454: element (1,k,0,0,1) of tide.bin is the
    constituent frequency in rad/day; deep inside the western sponge (relaxation coefficient 0.97 per step) every layer's
    current is the ramped target ramp * 0.1 cos(pi/2 - omega t)."""
    c = cases.tide_ridge(lm=200)
    hm, orc = make(c)
    nstp, _, _ = orc.counts()
    assert c.nlay == 7 and hm.params.ocrp > 0.5
    n1 = nstp // 6
    orc.advance(1, n1)
    ctim = n1 * hm.params.dt / 86400.0
    ramp = min(1.0, ctim / hm.params.dt_r)
    target = ramp * 0.1 * np.cos(np.pi / 2.0 - c.info["omega"] * ctim)
    assert abs(target) > 0.004
    for k in range(c.nlay):
        assert abs(_grid(orc, c, orc.array("u")[k])[1, 1] / target - 1.0) < 1.0e-3, k
    assert abs(orc.scalar("ramp") - ramp) < 1e-9


@pytest.mark.parametrize("name", ["sill_exchange2D", "sill_exchange2Dtides"])
def test_two_dimensional_sill_cases_run_and_keep_their_layers(name):
    """testcases/sill_exchange2D.m, sill_exchange2Dtides.m have no analytical reference: the guard of
    private_mod.f95:2798-2808 (no layer thinner than hmin / 2) and the exchange itself (dense water moving down the sill)."""
    from tests.conftest import SMALL
    c = cases.CASES[name](**SMALL[name])
    hm, orc = make(c)
    h_start = orc.array("hlay").copy()
    orc.advance(1, 4000)
    h = orc.array("hlay")
    wet = orc.array("mk_n")[0] > 0.5
    assert np.all(np.isfinite(h)) and h[:, wet].min() > 0.5 * hm.params.hmin
    assert np.max(np.abs(orc.array("u"))) > 1.0e-3 and np.max(np.abs(h - h_start)) > 0.1
    assert np.allclose(h[:, wet].sum(axis=0), h_start[:, wet].sum(axis=0), atol=0.5)  # the free surface barely moves


@pytest.mark.parametrize("name", ["stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D", "conservation", "soliton",
                                  "baines_ridge", "carrier_beach", "upwelling_seaward_wind", "mixed_open_bc", "morel_upwelling",
                                  "outcrop_seamount", "sill_exchange2D", "sill_exchange2Dtides", "tide_ridge", "wave_sponge"])
def test_oracle_reproduces_the_reference_vectors(name):
    """tests/golden/*.npz are outputs of the reference itself (its own sources translated to C++ by oracle/f95c and run in
    the development container, tests/golden/make_golden.py): the double-precision state after N steps and the last
    record of the eta_.bin it wrote.  The oracle must reproduce them bit for bit -- this runs wherever the fixtures are,
    /root/reference or not."""
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
