"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar: BIT-EXACT doubles.  The kernels are built -fmad=false and evaluate the reference's expressions in
the reference's order; the oracle is built -ffp-contract=off.  (cos() would differ in the last ulp, it
is only evaluated when tide.bin exists.)"""
import numpy as np
import pytest

from beom_b200 import model
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


def run_pair(case_factory, name, nsteps, fused, small=True, variant=0, extra=None, **kw):
    c, d, hm = case_factory(name, small=small, variant=variant, extra=extra, **kw)
    orc = Oracle(hm.params, d)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=fused))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, nsteps)
    orc.advance(1, nsteps)
    hl, u, v = gm.download_state()
    aux = gm.download_aux()
    path = gm.path
    gm.close()
    return c, hm, orc, (hl, u, v), aux, path


def assert_same(name, got, want, mask=None):
    want = want.reshape(got.shape)
    if mask is not None:
        got, want = got[..., mask], want[..., mask]
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        k = tuple(bad[0])
        rel = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
        raise AssertionError("%s differs at %d entries, first %s: got %r want %r (max rel %.3e)" %
                             (name, len(bad), k, got[k], want[k], rel))


@pytest.mark.parametrize("name,nsteps", [("stommel1948", 40), ("lock_exchange", 60), ("unstable_jet", 30),
                                         ("sill_exchange3D", 30), ("conservation", 30)])
@pytest.mark.parametrize("fused", [False, True])
def test_bit_exact_small(case_factory, name, nsteps, fused):
    c, hm, orc, (hl, u, v), aux, path = run_pair(case_factory, name, nsteps, fused)
    # the fused step covers closed / sponge domains and periodic ones whose aliases form a complete torus (deep
    # ghost cells): all five cases
    assert path == ("fused" if fused else "split")
    assert_same("hlay", hl, orc.array("hlay"))
    assert_same("u", u, orc.array("u"))
    assert_same("v", v, orc.array("v"))
    # frozen periodic duplicates carry no meaningful fluxes; compare the rest of the auxiliary state
    own = np.ones(c.ndeg + 1, dtype=bool)
    if hm.params.xper > 0.5 or hm.params.yper > 0.5:
        sub = hm.iarray("subc")
        own &= ~((sub[0] == c.lm + 1) | (sub[1] == c.mm + 1))
    h_u, h_v, rs_h, dmdx, dmdy = aux
    assert_same("h_u", h_u, orc.array("h_u"), own)
    assert_same("h_v", h_v, orc.array("h_v"), own)
    assert_same("rs_h", rs_h.transpose(0, 2, 1), orc.array("rs_h").transpose(0, 2, 1), own)
    assert_same("dmdx", dmdx.transpose(0, 2, 1), orc.array("dmdx").transpose(0, 2, 1), own)
    assert_same("dmdy", dmdy.transpose(0, 2, 1), orc.array("dmdy").transpose(0, 2, 1), own)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("nlay", [1, 3, 4])
def test_bit_exact_synthetic_basin(case_factory, fused, nlay):
    """The bench workload at a size the oracle finishes in seconds: several x strips and y chunks of
    the fused kernel, wind stress, Leith viscosity, generalized forward-backward (both u/v orders)."""
    c, hm, orc, (hl, u, v), aux, path = run_pair(case_factory, "synthetic_basin", 17, fused, small=False, n=300, mm=170, nlay=nlay)
    assert path == ("fused" if fused else "split")
    assert_same("hlay", hl, orc.array("hlay"))
    assert_same("u", u, orc.array("u"))
    assert_same("v", v, orc.array("v"))
    h_u, h_v, rs_h, dmdx, dmdy = aux
    assert_same("h_u", h_u, orc.array("h_u"))
    assert_same("h_v", h_v, orc.array("h_v"))
    assert_same("rs_h", rs_h, orc.array("rs_h"))
    assert_same("dmdx", dmdx, orc.array("dmdx"))
    assert_same("dmdy", dmdy, orc.array("dmdy"))


@pytest.mark.parametrize("name", ["stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D", "conservation", "soliton",
                                  "baines_ridge", "carrier_beach", "upwelling_seaward_wind", "mixed_open_bc", "morel_upwelling",
                                  "outcrop_seamount", "sill_exchange2D", "sill_exchange2Dtides", "tide_ridge", "wave_sponge"])
def test_golden_vectors(case_factory, name):
    """The CUDA path against outputs of THE REFERENCE ITSELF, without running any checker: tests/golden/*.npz is the state the
    reference's own sources (translated to C++ by oracle/f95c, tests/golden/make_golden.py) left after N steps of each of its
    sixteen test-case scripts.  Bit for bit -- except the two tidal cases, where the device's cos() may differ from glibc's
    in the last place (1e-11, as everywhere tides are compared)."""
    import os
    from tests.golden.make_golden import GOLDEN
    kw, nsteps = GOLDEN[name]
    want = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    c, d, hm = case_factory(name, small=False, **kw)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, nsteps)
    hl, u, v = gm.download_state()
    gm.close()
    if name in ("sill_exchange2Dtides", "tide_ridge"):
        for k, got in (("hlay", hl), ("u", u), ("v", v)):
            assert np.max(np.abs(got - want[k])) <= 1e-11 * max(1.0, np.max(np.abs(want[k]))), k
        return
    assert_same("hlay", hl, want["hlay"])
    assert_same("u", u, want["u"])
    assert_same("v", v, want["v"])


def test_fused_equals_split_at_scale_and_conserves_volume(case_factory):
    """A grid far too large for the oracle in a test (2048 x 1024 x 4 = 8.4M cell-layers): the two GPU
    paths agree bit for bit, and the closed basin keeps its volume (size-independent property)."""
    c, d, hm = case_factory("synthetic_basin", small=False, n=2048, mm=1024, nlay=4)
    res = {}
    for fused in (True, False):
        gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=fused))
        gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
        gm.advance(1, 24)
        res[fused] = gm.download_state() + gm.download_aux()
        assert gm.path == ("fused" if fused else "split")
        gm.close()
    for a, b, nm in zip(res[True], res[False], ("hlay", "u", "v", "h_u", "h_v", "rs_h", "dmdx", "dmdy")):
        assert_same(nm, a, b)
    wet = hm.array("mk_n")[0] > 0.5
    v0 = hm.array("hlay")[:, wet].sum(axis=1)
    v1 = res[True][0][:, wet].sum(axis=1)
    assert np.all(np.abs(v1 - v0) <= 1e-12 * np.abs(v0))
    assert np.all(np.isfinite(res[True][1])) and np.abs(res[True][1]).max() > 0


def _check_state(tag, got, orc, names=("hlay", "u", "v")):
    for nm, a in zip(names, got):
        assert_same(tag + ":" + nm, a, orc.array(nm))


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("extra", [
    dict(mcbc="0."),                                   # a8: no_gradient_obc active on the sponge boundaries
    dict(bdrg="3.e-3", qdrg="1."),                     # a7: quadratic bottom drag with outcropping (layb)
    dict(bdrg="1.e-4", qdrg="0.", tdrg="2.e-4"),       # a7: linear bottom + top drag (layu)
    dict(dt3d="0.001"),                                # stress / viscosity refreshed every n_3d > 1 steps
    dict(dt3d="0.001", dvis="0.", bvis="30."),         # n_3d > 1 with a constant viscosity: nothing depends on upst, fused
], ids=["obc", "quad_drag", "top_drag", "n3d", "n3d_const_visc"])
def test_sill_options_bit_exact(case_factory, fused, extra):
    c, hm, orc, st, aux, path = run_pair(case_factory, "sill_exchange3D", 40, fused, extra=extra)
    if "dt3d" in extra:
        assert orc.counts()[2] > 1 and path == ("fused" if (fused and "bvis" in extra) else "split")
    _check_state("sill" + str(extra), st, orc)
    assert_same("h_u", aux[0], orc.array("h_u"))
    assert_same("h_v", aux[1], orc.array("h_v"))


@pytest.mark.parametrize("variant,nlay,plum", [(1, 2, "0."), (2, 3, "0."), (3, 3, "0."), (3, 3, "1.")],
                         ids=["1d", "3d", "plume_off", "plume_on"])
def test_update_h_variants_bit_exact(case_factory, variant, nlay, plum):
    """a1': the water-mass-transformation epilogues of private_mod1d/3d/plumenew.f95 (split path)."""
    c, hm, orc, st, aux, path = run_pair(case_factory, "sponge_basin", 60, True, small=False, variant=variant,
                                         extra=dict(plum=plum), nlay=nlay)
    assert path == "split" and hm.params.variant == variant
    _check_state("variant%d" % variant, st, orc)
    std = run_pair(case_factory, "sponge_basin", 60, False, small=False, variant=0, nlay=nlay)
    assert not np.array_equal(std[3][0], st[0])  # the epilogue does change the answer


@pytest.mark.parametrize("name,extra", [("lock_exchange", dict(svis="50.")), ("sponge_basin", dict(svis="200.", bvis="1.0"))])
def test_biharmonic_viscosity_bit_exact(case_factory, name, extra):
    """a3 with svis > 0: masked Laplacians + thickness-weighted biharmonic fluxes (pm:2508-2599, 1471-1473)."""
    c, hm, orc, st, aux, path = run_pair(case_factory, name, 40, True, small=(name != "sponge_basin"), extra=extra)
    assert path == "split" and hm.params.svis > 0
    _check_state("svis", st, orc)


@pytest.mark.parametrize("name,kw,extra", [("lock_exchange", {}, dict(rgld="1.")),
                                           ("synthetic_basin", dict(n=48, mm=30, nlay=2), dict(rgld="1.", ocrp="1."))],
                         ids=["lock_exchange_rgld", "basin_rgld_ocrp"])
def test_rigid_lid_bit_exact(case_factory, name, kw, extra):
    """a9/a10: surf_pressure (hyperplane-ordered Gauss-Seidel = the reference's lexicographic sweep), the
    flux rebuilds and the single-precision column correction of update_h."""
    c, d, hm = case_factory(name, small=(name == "lock_exchange"), extra=extra, **kw)
    orc = Oracle(hm.params, d)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    assert gm.path == "split"
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, 25)
    orc.advance(1, 25)
    st = gm.download_state()
    pi_s = gm.download_pi_s()
    gm.close()
    _check_state("rgld", st, orc)
    assert_same("pi_s", pi_s, orc.array("pi_s")[0])
    if "ocrp" in extra:
        assert np.abs(pi_s).max() > 0


def test_rigid_lid_wavefront_solver_many_tiles(case_factory):
    """surf_pressure on every SM (rigid.cuh): 300 x 170 points = 5 x 6 tiles of 64 x 32, up to 15 sweeps in flight -- pi_s, the
    state and the number of sweeps equal the oracle's lexicographic Gauss-Seidel (pm:1756-1803) bit for bit."""
    c, d, hm = case_factory("synthetic_basin", small=False, extra=dict(rgld="1.", ocrp="1."), n=300, mm=170, nlay=2)
    orc = Oracle(hm.params, d)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    iters = []
    for t in range(1, 7):
        gm.advance(t, t)
        iters.append(gm.pi_iterations())
    orc.advance(1, 6)
    st = gm.download_state()
    pi_s = gm.download_pi_s()
    gm.close()
    _check_state("rgld_wave", st, orc)
    assert_same("pi_s", pi_s, orc.array("pi_s")[0])
    assert np.abs(pi_s).max() > 0 and all(1 <= k <= 1000 for k in iters) and max(iters) > 16  # (more sweeps than rotating arrays)


@pytest.mark.parametrize("nlay", [1, 4])
def test_fma_flavour_within_tolerance(nlay):
    """BEOM_FMA=1 (opt-in): the fused step compiled with FMA contraction, like the reference's own -Ofast build.
    Tolerance of BASELINE.json's north star: max relative field difference <= 1e-10 after N steps."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fma_worker.py"), str(nlay), "40"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fma max relative field difference" in r.stdout


OPTION_MATRIX = {
    "ekman_sponge": dict(),                      # wind inside a sponge: the Ekman term of the relaxation target
    "ramp": dict(dt_r=0.2),                      # dt_r ramp of the forcing (ramp < 1 during the whole test)
    "bodf": dict(bodf=True),                     # body force per layer
    "hdot": dict(hdot=True),                     # thickness source
    "beta": dict(beta=True),                     # fcor.bin (psi-point averaging in float32)
    "six_layers": dict(nlay=6),                  # the any-layer-count instantiation
    "outcrop_wind": dict(ocrp=1.0),              # wind stress spread over outcropping layers (every layer stages tt3d)
    "no_sponge": dict(sponge=False),             # wind only: the specialised instantiation with an island inside the tiles
    "no_wind": dict(wind=False),
    "tide": dict(tide=True),                     # tidal targets (cos per point and step): split path
}


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("opt", sorted(OPTION_MATRIX))
def test_option_matrix_bit_exact(case_factory, opt, fused):
    """cases.option_basin: the forcing files and switches no named config reaches, each against the oracle."""
    c, hm, orc, st, aux, path = run_pair(case_factory, "option_basin", 30, fused, small=False, **OPTION_MATRIX[opt])
    assert path == ("fused" if fused and opt != "tide" else "split")
    if opt == "tide":
        # cos() is the one libm-dependent operation of the path (CUDA's and glibc's differ in the last ulp, DESIGN.md
        # section 3): tolerance 1e-11 relative to the field's magnitude instead of bit-exact
        for nm, a in zip(("hlay", "u", "v"), st):
            w = orc.array(nm).reshape(a.shape)
            assert np.abs(a - w).max() <= 1.0e-11 * np.abs(w).max(), nm
        return
    _check_state("option_basin[%s]" % opt, st, orc)
    assert_same("h_u", aux[0], orc.array("h_u"))
    assert_same("h_v", aux[1], orc.array("h_v"))
    assert_same("rs_h", aux[2], orc.array("rs_h"))
    assert_same("dmdx", aux[3], orc.array("dmdx"))
    assert_same("dmdy", aux[4], orc.array("dmdy"))
