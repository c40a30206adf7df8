"""The oracle against THE REFERENCE ITSELF, bit for bit (CPU only; needs /root/reference, i.e. the development container).

No Fortran compiler exists in this image or on the GPU boxes, so oracle/f95c translates the reference's own sources
-- shared_mod.f95 with the case's parameter block in its user section, private_mod.f95 (or private_mod1d / 3d /
plumenew.f95), main.f95, read where they lie -- statement by statement into C++ (the translator knows the language, not the
model), g++ compiles it in strict IEEE mode into oracle/_ref/, and the program runs the reference's ``run()``:
read_input_data, integrate_time, its own output files.  At STOP the harness dumps every module array in raw form.

Each test runs that program and the hand-written oracle (oracle/beom_oracle.c) for the same number of steps on the same
input files and requires EVERY module array of private_mod.f95 to be identical bit for bit: the state (hlay, u, v), the
fluxes and histories (h_u, h_v, rs_h, dmdx, dmdy), the scratch fields of the last step (mont, rvor, pvor, dive, d2hx,
d2hy, v_cc, v_ll, UU4, VV4, delu, delv, tt3d, tb3d, tu3d), the masks and the neighbour table of index_grid_points, the
rest thickness, the forcing of read_input_file (nudg, fnud, taus, hdot, bodf, tide, fcor), the open-boundary segment table,
the rigid-lid operators, ctim / ramp / gene / invf -- and the files the reference wrote (eta_, u___, v___ records,
the diag records, h_0.bin, grid.bin) against the oracle's records.

Covered: all sixteen reference scripts, the option switches no script turns on (mcbc = 0, quadratic / linear / top drag,
dt3d > dt, constant and biharmonic viscosity, diag = 1, the rigid lid with and without outcropping, ramp, body force,
thickness source, beta plane, six layers, tides), a random coastline, and the three variant files."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refbuild, refcheck  # noqa: E402

pytestmark = pytest.mark.skipif(not refbuild.reference_available(), reason="needs the reference's sources (/root/reference)")

# (id, generator, kwargs or None = tests/conftest.SMALL, steps, extra parameter lines, variant)
PAIRS = [(n, n, None, 48, {}, 0) for n in (
    "stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D", "conservation", "soliton", "baines_ridge",
    "carrier_beach", "upwelling_seaward_wind", "mixed_open_bc", "morel_upwelling", "outcrop_seamount", "sill_exchange2D",
    "sill_exchange2Dtides", "tide_ridge", "wave_sponge", "random_coast")]
PAIRS += [
    ("sill-obc", "sill_exchange3D", None, 40, dict(mcbc="0."), 0),
    ("sill-quadratic-drag", "sill_exchange3D", None, 40, dict(bdrg="3.e-3", qdrg="1."), 0),
    ("sill-linear-and-top-drag", "sill_exchange3D", None, 40, dict(bdrg="1.e-4", qdrg="0.", tdrg="2.e-4"), 0),
    ("sill-n3d", "sill_exchange3D", None, 40, dict(dt3d="0.001"), 0),
    ("sill-n3d-constant-viscosity", "sill_exchange3D", None, 40, dict(dt3d="0.001", dvis="0.", bvis="30."), 0),
    ("sill-diag-records", "sill_exchange3D", None, 40, dict(diag="1."), 0),
    ("baines-obc", "baines_ridge", None, 40, dict(mcbc="0."), 0),
    ("lock-biharmonic", "lock_exchange", None, 40, dict(svis="50."), 0),
    ("sponge-biharmonic", "sponge_basin", {}, 40, dict(svis="200.", bvis="1.0"), 0),
    ("lock-rigid-lid", "lock_exchange", None, 40, dict(rgld="1."), 0),
    ("basin-rigid-lid-outcrop", "synthetic_basin", dict(n=48, mm=30, nlay=2), 25, dict(rgld="1.", ocrp="1."), 0),
    ("opt-ekman-sponge", "option_basin", {}, 40, {}, 0),
    ("opt-ramp", "option_basin", dict(dt_r=0.2), 40, {}, 0),
    ("opt-bodf", "option_basin", dict(bodf=True), 40, {}, 0),
    ("opt-hdot", "option_basin", dict(hdot=True), 40, {}, 0),
    ("opt-beta", "option_basin", dict(beta=True), 40, {}, 0),
    ("opt-six-layers", "option_basin", dict(nlay=6), 40, {}, 0),
    ("opt-outcrop-wind", "option_basin", dict(ocrp=1.0), 40, {}, 0),
    ("opt-tide", "option_basin", dict(tide=True), 40, {}, 0),
    ("variant-1d", "sponge_basin", dict(nlay=2), 60, dict(plum="0."), 1),
    ("variant-3d", "sponge_basin", dict(nlay=3), 60, dict(plum="0."), 2),
    ("variant-plume-off", "sponge_basin", dict(nlay=3), 60, dict(plum="0."), 3),
    ("variant-plume-on", "sponge_basin", dict(nlay=3), 60, dict(plum="1."), 3),
]
# the four named configs of BASELINE.json at the sizes their scripts ship with (ndeg 6 464 / 322 / 54 136 / 63 252), 200 steps:
# what tests/test_gpu_native_configs.py compares the CUDA path with
PAIRS += [("native-" + n, n, {}, 200, {}, 0) for n in ("stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D")]
IDS = [p[0] for p in PAIRS]


def _prepare(spec, root):
    """Write the inputs, set the step count, translate + compile the reference for this parameter block."""
    from beom_b200 import cases, model
    from tests.conftest import SMALL

    pid, gen, kw, nsteps, extra, variant = spec
    args = dict(SMALL.get(gen, {})) if kw is None else dict(kw)
    c = cases.CASES[gen](**args)
    c.params_text += "".join("%-10s = %s\n" % kv for kv in extra.items())
    d = os.path.join(root, pid)  # idir / odir are character(99) in shared_mod.f95: keep the path short
    blk = c.write(d)
    text = open(blk).read()
    p, _, _, _ = model.parse_params(text, variant)
    text = refcheck.block_for_steps(text, nsteps, p.dt)
    exe = refbuild.build_case(text, variant)
    # ... and run it here, next to the other cases' compiles and runs (the native-size sill run alone takes a minute)
    _, _, odir, _ = model.parse_params(text, variant)
    try:
        ran = refbuild.run_case(exe, odir, threads=1)
    except Exception as err:  # reported by the case's own test
        ran = err
    return text, exe, ran


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    """All translated references of this module, compiled and run side by side (g++ -O2 takes about six seconds per case)."""
    import tempfile

    root = tempfile.mkdtemp(prefix="rp", dir="/tmp")
    assert len(root) < 40
    out = {}
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        futs = {spec[0]: ex.submit(_prepare, spec, root) for spec in PAIRS}
        for pid, f in futs.items():
            try:
                out[pid] = f.result()
            except Exception as err:  # reported by the case's own test
                out[pid] = err
    return out


@pytest.mark.parametrize("spec", PAIRS, ids=IDS)
def test_oracle_equals_the_translated_reference(built, spec):
    from beom_b200 import model
    from oracle.pyoracle import Oracle

    pid, gen, kw, nsteps, extra, variant = spec
    got = built[pid]
    if isinstance(got, Exception):
        raise got
    text, exe, ran = got
    if isinstance(ran, Exception):
        raise ran
    p, idir, odir, _ = model.parse_params(text, variant)
    dump, _ = ran
    orc = Oracle(p, idir)
    orc.advance(1, nsteps)
    rows = refcheck.compare(dump, orc) + refcheck.compare_files(odir, orc, diag=p.diag > 0.5)
    orc.close()
    names = [r[0] for r in rows]
    for must in ("hlay", "u", "v", "h_u", "h_v", "rs_h", "dmdx", "dmdy", "mont", "pvor", "neig", "mk_u", "fnud", "eta_.bin"):
        assert must in names, "%s was not compared" % must
    bad = [r for r in rows if not r[1]]
    assert not bad, "%s: the oracle differs from the reference in\n%s" % (pid, refcheck.format_rows(bad))
    assert float(dump["ctim"]) > 0.0 and abs(dump["hlay"]).max() > 0.0  # the program really ran


def test_translator_subset_is_strict():
    """f95c stops on what it does not know instead of guessing."""
    sys.path.insert(0, os.path.join(os.path.dirname(refbuild.__file__), "f95c"))
    import f95c

    for bad in ("subroutine s()\n  goto 10\nend subroutine s\n",
                "subroutine s()\n  integer :: i\n  i = undeclared_thing + 1\nend subroutine s\n",
                "subroutine s()\n  real :: a(4)\n  a(1:4:2) = 0.\nend subroutine s\n"):
        with pytest.raises(f95c.FError):
            f95c.translate([("module m\ncontains\n" + bad + "end module m\nprogram p\nend program p\n", "t.f95", False)])


def test_translator_arithmetic_conventions(tmp_path):
    """Operator precedence, integer division, powers, single-precision literals and REAL() -- checked on a program
    whose results are known in closed form."""
    import subprocess
    sys.path.insert(0, os.path.join(os.path.dirname(refbuild.__file__), "f95c"))
    import f95c

    src = """module m
  implicit none
  integer, parameter :: r4 = selected_real_kind(6), r8 = selected_real_kind(12), rw = r8
  real(rw) :: a(0:3), b(4), c, d, e, f, g, h
  integer :: i, n, errc
  character(len=99) :: odir = './'
contains
subroutine s()
  implicit none
  a(:) = 2._rw
  a(1:2) = a(1:2) * 3._rw
  b(:) = a(:) - 1._rw
  c = sum(b(:)) / real(size(b(:)), rw)
  d = - a(0)**2 + 9/10 + 7/2
  e = 0.1
  f = real(0.1_rw)
  g = 2._rw**3 * 2._rw**(-1)
  h = maxval(b(:), mask = b(:) < 4._rw)
  n = 0
  do i = 10, 1, -3
    n = n + i
  end do
  n = n + i
  where (b(:) > 4._rw) b(:) = 0._rw
end subroutine s
end module m
program p
  use m
  call s()
end program p
"""
    cpp = tmp_path / "t.cpp"
    cpp.write_text(f95c.translate([(src, "t.f95", True)], dump_path_expr='std::string("%s/d.bin")' % tmp_path))
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-I",
                    os.path.join(os.path.dirname(refbuild.__file__), "f95c"), str(cpp), "-o", str(exe)], check=True)
    subprocess.run([str(exe)], check=True)
    d = refbuild.read_dump(str(tmp_path / "d.bin"))
    import numpy as np
    assert list(d["a"]) == [2.0, 6.0, 6.0, 2.0] and list(d["b"]) == [1.0, 0.0, 0.0, 1.0]
    assert d["c"] == 3.0                      # (1 + 5 + 5 + 1) / 4
    assert d["d"] == -4.0 + 0 + 3             # -(a**2), integer divisions
    assert d["e"] == float(np.float32(0.1)) and d["f"] == float(np.float32(0.1))
    assert d["g"] == 4.0 and d["h"] == 1.0
    assert d["n"] == 10 + 7 + 4 + 1 + (-2)    # the loop variable after the loop
