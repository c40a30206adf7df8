"""Host logic that needs no GPU: the shared_mod.f95 parser (Fortran literal semantics), the numpy
restatements of the reference's case generators, and the host read_input_data against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from beom_b200 import cases, model

f32 = lambda x: float(np.float32(x))  # noqa: E731


def test_default_real_literals_are_rounded_through_float32():
    text = """
    lm = 10
    mm = 5
    nlay = 2
    ndeg = 66
    dl = 400.
    cext = 82.8
    f0 = 1.409e-4
    rhon(nlay) = (/1027.47,1027.75/)
    topl(nlay) = (/0.000000,0.142800/)
    hmin = 0.500000
    dvis = 0.9_rw
    bdrg = 2.000000e-04
    tauw = (0.10,0.0)
    idir = '/tmp/in/'
    """
    p, idir, odir, desc = model.parse_params(text)
    assert (p.lm, p.mm, p.nlay, p.ndeg) == (10, 5, 2, 66)
    assert p.cext == f32(82.8) and p.cext != 82.8
    assert p.f0 == f32(1.409e-4)
    assert p.rhon[0] == f32(1027.47) and p.rhon[1] == f32(1027.75)
    assert p.topl[1] == f32(0.1428)
    assert p.dvis == 0.9  # kind suffix: a true double
    assert p.bdrg == f32(2.0e-4)
    assert p.tauw[0] == f32(0.10) and p.tauw[1] == 0.0
    assert idir == "/tmp/in/"
    # derived constants, shared_mod.f95:83-99
    assert p.grav == f32(9.8) and p.pi == f32(3.1415927) and p.hdry == f32(1.0e-3) and p.sor == f32(1.9)
    assert p.dt == 0.5 * 400.0 / f32(82.8)
    assert p.hsal == 10.0 * f32(0.5)
    assert p.rho0 == p.rhon[1]
    assert p.del1 == 0.5 + f32(0.088) + 2.0 * f32(0.013)
    assert p.del2 == 1.0 - p.del1 - f32(0.088) - f32(0.013)


def test_mixed_kind_arithmetic_and_integer_division():
    p, *_ = model.parse_params("lm=4\nmm=4\nnlay=1\nndeg=1\nrhon(nlay)=(/1000./)\ntopl(nlay)=(/0./)\n"
                               "dl = 2*500\ncext = 0.1*3.\nhmin = 1/2 + 0.25\ndvis = 0.1_r8*3.\n")
    assert p.dl == 1000.0
    assert p.cext == float(np.float32(0.1) * np.float32(3.0))  # r4 * r4 evaluated in float32
    assert p.hmin == 0.25  # integer division 1/2 = 0
    assert p.dvis == 0.1 * float(np.float32(3.0))  # r8 * r4 -> double


@pytest.mark.skipif(not os.path.exists("/root/reference/shared_mod.f95"), reason="reference tree not present on this machine")
def test_parses_the_reference_shared_mod_verbatim():
    with open("/root/reference/shared_mod.f95") as f:
        p, idir, odir, desc = model.parse_params(f.read())
    assert (p.lm, p.mm, p.nlay, p.ndeg) == (125, 501, 2, 63252)
    assert p.dl == 400.0 and p.cext == f32(82.8) and p.f0 == f32(1.409e-4)
    assert p.hmin == f32(0.3) and p.ocrp == 1.0 and p.mcbc == 1.0 and p.rgld == 0.0
    assert p.itmx == 99999 and p.nsal == 4
    assert desc == "Test-case: 3D sill exchange"
    assert p.dt == 0.5 * 400.0 / f32(82.8)


def test_print_params_reproduces_the_script_rounding():
    c = cases.stommel1948()
    t = c.params_text
    assert "lm         = 100\n" in t and "mm         = 63\n" in t and "ndeg       = 6464\n" in t
    assert "dl         = 100.e3\n" in t
    assert "cext       = 44.3\n" in t
    assert "f0         = 0.000000e+00\n" in t
    assert "rhon(nlay) = (/1027./)\n" in t
    assert "bdrg       = 2.000000e-04\n" in t
    assert "dt_s       = 40.000000\n" in t and "dt_r       = 0.\n" in t
    assert "tauw       = (0.00,0.00)\n" in t
    s = cases.sill_exchange3D().params_text
    assert "f0         = 1.409e-4\n" in s and "rhon(nlay) = (/1027.470,1027.750/)\n" in s
    assert "hmin       = 0.500000\n" in s and "dvis       = 0.900\n" in s and "dt_o       = 0.010000\n" in s


@pytest.mark.parametrize("name,lm,mm,nlay,ndeg", [("stommel1948", 100, 63, 1, 6464), ("lock_exchange", 160, 1, 2, 322),
                                                   ("unstable_jet", 201, 267, 1, 54136), ("sill_exchange3D", 125, 501, 2, 63252),
                                                   ("conservation", 61, 61, 2, 3844), ("soliton", 307, 153, 1, 47432),
                                                   ("baines_ridge", 751, 1, 2, 1504), ("carrier_beach", 515, 1, 1, 1032),
                                                   ("upwelling_seaward_wind", 200, 1, 2, 402), ("mixed_open_bc", 200, 100, 2, 20301),
                                                   ("morel_upwelling", 1, 221, 2, 444), ("outcrop_seamount", 121, 1, 5, 244),
                                                   ("sill_exchange2D", 2001, 1, 2, 4004), ("sill_exchange2Dtides", 1001, 1, 2, 2004),
                                                   ("tide_ridge", 501, 1, 7, 1004), ("wave_sponge", 91, 91, 2, 8464)])
def test_case_sizes_match_the_reference_scripts(name, lm, mm, nlay, ndeg):
    """lm, mm, nlay as the scripts compute them; ndeg from get_nbr_deg_freedom.m -- and the reference's own check
    (i_c == ndeg, private_mod.f95:604-610, restated by the host driver) accepts it."""
    c = cases.CASES[name]()
    assert (c.lm, c.mm, c.nlay, c.ndeg) == (lm, mm, nlay, ndeg)
    for key, arr in c.files.items():
        if key == "bodf":
            assert arr.shape == (nlay, 2)
        elif key == "tide":
            assert arr.shape == (2, 1, lm + 2, mm + 2, 3)
        else:
            assert arr.shape[:2] == (lm + 2, mm + 2)


def test_input_files_are_column_major_float32(tmp_path):
    c = cases.lock_exchange()
    c.write(str(tmp_path))
    raw = np.fromfile(tmp_path / "init.bin", dtype="<f4")
    assert raw.size == (c.lm + 2) * (c.mm + 2) * c.nlay * 3
    a = raw.reshape(3, c.nlay, c.mm + 2, c.lm + 2)  # Fortran order (i, j, k, c) read back in C order
    assert a[0, 1, 1, 0] == np.float32(0.5 * 20.0 - 4.0 * 0.05)   # eta of layer 2, west half
    assert a[0, 1, 1, -1] == np.float32(-0.5 * 20.0 + 4.0 * 0.05)  # east half
    assert np.all(a[1:] == 0)


@pytest.mark.parametrize("name", ["stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D", "conservation", "soliton",
                                  "baines_ridge", "carrier_beach", "upwelling_seaward_wind", "mixed_open_bc", "morel_upwelling",
                                  "outcrop_seamount", "sill_exchange2D", "sill_exchange2Dtides", "tide_ridge", "wave_sponge"])
def test_host_read_input_data_equals_oracle(case_factory, name):
    """Two independent restatements of read_input_data (C++ host driver, C oracle) agree bit for bit:
    connectivity, masks, rest thickness (Newton solve when ocrp = 1), sponge, segments, forcing."""
    from oracle.pyoracle import Oracle
    c, d, hm = case_factory(name)
    orc = Oracle(hm.params, d)
    for nm in ("neig", "subc"):
        assert np.array_equal(hm.iarray(nm), orc.iarray(nm)), nm
    assert hm.scalar("nseg") == orc.nseg()
    if orc.nseg():
        assert np.array_equal(hm.iarray("segm"), orc.iarray("segm"))
    for nm in ("mk_u", "mk_v", "mk_n", "mkpe", "mkpi", "fcor", "h_th", "nudg", "fnud", "hdot", "taus", "h_0", "hlay", "u", "v"):
        a, b = hm.array(nm), orc.array(nm)
        if a is None:
            assert not np.any(b), nm + " absent on the host but non-zero in the oracle"
        else:
            assert np.array_equal(a, b.reshape(a.shape)), nm
    assert hm.scalar("invf") == orc.scalar("invf")
    assert hm.counts() == orc.counts()
    assert bool(hm.scalar("flag_nudging")) == bool(orc.scalar("flag_nudging"))


def test_step_counts_of_the_named_configs():
    """nstp of SURVEY section 8 (S 3062, L 30240, J 4205, X 1073088): pins dt and the float32 literals."""
    want = {"stommel1948": 3062, "lock_exchange": 30240, "unstable_jet": 4205, "sill_exchange3D": 1073088}
    for name, nstp in want.items():
        c = cases.CASES[name]()
        p, *_ = model.parse_params(c.params_text)
        dtd8 = p.dt / 24.0 / 3600.0
        assert int(np.floor(p.dt_s / dtd8 + 0.5)) == nstp, name


def test_wrong_ndeg_is_reported_like_the_reference(tmp_path):
    c = cases.lock_exchange()
    c.params_text = c.params_text.replace("ndeg       = 322", "ndeg       = 321")
    blk = c.write(str(tmp_path))
    with pytest.raises(RuntimeError, match="Please set ndeg = 322"):
        model.HostModel.from_block(blk)
