"""beom_gpu_init_grids: read_input_data's grid-shaped work on the device (index_grid_points as a prefix sum, the rest
thickness incl. the per-column Newton solve of get_equilibrium_thickness_h_0, the forcing files; private_mod.f95:105-250,
309-502, 567-764, 840-964) against the host restatement (csrc/host/init.cc, itself equal to the oracle bit for bit): every
mask, the vector numbering, every static plane, the initial state -- and the state after a few steps -- must be identical.
Cases the device path does not cover must be refused with BEOM_GRIDS_UNSUPPORTED, not half done."""
import numpy as np
import pytest

from beom_b200 import cases, model

pytestmark = pytest.mark.gpu

FLAG_BITS = {"mk_n": 1, "mk_u": 2, "mk_v": 4, "mkpe": 8, "mkpi": 16}

COVERED = [
    ("synthetic_basin", dict(n=130, mm=70, nlay=4), {}),                   # the bench workload: flat basin, init.bin, taus.bin
    ("synthetic_basin", dict(n=90, mm=200, nlay=4, sponge=True), {}),      # bench --workload sill_like: nudg.bin + the Newton solve
    ("stommel1948", None, {}),                                             # fcor.bin (float32 psi-point average), taus.bin, one layer
    ("lock_exchange", None, {}),                                           # two layers, init.bin only
    ("sill_exchange3D", None, {}),                                         # h_bo.bin with land, sponges (mcbc = 1), outcropping
    ("sill_exchange2D", None, {}),
    ("carrier_beach", None, {}),                                           # one outcropping layer, sloping beach
    ("outcrop_seamount", None, {}),                                        # five layers, most of them grounded
    ("wave_sponge", None, {}),                                             # sponges on four sides (mcbc = 1)
    ("sill_exchange2Dtides", None, {}),                                    # tide.bin: amplitude / phase planes, w_ti
    ("tide_ridge", None, {}),                                              # seven layers, tide.bin, dt_r ramp
    ("random_coast", None, {}),                                            # random coastline: every mask combination
    ("conservation", None, {}),                                            # doubly periodic torus: aliases, images, displaced duplicates
    ("unstable_jet", None, {}),                                            # doubly periodic, one layer
    ("soliton", None, {}),                                                 # periodic in x only, fcor.bin
    ("baines_ridge", None, {}),                                            # one-row channel periodic in y, bodf.bin, nudged duplicates
    ("upwelling_seaward_wind", None, {}),                                  # periodic in y, no input file at all, dt_r ramp
    ("morel_upwelling", None, {}),                                         # lm = 1 periodic in x, outcropping, body force
    ("mixed_open_bc", None, {}),                                           # mcbc = 0: open-boundary segments on three sides
    ("baines_ridge", None, {"mcbc": "0."}),                                # ... and at the two open ends of the periodic channel
    ("sill_exchange3D", None, {"mcbc": "0."}),
    ("option_basin", dict(hdot=True, sponge=False, wind=True), {}),        # hdot.bin, island
    ("option_basin", dict(bodf=True, sponge=False, wind=False), {}),       # bodf.bin
]
REFUSED = [("lock_exchange", None, {"rgld": "1."})]


def _ids(rows):
    return ["%s%s%s" % (n, "-" + "-".join("%s" % v for v in (k or {}).values()) if k else "", "-" + "".join(e) if e else "") for n, k, e in rows]


@pytest.mark.parametrize("name,kwargs,extra", COVERED, ids=_ids(COVERED))
def test_device_side_initialisation_equals_read_input_data(case_factory, name, kwargs, extra):
    c, d, hm = case_factory(name, small=kwargs is None, extra=extra, **(kwargs or {}))
    gd = model.GpuModel.from_grids(hm.params, d, model.default_options(fused=True))
    assert gd is not None, "the device path should cover %s" % name
    nlay = hm.params.nlay
    # the vector numbering: grid coordinates of every point
    si, sj = gd.download_subc()
    sub = hm.iarray("subc")
    held = si > 0  # (a displaced periodic duplicate has no cell: 0, 0)
    assert np.array_equal(si[held], sub[0][1:][held]) and np.array_equal(sj[held], sub[1][1:][held])
    # the masks
    fl = gd.debug_static("flags")[1:].astype(np.int64)
    dup = np.zeros(c.ndeg, dtype=bool)  # displaced periodic duplicates have no cell on the device (their masks are 0, their state is
    if hm.params.xper > 0.5:            # kept on the host): the planes are compared at the other points, the state everywhere
        dup |= (sub[0] == c.lm + 1)[1:]
    if hm.params.yper > 0.5:
        dup |= (sub[1] == c.mm + 1)[1:]
    own = ~dup
    assert np.all(fl[own] & 32)
    for k, bit in FLAG_BITS.items():
        assert np.array_equal(((fl & bit) != 0)[own], (hm.array(k)[0][1:] > 0.5)[own]), k
    # statics
    for k, planes in (("fcor", 1), ("h_th", 1), ("h_0", nlay)):
        want = hm.array(k)
        for q in range(planes):
            assert np.array_equal(gd.debug_static(k, q)[1:][own], want[q][1:][own]), (k, q)
    fld = hm.fields()
    if fld.nudg and np.any(hm.array("nudg") != 0.0):
        for q in range(3):
            assert np.array_equal(gd.debug_static("nudg", q)[1:][own], hm.array("nudg")[q][1:][own]), ("nudg", q)
        want = hm.array("fnud")
        for q in range(3 * nlay):
            assert np.array_equal(gd.debug_static("fnud", q)[1:][own], want.reshape(3 * nlay, -1)[q][1:][own]), ("fnud", q)
    if np.any(np.abs(hm.array("taus")[:, 1:]) > 1e-7):
        for q in range(2):
            assert np.array_equal(gd.debug_static("taus", q)[1:][own], hm.array("taus")[q][1:][own]), ("taus", q)
    # the initial state, then a few steps on both
    st0 = gd.download_state()
    for nm, a in zip(("hlay", "u", "v"), st0):
        assert np.array_equal(a[:, 1:], hm.array(nm)[:, 1:]), nm
    gd.advance(1, 12)
    got = gd.download_state()
    variant = gd.fused_variant
    gd.close()
    gh = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    gh.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gh.advance(1, 12)
    want = gh.download_state()
    assert gh.fused_variant == variant
    gh.close()
    for nm, a, b in zip(("hlay", "u", "v"), got, want):
        assert np.array_equal(a, b), nm  # (tides: the same libm cos on both sides, so also exact)
    assert np.abs(got[1]).max() + np.abs(got[2]).max() > 0 or name in ("carrier_beach",)


@pytest.mark.parametrize("name,kwargs,extra", REFUSED, ids=_ids(REFUSED))
def test_what_the_device_path_does_not_cover_is_refused(case_factory, name, kwargs, extra):
    c, d, hm = case_factory(name, small=True, extra=extra, **(kwargs or {}))
    assert model.GpuModel.from_grids(hm.params, d, model.default_options(fused=True)) is None
    gh = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))  # the host path still serves it
    gh.close()
