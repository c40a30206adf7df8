"""Output side of the boundary on the GPU: the float32 diagnostic records of write_array (pvor, mont, v_cc;
private_mod.f95:2884-2974) against the oracle's restatement (bit-exact float32), and the conservation integrals of
testcases/conservation.m:116-211 against a numpy restatement on the downloaded state (rtol 1e-12: a sum, so the
order of additions differs)."""
import numpy as np
import pytest

from beom_b200 import model
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,nsteps,fused", [("sill_exchange3D", 25, True), ("conservation", 20, False), ("unstable_jet", 15, False),
                                               ("lock_exchange", 30, True)])
def test_diagnostic_records_bit_exact(case_factory, name, nsteps, fused):
    c, d, hm = case_factory(name)
    orc = Oracle(hm.params, d)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=fused))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, nsteps)
    orc.advance(1, nsteps)
    got = gm.download_diag()
    gm.close()
    # the duplicate column / row of a periodic domain is frozen on the device; conservation.m
    # discards those points too ("do not account twice for points along boundary", :196-201)
    own = np.ones(c.ndeg, dtype=bool)
    if hm.params.xper > 0.5 or hm.params.yper > 0.5:
        sub = hm.iarray("subc")
        own &= ~(((sub[0] == c.lm + 1) & (hm.params.xper > 0.5)) | ((sub[1] == c.mm + 1) & (hm.params.yper > 0.5)))[1:]
    for var in ("pvor", "mont", "v_cc"):
        want = orc.record(var).reshape(got[var].shape)
        both_nan = np.isnan(got[var]) & np.isnan(want)
        same = (got[var] == want) | both_nan | ~own[None, :]
        assert same.all(), "%s record of %s differs at %d of %d entries (first %s: %r vs %r)" % (
            var, name, (~same).sum(), same.size, np.argwhere(~same)[0], got[var][~same][0], want[~same][0])
    assert np.abs(got["mont"][np.isfinite(got["mont"])]).max() > 0


def conservation_numpy(hm, hlay, u, v, h_0):
    """conservation.m:116-211 on the vector layout: wet thickness, kinetic energy proxy, sum of eta_1^2."""
    neig = hm.iarray("neig")  # neig(8, 0:ndeg): E=1, N=3, W=5, S=7 (private_mod.f95:28-31)
    if neig.shape[0] != 8:
        neig = neig.T
    wet = hm.array("mk_n")[0] > 0.5
    nlay = hlay.shape[0]
    E, N, W, S = neig[0], neig[2], neig[4], neig[6]
    vol, ke = np.zeros(nlay), np.zeros(nlay)
    sub = hm.iarray("subc")
    for l in range(nlay):
        hw = np.where(wet, hlay[l], 0.0)
        hw[0] = 0.0
        vol[l] = hlay[l][wet].sum()
        U = u[l] ** 2 * (0.5 * (hw[W] + hw))
        V = v[l] ** 2 * (0.5 * (hw[S] + hw))
        U[0] = V[0] = 0.0
        ok = np.arange(hlay.shape[1]) > 0  # every vector point except the frozen periodic duplicates
        if hm.params.xper > 0.5:
            ok &= sub[0] != sub[0].max()
        if hm.params.yper > 0.5:
            ok &= sub[1] != sub[1].max()
        ke[l] = (0.5 * (0.5 * (U + U[E])) + 0.5 * (0.5 * (V + V[N])))[ok].sum()
    eta = (hlay - h_0).sum(axis=0)
    return vol, ke, float((eta[wet] ** 2).sum())


@pytest.mark.parametrize("name,nsteps", [("conservation", 20), ("sill_exchange3D", 20)])
def test_conservation_integrals(case_factory, name, nsteps):
    c, d, hm = case_factory(name)
    orc = Oracle(hm.params, d)
    h_0 = orc.array("h_0").reshape(hm.array("hlay").shape).copy()
    orc.close()
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    vol0, ke0, pe0 = gm.diagnostics(h_0)
    gm.advance(1, nsteps)
    hl, u, v = gm.download_state()
    vol, ke, pe = gm.diagnostics(h_0)
    gm.close()
    wv, wk, wp = conservation_numpy(hm, hl, u, v, h_0)
    np.testing.assert_allclose(vol, wv, rtol=1e-12)
    np.testing.assert_allclose(ke, wk, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(pe, wp, rtol=1e-11, atol=1e-300)
    if name == "conservation":  # closed/periodic, unforced: layer volumes are conserved (doc p.4-6)
        np.testing.assert_allclose(vol, vol0, rtol=1e-11)
        assert ke.sum() > 0


def vorticity_numpy(hm, hlay, u, v):
    """conservation.m:169-211 on the vector layout: potential enstrophy, relative vorticity and its square, summed over
    the vector points (frozen periodic duplicates excluded), with pvor as write_array evaluates it (pm:2951-2974)."""
    neig = hm.iarray("neig")
    if neig.shape[0] != 8:
        neig = neig.T
    W, SW, S = neig[4], neig[5], neig[6]
    mk_n, mkpe, mkpi, fcor = (hm.array(k)[0] for k in ("mk_n", "mkpe", "mkpi", "fcor"))
    dl, uadv = hm.params.dl, hm.params.uadv
    sub = hm.iarray("subc")
    n = hlay.shape[1]
    ok = np.arange(n) > 0
    if hm.params.xper > 0.5:
        ok &= sub[0] != sub[0].max()
    if hm.params.yper > 0.5:
        ok &= sub[1] != sub[1].max()
    has = (np.arange(n) > 0).astype(float)  # a neighbour that is a vector point (or a periodic alias of one) holds a value
    out = {k: np.zeros(hlay.shape[0]) for k in ("enst", "zeta", "zeta2")}
    for l in range(hlay.shape[0]):
        h, uu, vv = hlay[l].copy(), u[l].copy(), v[l].copy()
        h[0] = uu[0] = vv[0] = 0.0
        zr = ((vv - vv[W]) / dl - (uu - uu[S]) / dl) * mkpe
        msum = mk_n + mk_n[W] + mk_n[S] + mk_n[SW]
        with np.errstate(all="ignore"):
            pv = (fcor + zr * uadv) * mkpi * msum / (h + h[W] + h[SW] + h[S])
        hatp = (h + h[W] * has[W] + h[S] * has[S] + h[SW] * has[SW]) / (1.0 + has[W] + has[S] + has[SW])
        z = pv * hatp - fcor
        good = ok & ~np.isnan(pv)
        out["enst"][l] = (pv ** 2 * hatp * 0.5)[good].sum()
        out["zeta"][l] = z[good].sum()
        out["zeta2"][l] = (z ** 2)[good].sum()
    return out, float(ok.sum())


@pytest.mark.parametrize("name,nsteps", [("conservation", 20), ("unstable_jet", 20), ("sill_exchange3D", 20)])
def test_vorticity_integrals(case_factory, name, nsteps):
    """Potential enstrophy and relative vorticity on the device (beom_gpu_diagnostics_all) against the numpy restatement
    on the downloaded state; sums, so the order of additions differs (rtol 1e-11 of the sum of magnitudes)."""
    c, d, hm = case_factory(name)
    orc = Oracle(hm.params, d)
    h_0 = orc.array("h_0").reshape(hm.array("hlay").shape).copy()
    orc.close()
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=True))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, nsteps)
    hl, u, v = gm.download_state()
    got = gm.diagnostics_all(h_0)
    vol, ke, pe = gm.diagnostics(h_0)
    gm.close()
    want, npts = vorticity_numpy(hm, hl, u, v)
    assert got["npts"] == npts
    np.testing.assert_array_equal(got["vol"], vol)
    np.testing.assert_array_equal(got["ke"], ke)
    np.testing.assert_allclose(got["enst"], want["enst"], rtol=1e-11)
    np.testing.assert_allclose(got["zeta2"], want["zeta2"], rtol=1e-11)
    # the relative vorticity nearly cancels over a closed or periodic domain: compare against the scale of its terms
    scale = np.sqrt(want["zeta2"] * npts)
    assert np.all(np.abs(got["zeta"] - want["zeta"]) <= 1e-11 * scale)
    assert np.all(got["enst"] > 0)


@pytest.mark.parametrize("name,fused,diag", [("sill_exchange3D", True, True), ("lock_exchange", True, False), ("conservation", False, True),
                                             ("stommel1948", True, False)])
def test_device_side_records_are_the_oracles_and_overlap_the_steps(case_factory, name, fused, diag):
    """beom_gpu_records_begin / _wait (write_array on the device, pm:2848-2883): two record sets are begun 12 steps apart
    without waiting in between -- each must be the oracle's float32 record of ITS step (eta_ with the float32 bottom-up
    accumulation, u___, v___, and the diag records), with the min / max thickness of the report."""
    c, d, hm = case_factory(name)
    orc = Oracle(hm.params, d)
    h0 = orc.array("h_0").reshape(hm.array("hlay").shape)[:, 1:].astype(np.float32)
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=fused))
    gm.set_rest_thickness(h0)
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    want = []
    wet = hm.array("mk_n")[0] > 0.5
    t = 0
    for nsteps in (10, 12):
        gm.advance(t + 1, t + nsteps)
        orc.advance(t + 1, t + nsteps)
        t += nsteps
        gm.records_begin(with_diag=diag)
        w = {k: orc.record(v).copy() for k, v in (("eta", "eta_"), ("u", "u___"), ("v", "v___"))}
        if diag:
            w.update({k: orc.record(k).copy() for k in ("pvor", "mont", "v_cc")})
        hl = orc.array("hlay").reshape(hm.array("hlay").shape)
        w["hmin"], w["hmax"] = hl[:, wet].min(axis=1), hl[:, wet].max(axis=1)
        want.append(w)
    gm.advance(t + 1, t + 3)  # the copies are still owed: the compute stream goes on
    own = np.ones(c.ndeg, dtype=bool)
    if hm.params.xper > 0.5 or hm.params.yper > 0.5:
        sub = hm.iarray("subc")
        own &= ~(((sub[0] == c.lm + 1) & (hm.params.xper > 0.5)) | ((sub[1] == c.mm + 1) & (hm.params.yper > 0.5)))[1:]
    for w in want:
        got = gm.records_wait()
        assert (got["first_point"], got["count"], got["thin_layer"]) == (1, c.ndeg, 0)
        for k in ("eta", "u", "v"):  # the state records hold the frozen duplicates too (patched from the host side)
            a, b = got[k], w[k].reshape(got[k].shape)
            assert np.array_equal(a, b), "%s record of %s differs at %d entries" % (k, name, (a != b).sum())
        for k in ("pvor", "mont", "v_cc"):
            if diag:
                a, b = got[k], w[k].reshape(got[k].shape)
                same = (a == b) | (np.isnan(a) & np.isnan(b)) | ~own[None, :]
                assert same.all(), "%s record of %s differs at %d entries" % (k, name, (~same).sum())
            else:
                assert got[k] is None
        np.testing.assert_array_equal(got["hmin"], w["hmin"])
        np.testing.assert_array_equal(got["hmax"], w["hmax"])
    with pytest.raises(RuntimeError):
        gm.records_wait()  # nothing in flight any more
    gm.close()
    orc.close()
