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
