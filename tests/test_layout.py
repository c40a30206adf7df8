"""The dense layout beom_gpu_init derives from the reference's neighbour table (beom_b200/csrc/gpu/layout.h), checked on
the CPU through tools/layout_host.cc.  The property that makes the kernels' +-1 / +-NX neighbours right: after the
periodic images have been refreshed (local mirrors, then the slab / ring halo exchange), the cell next to every owned
vector point in each of the 8 directions shows exactly the point neig(k, p) names (0 = the discarded cell), with that
point's masks.  Single rank, y-slabs, and the ring-closed exchange of y-periodic domains."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from beom_b200 import cases, model
from tests.conftest import SMALL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F_N, F_U, F_V, F_PE, F_PI, F_ACT, F_GHOST = 1, 2, 4, 8, 16, 32, 64
DI = [1, 1, 0, -1, -1, -1, 0, 1]
DJ = [0, 1, 1, 1, 0, -1, -1, -1]


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("layout") / "liblayout_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tools", "layout_host.cc"), "-o", so],
                   check=True, capture_output=True, timeout=300)
    lib = C.CDLL(so)
    lib.layout_error.restype = C.c_char_p
    return lib


class Rank:
    def __init__(self, lib, hm, c, rank, nranks):
        p = hm.params
        nd1 = c.ndeg + 1
        self.sub, self.neig = hm.iarray("subc"), hm.iarray("neig")
        masks = [np.ascontiguousarray(hm.array(k)[0]) for k in ("mk_n", "mk_u", "mk_v", "mkpe", "mkpi")]
        ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        self.rc = lib.layout_run(c.lm, c.mm, c.ndeg, int(p.xper > 0.5), int(p.yper > 0.5), rank, nranks,
                                 self.sub.ctypes.data_as(ip), self.neig.ctypes.data_as(ip), *[m.ctypes.data_as(dp) for m in masks])
        self.error = lib.layout_error().decode()
        if self.rc:
            return
        d = (C.c_int * 13)()
        lib.layout_dims(d)
        (self.NX, self.NY, self.j0, self.j1, self.j_off, self.p_lo, self.p_hi, nmir, norph, torus, ring, self.G, self.GX0) = list(d)
        self.torus, self.ring = bool(torus), bool(ring)
        self.cell = np.zeros(nd1, dtype=np.int32)
        self.flags = np.zeros(self.NX * self.NY, dtype=np.uint8)
        self.poc = np.zeros(self.NX * self.NY, dtype=np.int32)
        self.mdst, self.msrc = np.zeros(max(nmir, 1), dtype=np.int32), np.zeros(max(nmir, 1), dtype=np.int32)
        self.orph = np.zeros(max(norph, 1), dtype=np.int32)
        lib.layout_copy(self.cell.ctypes.data_as(ip), self.flags.ctypes.data_as(C.POINTER(C.c_uint8)), self.poc.ctypes.data_as(ip),
                        self.mdst.ctypes.data_as(ip), self.msrc.ctypes.data_as(ip), self.orph.ctypes.data_as(ip))
        self.mdst, self.msrc, self.orph = self.mdst[:nmir], self.msrc[:nmir], self.orph[:norph]
        h = (C.c_int * 6)()
        lib.layout_halo(h)
        self.peer_lo, self.peer_hi, self.send_lo, self.send_hi, self.recv_lo, self.recv_hi = list(h)
        self.masks = masks
        # a field whose value at every vector point is the point's index
        self.plane = np.zeros((self.NY, self.NX))
        held = np.nonzero(self.cell >= 0)[0]
        self.plane.reshape(-1)[self.cell[held]] = held

    def mirror(self):
        flat = self.plane.reshape(-1)
        flat[self.mdst] = flat[self.msrc]


def exchange(ranks):
    """sync_fields: local mirrors, then one packed G-row message per neighbour (comm_exchange)."""
    for r in ranks:
        r.mirror()
    G = ranks[0].G
    msgs = {}
    for k, r in enumerate(ranks):
        if r.peer_lo >= 0:
            msgs[(k, r.peer_lo, "lo")] = r.plane[r.send_lo:r.send_lo + G].copy()  # what I send down arrives "from above"
        if r.peer_hi >= 0:
            msgs[(k, r.peer_hi, "hi")] = r.plane[r.send_hi:r.send_hi + G].copy()
    for k, r in enumerate(ranks):
        if r.peer_lo >= 0:
            r.plane[r.recv_lo:r.recv_lo + G] = msgs[(r.peer_lo, k, "hi")]
        if r.peer_hi >= 0:
            r.plane[r.recv_hi:r.recv_hi + G] = msgs[(r.peer_hi, k, "lo")]


def check_neighbours(ranks, c):
    """Every owned vector point of every rank: the 8 neighbouring cells show neig(k, p), with its masks."""
    checked = 0
    for r in ranks:
        si, sj = r.sub[0], r.sub[1]
        for p in range(1, c.ndeg + 1):
            if r.cell[p] < 0 or not (r.j0 <= sj[p] <= r.j1):
                continue
            for k in range(8):
                q = int(r.neig[p, k])
                y, x = sj[p] + DJ[k] + r.j_off, si[p] + DI[k] + r.GX0
                assert r.plane[y, x] == q, (r.j0, r.j1, p, k, q, r.plane[y, x])
                f = int(r.flags[y * r.NX + x])
                if q:
                    want = sum(bit for bit, m in zip((F_N, F_U, F_V, F_PE, F_PI), r.masks) if m[q] > 0.5)
                    assert f & 31 == want, (p, k, q)
                else:
                    assert f & 31 == 0 or r.plane[y, x] == 0
                checked += 1
    return checked


def build(lib, name, nranks, extra="", **kw):
    args = dict(SMALL.get(name, {}))
    args.update(kw)
    c = cases.CASES[name](**args)
    c.params_text += extra
    import tempfile
    d = tempfile.mkdtemp(prefix="beom_layout_")
    hm = model.HostModel.from_block(c.write(d))
    return c, hm, [Rank(lib, hm, c, r, nranks) for r in range(nranks)]


@pytest.mark.parametrize("name", ["stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D", "conservation", "soliton",
                                  "baines_ridge", "upwelling_seaward_wind", "morel_upwelling", "outcrop_seamount", "wave_sponge"])
def test_single_rank_layout_shows_the_reference_neighbours(shim, name):
    c, hm, (r,) = build(shim, name, 1)
    assert r.rc == 0, r.error
    r.mirror()
    assert check_neighbours([r], c) == 8 * np.count_nonzero(r.cell >= 0)
    # duplicates: masked, displaced, and nothing else lost its cell
    assert np.array_equal(np.sort(r.orph), np.nonzero(r.cell == -2)[0])
    for m in r.masks[:3]:
        assert not np.any(m[r.orph] > 0.5)
    per = hm.params.xper > 0.5 or hm.params.yper > 0.5
    assert (len(r.orph) > 0) == per and r.ring is False
    if name in ("unstable_jet", "conservation", "soliton"):
        assert r.torus  # complete torus: deep images for the fused step
        flat = r.plane.reshape(-1)
        assert np.all(flat[r.mdst] == flat[r.msrc]) and np.all(r.flags[r.mdst] & F_ACT == 0)


@pytest.mark.parametrize("name,nranks", [("synthetic_basin", 2), ("synthetic_basin", 8), ("sill_exchange3D", 2), ("soliton", 4)])
def test_y_slabs_show_the_reference_neighbours(shim, name, nranks):
    kw = dict(n=96, mm=70, nlay=1) if name == "synthetic_basin" else {}
    c, hm, ranks = build(shim, name, nranks, **kw)
    assert all(r.rc == 0 for r in ranks), [r.error for r in ranks]
    assert [r.j0 for r in ranks][0] == 1 and ranks[-1].j1 == c.mm + 1 and not any(r.ring for r in ranks)
    # periodic in x only: the images stay inside a row, so every slab is a complete torus of its own (fused step allowed)
    assert all(r.torus == (name == "soliton") for r in ranks)
    assert ranks[0].peer_lo == -1 and ranks[-1].peer_hi == -1
    exchange(ranks)
    owned = sum(np.count_nonzero((r.cell >= 0) & (r.sub[1] >= r.j0) & (r.sub[1] <= r.j1)) for r in ranks)
    assert check_neighbours(ranks, c) == 8 * owned


@pytest.mark.parametrize("name,nranks,extra", [("conservation", 2, ""), ("conservation", 3, ""), ("unstable_jet", 4, ""),
                                               ("conservation", 2, "xper       = 0.\n")])
def test_y_periodic_slabs_close_the_ring(shim, name, nranks, extra):
    """SURVEY.md section 8(e), 'periodic yper = 1: same messages, ring-closed': row 0 of the first rank shows row mm of
    the last, row mm+1 of the last shows row 1 of the first (one row below a plain slab boundary, because row mm+1 is
    the duplicate of row 1)."""
    c, hm, ranks = build(shim, name, nranks, extra)
    assert all(r.rc == 0 for r in ranks), [r.error for r in ranks]
    assert all(r.ring and r.torus for r in ranks)  # the seam's deep rows are flagged: the fused step may run
    first, last = ranks[0], ranks[-1]
    assert first.peer_lo == nranks - 1 and last.peer_hi == 0
    G = first.G
    assert (last.send_hi, last.recv_hi) == (G + (last.j1 - last.j0) - G, G + (last.j1 - last.j0))   # rows mm-G+1..mm / mm+1..
    assert (first.send_lo, first.recv_lo) == (G, 0)
    # the duplicate row of the last rank is displaced there, exactly as on one rank
    c1, hm1, (one,) = build(shim, name, 1, extra)
    dup_row = np.nonzero((one.sub[1] == c.mm + 1) & (np.arange(c.ndeg + 1) > 0))[0]
    assert set(dup_row) <= set(one.orph) and set(dup_row) <= set(last.orph)
    exchange(ranks)
    owned = sum(np.count_nonzero((r.cell >= 0) & (r.sub[1] >= r.j0) & (r.sub[1] <= r.j1)) for r in ranks)
    assert check_neighbours(ranks, c) == 8 * owned


def test_unsupported_slab_splits_fail_loudly(shim):
    # a one-row y-periodic channel cannot give every rank the G + 1 rows the ring needs
    c, hm, ranks = build(shim, "baines_ridge", 2)
    assert all(r.rc == -9 for r in ranks) and "rows per rank" in ranks[0].error
    # more ranks than rows
    c, hm, ranks = build(shim, "lock_exchange", 4)
    assert any(r.rc == -5 for r in ranks)
