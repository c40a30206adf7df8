"""Compare the hand-written oracle with the reference itself (TEST INFRASTRUCTURE ONLY).

``oracle/refbuild.py`` builds the reference from its own sources through the f95c translator; this module runs both
for the same number of steps on the same inputs and compares every module array of ``private_mod.f95`` -- state,
histories, scratch fields, masks, connectivity, forcing, open-boundary segments, scalars -- bit for bit, plus the
reference's own output files against the oracle's records.  Used by tests/test_reference_pin.py and
tests/golden/make_golden.py."""
from __future__ import annotations

import os
import re

import numpy as np

from . import refbuild
from .pyoracle import Oracle


def block_for_steps(block: str, nsteps: int, dt: float) -> str:
    """The same parameter block with dt_s (and dt_o) set so that integrate_time runs exactly ``nsteps`` steps and writes
    one record after the last (private_mod.f95:1852-1856: nstp = nint(dt_s / dtd8), notp = nint(dt_o / dtd8))."""
    dtd8 = dt / 24.0 / 3600.0
    val = "%.9e" % (nsteps * dtd8)
    f32 = float(np.float32(float(val)))
    assert int(np.floor(f32 / dtd8 + 0.5)) == nsteps, (val, f32 / dtd8)
    out = []
    for line in block.splitlines():
        m = re.match(r"^\s*(dt_s|dt_o)\s*=", line)
        if m:
            line = "%-10s = %s" % (m.group(1), val)
        out.append(line)
    return "\n".join(out) + "\n"


def _flat(orc: Oracle, name: str, n: int, integer: bool = False):
    p = (orc.lib.beom_oracle_iarray if integer else orc.lib.beom_oracle_array)(orc.h, name.encode())
    if not p:
        return None
    return np.ctypeslib.as_array(p, shape=(n,))


# module array -> axes permutation that turns the reference's (Fortran-order) array into the oracle's C-order one
_PERM = {
    "hlay": (1, 0), "u": (1, 0), "v": (1, 0), "h_u": (1, 0), "h_v": (1, 0), "v_cc": (1, 0), "v_ll": (1, 0),
    "uu4": (1, 0), "vv4": (1, 0), "delu": (1, 0), "delv": (1, 0), "hdot": (1, 0),
    "tt3d": (2, 1, 0), "tb3d": (2, 1, 0), "tu3d": (2, 1, 0), "taus": (1, 0), "nudg": (1, 0), "fnud": (2, 1, 0),
    "rs_h": (2, 1, 0), "dmdx": (2, 1, 0), "dmdy": (2, 1, 0), "bodf": (1, 0),
    "rvor": (0,), "pvor": (0,), "dive": (0,), "fcor": (0,), "mk_u": (0,), "mk_v": (0,), "mk_n": (0,), "mkpe": (0,),
    "mkpi": (0,), "mont": (0,), "d2hx": (0,), "d2hy": (0,), "h_bo": (0,), "h_to": (0,), "h_th": (0,), "ow": (0,),
    "os": (0,), "osum_": (0,), "pi_s": (0,),
}
_ORACLE_NAME = {"uu4": "UU4", "vv4": "VV4", "ow": "Ow", "os": "Os", "osum_": "Osum_"}
_INT = {"neig": (1, 0), "subc": (1, 0), "segm": (1, 0)}
_SCALARS = ("ctim", "invf", "ramp", "gene", "tres")


def compare(dump: dict, orc: Oracle, skip=()):
    """-> [(name, identical, max |difference|, max |reference|)] over everything both sides hold."""
    rows = []

    def one(name, ref, mine):
        ref = np.ascontiguousarray(ref)
        mine = np.ascontiguousarray(mine)
        if ref.dtype.kind == "f":
            same = ref.shape == mine.shape and np.array_equal(ref.view(np.int64), mine.view(np.int64))
        else:
            same = ref.shape == mine.shape and np.array_equal(ref, mine)
        diff = float(np.max(np.abs(ref.astype(np.float64) - mine.astype(np.float64)))) if (
            ref.shape == mine.shape and ref.size) else float("nan")
        rows.append((name, bool(same), diff, float(np.max(np.abs(ref))) if ref.size else 0.0))

    for name, perm in _PERM.items():
        if name in skip or name not in dump:
            continue
        ref = np.transpose(np.asarray(dump[name], dtype=np.float64), perm).reshape(-1)
        mine = _flat(orc, _ORACLE_NAME.get(name, name), ref.size)
        if mine is None:
            continue
        one(name, ref, mine)
    if "tide" in dump and "tide" not in skip:
        ref = np.transpose(dump["tide"][:, 0, :, :], (2, 1, 0)).reshape(-1)
        one("tide", ref, _flat(orc, "tide", ref.size))
        one("w_ti", np.asarray(dump["w_ti"], dtype=np.float64).reshape(-1)[:1], np.array([orc.scalar("w_ti")]))
    for name, perm in _INT.items():
        if name not in dump or name in skip:
            continue
        ref = np.transpose(dump[name], perm).reshape(-1).astype(np.int32)
        if name == "segm":
            if orc.nseg() != dump["segm"].shape[0]:
                rows.append(("segm", False, float("nan"), float(dump["segm"].shape[0])))
                continue
            if ref.size == 0:
                continue
        mine = _flat(orc, name, ref.size, integer=True)
        one(name, ref, mine)
    for name in _SCALARS:
        if name in dump:
            one(name, np.array([float(dump[name])]), np.array([orc.scalar(name)]))
    if "flag_nudging" in dump:
        one("flag_nudging", np.array([int(dump["flag_nudging"])]), np.array([int(orc.scalar("flag_nudging"))]))
    return rows


def read_records(path: str, ndeg: int, nlay: int):
    a = np.fromfile(path, dtype="<f4")
    return a.reshape(-1, nlay, ndeg)


def compare_files(odir: str, orc: Oracle, diag: bool = False):
    """The reference's own output files (last record) against the oracle's records; h_0.bin and grid.bin too."""
    rows = []
    nd, nl = orc.ndeg, orc.nlay
    for var in ("eta_", "u___", "v___") + (("pvor", "mont", "v_cc") if diag else ()):
        rec = read_records(os.path.join(odir, var + ".bin"), nd, nl)[-1]
        mine = orc.record(var)
        same = np.array_equal(rec.view(np.int32), mine.view(np.int32))
        rows.append((var + ".bin", bool(same), float(np.max(np.abs(rec.astype(np.float64) - mine))), float(np.max(np.abs(rec)))))
    h0 = np.fromfile(os.path.join(odir, "h_0.bin"), dtype="<f4").reshape(nl, nd)
    mine = orc.array("h_0")[:, 1:].astype(np.float32)
    rows.append(("h_0.bin", bool(np.array_equal(h0.view(np.int32), mine.view(np.int32))),
                 float(np.max(np.abs(h0.astype(np.float64) - mine))), float(np.max(np.abs(h0)))))
    grid = np.fromfile(os.path.join(odir, "grid.bin"), dtype="<i4").reshape(5, nd)
    posc = _flat(orc, "posc", nd + 1, integer=True)
    if posc is not None:
        mine = np.stack([posc[1:]] + [np.rint(orc.array(m)[0, 1:]).astype(np.int32) for m in ("mk_n", "mk_u", "mk_v", "mkpi")])
        # posc is stored 1-based from index 0 in some builds of the oracle: accept either alignment, report which
        if not np.array_equal(grid[0], mine[0]):
            mine[0] = _flat(orc, "posc", nd, integer=True)
        rows.append(("grid.bin", bool(np.array_equal(grid, mine)), float(np.max(np.abs(grid - mine))), float(grid.max())))
    return rows


def run_pair(block: str, idir: str, nsteps: int, variant: int = 0, keep_oracle: bool = False):
    """Build + run the translated reference for ``nsteps`` steps and the oracle beside it.
    -> (dump, rows of compare(), rows of compare_files(), oracle or None)."""
    from beom_b200 import model

    p, _, _, _ = model.parse_params(block, variant)
    blk = block_for_steps(block, nsteps, p.dt)
    p, pidir, podir, _ = model.parse_params(blk, variant)
    exe = refbuild.build_case(blk, variant)
    dump, _ = refbuild.run_case(exe, podir)
    orc = Oracle(p, pidir)
    orc.advance(1, nsteps)
    rows = compare(dump, orc)
    frows = compare_files(podir, orc, diag=p.diag > 0.5)
    if not keep_oracle:
        orc.close()
        orc = None
    return dump, rows, frows, orc


def format_rows(rows):
    return "\n".join("%-12s %-9s max|diff| %.3e  max|ref| %.3e" % (n, "identical" if s else "DIFFERENT", d, m)
                     for n, s, d, m in rows)
