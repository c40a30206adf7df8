"""Readers for the reference's output files (numpy restatement of testcases/get_metadata.m and
testcases/get_field.m).  Formats: SURVEY.md appendix A."""
from __future__ import annotations

import os
import re

import numpy as np


def get_metadata(outdir: str) -> dict:
    """param_basin.txt (Octave-evaluable `name = value ;` lines), time.txt, grid.bin, h_0.bin."""
    meta: dict = {}
    with open(os.path.join(outdir, "param_basin.txt")) as f:
        for line in f:
            line = line.strip()
            if len(line) <= 1:
                break
            m = re.match(r"(\w+)\s*=\s*(.*?);?\s*$", line)
            if not m:
                continue
            key, val = m.group(1), m.group(2).strip().rstrip(";").strip()
            if val.startswith("'"):
                meta[key] = val.strip("'")
            elif val.startswith("["):
                meta[key] = np.array([float(x) for x in val.strip("[]").split()])
            else:
                meta[key] = float(val) if re.search(r"[.eE]", val) else int(val)
    meta["taxi"] = np.atleast_1d(np.loadtxt(os.path.join(outdir, "time.txt")))
    grid = np.fromfile(os.path.join(outdir, "grid.bin"), dtype="<i4")
    n = grid.size // 5
    meta["posc"], meta["mk_n"], meta["mk_u"], meta["mk_v"], meta["mkpi"] = (grid[k * n:(k + 1) * n] for k in range(5))
    nlay = int(meta["nlay"])
    meta["h_0_vec"] = np.fromfile(os.path.join(outdir, "h_0.bin"), dtype="<f4").reshape(nlay, n)
    return meta


def get_field(vnam: str, irec: int, outdir: str, meta: dict | None = None) -> np.ndarray:
    """Record `irec` (1-based; 0 = last) of eta_/u___/v___/pvor/mont/v_cc as a (lm+2, mm+2, nlay) array
    with NaN on masked / land points (get_field.m:98-116)."""
    meta = meta or get_metadata(outdir)
    lm, mm, nlay = int(meta["lm"]), int(meta["mm"]), int(meta["nlay"])
    n = meta["posc"].size
    if irec == 0:
        irec = meta["taxi"].size
    mask = {"u___": meta["mk_u"], "v___": meta["mk_v"], "pvor": meta["mkpi"]}.get(vnam, meta["mk_n"]).astype(float)
    mask[mask == 0] = np.nan
    rec = np.fromfile(os.path.join(outdir, vnam + ".bin"), dtype="<f4", count=n * nlay, offset=4 * (irec - 1) * n * nlay)
    rec = rec.reshape(nlay, n).astype(float)
    out = np.full(((lm + 2) * (mm + 2), nlay), np.nan)
    for l in range(nlay):
        out[meta["posc"] - 1, l] = rec[l] * mask
    return out.reshape(mm + 2, lm + 2, nlay).transpose(1, 0, 2)  # posc = i+1 + j*(lm+2): i fastest


def vector_to_grid(vec: np.ndarray, subc: np.ndarray, lm: int, mm: int) -> np.ndarray:
    """A (0:ndeg) model vector -> (lm+2, mm+2) grid (NaN where no vector point)."""
    out = np.full((lm + 2, mm + 2), np.nan)
    out[subc[0, 1:], subc[1, 1:]] = vec[1:]
    return out


def get_h_0(meta: dict) -> np.ndarray:
    """h_0 as get_metadata.m:44-49 returns it: (lm+2, mm+2, nlay), 0 where there is no vector point."""
    lm, mm, nlay = int(meta["lm"]), int(meta["mm"]), int(meta["nlay"])
    out = np.zeros(((lm + 2) * (mm + 2), nlay))
    for l in range(nlay):
        out[meta["posc"] - 1, l] = meta["h_0_vec"][l]
    return out.reshape(mm + 2, lm + 2, nlay).transpose(1, 0, 2)


def conservation_integrals(outdir: str, grav: float = 9.8) -> dict:
    """The time series testcases/conservation.m:116-211 computes from a run's output files (needs ``diag = 1`` for
    pvor.bin): per record and layer the domain-mean thickness ``volu`` and its mean absolute change ``vstd``, the
    domain-mean potential enstrophy ``enst`` (+ ``estd``), the mean and standard deviation of the relative vorticity
    ``rvor`` / ``rstd``, the kinetic energy ``kine`` (J) and the barotropic potential energy ``pote`` (J, first column)."""
    meta = get_metadata(outdir)
    lm, mm, nlay = int(meta["lm"]), int(meta["mm"]), int(meta["nlay"])
    xper, yper = float(meta.get("xper", 0)) > 0.5, float(meta.get("yper", 0)) > 0.5
    dl, fcor = float(meta["dl"]), float(meta["f0"])
    rhon = np.atleast_1d(np.asarray(meta["rhon"], dtype=float))
    h_0 = get_h_0(meta)
    nrec = meta["taxi"].size
    out = {k: np.full((nrec, nlay), np.nan) for k in ("volu", "vstd", "enst", "estd", "rvor", "rstd", "pote", "kine")}
    out["taxi"] = meta["taxi"]
    hl_0 = en_0 = None
    for irec in range(1, nrec + 1):
        n = get_field("eta_", irec, outdir, meta)
        u = get_field("u___", irec, outdir, meta)
        v = get_field("v___", irec, outdir, meta)
        pvor = get_field("pvor", irec, outdir, meta)
        hlay = np.full((lm + 2, mm + 2, nlay), np.nan)
        hlay[:, :, nlay - 1] = h_0[:, :, nlay - 1] + n[:, :, nlay - 1]
        for l in range(nlay - 1):
            hlay[:, :, l] = h_0[:, :, l] + n[:, :, l] - n[:, :, l + 1]
        if hl_0 is None:
            hl_0 = hlay.copy()
        with np.errstate(all="ignore"):
            out["volu"][irec - 1] = np.nanmean(hlay.reshape(-1, nlay), axis=0)
            out["vstd"][irec - 1] = np.nanmean(np.abs(hlay - hl_0).reshape(-1, nlay), axis=0)
        out["pote"][irec - 1, 0] = 0.5 * rhon[0] * grav * dl ** 2 * np.nansum(n[:, :, 0] ** 2)
        if xper:
            hlay[0, :, :] = hlay[lm, :, :]
        if yper:
            hlay[:, 0, :] = hlay[:, mm, :]
        hzer, uzer, vzer = np.nan_to_num(hlay), np.nan_to_num(u), np.nan_to_num(v)
        m0 = (~np.isnan(hlay)).astype(float)
        mask = m0.copy()
        mask[1:, 1:, :] = m0[:-1, :-1, :] + m0[:-1, 1:, :] + m0[1:, :-1, :] + m0[1:, 1:, :]
        hatu, hatv, hatp = np.zeros_like(hzer), np.zeros_like(hzer), np.zeros_like(hzer)
        hatu[1:, :, :] = 0.5 * (hzer[:-1, :, :] + hzer[1:, :, :])
        hatv[:, 1:, :] = 0.5 * (hzer[:, :-1, :] + hzer[:, 1:, :])
        utmp, vtmp = uzer ** 2 * hatu, vzer ** 2 * hatv
        utmp = 0.5 * (utmp[:-1, :, :] + utmp[1:, :, :])
        vtmp = 0.5 * (vtmp[:, :-1, :] + vtmp[:, 1:, :])
        kine = 0.5 * utmp.reshape(-1, nlay).sum(axis=0) + 0.5 * vtmp.reshape(-1, nlay).sum(axis=0)
        out["kine"][irec - 1] = kine * rhon * dl ** 2
        mask[mask == 0] = 1.0
        hatp[1:, 1:, :] = (hzer[:-1, :-1, :] + hzer[:-1, 1:, :] + hzer[1:, :-1, :] + hzer[1:, 1:, :]) / mask[1:, 1:, :]
        if xper:
            pvor[-1, :, :] = np.nan
        if yper:
            pvor[:, -1, :] = np.nan
        ens = pvor ** 2 * hatp * 0.5
        if en_0 is None:
            en_0 = ens.copy()
        with np.errstate(all="ignore"):
            out["enst"][irec - 1] = np.nanmean(ens.reshape(-1, nlay), axis=0)
            out["estd"][irec - 1] = np.nanmean(np.abs(ens - en_0).reshape(-1, nlay), axis=0)
            zeta = (pvor * hatp - fcor).reshape(-1, nlay)
            out["rvor"][irec - 1] = np.nanmean(zeta, axis=0)
            out["rstd"][irec - 1] = np.nanstd(zeta, axis=0, ddof=1)  # Octave's nanstd normalises by N - 1
    return out
