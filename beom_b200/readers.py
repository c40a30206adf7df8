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
