"""Input generators for the reference's test cases, restated in numpy.

The reference ships Octave scripts (``testcases/*.m``) that write little-endian float32, column-major
``.bin`` inputs and print a parameter block (``testcases/print_params.m``) to paste into
``shared_mod.f95``.  Neither Octave nor MATLAB exists in this environment, so each generator below
follows its script line by line (cited) and writes byte-compatible files; ``print_params`` reproduces
the *printf rounding* of the script, which changes the numbers the model actually runs with
(e.g. ``cext`` is printed with ``%0.1f``).

Rounding: MATLAB/Octave ``round`` is half-away-from-zero (``_mround``), unlike Python's.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np


def _mround(x: float) -> int:
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def fwrite_r4(path: str, arr: np.ndarray) -> None:
    """fwrite(fid, arr, 'real*4', 0, 'ieee-le'): column-major float32 (written plane by plane)."""
    a = np.asarray(arr)
    if a.dtype != np.float32:
        a = a.astype(np.float64).astype("<f4")
    if a.ndim <= 2:
        a.flatten(order="F").tofile(path)
        return
    tail = a.shape[2:]
    with open(path, "wb") as f:
        for flat in range(int(np.prod(tail))):
            idx = np.unravel_index(flat, tail, order="F")
            np.ascontiguousarray(a[(slice(None), slice(None)) + tuple(idx)].T).tofile(f)


def get_nbr_deg_freedom(h_bo: np.ndarray) -> int:
    """testcases/get_nbr_deg_freedom.m:11-52."""
    hdry = 1.0e-3
    h = np.array(h_bo, dtype=np.float64, copy=True)
    h[h < 2.0 * hdry] = 0.0
    h[0, :] = 0.0
    h[-1, :] = 0.0
    h[:, 0] = 0.0
    h[:, -1] = 0.0
    lm, mm = h.shape[0] - 2, h.shape[1] - 2
    mask = np.zeros((lm + 4, mm + 4))
    mask[1:-1, 1:-1] = h > hdry
    neig = mask[1:-1, 1:-1] + mask[0:-2, 1:-1] + mask[1:-1, 0:-2] + mask[0:-2, 0:-2]
    return int(np.count_nonzero(neig > 0))


def print_params(lm, mm, nlay, ndeg, dl, cext, f0, rhon, topl, dt_s, dt_o, dt_r, dt3d, bvis, dvis, bdrg,
                 hmin, hsbl, hbbl, g_fb, uadv, qdrg, ocrp, rsta, xper, yper, diag, tauw, idir, odir, desc,
                 extra: dict | None = None) -> str:
    """testcases/print_params.m:13-92 -- returns the block of Fortran assignments it displays."""
    rhon = np.atleast_1d(np.asarray(rhon, dtype=np.float64))
    topl = np.atleast_1d(np.asarray(topl, dtype=np.float64))
    out = []
    out.append("lm         = %d" % lm)
    out.append("mm         = %d" % mm)
    out.append("nlay       = %d" % nlay)
    out.append("ndeg       = %d" % ndeg)
    if dl < 1.0e3:
        out.append("dl         = " + "%#0.0f" % dl)
    elif (dl - math.floor(dl / 1.0e3) * 1.0e3) > 1.0:
        out.append("dl         = " + "%g" % (dl / 1.0e3) + "e3")
    else:
        out.append("dl         = " + "%#0.0f" % (dl / 1.0e3) + "e3")
    out.append("cext       = " + "%0.1f" % cext)
    if abs(f0) < 1.0e-4:
        out.append("f0         = " + "%e" % f0)
    elif abs(f0) > 1.001e-4:
        out.append("f0         = " + "%0.3f" % (f0 / 1.0e-4) + "e-4")
    else:
        out.append("f0         = " + "%#0.0f" % (f0 / 1.0e-4) + "e-4")
    if np.all((rhon * 1.0e3 - np.floor(rhon) * 1.0e3) < 1.0):
        strg = "".join("%#0.0f," % r for r in rhon)
    else:
        strg = "".join("%0.3f," % r for r in rhon)
    out.append("rhon(nlay) = (/" + strg[:-1] + "/)")
    strg = "".join("%f," % t for t in topl)
    out.append("topl(nlay) = (/" + strg[:-1] + "/)")
    out.append("dt_s       = " + "%#f" % dt_s)
    out.append("dt_o       = " + "%#f" % dt_o)
    out.append("dt_r       = " + ("%#f" % dt_r if dt_r > 0.0 else "%#0.0f" % dt_r))
    out.append("dt3d       = " + ("%#f" % dt3d if dt3d > 0.0 else "%#0.0f" % dt3d))
    out.append("bvis       = " + ("%#f" % bvis if bvis > 0.0 else "%#0.0f" % bvis))
    out.append("dvis       = " + ("%#0.3f" % dvis if dvis > 0.0 else "%#0.0f" % dvis))
    out.append("bdrg       = " + ("%e" % bdrg if bdrg > 0.0 else "%#0.0f" % bdrg))
    out.append("hmin       = " + "%#f" % hmin)
    for name, val in (("hsbl", hsbl), ("hbbl", hbbl), ("g_fb", g_fb), ("uadv", uadv), ("qdrg", qdrg),
                      ("ocrp", ocrp), ("rsta", rsta), ("xper", xper), ("yper", yper), ("diag", diag)):
        out.append("%-10s = " % name + "%#0.0f" % val)
    strg = "".join("%#0.2f," % t for t in tauw)
    out.append("tauw       = (" + strg[:-1] + ")")
    out.append("idir       = '" + idir + "'")
    out.append("odir       = '" + odir + "'")
    out.append("desc       = '" + desc + "'")
    # parameters print_params.m has no slot for (rgld, mcbc, svis, tdrg, topt, plum): appended verbatim
    for k, v in (extra or {}).items():
        out.append("%-10s = %s" % (k, v))
    return "\n".join(out) + "\n"


@dataclass
class Case:
    name: str
    lm: int
    mm: int
    nlay: int
    ndeg: int
    params_text: str
    files: dict = field(default_factory=dict)  # name -> float array in the script's (i, j, ...) shape
    info: dict = field(default_factory=dict)

    def write(self, directory: str) -> str:
        """Write the .bin inputs and the parameter block (``shared_mod_block.f95``); returns its path."""
        os.makedirs(directory, exist_ok=True)
        for name, arr in self.files.items():
            fwrite_r4(os.path.join(directory, name + ".bin"), arr)
        d = directory if directory.endswith("/") else directory + "/"
        text = self.params_text.replace("@DIR@", d)
        path = os.path.join(directory, "shared_mod_block.f95")
        with open(path, "w") as f:
            f.write(text)
        return path


def _edge_fill(tmp: np.ndarray) -> np.ndarray:
    """tmp(1,:)=tmp(2,:); tmp(:,1)=tmp(:,2); tmp(end,:)=tmp(end-1,:); tmp(:,end)=tmp(:,end-1)."""
    tmp[0, ...] = tmp[1, ...]
    tmp[:, 0, ...] = tmp[:, 1, ...]
    tmp[-1, ...] = tmp[-2, ...]
    tmp[:, -1, ...] = tmp[:, -2, ...]
    return tmp


def stommel1948(dl: float = 100.0e3, dt_s: float = 40.0) -> Case:
    """testcases/stommel1948.m:17-88.  ``dl`` may be changed to get a smaller/larger grid."""
    lam = 1.0e7
    b = 2.0 * math.pi * 1.0e6
    lm = _mround(lam / dl)
    mm = _mround(b / dl)
    rhon = 1027.0
    hfla = 200.0
    F = 0.1 / rhon
    R = 2.0e-4
    beta = 1.0e-11
    fmin = 0.0
    grav = 9.8
    xs = (np.arange(lm) + 0.5) * dl
    ys = (np.arange(mm) + 0.5) * dl
    yy = np.broadcast_to(ys[None, :], (lm, mm))
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = hfla
    ndeg = get_nbr_deg_freedom(h_bo)
    fcor = np.empty((lm, mm))
    for j in range(1, mm + 1):
        fcor[:, j - 1] = fmin + (j - 0.5) * dl * beta
    tausx = -rhon * F * np.cos(math.pi * yy / b)
    tausy = np.zeros_like(tausx)
    tmp = np.full((lm + 2, mm + 2), np.nan)
    tmp[1:-1, 1:-1] = fcor
    fcor_file = _edge_fill(tmp)
    tmp = np.full((lm + 2, mm + 2, 2), np.nan)
    tmp[1:-1, 1:-1, 0] = tausx
    tmp[1:-1, 1:-1, 1] = tausy
    taus_file = _edge_fill(tmp)
    cext = math.sqrt(grav * h_bo.max())
    text = print_params(lm, mm, 1, ndeg, dl, cext, 0.0, [rhon], [0.0], dt_s, 1.0, 0.0, 0.0, 0.0, 0.0, R, 1.0, 10.0, 10.0,
                        1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "Test-case for Stommel 1948")
    # analytical solution, stommel1948.m:31-36, 98-104
    alpha = hfla * beta / R
    gamma = F * math.pi / R / b
    A = -0.5 * alpha + math.sqrt(0.25 * alpha ** 2 + (math.pi / b) ** 2)
    B = -0.5 * alpha - math.sqrt(0.25 * alpha ** 2 + (math.pi / b) ** 2)
    p = (1.0 - math.exp(B * lam)) / (math.exp(A * lam) - math.exp(B * lam))
    q = 1.0 - p
    xx = np.broadcast_to(xs[:, None], (lm, mm))
    eta = (-F / grav / hfla * (np.exp(A * xx) * p / A + np.exp(B * xx) * q / B)
           - (b / math.pi) ** 2 * F / grav / hfla * (p * A * np.exp(A * xx) + q * B * np.exp(B * xx)) * (np.cos(math.pi * yy / b) - 1.0)
           - (fcor * gamma / grav * (b / math.pi) ** 2 * np.sin(math.pi * yy / b)
              + beta * gamma / grav * (b / math.pi) ** 3 * (np.cos(math.pi * yy / b) - 1.0))
           * (p * np.exp(A * xx) + q * np.exp(B * xx) - 1.0))
    return Case("stommel1948", lm, mm, 1, ndeg, text, {"fcor": fcor_file, "taus": taus_file},
                {"eta_analytic": eta, "hfla": hfla})


def lock_exchange(dt_s: float = 5.0) -> Case:
    """testcases/lock_exchange.m:11-64."""
    hmax, nlay = 20.0, 2
    rhon = [1025.0, 1030.0]
    topl = [0.0, 0.5]
    grav, dl, lx, mm = 9.8, 400.0, 64.0e3, 1
    hmin = 0.005
    dt_o = 1.0 / 24.0
    hsal = 10.0 * hmin
    lm = _mround(lx / dl)
    h_bo = hmax * np.ones((lm + 2, mm + 2))
    h_bo[:, 0] = 0.0
    h_bo[:, -1] = 0.0
    h_bo[0, :] = 0.0
    h_bo[-1, :] = 0.0
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    cint = 0.5 * math.sqrt(grav * hmax * (rhon[1] - rhon[0]) / rhon[1])
    n = np.zeros((lm + 2, mm + 2, nlay))
    u = np.zeros_like(n)
    v = np.zeros_like(n)
    half = _mround(0.5 * (lm + 2))
    n[:half, :, 1] = 0.5 * hmax - 4.0 * hsal
    n[half:, :, 1] = -0.5 * hmax + 4.0 * hsal
    init = np.stack([n, u, v], axis=3)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, 0.0, rhon, topl, dt_s, dt_o, 0.0, 0.0, 0.0, 0.03, 0.0, hmin, 10.0,
                        10.0, 1.0, 1.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "Test-case for lock-exchange")
    return Case("lock_exchange", lm, mm, nlay, ndeg, text, {"init": init}, {"cint": cint, "hsal": hsal, "hmax": hmax})


def unstable_jet(dl: float = 15.0e3, dt_s: float = 50.0) -> Case:
    """testcases/unstable_jet.m:8-66 (single layer, doubly periodic)."""
    hshf = 5.0
    lx, ly = 3000.0e3, 4000.0e3
    rhon = 1030.0
    lm = _mround(lx / dl)
    lm += (lm % 2 == 0)
    mm = _mround(ly / dl)
    mm += (mm % 2 == 0)
    nlay, grav, fcor = 1, 9.8, 0.5e-4
    xs = np.arange(1, lm + 3) - 1.5
    ys = np.arange(1, mm + 3) - 1.5
    xx = np.broadcast_to(xs[:, None], (lm + 2, mm + 2)).copy()
    yy = np.broadcast_to(ys[None, :], (lm + 2, mm + 2)).copy()
    xx = xx - xx.mean()
    yy = yy - yy.mean()
    h_bo = hshf + 0.1 * hshf * np.cos(4.0 * math.pi * xx / lm)
    h_bo[h_bo < 1.0] = 0.0
    h_bo[0, :] = 0.0
    h_bo[-1, :] = 0.0
    h_bo[:, 0] = 0.0
    h_bo[:, -1] = 0.0
    cext = math.sqrt(grav * h_bo.max())
    ndeg = get_nbr_deg_freedom(h_bo)
    n = np.zeros((lm + 2, mm + 2, nlay))
    u = np.zeros_like(n)
    v = np.zeros_like(n)
    n[:, :, 0] = 1.0 * np.exp(-yy ** 2 / (0.1 * mm) ** 2)
    for iy in range(2, mm + 2):  # 1-based iy = 2 : mm + 1
        u[:, iy - 1, :] = (n[:, iy, :] - n[:, iy - 2, :]) / (2.0 * dl) * grav / abs(fcor) * (-1.0)
    init = np.stack([n, u, v], axis=3)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, [rhon], [0.0], dt_s, 2.0, 0.0, 0.0, 0.0, 0.2, 0.0, 0.1, 10.0,
                        10.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0, [0, 0], "@DIR@", "@DIR@",
                        "Test-case for barotropic instability")
    return Case("unstable_jet", lm, mm, nlay, ndeg, text, {"h_bo": h_bo, "init": init}, {})


def sill_exchange3D(lx: float = 50.0e3, ly: float = 200.0e3, dt_s: float = 30.0) -> Case:
    """testcases/sill_exchange3D.m:6-155."""
    ocrp = 1
    hmax, dt_r, bdrg, nlay = 700.0, 0.0, 0.0, 2
    fcor = 0.00014087
    dl = 400.0
    lm = _mround(lx / dl)
    mm = _mround(ly / dl)
    lm += (lm % 2 == 0)
    mm += (mm % 2 == 0)
    npts = 15
    grav = 9.8
    rhon = [1027.47, 1027.75]
    hsill = 400.0
    topl = [0.0, 0.1428]
    xs = (np.arange(1, lm + 3) - 1.5) * dl
    ys = (np.arange(1, mm + 3) - 1.5) * dl
    yy = np.broadcast_to(ys[None, :], (lm + 2, mm + 2)).copy()
    xx = np.broadcast_to(xs[:, None], (lm + 2, mm + 2)).copy()
    xx = xx - xx.mean()
    yy = yy - yy.mean()
    h_bo = hsill * np.exp(-(yy / (50.0 * dl)) ** 2)
    h_bo = hmax - h_bo
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    n = np.zeros((lm + 2, mm + 2, nlay))
    u = np.zeros_like(n)
    v = np.zeros_like(n)
    hsal = 5.0
    hmin = hsal / 10.0
    k = _mround(0.6 * (mm + 2))
    n[:, :k, 1] = 0.0
    n[:, k:, 1] = np.minimum(-h_bo[:, k:] + hmax - 100.0 + 4.0 * hsal, 0.0)
    nort = np.zeros((lm + 2, mm + 2, 3))
    sout = np.zeros((lm + 2, mm + 2, 3))
    dt = 0.5 * dl / cext
    widt = npts * dl
    for j in range(1, mm + 3):
        xpos = j - 1.5 + npts - mm
        xpos = max(xpos, 0.0)
        xpos = min(xpos, npts - 0.5)
        nort[:, j - 1, 1:3] = dt * cext / widt * xpos / (npts - xpos)
        nort[:, j - 1, 0] = dt / (31.0 * 24.0 * 3600.0) * xpos / npts
    for j in range(1, mm + 3):
        xpos = npts - (j - 1.5)
        xpos = max(xpos, 0.0)
        xpos = min(xpos, npts - 0.5)
        sout[:, j - 1, 1:3] = dt * cext / widt * xpos / (npts - xpos)
        sout[:, j - 1, 0] = dt / (31.0 * 24.0 * 3600.0) * xpos / npts
    nudg = np.maximum(np.maximum(np.zeros_like(nort), nort), sout)
    nudg[0, :, :] = 0.0
    nudg[-1, :, :] = 0.0
    init = np.stack([n, u, v], axis=3)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, rhon, topl, dt_s, 0.01, dt_r, 0.0, 0.0, 0.9, bdrg, hmin, 5.0, 5.0,
                        1.0, 1.0, 1.0, ocrp, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "Test-case: 3D sill exchange")
    return Case("sill_exchange3D", lm, mm, nlay, ndeg, text, {"init": init, "h_bo": h_bo, "nudg": nudg}, {"hsal": hsal})


def conservation(outc: float = 0.0, topo: bool = True, lx: float = 600.0e3, dl: float = 10.0e3, dt_s: float = 50.0) -> Case:
    """testcases/conservation.m:8-98 (two layers, doubly periodic, no forcing, plain forward-backward)."""
    xper = yper = 1
    nlay = 2
    fcor = 1.0e-4
    rhon = [1000.0 + 30.0 * k / nlay for k in range(1, nlay + 1)]
    topl = [(k - 1) / nlay for k in range(1, nlay + 1)]
    hfla, grav = 200.0, 9.8
    lm = _mround(lx / dl)
    lm += (lm % 2 == 0)
    mm = lm
    a = np.arange(1, lm + 3) - 1.5
    xs = (a - a.mean()) * dl
    ys = xs.copy()
    xx = np.broadcast_to(xs[:, None], (lm + 2, mm + 2))
    yy = np.broadcast_to(ys[None, :], (lm + 2, mm + 2))
    h_bo = hfla * np.ones((lm + 2, mm + 2))
    if topo:
        if outc:
            h_bo = hfla - 0.5 * (2.0 - topl[-1] - topl[-2]) * hfla * np.exp(-(xx ** 2 + yy ** 2) / (0.25 * lx) ** 2)
        else:
            h_bo = hfla - 0.5 * (1.0 - topl[-1]) * hfla * np.exp(-(xx ** 2 + yy ** 2) / (0.25 * lx) ** 2)
    h_bo = np.array(h_bo)
    h_bo[0, :] = 0.0
    h_bo[-1, :] = 0.0
    h_bo[:, 0] = 0.0
    h_bo[:, -1] = 0.0
    ndeg = get_nbr_deg_freedom(h_bo)
    n = np.zeros((lm + 2, mm + 2, nlay))
    n[:, :, 0] = 1.0 * np.exp(-(xx ** 2 + yy ** 2) / (lx / 12.0) ** 2)
    init = np.stack([n, np.zeros_like(n), np.zeros_like(n)], axis=3)
    cext = math.sqrt(grav * h_bo.max())
    files = {"init": init}
    if topo:
        files["h_bo"] = h_bo
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, rhon, topl, dt_s, 0.5, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 10.0, 10.0,
                        0.0, 1.0, 0.0, outc, 0.0, xper, yper, 1.0, [0, 0], "@DIR@", "@DIR@",
                        "Test-case for integral conservation of properties")
    return Case("conservation", lm, mm, nlay, ndeg, text, files, {})


def soliton(dl: float = 20.0e3, dt_s: float = 60.0) -> Case:
    """testcases/soliton.m:8-93 -- the equatorial Rossby soliton of Lavelle & Thacker (2008) after Boyd (1980): one layer
    of 1 m equivalent depth on an equatorial beta plane (fcor.bin), periodic in x, no dissipation.  ``info`` carries the
    asymptotic westward phase speed -(1/3 + 0.395 B^2) sqrt(g H) the soliton should travel at."""
    rhon = 1029.0
    lx = 0.5 * 12.24e6
    ly = 1.5 * 2.04e6
    H = 1.0
    grav = 9.8
    a_rd = 6371.0e3
    omeg = 2.0 * math.pi / (24.0 * 3600.0)
    x_0 = 0.0
    parB = 0.394
    parA = 0.772 * parB ** 2
    Elam = 4.0 * omeg ** 2 * a_rd ** 2 / grav / H
    L_ls = a_rd / Elam ** 0.25
    lm = _mround(lx / dl)
    mm = _mround(ly / dl)
    if lm % 2 == 0:
        lm += 1
    if mm % 2 == 0:
        mm += 1
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = H
    ndeg = get_nbr_deg_freedom(h_bo)
    xs = (np.arange(1, lm + 3) - 1.5) * dl
    ys = (np.arange(1, mm + 3) - 1.5) * dl
    xx = np.broadcast_to(xs[:, None], (lm + 2, mm + 2)).copy()
    yy = np.broadcast_to(ys[None, :], (lm + 2, mm + 2)).copy()
    xx -= xx.mean()
    yy -= yy.mean()
    sech2 = 1.0 / np.cosh(parB * (xx - x_0) / L_ls) ** 2
    gauss = np.exp(-yy ** 2 / (2.0 * L_ls ** 2))
    n = parA * H * sech2 * (6.0 * yy ** 2 + 3.0 * L_ls ** 2) / (4.0 * L_ls ** 2) * gauss
    u = parA * math.sqrt(grav * H) * sech2 * (6.0 * yy ** 2 - 9.0 * L_ls ** 2) / (4.0 * L_ls ** 2) * gauss
    v = -2.0 * parA * parB * math.sqrt(grav * H) * np.tanh(parB * (xx - x_0) / L_ls) * sech2 * 2.0 * yy / L_ls * gauss
    init = np.zeros((lm + 2, mm + 2, 1, 3))
    init[:, :, 0, 0], init[:, :, 0, 1], init[:, :, 0, 2] = n, u, v
    cext = math.sqrt(grav * (h_bo + n).max())
    deld = 0.25 * ly / 40.0e6
    beta = 2.0 * omeg * math.sin(math.radians(deld)) - 2.0 * omeg * math.sin(math.radians(-deld))
    beta = beta / (2.0 * deld * 40.0e6 / 360.0)
    fcor = beta * yy
    text = print_params(lm, mm, 1, ndeg, dl, cext, 0.0, [rhon], [0.0], dt_s, 2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.05, 10.0, 10.0,
                        1.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "Test-case for equatorial soliton")
    speed = -(1.0 / 3.0 + 0.395 * parB ** 2) * math.sqrt(grav * H)  # Boyd (1980), to first order in the amplitude
    return Case("soliton", lm, mm, 1, ndeg, text, {"h_bo": h_bo, "fcor": fcor, "init": init},
                {"speed": speed, "amplitude": float(n.max()), "dl": dl, "x0_index": int(np.argmax(n[:, (mm + 2) // 2]))})


def _frs(n: int, npts: int, dt: float, cext: float, widt: float, upper: bool) -> np.ndarray:
    """The flow-relaxation coefficient of the scripts' sponge loops (e.g. wave_sponge.m:57-86), Eq. 29 of Modave et al.
    2010: ``dt cext / widt * xpos / (npts - xpos)`` with ``xpos`` the clipped depth (grid points) into the sponge, for
    the 1-based index 1..n+2 along one axis; ``upper`` = the sponge sits at the high-index end."""
    i = np.arange(1, n + 3, dtype=np.float64)
    xpos = (i - 1.5 + npts - n) if upper else (npts - (i - 1.5))
    xpos = np.minimum(np.maximum(xpos, 0.0), npts - 0.5)
    return dt * cext / widt * xpos / (npts - xpos)


def _weak(n: int, npts: int, dt: float, upper: bool) -> np.ndarray:
    """The one-month relaxation ``dt / (31 d) * xpos / npts`` of mixed_open_bc.m:64, 71, 78 / sill_exchange3D.m."""
    i = np.arange(1, n + 3, dtype=np.float64)
    xpos = (i - 1.5 + npts - n) if upper else (npts - (i - 1.5))
    xpos = np.minimum(np.maximum(xpos, 0.0), npts - 0.5)
    return dt / (31.0 * 24.0 * 3600.0) * xpos / npts


def _centred_axes(lm: int, mm: int, dl: float):
    """[xx, yy] = meshgrid((1:lm+2)' - 1.5, (1:mm+2) - 1.5); xx = xx' * dl; ...; xx = xx - mean(xx(:))."""
    xs = (np.arange(1, lm + 3) - 1.5) * dl
    ys = (np.arange(1, mm + 3) - 1.5) * dl
    xx = np.broadcast_to(xs[:, None], (lm + 2, mm + 2)).copy()
    yy = np.broadcast_to(ys[None, :], (lm + 2, mm + 2)).copy()
    return xx - xx.mean(), yy - yy.mean()


def _dry_margins(h_bo: np.ndarray) -> np.ndarray:
    h_bo[:, 0] = 0.0
    h_bo[:, -1] = 0.0
    h_bo[0, :] = 0.0
    h_bo[-1, :] = 0.0
    return h_bo


def baines_ridge(scale: float = 1.0, dt_s: float = 10.0) -> Case:
    """testcases/baines_ridge.m:8-153 -- rotating two-layer flow over a cosine ridge (Baines & Leonard 1989): channel
    periodic in y (mm = 1), uniform inflow U_0 held by east/west sponges, body force balancing f U_0.  ``info`` holds the
    steady analytical lower-layer thickness (their Eq. 5.1-5.5).  ``scale`` < 1 shortens the domain (150 Rossby radii)."""
    hmax, U_0, nlay = 110.0, 1.2, 2
    rhon = [1025.0, 1030.0]
    topl = [0.0, 1.0 / 1.1]
    fcor, grav = 1.0e-4, 9.8
    gp = grav * (rhon[1] - rhon[0]) / rhon[1]
    d_0 = hmax * (1.0 - topl[1])
    Lros = math.sqrt(gp * d_0) / abs(fcor)
    lx = 150.0 * scale * Lros
    dl = Lros / 5.0
    lm = _mround(lx / dl)
    lm += (lm % 2 == 0)
    npts, mm = 15, 1
    H_m = 0.1
    F_0 = U_0 / math.sqrt(gp * d_0)
    dksi = dl / Lros
    xx, yy = _centred_axes(lm, mm, dl)
    X = xx[:, 1] / Lros
    h_bo = H_m * d_0 * np.cos(math.pi * xx / (10.0 * Lros))
    h_bo[(xx < -5.0 * Lros) | (xx > 5.0 * Lros)] = 0.0
    H = h_bo[:, 1] / d_0
    h_bo = _dry_margins(hmax - h_bo)
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    n = np.zeros((lm + 2, mm + 2, nlay))
    u = np.ones((lm + 2, mm + 2, nlay)) * U_0
    v = np.zeros_like(n)
    bodf = np.zeros((nlay, 2))
    bodf[:, 1] = fcor * U_0
    dt = 0.5 * dl / cext
    widt = npts * dl
    nudg = np.zeros((lm + 2, mm + 2, 3))
    prof = np.maximum(_frs(lm, npts, dt, cext, widt, True), _frs(lm, npts, dt, cext, widt, False))
    nudg[:, :, 0] = prof[:, None]
    nudg[:, :, 1] = prof[:, None]
    # Eq. 5.1-5.5 of Baines & Leonard (1989), baines_ridge.m:100-138
    Dtil = np.zeros(lm + 2)
    if F_0 < 1.0:
        r = math.sqrt(1.0 - F_0 ** 2)
        for ix in range(lm + 2):
            Dtil[ix] = (-H[ix] / (1.0 - F_0 ** 2) + 0.5 * (1.0 - F_0 ** 2) ** (-1.5)
                        * (dksi * np.sum(np.exp((X[ix] - X[ix:]) / r) * H[ix:])
                           + dksi * np.sum(np.exp(-(X[ix] - X[:ix + 1]) / r) * H[:ix + 1])))
    elif F_0 > 1.0:
        r = math.sqrt(F_0 ** 2 - 1.0)
        for ix in range(lm + 2):
            Dtil[ix] = H[ix] / (F_0 ** 2 - 1.0) + (F_0 ** 2 - 1.0) ** (-1.5) * dksi * np.sum(H[:ix + 1] * np.sin((X[:ix + 1] - X[ix]) / r))
    d_an = (1.0 + Dtil) * d_0
    u_an = U_0 * d_0 / d_an
    init = np.stack([n, u, v], axis=3)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, rhon, topl, dt_s, 0.2, 0.0, 0.0, 0.0, 0.0, 0.0, 0.1, 10.0, 10.0,
                        1.0, 1.0, 0.0, 0.0, 0, 0.0, 1.0, 0.0, [0, 0], "@DIR@", "@DIR@", "Test-case for flow over a ridge")
    return Case("baines_ridge", lm, mm, nlay, ndeg, text, {"h_bo": h_bo, "init": init, "nudg": nudg, "bodf": bodf},
                {"d_an": d_an, "u_an": u_an, "F_0": F_0, "Lros": Lros, "X": X, "d_0": d_0, "npts": npts})


def carrier_beach(dt_s: float = 0.08, mesh: float = 50.0) -> Case:
    """testcases/carrier_beach.m:11-172 -- wetting and drying: a mound of water released on a sloping beach (Carrier &
    Greenspan 1958).  One layer with outcropping (``ocrp = 1``: the shoreline is where the layer thins to Salmon's
    thickness), sponge on the western open boundary.  ``info`` holds the analytical shoreline track (their Eq. 3.23-3.29):
    ``t_sl`` (s) and ``x_sl`` (m from the mean shoreline).  ``mesh`` = grid points per length scale l_0 (50 in the script)."""
    alph, l_0, epsi = 1.0e-3, 3.0e3, 0.1
    dl = l_0 / mesh
    hhti = 3.0 * epsi * alph * l_0
    l_x = 10.0 * l_0 + hhti / alph
    lm = _mround(l_x / dl)
    mm = 1
    hsal = 0.2 * alph * dl
    hmin = hsal / 10.0
    ix_0 = int(math.ceil(0.5 * (mm + 2)))  # 1-based
    nlay, grav = 1, 9.8
    v_0 = math.sqrt(grav * l_0 * alph)
    T = v_0 / (alph * grav)
    p = 1.0 / 8.0 / (1.0 + epsi)
    npts = 15
    h_bo = np.repeat((np.arange(lm + 1, -1, -1, dtype=np.float64) * dl * alph)[:, None], mm + 2, axis=1)
    h_bo[:npts, ix_0 - 1] = h_bo[npts - 1, ix_0 - 1]
    h_bo[:, 0] = 0.0
    h_bo[:, -1] = 0.0
    h_bo[-1, :] = 0.0
    h_bo[0, :] = 0.0
    ndeg = get_nbr_deg_freedom(h_bo)
    col = h_bo[:, ix_0 - 1]
    i_sl = int(np.argmin(np.abs(col - hhti)))  # first index of the minimum, 0-based
    hhti = float(col[i_sl])
    topl = hhti / h_bo.max()
    cext = math.sqrt(grav * h_bo.max())
    xref = np.arange(1, lm + 3, dtype=np.float64) * dl
    xref = xref - xref[i_sl]
    sigm = 10.0 - 0.1 * np.arange(101, dtype=np.float64)  # (10 : -1.e-1 : 0)'
    x = 0.25 * epsi * math.exp(2.0) * p ** 2 * sigm ** 4 * np.exp(-sigm ** 2 * p) - sigm ** 2 / 16.0
    eta = 0.25 * epsi * p ** 2 * math.exp(2.0) * sigm ** 4 * np.exp(-sigm ** 2 * p)
    n1 = np.interp(xref, x * l_0, eta * alph * l_0, left=np.nan, right=np.nan)  # x is increasing as sigm decreases
    n1 = np.nan_to_num(n1, nan=0.0)
    n = np.repeat(n1[:, None], mm + 2, axis=1)[:, :, None]
    # shoreline, Eq. 3.23-3.29 (carrier_beach.m:93-114)
    dlam = 1.0e-3
    lamb = np.arange(0, 20001, dtype=np.float64) * dlam
    lmid = 0.5 * (lamb[:-1] + lamb[1:])
    E_la = np.concatenate([[0.0], np.cumsum(np.exp(0.25 * lmid ** 2)) * dlam])
    f_la = (-lamb ** 2 + 0.5 * lamb ** 4 + np.exp(-0.25 * lamb ** 2) * E_la * (lamb + lamb ** 3 - 0.25 * lamb ** 5)) / 16.0
    dfdl = (-lamb + 3.0 * lamb ** 3 - 0.25 * lamb ** 5
            + np.exp(-0.25 * lamb ** 2) * E_la * (1.0 + 2.5 * lamb ** 2 - 7.0 * lamb ** 4 / 4.0 + 0.125 * lamb ** 6)) / 16.0
    v_sl = math.sqrt(math.pi * p) * epsi * math.exp(2.0) * dfdl
    t_sl = 0.25 * lamb / math.sqrt(p) - math.sqrt(math.pi * p) * epsi * math.exp(2.0) * dfdl
    x_sl = -v_sl ** 2 / 16.0 + epsi * math.exp(2.0) * math.sqrt(math.pi) * 0.25 * f_la
    dt = 0.5 * dl / cext
    widt = npts * dl
    nudg = np.zeros((lm + 2, mm + 2, 3))
    prof = _frs(lm, npts, dt, cext, widt, False)
    nudg[:, :, 0] = prof[:, None]
    nudg[:, :, 1] = prof[:, None]
    init = np.stack([n, np.zeros_like(n), np.zeros_like(n)], axis=3)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, 0.0, [1030.0], [topl], dt_s, 5.0e-4, 0.0, 0.0, 0.0, 0.0, 0.0, hmin, 10.0,
                        10.0, 1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@",
                        "Test-case for wave on sloping beach")
    return Case("carrier_beach", lm, mm, nlay, ndeg, text, {"h_bo": h_bo, "init": init, "nudg": nudg},
                {"t_sl": t_sl * T, "x_sl": x_sl * l_0, "xref": xref, "hsal": hsal, "hhti": hhti, "T": T, "l_0": l_0})


def _millot_crepon(lm: int, hfla: float, topl, rhon, tauw, f0: float, dl: float, grav: float = 9.8) -> dict:
    """Steady two-layer response to a seaward wind (Millot & Crepon 1981), upwelling_seaward_wind.m:57-81: interface
    displacement and along-shore velocity against the distance from the coast, for undisturbed thicknesses from topl."""
    h1 = (topl[1] - topl[0]) * hfla
    h2 = (1.0 - topl[1]) * hfla
    gpri = grav * (rhon[1] - rhon[0]) / rhon[1]
    r_1 = math.sqrt(grav * hfla) / f0
    r_2 = math.sqrt(gpri * h1 * h2 / hfla) / abs(f0)
    x = (np.arange(1, lm + 3) - 1.5) * dl
    t_eta2 = np.exp(-x / r_2) * (-1.0) * tauw[0] / rhon[0] / (r_1 * f0 ** 2) * h2 / hfla * (-1.0) * r_1 / r_2
    t_eta1 = (-1.0) * tauw[0] / rhon[0] / (r_1 * f0 ** 2) * (np.exp(-x / r_1) + h2 / h1 * r_2 / r_1 * np.exp(-x / r_2))
    t_v1 = tauw[0] / rhon[0] / f0 / hfla * h2 / h1 * (np.exp(-x / r_2) - 1.0)
    t_v2 = t_v1 * (h1 / h2) * (-1.0)
    return {"x": x, "t_eta": np.stack([t_eta1, t_eta2]), "t_v": np.stack([t_v1, t_v2]), "r_1": r_1, "r_2": r_2}


def upwelling_seaward_wind(lm: int = 200, dt_s: float = 6.0) -> Case:
    """testcases/upwelling_seaward_wind.m:10-36 -- two layers, coast on the western edge, uniform seaward wind ``tauw``
    ramped over ``dt_r`` = 4 days, channel periodic in y (mm = 1).  No input file at all: parameters only."""
    mm, nlay, dl, f0 = 1, 2, 1.0e3, 1.0e-4
    dt_o, hfla = 0.16667, 40.0
    topl = [0.0, 0.5]
    rhon = [1028.95, 1030.0]
    tauw = [0.1, 0.0]
    grav = 9.8
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = hfla
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    text = print_params(lm, mm, nlay, ndeg, dl, cext, f0, rhon, topl, dt_s, dt_o, 4.0, 0.0, 0.0, 0.0, 0.0, 1.0, 10.0, 10.0,
                        1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, tauw, "@DIR@", "@DIR@", "Test-case for upwelling seaward wind")
    return Case("upwelling_seaward_wind", lm, mm, nlay, ndeg, text, {}, _millot_crepon(lm, hfla, topl, rhon, tauw, f0, dl))


def mixed_open_bc(lm: int = 200, mm: int = 100, dt_s: float = 6.0, mcbc: float = 0.0) -> Case:
    """testcases/mixed_open_bc.m:12-118 -- the seaward-wind upwelling in a basin with a coast (west), a wave sponge
    (east) and weakly relaxed open boundaries (north, south).  The script predates the ``mcbc`` switch of shared_mod.f95:71
    (it was written for a model that always applied no_gradient_obc, private_mod.f95:2201, 2285): with the shipped
    ``mcbc = 1`` the north/south walls are closed, the coastal jet piles up in the north-west corner and the top layer
    drains after about three days, so the case sets ``mcbc = 0`` (appended to the block; print_params.m has no slot)."""
    nlay, dl, f0 = 2, 1.0e3, 1.0e-4
    dt_o, hfla = 0.16667, 40.0
    topl = [0.0, 0.5]
    rhon = [1028.95, 1030.0]
    tauw = [0.1, 0.0]
    grav, npts = 9.8, 15
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = hfla
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    dt = 0.5 * dl / cext
    widt = npts * dl
    east = np.zeros((lm + 2, mm + 2, 3))
    e = _frs(lm, npts, dt, cext, widt, True)
    east[:, :, 0] = e[:, None]
    east[:, :, 1] = e[:, None]
    east[:, :, 2] = _weak(lm, npts, dt, True)[:, None]
    nort = np.zeros((lm + 2, mm + 2, 3))
    sout = np.zeros((lm + 2, mm + 2, 3))
    nort[:, :, :] = _weak(mm, npts, dt, True)[None, :, None]
    sout[:, :, :] = _weak(mm, npts, dt, False)[None, :, None]
    nudg = np.maximum(np.maximum(east, nort), sout)
    nudg[0, :, :] = 0.0
    text = print_params(lm, mm, nlay, ndeg, dl, cext, f0, rhon, topl, dt_s, dt_o, 4.0, 0.0, 0.0, 0.0, 0.0, 0.001, 1.0, 1.0,
                        1.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, tauw, "@DIR@", "@DIR@",
                        "Test-case: Upwelling seaward wind with mixed open boundary conditions",
                        extra={"mcbc": "%#0.0f" % mcbc})
    return Case("mixed_open_bc", lm, mm, nlay, ndeg, text, {"nudg": nudg}, _millot_crepon(lm, hfla, topl, rhon, tauw, f0, dl))


def morel_upwelling(dl: float = 1.0e3) -> Case:
    """testcases/morel_upwelling.m:13-63 -- upwelling driven by an along-shore wind up to and past the outcropping of the
    interface (Morel, Darr & Talandier 2006): two layers, periodic in x (lm = 1), coast at y = y_max, the wind applied
    as a body force on the top layer (bodf.bin), ``ocrp = 1``.  ``info`` has the constants of the analytical solution."""
    nlay = 2
    topl = [0.0, 0.5]
    rhon = [1015.0, 1030.0]
    hfla, hsal, f0, grav = 50.0, 0.5, 1.0e-4, 9.8
    rext = math.sqrt(grav * hfla) / abs(f0)
    ly = 1.0 * rext
    mm = _mround(ly / dl)
    lm = 1
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = hfla
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    hmin = hsal / 10.0
    H_1 = topl[1] * hfla
    H_2 = (1.0 - topl[1]) * hfla
    R_d = math.sqrt(grav * (rhon[1] - rhon[0]) / rhon[1] * H_1 * H_2 / hfla) / abs(f0)
    delt = H_1 / H_2
    T_w = 0.05 / rhon[0] / H_1
    t_o = abs(f0) * R_d * (1.0 + delt) / T_w
    dt_s = float(_mround(3.0 * t_o / 3600.0 / 24.0))
    bodf = np.zeros((nlay, 2))
    bodf[0, 0] = T_w
    text = print_params(lm, mm, nlay, ndeg, dl, cext, f0, rhon, topl, dt_s, 0.05 * dt_s, 0.0, 0.0, 0.0, 0.0, 0.0, hmin, 10.0,
                        10.0, 1.0, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, [0, 0], "@DIR@", "@DIR@",
                        "Test-case: Upwelling in presence of outcrop")
    return Case("morel_upwelling", lm, mm, nlay, ndeg, text, {"bodf": bodf},
                {"H_1": H_1, "H_2": H_2, "R_d": R_d, "delt": delt, "T_w": T_w, "t_o": t_o, "f0": f0, "dl": dl})


def outcrop_seamount(lx: float = 600.0e3, dl: float = 5.0e3, nlay: int = 5, threed: bool = False, dt_s: float = 15.0) -> Case:
    """testcases/outcrop_seamount.m:9-104 -- a stratified state of rest over a seamount and sloping coasts with
    isopycnals outcropping into the bottom (Salmon 2002): h_0 from the Newton solve must stay at rest (du/dt = 0).
    ``threed`` selects the script's commented 3-D option (mm = lm)."""
    fcor = 1.0e-4
    if nlay == 1:
        rhon, topl = [1000.0], [0.0]
    else:
        rhon = [1000.0 + 30.0 * k / (nlay - 1) for k in range(nlay)]
        topl = [k / nlay for k in range(nlay)]
    grav, hmax = 9.8, 300.0
    lm = _mround(lx / dl)
    lm += (lm % 2 == 0)
    mm = lm if threed else 1
    xi = np.arange(1, lm + 3, dtype=np.float64)
    yi = np.arange(1, mm + 3, dtype=np.float64)
    xx = np.broadcast_to(xi[:, None], (lm + 2, mm + 2)).copy()
    yy = np.broadcast_to(yi[None, :], (lm + 2, mm + 2)).copy()
    xx = (xx - xx.mean()) * dl
    yy = (yy - yy.mean()) * dl
    ix_0 = int(math.ceil(0.5 * (mm + 2))) - 1
    h_bo = np.sqrt(xx ** 2 + yy ** 2) / dl
    h_bo = h_bo / h_bo[:, ix_0].max() * hmax
    h_bo = h_bo[:, ix_0].max() - h_bo
    smnt = hmax - np.exp(-(xx ** 2 + yy ** 2) / 50.0e3 ** 2) * 0.75 * hmax
    h_bo = np.minimum(h_bo, smnt)
    h_bo[h_bo < h_bo[:, ix_0].min()] = 0.0
    dhdx = np.zeros((lm + 2, mm + 2))
    dhdy = np.zeros((lm + 2, mm + 2))
    dhdx[1:-1, :] = (h_bo[2:, :] - h_bo[:-2, :]) / (2.0 * dl)
    dhdy[:, 1:-1] = (h_bo[:, 2:] - h_bo[:, :-2]) / (2.0 * dl)
    slop = np.sqrt(dhdx ** 2 + dhdy ** 2)
    hsal = 1.0 * slop[h_bo > 1.0e-3].max() * dl * (rhon[1] - rhon[0]) / rhon[1] if nlay > 1 else 1.0
    hmin = hsal / 10.0
    dryd = 10.0 * hsal + (nlay - 1) * hsal
    h_bo[h_bo < dryd] = 0.0
    h_bo[0, :] = 0.0
    h_bo[-1, :] = 0.0
    h_bo[:, -1] = 0.0
    h_bo[:, 0] = 0.0
    cext = math.sqrt(grav * h_bo.max())
    ndeg = get_nbr_deg_freedom(h_bo)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, rhon, topl, dt_s, 0.2, 0.0, 0.0, 0.0, 0.0, 0.0, hmin, 10.0, 10.0,
                        1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@",
                        "Test-case for state of rest allowing isopycnal outcrop")
    return Case("outcrop_seamount", lm, mm, nlay, ndeg, text, {"h_bo": h_bo}, {"hsal": hsal})


def _m2_tide(lm: int, mm: int) -> np.ndarray:
    """tide.bin of sill_exchange2Dtides.m:84-97 / tide_ridge.m:73-92: an M2 current of 0.1 m/s along x starting from
    rest (phase pi/2); the constituent frequency (rad/day) sits in element (1,1,1,1,1)."""
    tide = np.zeros((2, 1, lm + 2, mm + 2, 3))
    tide[0, 0, :, :, 1] = 0.1
    tide[1, 0, :, :, 1] = math.pi / 2.0
    tide[1, 0, :, :, 2] = math.pi / 2.0
    tide[0, 0, 0, 0, 0] = 2.0 * math.pi / (12.4206012 / 24.0)
    return tide


def _sill2d(tides: bool, lx: float, dl: float, dt_s: float) -> Case:
    ocrp, hmax, dt_r, bdrg, nlay = 1, 700.0, 0.0, 0.0, 2
    fcor = 0.00014087
    mm = 1
    lm = _mround(lx / dl)
    lm += (lm % 2 == 0)
    npts = 15 if tides else 200
    grav = 9.8
    rhon = [1027.47, 1027.75]
    hsill = 400.0
    topl = [0.0, 0.5] if tides else [0.0, 0.1428]
    xx, yy = _centred_axes(lm, mm, dl)
    h_bo = _dry_margins(hmax - hsill * np.exp(-(xx / (200.0 * dl)) ** 2))
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    dhdx = np.zeros((lm + 2, mm + 2))
    dhdx[2:-2, :] = h_bo[3:-1, :] - h_bo[1:-3, :]
    dhdx = dhdx / (2.0 * dl)
    n = np.zeros((lm + 2, mm + 2, nlay))
    hsal = 20.0 * dhdx.max() * dl * (rhon[-1] - rhon[0]) / rhon[0]
    hmin = hsal / 10.0
    if tides:  # sill_exchange2Dtides.m:60-66
        k = _mround(0.5 * (lm + 2))
        n[:k, :, 1] = 0.5 * hmax - 4.0 * hsal - 100.0
        n[k:, :, 1] = (-0.5 * hmax + 4.0 * hsal + 450.0 + (hmax - h_bo[k:, 1]))[:, None]
        n[k:, :, 1] = np.minimum(n[k:, :, 1], 0.5 * hmax - 4.0 * hsal - 100.0)
    else:      # sill_exchange2D.m:67-69
        k = _mround(0.6 * (lm + 2))
        n[:k, :, 1] = 0.0
        n[k:, :, 1] = (-h_bo[k:, 1] + 500.0 + 4.0 * hsal)[:, None]
    dt = 0.5 * dl / cext
    widt = npts * dl
    nudg = np.zeros((lm + 2, mm + 2, 3))
    prof = np.maximum(_frs(lm, npts, dt, cext, widt, True), _frs(lm, npts, dt, cext, widt, False))
    nudg[:, :, 0] = prof[:, None]
    nudg[:, :, 1] = prof[:, None]
    init = np.stack([n, np.zeros_like(n), np.zeros_like(n)], axis=3)
    files = {"init": init, "h_bo": h_bo, "nudg": nudg}
    if tides:
        files["tide"] = _m2_tide(lm, mm)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, rhon, topl, dt_s, 0.01, dt_r, 0.0, 0.0, 0.5 if tides else 0.9, bdrg,
                        hmin, 5.0, 5.0, 1.0, 1.0, 1.0, ocrp, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@",
                        "Test-case for tidal flow over a ridge" if tides else "Test-case: 2D sill exchange")
    return Case("sill_exchange2Dtides" if tides else "sill_exchange2D", lm, mm, nlay, ndeg, text, files, {"hsal": hsal})


def sill_exchange2D(lx: float = 200.0e3, dl: float = 100.0, dt_s: float = 3.0) -> Case:
    """testcases/sill_exchange2D.m:7-137 -- two-layer exchange over a Gaussian sill in an x-z channel (mm = 1) with
    200-point sponges at both ends, outcropping, quadratic drag switch on (bdrg = 0)."""
    return _sill2d(False, lx, dl, dt_s)


def sill_exchange2Dtides(lx: float = 100.0e3, dl: float = 100.0, dt_s: float = 10.0) -> Case:
    """testcases/sill_exchange2Dtides.m:7-166 -- the 2-D sill with a lock-exchange initial state, 15-point sponges and an
    M2 tidal current in the relaxation target (tide.bin)."""
    return _sill2d(True, lx, dl, dt_s)


def tide_ridge(lm: int = 500, ocrp: int = 1, dt_s: float = 3.0) -> Case:
    """testcases/tide_ridge.m:8-156 -- semidiurnal tidal flow over a Gaussian ridge in 30 m of water: seven layers cut from
    an exponential density profile (three without outcropping), no rotation, M2 current in the sponges' target ramped
    over 0.75 dt_s days."""
    hmax = 30.0
    dt_r = 0.75 * dt_s
    bdrg = 0.0
    nlay = 7 if ocrp == 1 else 3
    fcor, dl, mm = 0.0, 100.0, 1
    lm += (lm % 2 == 0)
    npts, grav = 15, 9.8
    z = np.arange(0.0, hmax + 0.5, 1.0)
    rhop = 1028.0 - 6.0 * np.exp(-z / (0.2 * hmax))
    drho = (rhop.max() - rhop.min()) / nlay
    rhon = rhop.min() + (np.arange(nlay) + 0.5) * drho
    xx, yy = _centred_axes(lm, mm, dl)
    h_bo = _dry_margins(hmax - 0.75 * hmax * np.exp(-(xx / (75.0 * dl)) ** 2))
    ndeg = get_nbr_deg_freedom(h_bo)
    cext = math.sqrt(grav * h_bo.max())
    dhdx = np.zeros((lm + 2, mm + 2))
    dhdx[2:-2, :] = h_bo[3:-1, :] - h_bo[1:-3, :]
    dhdx = dhdx / (2.0 * dl)
    hsal = 20.0 * dhdx.max() * dl * (rhon[-1] - rhon[0]) / rhon[0]
    hmin = hsal / 10.0
    topl = np.interp(rhop.min() + np.arange(nlay) * drho, rhop, z)
    if topl[1] < 10.0 * hsal:
        topl[1:] = topl[1:] + (10.0 * hsal - topl[1])
    topl = topl / hmax
    dt = 0.5 * dl / cext
    widt = npts * dl
    nudg = np.zeros((lm + 2, mm + 2, 3))
    prof = np.maximum(_frs(lm, npts, dt, cext, widt, True), _frs(lm, npts, dt, cext, widt, False))
    nudg[:, :, 0] = prof[:, None]
    nudg[:, :, 1] = prof[:, None]
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, rhon, topl, dt_s, 0.01, dt_r, 0.0, 0.0, 0.5, bdrg, hmin, 5.0, 5.0,
                        1.0, 1.0, 1.0, ocrp, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "Test-case for tidal flow over a ridge")
    return Case("tide_ridge", lm, mm, nlay, ndeg, text, {"h_bo": h_bo, "nudg": nudg, "tide": _m2_tide(lm, mm)},
                {"hsal": hsal, "omega": 2.0 * math.pi / (12.4206012 / 24.0)})


def wave_sponge(lx: float = 600.0e3, dl: float = 10.0e3, npts: int = 15, dt_s: float = 0.25) -> Case:
    """testcases/wave_sponge.m:12-116 -- a Gaussian mound adjusting in a rotating two-layer basin whose four sides are
    flow-relaxation sponges (eta and the normal velocity): the radiated gravity waves must leave without reflection."""
    ly = lx
    nlay = 2
    fcor = 1.0e-4
    rhon = [1000.0, 1030.0]
    topl = [0.0, 0.5]
    hfla, grav = 200.0, 9.8
    lm = _mround(lx / dl)
    mm = _mround(ly / dl)
    lm += (lm % 2 == 0)
    mm += (mm % 2 == 0)
    lm += 2 * npts
    mm += 2 * npts
    n = np.zeros((lm + 2, mm + 2, nlay))
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = hfla
    cext = math.sqrt(grav * hfla)
    ndeg = get_nbr_deg_freedom(h_bo)
    xs = np.arange(1, lm + 3) - 1.5
    ys = np.arange(1, mm + 3) - 1.5
    xx = np.broadcast_to((xs - np.median(xs))[:, None], (lm + 2, mm + 2)) * dl
    yy = np.broadcast_to((ys - np.median(ys))[None, :], (lm + 2, mm + 2)) * dl
    n[:, :, 0] = 1.0 * np.exp(-(xx ** 2 + yy ** 2) / 50.0e3 ** 2)
    dt = 0.5 * dl / cext
    widt = npts * dl
    ew = np.maximum(_frs(lm, npts, dt, cext, widt, True), _frs(lm, npts, dt, cext, widt, False))
    ns = np.maximum(_frs(mm, npts, dt, cext, widt, True), _frs(mm, npts, dt, cext, widt, False))
    nudg = np.zeros((lm + 2, mm + 2, 3))
    nudg[:, :, 0] = np.maximum(ew[:, None], ns[None, :])
    nudg[:, :, 1] = ew[:, None]
    nudg[:, :, 2] = ns[None, :]
    init = np.stack([n, np.zeros_like(n), np.zeros_like(n)], axis=3)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, fcor, rhon, topl, dt_s, 4.17e-3, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 10.0, 10.0,
                        1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "Test-case for wave sponge")
    return Case("wave_sponge", lm, mm, nlay, ndeg, text, {"init": init, "nudg": nudg}, {"npts": npts, "rhon": rhon, "hfla": hfla})


def synthetic_basin(n: int = 8192, nlay: int = 4, seed: int = 20261018, dt_s: float = 1.0, wind: bool = True,
                    mm: int | None = None, sponge: bool = False) -> Case:
    """The throughput workload of BASELINE.json / SURVEY.md section 8(d): an n x n closed flat basin,
    1 km mesh, ``nlay`` layers, Leith viscosity every step, generalized forward-backward, wind stress
    0.1 cos(pi y / L) Pa, initial surface bump + interface noise.  Not a reference script.
    ``sponge``: the option set of sill_exchange3D on the same grid instead of the wind -- sponges on eta, u, v over the
    64 southernmost and northernmost rows (nudg.bin, cosine ramp up to 0.02 per step) and outcropping switched on (ocrp = 1)."""
    lm = n
    mm = n if mm is None else mm
    dl = 1000.0
    rng = np.random.default_rng(seed)
    rhon = [1025.0 + k for k in range(nlay)]
    topl = [0.0, 0.1, 0.25, 0.5, 0.6, 0.7, 0.8, 0.9][:nlay]
    h_bo = np.zeros((lm + 2, mm + 2), dtype=np.float32)
    h_bo[1:-1, 1:-1] = 1.0
    ndeg = (lm + 1) * (mm + 1)  # full rectangle: get_nbr_deg_freedom(h_bo) == (lm+1)(mm+1)
    L = lm * dl
    xs = ((np.arange(lm + 2) - 0.5) * dl - 0.5 * L).astype(np.float64)
    ys = ((np.arange(mm + 2) - 0.5) * dl - 0.5 * mm * dl).astype(np.float64)
    init = np.zeros((lm + 2, mm + 2, nlay, 3), dtype=np.float32)
    r2 = xs[:, None] ** 2 + ys[None, :] ** 2
    init[:, :, 0, 0] = (0.5 * np.exp(-r2 / (0.1 * L) ** 2)).astype(np.float32)
    for k in range(nlay):
        init[:, :, k, 0] += rng.uniform(-0.01, 0.01, size=(lm + 2, mm + 2)).astype(np.float32)
    files = {"init": init}
    if sponge:
        wind = False
        w = max(2, min(64, mm // 4))
        j = np.arange(mm + 2)
        dist = np.minimum(j, mm + 1 - j).astype(np.float64)  # rows from the nearest of the two ends
        coef = np.where(dist < w, 0.01 * (1.0 + np.cos(math.pi * dist / w)), 0.0).astype(np.float32)
        nudg = np.zeros((lm + 2, mm + 2, 3), dtype=np.float32)
        nudg[1:-1, :, :] = coef[None, :, None]
        files["nudg"] = nudg
    if wind:
        taus = np.zeros((lm + 2, mm + 2, 2), dtype=np.float32)
        taus[:, :, 0] = (0.1 * np.cos(math.pi * (ys + 0.5 * mm * dl) / (mm * dl)))[None, :].astype(np.float32)
        files["taus"] = taus
    text = print_params(lm, mm, nlay, ndeg, dl, 198.0, 1.0e-4, rhon, topl, dt_s, dt_s, 0.0, 0.0, 0.0, 0.2, 0.0, 1.0, 10.0, 10.0,
                        1.0, 1.0, 0.0, 1.0 if sponge else 0.0, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@",
                        "synthetic %dx%dx%d basin%s" % (lm, mm, nlay, " with N/S sponges and outcropping" if sponge else ""))
    return Case("synthetic_basin", lm, mm, nlay, ndeg, text, files, {})


def sponge_basin(nlay: int = 3, lm: int = 40, mm: int = 14, dt_s: float = 0.2, ocrp: float = 0.0) -> Case:
    """A small closed basin with a thickness sponge everywhere and tilted interfaces.  Not a reference
    script: it exists to exercise the update_h epilogues of private_mod1d/3d/plumenew.f95, which relax
    toward fixed thicknesses east of lm/2 and toward the initial state west of it."""
    dl = 2000.0
    depth = 900.0
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = depth
    ndeg = get_nbr_deg_freedom(h_bo)
    xs = (np.arange(lm + 2) - 0.5) / lm
    init = np.zeros((lm + 2, mm + 2, nlay, 3))
    for k in range(1, nlay):
        init[:, :, k, 0] = (20.0 * k * (xs - 0.5))[:, None] + 3.0 * np.cos(np.arange(mm + 2) * 0.7)[None, :]
    init[:, :, 0, 0] = 0.05 * np.sin(6.0 * xs)[:, None]
    nudg = np.zeros((lm + 2, mm + 2, 3))
    nudg[1:-1, 1:-1, 0] = 0.02
    nudg[0:2, 1:-1, 1] = 0.01  # a nudged western face: the reference insists on one open-boundary segment (pm:1226-1231)
    rhon = [1026.0 + 0.5 * k for k in range(nlay)]
    topl = [0.0, 0.3, 0.6, 0.8][:nlay]
    cext = math.sqrt(9.8 * depth)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, 1.0e-4, rhon, topl, dt_s, dt_s, 0.0, 0.0, 0.0, 0.2, 0.0, 1.0, 10.0, 10.0,
                        1.0, 1.0, 0.0, ocrp, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "sponge basin (variants)")
    return Case("sponge_basin", lm, mm, nlay, ndeg, text, {"h_bo": h_bo, "init": init, "nudg": nudg}, {})


def option_basin(nlay: int = 3, lm: int = 36, mm: int = 22, dt_s: float = 0.05, wind: bool = True, sponge: bool = True,
                 bodf: bool = False, hdot: bool = False, tide: bool = False, beta: bool = False, ocrp: float = 0.0,
                 dt_r: float = 0.0, f0: float = 1.0e-4) -> Case:
    """A small closed basin with switchable forcing files.  Not a reference script: it exists to drive the hot-path
    branches no named config reaches -- wind inside a sponge (the Ekman term of the relaxation target,
    private_mod.f95:1449-1452, 1534-1537), the dt_r ramp (:1898-1901), body force (bodf.bin), thickness source
    (hdot.bin), tidal targets (tide.bin, :1453-1454, 1538-1539, 1633-1634), a beta plane (fcor.bin), more than four
    layers, and wind stress spread over outcropping layers."""
    dl = 2000.0
    depth = 600.0
    h_bo = np.zeros((lm + 2, mm + 2))
    h_bo[1:-1, 1:-1] = depth
    h_bo[lm // 2 - 2:lm // 2 + 2, mm // 2 - 1:mm // 2 + 2] = 0.0  # an island: coasts inside the tiles
    ndeg = get_nbr_deg_freedom(h_bo)
    xs = (np.arange(lm + 2) - 0.5) / lm
    ys = (np.arange(mm + 2) - 0.5) / mm
    init = np.zeros((lm + 2, mm + 2, nlay, 3))
    init[:, :, 0, 0] = 0.04 * np.sin(5.0 * xs)[:, None] * np.cos(3.0 * ys)[None, :]
    for k in range(1, nlay):
        init[:, :, k, 0] = (6.0 * k * (xs - 0.5))[:, None] + 1.5 * np.cos(4.0 * ys + k)[None, :]
    files = {"h_bo": h_bo, "init": init}
    if sponge:
        nudg = np.zeros((lm + 2, mm + 2, 3))
        ramp = np.clip((np.arange(lm + 2) - (lm - 7)) / 8.0, 0.0, 1.0)  # an eastern sponge, eta, u and v
        for c3 in range(3):
            nudg[:, 1:-1, c3] = (0.03 * ramp)[:, None]
        nudg[0:2, 1:-1, 1] = 0.01  # the reference insists on one open-boundary segment (pm:1226-1231)
        files["nudg"] = nudg
    if wind:
        taus = np.zeros((lm + 2, mm + 2, 2))
        taus[:, :, 0] = (0.08 * np.cos(math.pi * ys))[None, :]
        taus[:, :, 1] = (0.03 * np.sin(2.0 * math.pi * xs))[:, None]
        files["taus"] = taus
    if bodf:
        b = np.zeros((nlay, 2))
        b[:, 0] = [1.0e-6 * (k + 1) for k in range(nlay)]
        b[:, 1] = [-5.0e-7 * (k + 1) for k in range(nlay)]
        files["bodf"] = b
    if hdot:
        hd = np.zeros((lm + 2, mm + 2, nlay))
        hd[3:9, 3:9, 0] = 2.0e-5
        hd[3:9, 3:9, nlay - 1] = -2.0e-5
        files["hdot"] = hd
    if tide:
        td = np.zeros((2, 1, lm + 2, mm + 2, 3))
        td[0, 0, :, :, 0] = 0.02  # amplitude of eta, u, v; phase below
        td[0, 0, :, :, 1] = 0.004
        td[0, 0, :, :, 2] = 0.002
        td[1, 0, :, :, 0] = (2.0 * math.pi * xs)[:, None]
        td[1, 0, :, :, 1] = (1.0 + 2.0 * math.pi * ys)[None, :]
        td[1, 0, :, :, 2] = 0.5
        td[0, 0, 0, 0, 0] = 2.0 * math.pi * 24.0 / 12.42  # omega, rad/day (element (1,k,0,0,1), tide_ridge.m:79-93)
        files["tide"] = td
    if beta:
        fc = np.zeros((lm + 2, mm + 2))
        fc[:, :] = (f0 + 2.0e-11 * (ys - 0.5) * mm * dl)[None, :]
        files["fcor"] = fc
    rhon = [1026.0 + 0.4 * k for k in range(nlay)]
    topl = [k / (nlay + 0.5) for k in range(nlay)]
    cext = math.sqrt(9.8 * depth)
    text = print_params(lm, mm, nlay, ndeg, dl, cext, f0, rhon, topl, dt_s, dt_s, dt_r, 0.0, 0.0, 0.2, 0.0, 1.0, 10.0, 10.0,
                        1.0, 1.0, 0.0, ocrp, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "option basin")
    return Case("option_basin", lm, mm, nlay, ndeg, text, files, {})


def random_coast(seed: int = 1, lm: int = 70, mm: int = 44, nlay: int = 3, land: float = 0.25, wind: bool = True, sponge: bool = False,
                 ocrp: float = 0.0, dt_s: float = 0.05) -> Case:
    """A basin with a random coastline (smoothed noise thresholded at the ``land`` quantile: bays, islands, one-cell
    channels, isolated lakes), tilted interfaces and optional wind / eastern sponge.  Not a reference script: it exists to
    drive every mask combination (mk_n, mk_u, mk_v, mkpe, mkpi) through the kernels' masked rows."""
    rng = np.random.default_rng(seed)
    dl, depth = 2000.0, 500.0
    z = rng.standard_normal((lm + 2, mm + 2))
    for _ in range(3):  # cheap smoothing: bays and islands a few cells wide, with ragged one-cell features left over
        z = 0.25 * (np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) + 0.35 * z
    wet = z > np.quantile(z, land)
    h_bo = np.where(wet, depth, 0.0)
    if sponge:
        h_bo[lm - 9:lm + 1, 1:-1] = depth  # open water under the sponge
    h_bo = _dry_margins(h_bo)
    ndeg = get_nbr_deg_freedom(h_bo)
    xs = (np.arange(lm + 2) - 0.5) / lm
    ys = (np.arange(mm + 2) - 0.5) / mm
    init = np.zeros((lm + 2, mm + 2, nlay, 3))
    init[:, :, 0, 0] = 0.05 * np.sin(7.0 * xs)[:, None] * np.cos(5.0 * ys)[None, :]
    for k in range(1, nlay):
        init[:, :, k, 0] = (8.0 * k * (xs - 0.5))[:, None] + 2.0 * np.cos(6.0 * ys + k)[None, :]
    init[:, :, :, 1] = 0.02 * rng.standard_normal((lm + 2, mm + 2, nlay))
    init[:, :, :, 2] = 0.02 * rng.standard_normal((lm + 2, mm + 2, nlay))
    files = {"h_bo": h_bo, "init": init}
    if wind:
        taus = np.zeros((lm + 2, mm + 2, 2))
        taus[:, :, 0] = (0.08 * np.cos(math.pi * ys))[None, :]
        taus[:, :, 1] = (0.04 * np.sin(2.0 * math.pi * xs))[:, None]
        files["taus"] = taus
    if sponge:
        nudg = np.zeros((lm + 2, mm + 2, 3))
        ramp = np.clip((np.arange(lm + 2) - (lm - 8)) / 8.0, 0.0, 1.0)
        for c3 in range(3):
            nudg[:, 1:-1, c3] = (0.03 * ramp)[:, None]
        files["nudg"] = nudg
    rhon = [1026.0 + 0.5 * k for k in range(nlay)]
    topl = [k / (nlay + 0.5) for k in range(nlay)]
    text = print_params(lm, mm, nlay, ndeg, dl, math.sqrt(9.8 * depth), 1.0e-4, rhon, topl, dt_s, dt_s, 0.0, 0.0, 0.0, 0.2, 1.0e-3, 1.0,
                        10.0, 10.0, 1.0, 1.0, 1.0, ocrp, 0.0, 0.0, 0.0, 0.0, [0, 0], "@DIR@", "@DIR@", "random coastline %d" % seed)
    return Case("random_coast", lm, mm, nlay, ndeg, text, files, {"wet_fraction": float(wet.mean())})


CASES = {
    "stommel1948": stommel1948,
    "lock_exchange": lock_exchange,
    "unstable_jet": unstable_jet,
    "sill_exchange3D": sill_exchange3D,
    "conservation": conservation,
    "soliton": soliton,
    "baines_ridge": baines_ridge,
    "carrier_beach": carrier_beach,
    "upwelling_seaward_wind": upwelling_seaward_wind,
    "mixed_open_bc": mixed_open_bc,
    "morel_upwelling": morel_upwelling,
    "outcrop_seamount": outcrop_seamount,
    "sill_exchange2D": sill_exchange2D,
    "sill_exchange2Dtides": sill_exchange2Dtides,
    "tide_ridge": tide_ridge,
    "wave_sponge": wave_sponge,
    "synthetic_basin": synthetic_basin,
    "sponge_basin": sponge_basin,
    "option_basin": option_basin,
    "random_coast": random_coast,
}
