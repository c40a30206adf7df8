"""Host-side mirror of the reference's `run` for Python callers: parameter block -> read_input_data
(C++ host driver) -> GPU library.  Everything numerical happens in the two native libraries; this
module only moves pointers."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Fields, Options, Params


def parse_params(text: str, variant: int = 0):
    """shared_mod.f95 text (or a print_params block) -> (Params, idir, odir, desc)."""
    lib = _lib.host_lib()
    p = Params()
    bufs = [C.create_string_buffer(1024) for _ in range(3)]
    rc = lib.beom_params_parse(text.encode(), C.byref(p), bufs[0], bufs[1], bufs[2], 1024)
    if rc:
        raise ValueError(_lib.host_error())
    p.variant = variant
    return p, bufs[0].value.decode(), bufs[1].value.decode(), bufs[2].value.decode()


class HostModel:
    """read_input_data (private_mod.f95:105-250) done by the C++ host driver."""

    _PLANES = {"mk_u": 1, "mk_v": 1, "mk_n": 1, "mkpe": 1, "mkpi": 1, "fcor": 1, "h_th": 1, "nudg": 3, "taus": 2,
               "Ow": 1, "Os": 1, "Osum_": 1, "pi_s": 1}

    def __init__(self, params: Params, idir: str = "", odir: str = "", desc: str = ""):
        self.lib = _lib.host_lib()
        self.params = params
        self.nlay, self.ndeg = params.nlay, params.ndeg
        self.h = self.lib.beom_host_create(C.byref(params), idir.encode(), odir.encode(), desc.encode())
        if not self.h:
            raise RuntimeError(_lib.host_error())

    @classmethod
    def from_block(cls, path: str, variant: int = 0, write_outputs: bool = False):
        with open(path) as f:
            p, idir, odir, desc = parse_params(f.read(), variant)
        return cls(p, idir, odir if write_outputs else "", desc)

    def close(self):
        if self.h:
            self.lib.beom_host_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def array(self, name: str):
        p = self.lib.beom_host_array(self.h, name.encode())
        if not p:
            return None
        nd1 = self.ndeg + 1
        if name in ("h_0", "hlay", "u", "v", "hdot"):
            planes = self.nlay
        elif name == "fnud":
            planes = 3 * self.nlay
        elif name == "tide":
            return np.ctypeslib.as_array(p, shape=(3, nd1, 2))
        elif name == "bodf":
            return np.ctypeslib.as_array(p, shape=(2, self.nlay))
        else:
            planes = self._PLANES[name]
        return np.ctypeslib.as_array(p, shape=(planes, nd1))

    def iarray(self, name: str):
        p = self.lib.beom_host_iarray(self.h, name.encode())
        if not p:
            return None
        nd1 = self.ndeg + 1
        if name == "neig":
            return np.ctypeslib.as_array(p, shape=(nd1, 8))
        if name == "subc":
            return np.ctypeslib.as_array(p, shape=(2, nd1))
        if name == "segm":
            return np.ctypeslib.as_array(p, shape=(18, int(self.scalar("nseg"))))
        return np.ctypeslib.as_array(p, shape=(nd1,))

    def scalar(self, name: str) -> float:
        return self.lib.beom_host_scalar(self.h, name.encode())

    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.beom_host_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def fields(self) -> Fields:
        f = Fields()
        self.lib.beom_host_fields(self.h, C.byref(f))
        return f

    def run(self, options: Options | None = None, max_steps: int = 0) -> int:
        """run() of private_mod.f95:99-103: time loop + outputs, steps executed on the GPU."""
        opt = options or default_options()
        rc = self.lib.beom_host_run(self.h, C.byref(opt), max_steps)
        if rc:
            raise RuntimeError(_lib.host_error())
        return rc


def default_options(fused: bool = True, rank: int = 0, nranks: int = 1, device: int = -1) -> Options:
    o = Options()
    _lib.gpu_lib().beom_gpu_default_options(C.byref(o))
    o.fused = 1 if fused else 0
    o.rank, o.nranks, o.device = rank, nranks, device
    return o


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_double_p)


class GpuModel:
    """The C ABI of include/beom_gpu.h for Python callers (tests, bench)."""

    def __init__(self, params: Params, fields: Fields, options: Options | None = None):
        self.lib = _lib.gpu_lib()
        self.params = params
        self.nlay, self.ndeg = params.nlay, params.ndeg
        self.opt = options or default_options()
        rc = self.lib.beom_gpu_init(C.byref(params), C.byref(fields), C.byref(self.opt))
        if rc:
            raise RuntimeError("beom_gpu_init: %s" % _lib.gpu_error())

    GRID_FILES = ("h_bo", "init", "nudg", "taus", "fcor", "hdot", "bodf", "tide")

    @classmethod
    def from_grids(cls, params: Params, idir: str, options: Options | None = None):
        """read_input_data's grid-shaped work on the device (beom_gpu_init_grids): the raw files of ``idir`` are memory-mapped
        and handed over as they are; masks, vector numbering, rest thickness, targets, forcing and the initial state are
        built in HBM.  Returns None when the case needs the host path (periodic, rigid lid, tides, ...): the caller then uses
        HostModel + GpuModel(...) + upload_state as before."""
        self = cls.__new__(cls)
        self.lib = _lib.gpu_lib()
        self.params = params
        self.nlay, self.ndeg = params.nlay, params.ndeg
        self.opt = options or default_options()
        gr = _lib.Grids()
        keep = []
        for k in cls.GRID_FILES:
            path = os.path.join(idir, k + ".bin")
            if os.path.exists(path):
                a = np.memmap(path, dtype="<f4", mode="r")
                keep.append(a)
                setattr(gr, k, C.cast(a.ctypes.data, C.POINTER(C.c_float)))
        gr.has_h_to = int(os.path.exists(os.path.join(idir, "h_to.bin")))
        rc = self.lib.beom_gpu_init_grids(C.byref(params), C.byref(gr), C.byref(self.opt))
        del keep
        if rc == 1:  # BEOM_GRIDS_UNSUPPORTED
            return None
        if rc:
            raise RuntimeError("beom_gpu_init_grids: %s" % _lib.gpu_error())
        return self

    def download_subc(self):
        """Grid coordinates (i, j) of the vector points this rank holds (point_range: first .. first + count - 1)."""
        first, count, _, _ = self.point_range()
        si, sj = np.zeros(count, dtype=np.int32), np.zeros(count, dtype=np.int32)
        self._ck(self.lib.beom_gpu_download_subc(si.ctypes.data_as(C.POINTER(C.c_int32)), sj.ctypes.data_as(C.POINTER(C.c_int32))), "download_subc")
        return si, sj

    def debug_static(self, name: str, index: int = 0):
        """A static plane in the reference's vector layout (0:ndeg) -- for checks of the device-side initialisation."""
        out = np.zeros(self.ndeg + 1)
        self._ck(self.lib.beom_gpu_debug_static(name.encode(), index, _dp(out)), "debug_static")
        return out

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError("%s: %s" % (what, _lib.gpu_error()))

    @property
    def path(self) -> str:
        return self.lib.beom_gpu_path().decode()

    @property
    def fused_variant(self) -> str:
        return self.lib.beom_gpu_fused_variant().decode()

    def point_range(self):
        """(first, count, own_first, own_count) of the vector points on this rank."""
        v = [C.c_int() for _ in range(4)]
        self._ck(self.lib.beom_gpu_point_range(*[C.byref(x) for x in v]), "point_range")
        return tuple(x.value for x in v)

    def set_window(self, first, count):
        self._ck(self.lib.beom_gpu_set_window(first, count), "set_window")
        self._win = count

    def upload_state(self, hlay, u, v):
        npts = getattr(self, "_win", 0) or (self.ndeg + 1)
        for a in (hlay, u, v):
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.nlay * npts
        self._ck(self.lib.beom_gpu_upload_state(_dp(hlay), _dp(u), _dp(v)), "upload_state")

    def stress(self):
        self._ck(self.lib.beom_gpu_stress(), "stress")

    def step(self, tstp, ctim, ramp, gene, upst, first_three):
        self._ck(self.lib.beom_gpu_step(tstp, ctim, ramp, gene, int(upst), int(first_three)), "step")

    def advance(self, tstp0, tstp1, tres=0.0):
        self._ck(self.lib.beom_gpu_advance(tstp0, tstp1, tres), "advance")

    def sync(self):
        self._ck(self.lib.beom_gpu_sync(), "sync")

    def download_state(self, out=None):
        shape = (self.nlay, getattr(self, "_win", 0) or (self.ndeg + 1))
        out = out or tuple(np.zeros(shape) for _ in range(3))
        self._ck(self.lib.beom_gpu_download_state(_dp(out[0]), _dp(out[1]), _dp(out[2])), "download_state")
        return out

    def download_aux(self):
        nd1 = self.ndeg + 1
        h_u, h_v = np.zeros((self.nlay, nd1)), np.zeros((self.nlay, nd1))
        rs_h, dmdx, dmdy = np.zeros((self.nlay, nd1, 2)), np.zeros((self.nlay, nd1, 3)), np.zeros((self.nlay, nd1, 3))
        self._ck(self.lib.beom_gpu_download_aux(_dp(h_u), _dp(h_v), _dp(rs_h), _dp(dmdx), _dp(dmdy)), "download_aux")
        return h_u, h_v, rs_h, dmdx, dmdy

    def download_diag(self, which=("pvor", "mont", "v_cc")):
        """The float32 diagnostic records of write_array (private_mod.f95:2884-2974): dict of (nlay, ndeg) arrays."""
        out = {k: np.zeros((self.nlay, self.ndeg), dtype=np.float32) for k in which}
        ptr = lambda k: out[k].ctypes.data_as(C.POINTER(C.c_float)) if k in out else None
        self._ck(self.lib.beom_gpu_download_diag(ptr("pvor"), ptr("mont"), ptr("v_cc")), "download_diag")
        return out

    def diagnostics(self, h_0):
        """Conservation integrals (testcases/conservation.m:116-211): (vol[nlay], ke[nlay], sum of eta_1^2)."""
        h_0 = np.ascontiguousarray(h_0, dtype=np.float64)
        vol, ke, pe = np.zeros(self.nlay), np.zeros(self.nlay), np.zeros(1)
        self._ck(self.lib.beom_gpu_diagnostics(_dp(h_0), _dp(vol), _dp(ke), _dp(pe)), "diagnostics")
        return vol, ke, float(pe[0])

    def diagnostics_all(self, h_0):
        """diagnostics() plus the vorticity integrals of conservation.m:169-211: a dict with vol, ke, enst, zeta, zeta2
        (per layer), pe (sum of eta_1^2) and npts (points counted); enst / npts, zeta / npts are the script's `enst`, `rvor`."""
        h_0 = np.ascontiguousarray(h_0, dtype=np.float64)
        o = {k: np.zeros(self.nlay) for k in ("vol", "ke", "enst", "zeta", "zeta2")}
        pe, npts = np.zeros(1), np.zeros(1)
        self._ck(self.lib.beom_gpu_diagnostics_all(_dp(h_0), _dp(o["vol"]), _dp(o["ke"]), _dp(pe), _dp(o["enst"]), _dp(o["zeta"]),
                                                   _dp(o["zeta2"]), _dp(npts)), "diagnostics_all")
        o["pe"], o["npts"] = float(pe[0]), float(npts[0])
        return o

    def set_rest_thickness(self, h_0_r4):
        """h_0.bin's content ([nlay][ndeg] float32): needed once before records_begin()."""
        a = np.ascontiguousarray(h_0_r4, dtype=np.float32).reshape(self.nlay, self.ndeg)
        self._ck(self.lib.beom_gpu_set_rest_thickness(a.ctypes.data_as(C.POINTER(C.c_float))), "set_rest_thickness")

    def records_begin(self, with_diag=False):
        """Enqueue the float32 output records of the current state (eta_, u___, v___ [, pvor, mont, v_cc]) and their
        asynchronous copy to the host; stepping may go on at once."""
        self._ck(self.lib.beom_gpu_records_begin(1 if with_diag else 0), "records_begin")

    def records_wait(self):
        """The oldest begun record set: dict of [nlay][count] float32 arrays (copies) + first_point, hmin, hmax, thin_layer."""
        r = _lib.Records()
        self._ck(self.lib.beom_gpu_records_wait(C.byref(r)), "records_wait")
        out = {"first_point": r.first_point, "count": r.count, "thin_layer": r.thin_layer,
               "hmin": np.array(r.hmin[:self.nlay]), "hmax": np.array(r.hmax[:self.nlay])}
        for k in ("eta", "u", "v", "pvor", "mont", "v_cc"):
            ptr = getattr(r, k)
            out[k] = np.ctypeslib.as_array(ptr, shape=(self.nlay, r.count)).copy() if ptr else None
        return out

    def pi_iterations(self) -> int:
        """Sweeps of the last rigid-lid pressure solve."""
        n = C.c_int(0)
        self._ck(self.lib.beom_gpu_pi_iterations(C.byref(n)), "pi_iterations")
        return int(n.value)

    def download_pi_s(self):
        out = np.zeros(self.ndeg + 1)
        self._ck(self.lib.beom_gpu_download_pi_s(_dp(out)), "download_pi_s")
        return out

    def mark(self, which):
        self._ck(self.lib.beom_gpu_mark(which), "mark")

    def elapsed_ms(self) -> float:
        ms = C.c_double()
        self._ck(self.lib.beom_gpu_elapsed_ms(C.byref(ms)), "elapsed_ms")
        return ms.value

    def launch_count(self) -> int:
        return int(self.lib.beom_gpu_launch_count())

    def graph_launch_count(self) -> int:
        """Steps that ran as one CUDA-graph launch (beom_gpu_graph_launch_count)."""
        return int(self.lib.beom_gpu_graph_launch_count())

    def pinned(self, shape) -> np.ndarray:
        """A float64 array in page-locked host memory (beom_gpu_host_alloc); freed at close()."""
        n = int(np.prod(shape))
        p = self.lib.beom_gpu_host_alloc(n * 8)
        if not p:
            raise MemoryError(_lib.gpu_error())
        self._pinned = getattr(self, "_pinned", []) + [p]
        return np.ctypeslib.as_array(C.cast(p, _lib.c_double_p), shape=(n,)).reshape(shape)

    def close(self):
        for p in getattr(self, "_pinned", []):
            self.lib.beom_gpu_host_free(p)
        self._pinned = []
        self.lib.beom_gpu_finalize()
