"""ctypes binding of the CPU oracle (oracle/beom_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
beom_b200/."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(omp):
    # omp: False = strict IEEE (the parity reference), "strict" = the same bits with OpenMP, True = the -O3 timing build
    name = "libbeom_oracle_omp_strict.so" if omp == "strict" else "libbeom_oracle_omp.so" if omp else "libbeom_oracle.so"
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        raise OSError("%s is not built: run `python -m beom_b200.build`" % path)
    lib = C.CDLL(path)
    lib.beom_oracle_create.restype = C.c_void_p
    lib.beom_oracle_create.argtypes = [C.c_void_p, C.c_char_p]
    lib.beom_oracle_destroy.argtypes = [C.c_void_p]
    lib.beom_oracle_destroy.restype = None
    lib.beom_oracle_error.restype = C.c_char_p
    lib.beom_oracle_advance.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.beom_oracle_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
    lib.beom_oracle_counts.restype = None
    lib.beom_oracle_array.restype = C.POINTER(C.c_double)
    lib.beom_oracle_array.argtypes = [C.c_void_p, C.c_char_p]
    lib.beom_oracle_iarray.restype = C.POINTER(C.c_int32)
    lib.beom_oracle_iarray.argtypes = [C.c_void_p, C.c_char_p]
    lib.beom_oracle_scalar.restype = C.c_double
    lib.beom_oracle_scalar.argtypes = [C.c_void_p, C.c_char_p]
    lib.beom_oracle_nseg.argtypes = [C.c_void_p]
    lib.beom_oracle_fields.argtypes = [C.c_void_p, C.c_void_p]
    lib.beom_oracle_fields.restype = None
    lib.beom_oracle_record.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_float)]
    for n in ("distribute_stress", "update_h", "surf_pressure"):
        getattr(lib, "beom_oracle_" + n).argtypes = [C.c_void_p]
        getattr(lib, "beom_oracle_" + n).restype = None
    for n in ("update_mont", "update_viscosity", "update_u", "update_v", "no_gradient_obc"):
        getattr(lib, "beom_oracle_" + n).argtypes = [C.c_void_p, C.c_int]
        getattr(lib, "beom_oracle_" + n).restype = None
    lib.beom_oracle_set_scalars.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    lib.beom_oracle_set_scalars.restype = None
    return lib


_SHAPES = {  # name -> (planes per point-array as a function of nlay, trailing history length)
    "h_u": ("L", 1), "h_v": ("L", 1), "u": ("L", 1), "v": ("L", 1), "hlay": ("L", 1), "h_0": ("L", 1), "hdot": ("L", 1),
    "v_cc": ("L", 1), "v_ll": ("L", 1), "UU4": ("L", 1), "VV4": ("L", 1), "delu": ("L", 1), "delv": ("L", 1),
    "tt3d": ("2L", 1), "tb3d": ("2L", 1), "tu3d": ("2L", 1), "taus": (2, 1), "nudg": (3, 1), "fnud": ("3L", 1),
    "rs_h": ("L", 2), "dmdx": ("L", 3), "dmdy": ("L", 3), "tide": (3, 2),
    "rvor": (1, 1), "pvor": (1, 1), "dive": (1, 1), "fcor": (1, 1), "mk_u": (1, 1), "mk_v": (1, 1), "mk_n": (1, 1),
    "mkpe": (1, 1), "mkpi": (1, 1), "mont": (1, 1), "d2hx": (1, 1), "d2hy": (1, 1), "h_th": (1, 1), "Ow": (1, 1),
    "Os": (1, 1), "Osum_": (1, 1), "pi_s": (1, 1),
}


class Oracle:
    """read_input_data + integrate_time of the reference, on the CPU."""

    def __init__(self, params, idir: str, omp=False):
        self.lib = _load(omp)
        self.params = params
        self.nlay, self.ndeg = params.nlay, params.ndeg
        self.h = self.lib.beom_oracle_create(C.byref(params), (idir or "").encode())
        if not self.h:
            raise RuntimeError("oracle: " + self.lib.beom_oracle_error().decode())

    def close(self):
        if self.h:
            self.lib.beom_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.beom_oracle_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def advance(self, tstp0: int, tstp1: int):
        self.lib.beom_oracle_advance(self.h, tstp0, tstp1)

    def array(self, name: str) -> np.ndarray:
        """View (no copy) of a module array in the reference layout: [planes, ndeg+1(, hist)]."""
        planes, hist = _SHAPES[name]
        if isinstance(planes, str):
            planes = int(planes[:-1] or 1) * self.nlay
        p = self.lib.beom_oracle_array(self.h, name.encode())
        if not p:
            raise KeyError(name)
        n = planes * (self.ndeg + 1) * hist
        a = np.ctypeslib.as_array(p, shape=(n,))
        return a.reshape(planes, self.ndeg + 1, hist) if hist > 1 else a.reshape(planes, self.ndeg + 1)

    def iarray(self, name: str) -> np.ndarray:
        p = self.lib.beom_oracle_iarray(self.h, name.encode())
        if not p:
            raise KeyError(name)
        nd1 = self.ndeg + 1
        if name == "neig":
            return np.ctypeslib.as_array(p, shape=(nd1, 8))
        if name == "subc":
            return np.ctypeslib.as_array(p, shape=(2, nd1))
        if name == "segm":
            return np.ctypeslib.as_array(p, shape=(18, self.nseg()))
        return np.ctypeslib.as_array(p, shape=(nd1,))

    def scalar(self, name: str) -> float:
        return self.lib.beom_oracle_scalar(self.h, name.encode())

    def nseg(self) -> int:
        return self.lib.beom_oracle_nseg(self.h)

    def record(self, var: str) -> np.ndarray:
        out = np.empty((self.nlay, self.ndeg), dtype=np.float32)
        rc = self.lib.beom_oracle_record(self.h, var.encode(), out.ctypes.data_as(C.POINTER(C.c_float)))
        if rc:
            raise KeyError(var)
        return out
