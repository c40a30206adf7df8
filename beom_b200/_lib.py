"""ctypes view of the C ABI (include/beom_gpu.h) and of the host driver (csrc/host/beom_host.h).

The shared libraries are built in-tree by ``beom_b200.build`` (``__graft_entry__.build()``).  There is
no fallback of any kind: if ``libbeom_gpu.so`` is missing, loading fails with an instruction to build.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.environ.get("BEOM_LIBDIR") or os.path.join(HERE, "lib")
MAXLAY = 16

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class Params(C.Structure):
    """struct beom_params"""
    _fields_ = [
        ("lm", C.c_int32), ("mm", C.c_int32), ("nlay", C.c_int32), ("ndeg", C.c_int32),
        ("dl", C.c_double), ("cext", C.c_double), ("f0", C.c_double),
        ("rhon", C.c_double * MAXLAY), ("topl", C.c_double * MAXLAY),
        ("dt_s", C.c_double), ("dt_o", C.c_double), ("dt_r", C.c_double), ("dt3d", C.c_double),
        ("bvis", C.c_double), ("dvis", C.c_double), ("bdrg", C.c_double), ("hmin", C.c_double),
        ("hsbl", C.c_double), ("hbbl", C.c_double),
        ("g_fb", C.c_double), ("uadv", C.c_double), ("qdrg", C.c_double), ("ocrp", C.c_double),
        ("rsta", C.c_double), ("xper", C.c_double), ("yper", C.c_double), ("diag", C.c_double),
        ("rgld", C.c_double), ("mcbc", C.c_double),
        ("tauw", C.c_double * 2),
        ("svis", C.c_double), ("tdrg", C.c_double), ("topt", C.c_double), ("plum", C.c_double),
        ("dt", C.c_double), ("hsal", C.c_double), ("hdry", C.c_double), ("tole", C.c_double),
        ("pi", C.c_double), ("grav", C.c_double), ("rho0", C.c_double), ("beta", C.c_double),
        ("epsi", C.c_double), ("gamm", C.c_double), ("del1", C.c_double), ("del2", C.c_double),
        ("sor", C.c_double),
        ("itmx", C.c_int32), ("nsal", C.c_int32), ("variant", C.c_int32), ("reserved_", C.c_int32),
    ]


class Fields(C.Structure):
    """struct beom_fields"""
    _fields_ = [
        ("neig", c_int32_p), ("subc", c_int32_p),
        ("mk_u", c_double_p), ("mk_v", c_double_p), ("mk_n", c_double_p), ("mkpe", c_double_p), ("mkpi", c_double_p),
        ("fcor", c_double_p), ("h_th", c_double_p),
        ("nudg", c_double_p), ("fnud", c_double_p), ("hdot", c_double_p), ("taus", c_double_p),
        ("tide", c_double_p), ("bodf", c_double_p),
        ("segm", c_int32_p),
        ("Ow", c_double_p), ("Os", c_double_p), ("Osum_", c_double_p), ("pi_s", c_double_p),
        ("nseg", C.c_int32), ("flag_nudging", C.c_int32),
        ("invf", C.c_double), ("w_ti", C.c_double),
    ]


class Options(C.Structure):
    """struct beom_gpu_options"""
    _fields_ = [
        ("device", C.c_int32), ("fused", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
        ("strict", C.c_int32), ("reserved_", C.c_int32 * 3),
    ]


BEOM_MAXLAY = 16


class Records(C.Structure):  # beom_records of include/beom_gpu.h
    _fields_ = [
        ("eta", C.POINTER(C.c_float)), ("u", C.POINTER(C.c_float)), ("v", C.POINTER(C.c_float)),
        ("pvor", C.POINTER(C.c_float)), ("mont", C.POINTER(C.c_float)), ("v_cc", C.POINTER(C.c_float)),
        ("first_point", C.c_int), ("count", C.c_int),
        ("hmin", C.c_double * BEOM_MAXLAY), ("hmax", C.c_double * BEOM_MAXLAY), ("thin_layer", C.c_int),
    ]


class Grids(C.Structure):  # beom_grids of include/beom_gpu.h: the raw input files (float32), NULL = absent
    _fields_ = [(k, C.POINTER(C.c_float)) for k in ("h_bo", "init", "nudg", "taus", "fcor", "hdot", "bodf", "tide")] + \
               [("has_h_to", C.c_int32)]


GPU_SYMBOLS = [
    "beom_gpu_version", "beom_gpu_abi_version", "beom_gpu_last_error", "beom_gpu_default_options",
    "beom_gpu_init", "beom_gpu_init_grids", "beom_gpu_download_subc", "beom_gpu_download_grid_files", "beom_gpu_debug_static", "beom_gpu_upload_state", "beom_gpu_stress", "beom_gpu_step", "beom_gpu_advance",
    "beom_gpu_download_state", "beom_gpu_download_aux", "beom_gpu_download_diag", "beom_gpu_download_pi_s", "beom_gpu_pi_iterations",
    "beom_gpu_diagnostics", "beom_gpu_diagnostics_all", "beom_gpu_set_rest_thickness", "beom_gpu_records_begin",
    "beom_gpu_records_wait", "beom_gpu_sync", "beom_gpu_mark", "beom_gpu_elapsed_ms", "beom_gpu_launch_count", "beom_gpu_graph_launch_count",
    "beom_gpu_path", "beom_gpu_fused_variant", "beom_gpu_point_range", "beom_gpu_set_window", "beom_gpu_host_alloc", "beom_gpu_host_free", "beom_gpu_comm_unique_id", "beom_gpu_comm_init", "beom_gpu_comm_finalize", "beom_gpu_finalize",
]

_gpu = None
_host = None


def _missing(path: str) -> OSError:
    return OSError(
        "%s not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `python -m beom_b200.build`) from the repository root. There is no CPU fallback." % path)


def gpu_lib() -> C.CDLL:
    """libbeom_gpu.so (hand-written sm_100a kernels behind the C ABI).  Raises if it is not built."""
    global _gpu
    if _gpu is not None:
        return _gpu
    path = os.path.join(LIBDIR, "libbeom_gpu.so")
    if not os.path.exists(path):
        raise _missing(path)
    _gpu = bind_gpu(C.CDLL(path, mode=C.RTLD_GLOBAL))
    return _gpu


def bind_gpu(lib: C.CDLL) -> C.CDLL:
    """Declares the prototypes of include/beom_gpu.h on a loaded library."""
    lib.beom_gpu_version.restype = C.c_char_p
    lib.beom_gpu_path.restype = C.c_char_p
    lib.beom_gpu_last_error.argtypes = [C.c_char_p, C.c_int]
    lib.beom_gpu_default_options.argtypes = [C.POINTER(Options)]
    lib.beom_gpu_default_options.restype = None
    lib.beom_gpu_init.argtypes = [C.POINTER(Params), C.POINTER(Fields), C.POINTER(Options)]
    lib.beom_gpu_upload_state.argtypes = [c_double_p] * 3
    lib.beom_gpu_step.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.beom_gpu_advance.argtypes = [C.c_int, C.c_int, C.c_double]
    lib.beom_gpu_download_state.argtypes = [c_double_p] * 3
    lib.beom_gpu_download_aux.argtypes = [c_double_p] * 5
    lib.beom_gpu_download_diag.argtypes = [c_float_p] * 3
    lib.beom_gpu_download_pi_s.argtypes = [c_double_p]
    lib.beom_gpu_pi_iterations.argtypes = [C.POINTER(C.c_int)]
    lib.beom_gpu_init_grids.argtypes = [C.POINTER(Params), C.POINTER(Grids), C.POINTER(Options)]
    lib.beom_gpu_download_subc.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.beom_gpu_debug_static.argtypes = [C.c_char_p, C.c_int, c_double_p]
    lib.beom_gpu_diagnostics.argtypes = [c_double_p] * 4
    lib.beom_gpu_diagnostics_all.argtypes = [c_double_p] * 8
    lib.beom_gpu_set_rest_thickness.argtypes = [c_float_p]
    lib.beom_gpu_records_begin.argtypes = [C.c_int]
    lib.beom_gpu_records_wait.argtypes = [C.POINTER(Records)]
    lib.beom_gpu_mark.argtypes = [C.c_int]
    lib.beom_gpu_elapsed_ms.argtypes = [c_double_p]
    lib.beom_gpu_launch_count.restype = C.c_longlong
    lib.beom_gpu_graph_launch_count.restype = C.c_longlong
    lib.beom_gpu_fused_variant.restype = C.c_char_p
    lib.beom_gpu_point_range.argtypes = [C.POINTER(C.c_int)] * 4
    lib.beom_gpu_set_window.argtypes = [C.c_int, C.c_int]
    lib.beom_gpu_host_alloc.argtypes = [C.c_size_t]
    lib.beom_gpu_host_alloc.restype = C.c_void_p
    lib.beom_gpu_host_free.argtypes = [C.c_void_p]
    lib.beom_gpu_host_free.restype = None
    lib.beom_gpu_comm_unique_id.argtypes = [C.c_char_p]
    lib.beom_gpu_comm_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    return lib


def host_lib() -> C.CDLL:
    """libbeom_host.so (parameter parser, read_input_data, write_outputs, integrate_time)."""
    global _host
    if _host is not None:
        return _host
    gpu_lib()  # libbeom_host.so calls into libbeom_gpu.so
    path = os.path.join(LIBDIR, "libbeom_host.so")
    if not os.path.exists(path):
        raise _missing(path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.beom_params_parse.argtypes = [C.c_char_p, C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    lib.beom_params_parse_file.argtypes = [C.c_char_p, C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    lib.beom_params_derive.argtypes = [C.POINTER(Params)]
    lib.beom_params_derive.restype = None
    lib.beom_params_defaults.argtypes = [C.POINTER(Params)]
    lib.beom_params_defaults.restype = None
    lib.beom_host_last_error.argtypes = [C.c_char_p, C.c_int]
    lib.beom_host_create.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_char_p]
    lib.beom_host_create.restype = C.c_void_p
    lib.beom_host_destroy.argtypes = [C.c_void_p]
    lib.beom_host_destroy.restype = None
    lib.beom_host_array.argtypes = [C.c_void_p, C.c_char_p]
    lib.beom_host_array.restype = c_double_p
    lib.beom_host_iarray.argtypes = [C.c_void_p, C.c_char_p]
    lib.beom_host_iarray.restype = c_int32_p
    lib.beom_host_scalar.argtypes = [C.c_void_p, C.c_char_p]
    lib.beom_host_scalar.restype = C.c_double
    lib.beom_host_fields.argtypes = [C.c_void_p, C.POINTER(Fields)]
    lib.beom_host_fields.restype = None
    lib.beom_host_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.beom_host_counts.restype = None
    lib.beom_host_write_outputs.argtypes = [C.c_void_p, C.c_double]
    lib.beom_host_read_restart.argtypes = [C.c_void_p]
    lib.beom_host_run.argtypes = [C.c_void_p, C.POINTER(Options), C.c_int]
    _host = lib
    return lib


def gpu_error() -> str:
    buf = C.create_string_buffer(2048)
    gpu_lib().beom_gpu_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")


def host_error() -> str:
    buf = C.create_string_buffer(2048)
    host_lib().beom_host_last_error(buf, len(buf))
    return buf.value.decode(errors="replace")
