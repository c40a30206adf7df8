"""Runs in its own process: one reference test case (small version) on the CPU EMULATION of the library (tools/emu: the
kernels' source compiled by g++; a launch is a serial loop over the grid, or, for the kernels whose threads cooperate --
the fused step above all -- a run on the SIMT emulator) against the CPU oracle.  Test infrastructure: the emulated library
is loaded here explicitly, under its own file name; the product never sees it.

usage: emu_worker.py <libbeom_gpu_emu.so> <case> <nsteps> <fused 0|1> ['{"param": "value", ...}' ['{"kwarg": ...}' [variant]]]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from beom_b200 import _lib  # noqa: E402

emu = _lib.bind_gpu(C.CDLL(sys.argv[1], mode=C.RTLD_LOCAL))
assert b"cpu-emulation" in emu.beom_gpu_version()
_lib.host_lib()      # the host driver (read_input_data) first: it links the real libbeom_gpu.so, which stays unused here
_lib._gpu = emu      # from here on model.GpuModel talks to the emulation

sys.argv = [sys.argv[0]] + sys.argv[2:]
import runpy  # noqa: E402

runpy.run_path(os.path.join(ROOT, "tests", "case_worker.py"), run_name="__main__")
