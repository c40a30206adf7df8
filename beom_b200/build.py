"""In-tree build of the native code (no torch, no cmake):

  beom_b200/lib/libbeom_gpu.so    hand-written CUDA kernels + the C ABI (nvcc, sm_100a only)
  beom_b200/lib/libbeom_host.so   C++ host driver (parameter parser, read_input_data, outputs, time loop)
  beom_b200/lib/beom_run          executable equivalent of main.f95

Run as ``python -m beom_b200.build`` or through ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.environ.get("BEOM_LIBDIR") or os.path.join(HERE, "lib")  # BEOM_LIBDIR / BEOM_NVCC_DEFS: experiment builds
GPU_SRC = os.path.join(HERE, "csrc", "gpu")
HOST_SRC = os.path.join(HERE, "csrc", "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # no FMA contraction: results are bit-identical to a strict IEEE evaluation
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.run(cmd, check=True)


def _sources(d: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_gpu(force: bool = False, verbose_ptxas: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libbeom_gpu.so")
    cus = _sources(GPU_SRC, (".cu",))
    hdrs = _sources(GPU_SRC, (".cuh", ".h")) + [os.path.join(ROOT, "include", "beom_gpu.h")]
    objdir = os.path.join(ROOT, "build", "gpu")
    os.makedirs(objdir, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose_ptxas else []) + os.environ.get("BEOM_NVCC_DEFS", "").split()
    if os.environ.get("BEOM_LIBDIR"):
        objdir = os.path.join(LIBDIR, "obj")
        os.makedirs(objdir, exist_ok=True)
    # one nvcc per translation unit, in parallel (the fused-step instantiations dominate the build time)
    jobs, objs = [], []
    for cu in cus:
        obj = os.path.join(objdir, os.path.basename(cu)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [cu] + hdrs):
            flags = NVCC_FLAGS
            if cu.endswith("_fma.cu"):  # the opt-in FMA-contracted copies of the fused step (BEOM_FMA=1)
                flags = [f for f in NVCC_FLAGS if f != "-fmad=false"] + ["-fmad=true"]
            cmd = [nvcc] + flags + extra + ["-c", cu, "-o", obj]
            print("+", " ".join(cmd), file=sys.stderr, flush=True)
            log = open(obj + ".log", "w") if verbose_ptxas else None
            jobs.append((cmd, subprocess.Popen(cmd, stderr=log) if log else subprocess.Popen(cmd), log))
    failed = []
    for cmd, pr, log in jobs:
        if pr.wait() != 0:
            failed.append(cmd)
        if log:
            log.close()
    if failed:
        raise subprocess.CalledProcessError(1, failed[0])
    if force or jobs or _newer(out, objs):
        _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", out, "-ldl"])
    return out


def build_host(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libbeom_host.so")
    ccs = [os.path.join(HOST_SRC, f) for f in ("params.cc", "init.cc", "io.cc", "run.cc")]
    deps = ccs + _sources(HOST_SRC, (".h",)) + [os.path.join(ROOT, "include", "beom_gpu.h"), os.path.join(LIBDIR, "libbeom_gpu.so")]
    common = ["-O2", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-fopenmp", "-ffp-contract=off"]
    if force or _newer(out, deps):
        _run(["g++"] + common + ["-shared"] + ccs + ["-o", out, "-L" + LIBDIR, "-lbeom_gpu", "-Wl,-rpath,$ORIGIN"])
    exe = os.path.join(LIBDIR, "beom_run")
    if force or _newer(exe, deps + [os.path.join(HOST_SRC, "main.cc"), out]):
        _run(["g++"] + common + [os.path.join(HOST_SRC, "main.cc"), "-o", exe, "-L" + LIBDIR, "-lbeom_host", "-lbeom_gpu",
                                 "-Wl,-rpath,$ORIGIN"])
    return out


def build_all(force: bool = False) -> None:
    if os.environ.get("BEOM_LIBDIR") and not force and all(
            os.path.exists(os.path.join(LIBDIR, f)) for f in ("libbeom_gpu.so", "libbeom_host.so")):
        return  # an experiment build (tools/ab.sh): use it as it is, never rebuild it from the current sources
    build_gpu(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
