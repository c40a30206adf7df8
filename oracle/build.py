"""Build recipe of the CPU checker (TEST INFRASTRUCTURE, see beom_oracle.h):

  oracle/libbeom_oracle.so        strict IEEE (-O2 -ffp-contract=off): the parity reference
  oracle/libbeom_oracle_omp.so    the same source, -O3 -fopenmp: the CPU baseline that bench.py times (contracts a*b+c into
                                  FMAs like the reference's own -Ofast build: NOT bit-identical to the strict one)
  oracle/libbeom_oracle_omp_strict.so  strict IEEE + OpenMP (-O2 -ffp-contract=off -fopenmp): the parallel loops hold no
                                  reductions, so it returns the strict library's bits; for parity tests at native sizes

There is no oracle/_ref: the reference is Fortran and no Fortran compiler exists in this environment
(SURVEY.md section 0), so the reference's own sources cannot be compiled here.
"""
from __future__ import annotations

import os
import subprocess
import sys

ORACLE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(ORACLE)


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.run(cmd, check=True)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ORACLE, "beom_oracle.c")
    deps = [src, os.path.join(ORACLE, "beom_oracle.h"), os.path.join(ROOT, "include", "beom_gpu.h")]
    out = os.path.join(ORACLE, "libbeom_oracle.so")
    if force or _newer(out, deps):
        _run(["gcc", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-fPIC", "-shared", src, "-o", out, "-lm"])
    omp = os.path.join(ORACLE, "libbeom_oracle_omp.so")
    if force or _newer(omp, deps):
        _run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-Wall", "-fPIC", "-shared", src, "-o", omp, "-lm"])
    strict = os.path.join(ORACLE, "libbeom_oracle_omp_strict.so")
    if force or _newer(strict, deps):
        _run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-Wall", "-fPIC", "-shared", src, "-o", strict, "-lm"])
    return out




if __name__ == "__main__":
    build_oracle(force="--force" in sys.argv)
