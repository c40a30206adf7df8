"""oracle/_ref: the reference itself, built here without a Fortran compiler (TEST INFRASTRUCTURE ONLY).

BEOM is configured at compile time: a user pastes the parameter block a test-case script prints into
``shared_mod.f95`` and rebuilds.  ``build_case`` does exactly that for one parameter block:

1. reads ``/root/reference/shared_mod.f95`` and sets the values of the user-modifiable section from the block
   (the three parameters ``private_mod.f95`` uses but the shipped ``shared_mod.f95`` never declares -- ``svis``,
   ``tdrg``, ``topt`` -- are declared next to their neighbours, as a user of this fork must do to compile it at all);
2. translates that text + ``/root/reference/private_mod.f95`` (or one of the 1d / 3d / plume variants) +
   ``/root/reference/main.f95`` to C++ with oracle/f95c (a language translator, not a restatement: it has no notion
   of what the statements mean);
3. compiles the result with ``g++ -O2 -ffp-contract=off`` (strict IEEE, the oracle's own convention) into
   ``oracle/_ref/<hash>/beom_ref`` and, with ``omp=True``, ``-O3 -fopenmp`` using the reference's own PARALLEL DO directives.

Nothing is copied from the reference into the repository: translated sources and binaries live under ``oracle/_ref/``
(git-ignored).  ``/root/reference`` exists only in the development container; on the GPU box the prebuilt binaries of
``oracle/_ref/`` travel with the snapshot and ``build_case`` returns them without looking at the sources.

``run_case`` runs the binary (the reference's ``run()``: read_input_data + integrate_time + its own output files) and
returns the harness's raw dump of every module array at STOP (``ref_dump.bin``, written by the translated program's
epilogue -- the one thing the harness adds) as numpy arrays in the reference's own shapes.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BEOM_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
VARIANT_FILES = {0: "private_mod.f95", 1: "private_mod1d.f95", 2: "private_mod3d.f95", 3: "private_modplumenew.f95"}  # beom_gpu.h: BEOM_VARIANT_*
# private_modplumenew.f95 reads its local Lnud (:1643, :1666) before the only assignment to it (:1645, inside the
# branch the first read guards): undefined in Fortran.  The oracle takes the assigned value, 5000 m, throughout
# (oracle/README.md); the translated build is given the same value for the unassigned local so that the two can be compared.
UNINIT = {3: {"lnud": "5000."}}
UNDECLARED = ("svis", "tdrg", "topt")  # used by private_mod.f95, absent from the shipped shared_mod.f95


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF, "private_mod.f95"))


def parse_block(block: str) -> dict:
    """'name = value' lines of a print_params block -> {name: value text}."""
    vals = {}
    for line in block.splitlines():
        line = line.strip()
        if not line or line.startswith("!"):
            continue
        m = re.match(r"^([a-z_0-9]+)\s*(?:\(\s*nlay\s*\))?\s*=\s*(.*?)\s*$", line, re.I)
        if not m:
            raise ValueError("parameter block line %r" % line)
        vals[m.group(1).lower()] = m.group(2)
    return vals


def shared_mod_text(block: str, variant: int = 0) -> str:
    """The reference's shared_mod.f95 with the block's values in its user-modifiable section."""
    vals = parse_block(block)
    src = open(os.path.join(REF, "shared_mod.f95")).read().split("\n")
    begin = next(i for i, s in enumerate(src) if "BEGINNING OF USER-MODIFIABLE SECTION" in s)
    end = next(i for i, s in enumerate(src) if "END OF USER-MODIFIABLE SECTION" in s)
    seen = set()
    pat = re.compile(r"^(\s*)([a-z_0-9]+)(\s*\(\s*nlay\s*\))?(\s*=\s*)", re.I)
    last_real = None
    for i in range(begin, end):
        m = pat.match(src[i])
        if not m or src[i].lstrip().startswith("!"):
            continue
        name = m.group(2).lower()
        if name == "mcbc":
            last_real = i
        if name in vals:
            seen.add(name)
            code, q = [], None  # the statement part of the line (comment stripped, quotes respected)
            for ch in src[i]:
                if q:
                    q = None if ch == q else q
                elif ch in "'\"":
                    q = ch
                elif ch == "!":
                    break
                code.append(ch)
            cont = "".join(code).rstrip().endswith("&")
            src[i] = "%s%s%s%s%s%s" % (m.group(1), m.group(2), m.group(3) or "", m.group(4), vals[name],
                                       ", &" if cont else "")
    if last_real is None:
        raise ValueError("shared_mod.f95: user section not recognised")
    extra = list(UNDECLARED) + (["plum"] if variant == 3 else [])  # private_modplumenew.f95 also uses plum
    lines = []
    for n in extra:
        lines.append("    %-10s = %s, &" % (n, vals.get(n, "0.")))
        seen.add(n)
    src[last_real:last_real] = lines
    missing = sorted(set(vals) - seen - {"plum"})
    if missing:
        raise ValueError("parameter block names with no slot in shared_mod.f95: %s" % missing)
    return "\n".join(src)


def _translate(block: str, variant: int, timing: bool = False) -> str:
    sys.path.insert(0, os.path.join(HERE, "f95c"))
    try:
        import f95c
    finally:
        sys.path.pop(0)
    pm = os.path.join(REF, VARIANT_FILES[variant])
    srcs = [(shared_mod_text(block, variant), os.path.join(REF, "shared_mod.f95"), False),
            (open(pm).read(), pm, True),
            (open(os.path.join(REF, "main.f95")).read(), os.path.join(REF, "main.f95"), False)]
    if timing:  # bench.py --impl reference: directories and run length come from the environment, steps are time-stamped
        return f95c.translate(srcs, uninit=UNINIT.get(variant), env_params=("idir", "odir", "dt_s", "dt_o"),
                              trace=("gener_forward_backward",))
    return f95c.translate(srcs, uninit=UNINIT.get(variant))


def case_key(block: str, variant: int, omp: bool, timing: bool = False) -> str:
    # idir / odir are part of the compiled text, like in the reference
    h = hashlib.sha256()
    h.update(block.encode())
    h.update(b"|%d|%d|%d" % (variant, int(omp), int(timing)))
    for f in ("f95c/f95c.py", "f95c/f95rt.h"):
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def build_case(block: str, variant: int = 0, omp: bool = False, name: str | None = None, timing: bool = False) -> str:
    """-> path of the binary.  ``name`` gives the binary a stable directory (bench); else it is keyed by content."""
    # named builds travel to the GPU box; content-keyed ones (the pin tests') are a local cache (.gpurunignore)
    d = os.path.join(OUT, name) if name else os.path.join(OUT, "cache", case_key(block, variant, omp, timing))
    exe = os.path.join(d, "beom_ref")
    stamp = os.path.join(d, "stamp")
    want = case_key(block, variant, omp, timing)
    if os.path.exists(exe) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return exe
    if not reference_available():
        if os.path.exists(exe) and name:
            return exe  # prebuilt, travelled with the snapshot
        raise FileNotFoundError("the reference sources (%s) are not here and %s is not prebuilt" % (REF, exe))
    os.makedirs(d, exist_ok=True)
    cpp = os.path.join(d, "beom_ref.cpp")
    with open(cpp, "w") as f:
        f.write(_translate(block, variant, timing))
    flags = ["-std=c++17", "-I", os.path.join(HERE, "f95c"), "-ffp-contract=off", "-fno-fast-math", "-w"]
    flags += ["-O3", "-march=x86-64-v3", "-fopenmp"] if omp else ["-O2"]
    r = subprocess.run(["g++"] + flags + [cpp, "-o", exe], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("g++ failed on the translated reference:\n" + r.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(want)
    return exe


_DT = {1: np.int32, 2: np.bool_, 4: np.float32, 8: np.float64}


def read_dump(path: str) -> dict:
    """ref_dump.bin -> {name: array in the reference's shape (Fortran order; index 0 = the declared lower bound)}."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        name = data[pos:pos + 32].split(b"\0", 1)[0].decode()
        pos += 32
        ty, rank = np.frombuffer(data, np.int32, 2, pos)
        pos += 8
        lo = np.frombuffer(data, np.int64, rank, pos)
        pos += 8 * rank
        ext = np.frombuffer(data, np.int64, rank, pos)
        pos += 8 * rank
        n = int(np.prod(ext)) if rank else 1
        dt = np.dtype(_DT[int(ty)])
        a = np.frombuffer(data, dt, n, pos)
        pos += n * dt.itemsize
        out[name] = a.reshape(tuple(int(e) for e in ext), order="F") if rank else a[0]
        out[name + "__lb"] = tuple(int(x) for x in lo)
    return out


def run_case(exe: str, odir: str, threads: int | None = None, timeout: float = 3600.0):
    """Run the translated reference (its directories are compiled in); -> (dump dict, stdout)."""
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode:
        raise RuntimeError("the translated reference stopped with code %d:\n%s\n%s" % (r.returncode, r.stdout[-2000:],
                                                                                    r.stderr[-2000:]))
    return read_dump(os.path.join(odir, "ref_dump.bin")), r.stdout


if __name__ == "__main__":
    blk = open(sys.argv[1]).read()
    print(build_case(blk, int(sys.argv[2]) if len(sys.argv) > 2 else 0))


# ------------------------------------------------------------------------------------------------ timing builds
# bench.py --impl reference and __graft_entry__.smoke() use binaries built HERE (development container) under fixed names,
# because the GPU box has no /root/reference: the grid size, layer count and every switch are compiled in (as in the
# reference); the directories and the run length come from the environment (F95_IDIR, F95_ODIR, F95_DT_S, F95_DT_O).

def named_block(case) -> str:
    """The parameter block of a generated case with a placeholder directory (the run sets the real one)."""
    return case.params_text.replace("@DIR@", "/tmp/")


def time_case(exe: str, block: str, workdir: str, steps: int, warm: int, threads: int, dump: bool = False):
    """Run a timing build for 3 + warm + steps + 1 time steps in ``workdir`` (inputs already written there) and return
    the seconds per generalized forward-backward step, from the time stamps the harness takes at the entries of
    gener_forward_backward (private_mod.f95:2225) -- read_input_data, the three start-up steps and the warm-up steps
    are outside the timed span, the reference's own per-record output (one record after the last step) too."""
    import time

    from beom_b200 import model

    p, _, _, _ = model.parse_params(block)
    dtd8 = p.dt / 24.0 / 3600.0
    nstp = 3 + warm + steps + 1
    val = "%.9e" % (nstp * dtd8)
    d = workdir if workdir.endswith("/") else workdir + "/"
    env = dict(os.environ, F95_IDIR=d, F95_ODIR=d, F95_DT_S=val, F95_DT_O=val, OMP_NUM_THREADS=str(threads))
    if not dump:
        env["F95_NO_DUMP"] = "1"
    t0 = time.perf_counter()
    r = subprocess.run([exe], env=env, capture_output=True, text=True)
    total = time.perf_counter() - t0
    if r.returncode:
        raise RuntimeError("translated reference: exit code %d\n%s" % (r.returncode, (r.stdout + r.stderr)[-2000:]))
    marks = np.loadtxt(os.path.join(d, "ref_trace.txt"))
    ts = marks[:, 1]
    if len(ts) < warm + steps + 1:
        raise RuntimeError("translated reference: %d step marks, expected %d" % (len(ts), warm + steps + 2))
    per_step = float(ts[warm + steps] - ts[warm]) / steps  # mark k = entry of time step 4 + k
    return {"seconds_per_step": per_step, "seconds_total": total, "seconds_before_first_gfb_step": float(ts[0] - (ts[-1] - total)),
            "steps": steps, "warmup_steps": 3 + warm, "threads": threads}


def build_named(name: str, case, omp: bool):
    return build_case(named_block(case), omp=omp, name=name, timing=True)


# cases of tests/test_gpu_dropin.py: name -> (generator, kwargs, steps)
DROPIN_CASES = {
    "lock_exchange": ("lock_exchange", {}, 200),                                  # native size, two layers, Leith
    "sill_exchange3D": ("sill_exchange3D", dict(lx=6.0e3, ly=100.0e3), 80),       # sponges, outcropping, open-boundary segments
    "stommel1948": ("stommel1948", dict(dl=250.0e3), 120),                        # wind, linear drag, beta plane, no viscosity
    "basin": ("synthetic_basin", dict(n=300, mm=170, nlay=4), 40),                # the bench workload, small
}


def build_prebuilt():
    """What travels to the GPU box: the bench sample (2048 x 2048 x 4 basin, OpenMP), the smoke case (40 x 25 x 2), and for
    every DROPIN_CASES entry the pure translated reference and its drop-in flavour linked against libbeom_gpu.so."""
    from concurrent.futures import ThreadPoolExecutor

    from beom_b200 import cases

    root = os.path.dirname(HERE)
    jobs = {"bench_2048x4": lambda: build_named("bench_2048x4", cases.synthetic_basin(n=2048, nlay=4), omp=True),
            "smoke_40x25x2": lambda: build_named("smoke_40x25x2", cases.synthetic_basin(n=40, mm=25, nlay=2), omp=False)}
    for name, (gen, kw, _) in DROPIN_CASES.items():
        blk = named_block(cases.CASES[gen](**kw))
        jobs["pure_" + name] = (lambda b=blk, n=name: build_case(b, name="pure_" + n, timing=True))
        jobs["dropin_" + name] = (lambda b=blk, n=name: build_dropin(
            b, os.path.join(root, "beom_b200", "lib"), "beom_gpu", name="dropin_" + n, rpath="$ORIGIN/../../../beom_b200/lib"))
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        futs = {k: ex.submit(f) for k, f in jobs.items()}
        return {k: f.result() for k, f in futs.items()}


def prebuilt(name: str):
    exe = os.path.join(OUT, name, "beom_ref")
    return exe if os.path.exists(exe) else None


# ------------------------------------------------------------------------------------------------ the drop-in flavour
# The reference's own program with the edits of INTEGRATION.md section 2 applied to its text (in memory, at build time):
# its read_input_data, time loop and output routines stay, the step routines are replaced one for one by calls into the
# C ABI of include/beom_gpu.h (oracle/f95c/dropin_hooks.h holds the C++ twin of the Fortran binding lines).

DROPIN_EXTERNS = ("gpu_setup", "gpu_stress", "gpu_step", "gpu_download", "gpu_finalize")


def _sub_once(pattern, repl, text, count=1, where=""):
    new, n = re.subn(pattern, repl, text, flags=re.I | re.M)
    if n != count:
        raise ValueError("drop-in edit %r: %d matches, expected %d" % (where or pattern, n, count))
    return new


def dropin_sources(block: str):
    """-> the (text, file, dump) list of the translator with INTEGRATION.md's edits applied to private_mod.f95 / main.f95."""
    pm = open(os.path.join(REF, "private_mod.f95")).read()
    mainf = open(os.path.join(REF, "main.f95")).read()
    # end of read_input_data (private_mod.f95:238)
    pm = _sub_once(r"^(\s*)if \( rsta > 0\.5_rw \) call read_restart_record\s*$",
                   r"\g<0>\n\1call gpu_setup()", pm, where="gpu_setup")
    a = pm.lower().index("subroutine integrate_time")
    b = pm.lower().index("end subroutine integrate_time")
    body = pm[a:b]
    body = _sub_once(r"call distribute_stress\(\)", "call gpu_stress()", body, 2, "distribute_stress")
    body = _sub_once(r"call first_three_timesteps\( tstp \)", "call gpu_step( tstp, ctim, ramp, gene, .true., 1 )", body, 3,
                     "first_three_timesteps")
    body = _sub_once(r"call gener_forward_backward\( tstp, upst \)", "call gpu_step( tstp, ctim, ramp, gene, upst, 0 )", body, 1,
                     "gener_forward_backward")
    body = _sub_once(r"^(\s*)call write_outputs\(\)\s*$", r"\1call gpu_download()\n\g<0>", body, 1, "write_outputs")
    pm = pm[:a] + body + pm[b:]
    mainf = _sub_once(r"^(\s*)call quit\(\)", r"\1call gpu_finalize()\n\g<0>", mainf, 1, "gpu_finalize")
    return [(shared_mod_text(block, 0), os.path.join(REF, "shared_mod.f95"), False),
            (pm, os.path.join(REF, "private_mod.f95") + " (+ INTEGRATION.md edits)", True),
            (mainf, os.path.join(REF, "main.f95") + " (+ INTEGRATION.md edits)", False)]


def build_dropin(block: str, libdir: str, libname: str = "beom_gpu", name: str | None = None, rpath: str | None = None) -> str:
    """The drop-in flavour for one parameter block, linked against ``lib<libname>.so`` in ``libdir`` (libbeom_gpu.so for a
    B200, the emulated library of tools/emu for the CPU suite).  Directories and run length come from the environment
    (F95_IDIR, F95_ODIR, F95_DT_S, F95_DT_O), like the timing builds."""
    sys.path.insert(0, os.path.join(HERE, "f95c"))
    try:
        import f95c
    finally:
        sys.path.pop(0)
    h = hashlib.sha256()
    h.update(("dropin|" + block + "|" + libdir + "|" + libname).encode())
    for f in ("f95c/f95c.py", "f95c/f95rt.h", "f95c/dropin_hooks.h", "../include/beom_gpu.h"):
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    want = h.hexdigest()[:16]
    d = os.path.join(OUT, name) if name else os.path.join(OUT, "cache", "dropin_" + want)
    exe, stamp = os.path.join(d, "beom_ref"), os.path.join(d, "stamp")
    if os.path.exists(exe) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return exe
    if not reference_available():
        if os.path.exists(exe) and name:
            return exe
        raise FileNotFoundError("the reference sources (%s) are not here and %s is not prebuilt" % (REF, exe))
    os.makedirs(d, exist_ok=True)
    cpp = os.path.join(d, "beom_ref.cpp")
    with open(cpp, "w") as f:
        f.write(f95c.translate(dropin_sources(block), env_params=("idir", "odir", "dt_s", "dt_o"), externs=DROPIN_EXTERNS,
                               hooks_include="dropin_hooks.h"))
    inc = os.path.join(os.path.dirname(HERE), "include")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", os.path.join(HERE, "f95c"), "-I", inc,
           cpp, "-o", exe, "-L", libdir, "-l" + libname, "-Wl,-rpath," + (rpath or libdir)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("g++ failed on the drop-in build:\n" + r.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(want)
    return exe


def run_with_env(exe: str, workdir: str, block: str, nsteps: int, extra_env=None, timeout: float = 1800.0):
    """Run a build whose directories / run length come from the environment for ``nsteps`` steps (one record after the
    last); -> (dump, stdout)."""
    from beom_b200 import model

    p, _, _, _ = model.parse_params(block)
    val = "%.9e" % (nsteps * (p.dt / 24.0 / 3600.0))
    d = workdir if workdir.endswith("/") else workdir + "/"
    env = dict(os.environ, F95_IDIR=d, F95_ODIR=d, F95_DT_S=val, F95_DT_O=val)
    env.update(extra_env or {})
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode:
        raise RuntimeError("exit code %d:\n%s\n%s" % (r.returncode, r.stdout[-3000:], r.stderr[-3000:]))
    return read_dump(os.path.join(d, "ref_dump.bin")), r.stdout
