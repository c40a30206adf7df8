"""The library's split path on the CPU: tools/emu compiles the SOURCE of beom_gpu.cu, split.cuh and diag.cuh with g++
against a stand-in CUDA runtime in which a kernel launch is a serial loop over its grid, and the result is compared
with the oracle bit for bit (tides: same libm on both sides here, so also exact).

What this shows without a GPU: the logic of the kernels and of the host plumbing around them -- indexing in the dense
layout, masks, option switches, upload / download, periodic images and displaced duplicates (incl. the nudged ones of
baines_ridge), open-boundary segments -- for every reference script and the option matrix.  What it cannot show:
anything about races, barriers, TMA, the fused step (stubbed out) or speed.  The cases that HAVE run on a B200
(tests/test_gpu_parity.py) were bit-identical to the same oracle, so agreement here means the emulation reproduces
the hardware's results on them; the cases written after the GPU budget was spent get their first end-to-end check here.

The emulated library is test infrastructure: built into a temporary directory under its own name, loaded explicitly by
tests/emu_worker.py, reporting itself as "cpu-emulation"; nothing under beom_b200/ knows about it."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_sanitizer_builds = {}


@pytest.fixture(scope="module")
def emu_so(tmp_path_factory):
    from concurrent.futures import ThreadPoolExecutor
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tools", "emu", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    so = mod.build(str(tmp_path_factory.mktemp("emu")))
    # the two sanitizer variants are compiled in the background while the emulation jobs run
    pool = ThreadPoolExecutor(max_workers=2)
    _sanitizer_builds["asan"] = pool.submit(mod.build, str(tmp_path_factory.mktemp("emu_asan")), True, False)
    _sanitizer_builds["tsan"] = pool.submit(mod.build, str(tmp_path_factory.mktemp("emu_tsan")), False, True)
    # the one-command tests of this module (EARLY, filled next to each of them) start now, three at a time, beside the
    # emulation jobs; the tests pick their results up (_sp_run) -- a few minutes of single-threaded runs that would
    # otherwise follow one another
    early_pool = ThreadPoolExecutor(max_workers=3)
    for argv, env in EARLY:
        cmd = [a.replace("@EMU@", so) for a in argv]
        env = {k: v.replace("@EMU@", so) for k, v in env.items()}
        _early_futs[_ekey(cmd, env)] = early_pool.submit(subprocess.run, cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT,
                                                         env=dict(os.environ, **env))
    yield so
    early_pool.shutdown(wait=False, cancel_futures=True)


EARLY = []        # (argv with "@EMU@" for the emulated library's path, environment additions)
_early_futs = {}


def _ekey(cmd, env):
    return json.dumps([list(cmd), sorted((env or {}).items())])


def early(argv, env=None):
    EARLY.append((list(argv), dict(env or {})))


def _sp_run(cmd, env=None, timeout=900):
    """subprocess.run of a command of this module: the result of its early start if it had one (same argv, same environment)."""
    fut = _early_futs.pop(_ekey(cmd, env), None)
    if fut is not None and not fut.cancel():  # (still queued: run it here, now, instead of waiting for its turn)
        return fut.result()
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ, **(env or {})))


def _ranks_cmd(emu_so, name, nsteps, nranks, extra=None, kwargs=None, fused=None):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emu_ranks_worker.py"), emu_so, name, str(nsteps), str(nranks)]
    if extra is not None or kwargs is not None or fused is not None:
        cmd.append(json.dumps(extra if extra is not None else {}))
    if kwargs is not None or fused is not None:
        cmd.append(json.dumps(kwargs))
    if fused is not None:
        cmd.append(str(fused))
    return cmd


def _cmd(emu_so, name, nsteps, extra, kwargs, variant, fused):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emu_worker.py"), emu_so, name, str(nsteps), str(fused), json.dumps(extra or {})]
    if kwargs is not None or variant:
        from tests.conftest import SMALL
        cmd += [json.dumps(kwargs if kwargs is not None else SMALL.get(name, {})), str(variant)]
    return cmd


# Every single-rank job of this module is known when it is imported (JOBS, filled next to each parametrisation): the first
# test that needs one runs them all, four worker processes at a time, and the tests then look their verdicts up -- the
# wall-clock of ~100 short processes is start-up, not emulation.
JOBS = []
_done = {}


def _key(name, nsteps, extra, kwargs, variant, fused):
    return json.dumps([name, nsteps, extra or {}, kwargs, variant, fused], sort_keys=True)


def job(name, nsteps, extra=None, kwargs=None, variant=0, fused=0):
    JOBS.append((name, nsteps, extra, kwargs, variant, fused))


def _run_one(emu_so, j):
    r = subprocess.run(_cmd(emu_so, *j), capture_output=True, text=True, timeout=900, cwd=ROOT)
    return _key(*j), r


def run(emu_so, name, nsteps, extra=None, kwargs=None, variant=0, fused=0, path="split"):
    k = _key(name, nsteps, extra, kwargs, variant, fused)
    if k not in _done:
        from concurrent.futures import ThreadPoolExecutor
        todo = [j for j in JOBS if _key(*j) not in _done] or [(name, nsteps, extra, kwargs, variant, fused)]
        if k not in {_key(*j) for j in todo}:
            todo.append((name, nsteps, extra, kwargs, variant, fused))
        with ThreadPoolExecutor(max_workers=max(4, min(16, os.cpu_count() or 4))) as pool:
            for kk, r in pool.map(lambda j: _run_one(emu_so, j), todo):
                _done[kk] = r
    r = _done[k]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert r.returncode == 0 and not res["bad"], res
    assert res["path"] == path and res["worst"] == 0.0, res
    return res


SCRIPTS = ["stommel1948", "lock_exchange", "unstable_jet", "sill_exchange3D", "conservation", "soliton", "baines_ridge",
           "carrier_beach", "upwelling_seaward_wind", "mixed_open_bc", "morel_upwelling", "outcrop_seamount", "sill_exchange2D",
           "sill_exchange2Dtides", "tide_ridge", "wave_sponge"]


for _n in SCRIPTS:
    job(_n, 16)


@pytest.mark.parametrize("name", SCRIPTS)
def test_every_reference_script_on_the_emulated_split_path(emu_so, name):
    run(emu_so, name, 16)


SPLIT_OPTIONS = [("baines_ridge", {"mcbc": "0."}), ("wave_sponge", {"mcbc": "0."}),
                 ("sill_exchange3D", {"mcbc": "0."}), ("sill_exchange3D", {"bdrg": "2.e-3", "qdrg": "1."}),
                 ("sill_exchange3D", {"bdrg": "1.e-3", "qdrg": "0."}), ("sill_exchange3D", {"tdrg": "1.e-3"}),
                 ("sill_exchange3D", {"dt3d": "0.002"}), ("lock_exchange", {"svis": "1.e6"}),
                 ("stommel1948", {"rgld": "0.", "g_fb": "0."})]
for _n, _e in SPLIT_OPTIONS:
    job(_n, 24, _e)


@pytest.mark.parametrize("name,extra", SPLIT_OPTIONS)
def test_options_on_the_emulated_split_path(emu_so, name, extra):
    run(emu_so, name, 24, extra)


OPTION_MATRIX = {
    "ekman_sponge": dict(), "ramp": dict(dt_r=0.2), "bodf": dict(bodf=True), "hdot": dict(hdot=True), "beta": dict(beta=True),
    "six_layers": dict(nlay=6), "outcrop_wind": dict(ocrp=1.0), "no_sponge": dict(sponge=False), "no_wind": dict(wind=False),
    "tide": dict(tide=True),
}


for _o in OPTION_MATRIX:
    job("option_basin", 20, kwargs=OPTION_MATRIX[_o])
    if _o != "tide":
        job("option_basin", 16, kwargs=OPTION_MATRIX[_o], fused=1)


@pytest.mark.parametrize("opt", sorted(OPTION_MATRIX))
def test_option_matrix_on_the_emulated_split_path(emu_so, opt):
    run(emu_so, "option_basin", 20, kwargs=OPTION_MATRIX[opt])


VARIANTS = [(1, 2, None), (2, 3, None), (3, 3, "0."), (3, 3, "1.")]
for _v, _l, _p in VARIANTS:
    job("sponge_basin", 20, {"plum": _p} if _p else {}, kwargs=dict(nlay=_l), variant=_v)


@pytest.mark.parametrize("variant,nlay,plum", VARIANTS)
def test_update_h_variants_on_the_emulated_split_path(emu_so, variant, nlay, plum):
    run(emu_so, "sponge_basin", 20, {"plum": plum} if plum else {}, kwargs=dict(nlay=nlay), variant=variant)


def test_the_emulation_is_not_part_of_the_product():
    for base, _, files in os.walk(os.path.join(ROOT, "beom_b200")):
        if os.path.basename(base) in ("lib", "__pycache__"):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                with open(os.path.join(base, fn), errors="replace") as f:
                    text = f.read()
                assert "tools/emu" not in text and "libbeom_gpu_emu" not in text and "BEOM_CUDA_EMULATION" not in text, fn


RANKS = [
    ("synthetic_basin", 9, 2, {}, dict(n=60, mm=40, nlay=2)),   # plain y-slabs (the configuration that HAS run on 2 B200s)
    ("sill_exchange3D", 12, 2, {}, None),
    ("wave_sponge", 12, 4, {"mcbc": "0."}, None),               # open-boundary segments spread over four ranks
    ("sill_exchange3D", 12, 3, {"mcbc": "0."}, None),           # ... and a middle rank that owns none (the exchanges stay collective)
    ("soliton", 12, 3, {}, None),                                # x-periodic: images stay inside a rank
    ("conservation", 12, 2, {}, None),                           # doubly periodic: the halo exchange is ring-closed
    ("conservation", 12, 3, {}, None),
    ("conservation", 12, 2, {"xper": "0."}, None),               # periodic in y only
    ("unstable_jet", 12, 4, {}, None),
    # the rigid lid across y-slabs: surf_pressure in rounds, batches of 15 sweeps (steps 3 and 4 take 18 and 55)
    ("synthetic_basin", 4, 3, {"rgld": "1.", "ocrp": "1."}, dict(n=300, mm=170, nlay=2)),
]


for _r in RANKS:
    early(_ranks_cmd("@EMU@", _r[0], _r[1], _r[2], _r[3], _r[4]))


@pytest.mark.parametrize("name,nsteps,nranks,extra,kwargs", RANKS, ids=["%s-%d%s" % (r[0], r[2], "-" + "".join(r[3]) if r[3] else "") for r in RANKS])
def test_y_slab_ranks_on_the_emulated_split_path(emu_so, name, nsteps, nranks, extra, kwargs):
    """One emulated library instance per rank (own file, own globals, own thread); the packed halo exchange goes through
    a callback that pairs the messages like NCCL.  Every rank's own slab -- periodic duplicates included -- must equal the
    oracle bit for bit, and the own ranges must tile 1..ndeg (output records are their concatenation)."""
    r = _sp_run(_ranks_cmd(emu_so, name, nsteps, nranks, extra, kwargs))
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert r.returncode == 0, res
    ranks = res["ranks"]
    assert all(not k["bad"] and k["path"] == "split" and k["exchanges"] > 3 * nsteps for k in ranks), res
    assert ranks[0]["points"][0] == 1 and all(a["points"][1] + 1 == b["points"][0] for a, b in zip(ranks, ranks[1:]))
    assert res["moved"] > 0


early(_ranks_cmd("@EMU@", "baines_ridge", 5, 2))


def test_ranks_that_cannot_hold_the_seam_fail_loudly(emu_so):
    r = _sp_run(_ranks_cmd(emu_so, "baines_ridge", 5, 2))
    assert r.returncode == 1 and "rows per rank" in r.stdout


@pytest.mark.parametrize("name,dt_o", [("lock_exchange", 0.005), ("conservation", 0.1), ("baines_ridge", 0.01), ("sill_exchange3D", 0.0005)])
def test_the_executable_end_to_end_on_the_emulation(emu_so, tmp_path, name, dt_o):
    """beom_run (= main.f95: read_input_data, integrate_time, write_outputs) with the emulated library preloaded in place
    of libbeom_gpu.so: the output files it leaves behind -- eta_/u___/v___.bin and, with diag = 1, pvor/mont/v_cc.bin,
    time.txt -- against the oracle's write_array records at the same steps, bit for bit (float32)."""
    import numpy as np
    from beom_b200 import cases, model
    from oracle.pyoracle import Oracle
    from tests.conftest import SMALL
    c = cases.CASES[name](**SMALL.get(name, {}))
    text = c.params_text
    import re
    text = re.sub(r"dt_o       = \S+", "dt_o       = %f" % dt_o, text)
    text = re.sub(r"diag       = \S+", "diag       = 1.", text)
    c.params_text = text
    blk = c.write(str(tmp_path))
    hm = model.HostModel.from_block(blk)
    nstp, notp, _ = hm.counts()
    assert 5 <= notp <= 200, notp
    nrec = 3
    steps = nrec * notp + 5
    exe = os.path.join(ROOT, "beom_b200", "lib", "beom_run")
    env = dict(os.environ, LD_PRELOAD=emu_so)
    r = subprocess.run([exe, blk, "--steps", str(steps), "--split"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "record = %d" % (nrec + 1) in r.stdout
    orc = Oracle(hm.params, str(tmp_path))
    n, nlay = c.ndeg, c.nlay
    sub = hm.iarray("subc")
    dup = np.zeros(n, dtype=bool)  # diag records are not reproduced at the periodic duplicates (conservation.m discards them)
    if hm.params.xper > 0.5:
        dup |= (sub[0] == c.lm + 1)[1:]
    if hm.params.yper > 0.5:
        dup |= (sub[1] == c.mm + 1)[1:]
    times = np.atleast_1d(np.loadtxt(tmp_path / "time.txt"))
    assert times.size == nrec + 1
    for k in range(nrec + 1):
        if k:
            orc.advance((k - 1) * notp + 1, k * notp)
        assert abs(times[k] - k * notp * hm.params.dt / 86400.0) < 1e-12
        for var in ("eta_", "u___", "v___", "pvor", "mont", "v_cc"):
            raw = np.fromfile(tmp_path / (var + ".bin"), dtype="<f4", count=n * nlay, offset=4 * k * n * nlay).reshape(nlay, n)
            want = orc.record(var)
            keep = ~dup if var in ("pvor", "mont", "v_cc") else np.ones(n, dtype=bool)
            same = (raw == want) | (np.isnan(raw) & np.isnan(want)) | ~keep[None, :]
            assert same.all(), (name, var, k, int((~same).sum()))


@pytest.mark.parametrize("name", ["sill_exchange3D", "stommel1948", "sill_exchange2Dtides"])
def test_the_executable_initialised_on_the_device_writes_the_same_files(emu_so, tmp_path, name):
    """beom_run initialises on the device where beom_gpu_init_grids covers the case (csrc/host/init.cc: beom_host_create_on_device):
    every file it leaves behind -- grid.bin, h_0.bin, param_basin.txt, the records, time.txt -- must be byte for byte what the run
    initialised by the host's read_input_data (--host-init) writes."""
    import re
    from beom_b200 import cases
    from tests.conftest import SMALL
    outs = {}
    for how in ("device", "host"):
        d = tmp_path / how
        d.mkdir()
        c = cases.CASES[name](**SMALL.get(name, {}))
        c.params_text = re.sub(r"diag       = \S+", "diag       = 1.", c.params_text)
        blk = c.write(str(d))
        exe = os.path.join(ROOT, "beom_b200", "lib", "beom_run")
        r = subprocess.run([exe, blk, "--steps", "40", "--fused"] + (["--host-init"] if how == "host" else []), capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, LD_PRELOAD=emu_so), cwd=str(d))
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert ("read_input_data: on the device" in r.stdout) == (how == "device"), r.stdout[:600]
        outs[how] = {f: open(os.path.join(d, f), "rb").read() for f in sorted(os.listdir(d))
                     if f in ("grid.bin", "h_0.bin", "eta_.bin", "u___.bin", "v___.bin", "pvor.bin", "mont.bin", "v_cc.bin", "time.txt")}
        assert "grid.bin" in outs[how] and "eta_.bin" in outs[how] and len(outs[how]["eta_.bin"]) > 0
    assert outs["device"].keys() == outs["host"].keys()
    for f in outs["host"]:
        assert outs["device"][f] == outs["host"][f], f


def test_restart_continues_the_record_files_on_the_emulation(emu_so, tmp_path):
    """rsta = 1 (private_mod.f95:236-243, 1299-1420): a second beom_run picks the state up from the last float32 record,
    keeps counting time from there and appends its records.  The oracle, started from the same float32-rounded state,
    must produce the appended records bit for bit."""
    import numpy as np
    import re
    from beom_b200 import cases, model
    from oracle.pyoracle import Oracle
    c = cases.lock_exchange()
    c.params_text = re.sub(r"dt_o       = \S+", "dt_o       = 0.005000", c.params_text)
    blk = c.write(str(tmp_path))
    hm = model.HostModel.from_block(blk)
    _, notp, _ = hm.counts()
    exe = os.path.join(ROOT, "beom_b200", "lib", "beom_run")
    env = dict(os.environ, LD_PRELOAD=emu_so)
    r = subprocess.run([exe, blk, "--steps", str(2 * notp + 3), "--split"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    n, nlay = c.ndeg, c.nlay
    assert os.path.getsize(tmp_path / "eta_.bin") == 3 * n * nlay * 4
    # second run: restart from record 3
    with open(blk) as f:
        text = f.read()
    with open(blk, "w") as f:
        f.write(re.sub(r"rsta       = \S+", "rsta       = 1.", text))
    # what read_restart_record makes of record 3 (float32 eta -> hlay = h_0 + eta_k - eta_k+1), for the oracle below
    hm2 = model.HostModel.from_block(blk, write_outputs=True)   # rsta = 1: the directory is left as it is
    assert hm2.lib.beom_host_read_restart(hm2.h) == 0 and hm2.scalar("irec") == 4
    start = {k: hm2.array(k).copy() for k in ("hlay", "u", "v")}
    r = subprocess.run([exe, blk, "--steps", str(2 * notp), "--split"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "Restarting from record number 3" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    assert os.path.getsize(tmp_path / "eta_.bin") == 5 * n * nlay * 4
    times = np.loadtxt(tmp_path / "time.txt")
    dtd8 = hm.params.dt / 86400.0
    assert times.size == 5 and abs(times[4] - (times[2] + 2 * notp * dtd8)) < 1e-12
    # the oracle from the restart state
    orc = Oracle(hm2.params, str(tmp_path))
    for k in ("hlay", "u", "v"):
        orc.array(k)[:] = start[k]
    for rec in (3, 4):
        orc.advance((rec - 3) * notp + 1, (rec - 2) * notp)
        for var in ("eta_", "u___", "v___"):
            raw = np.fromfile(tmp_path / (var + ".bin"), dtype="<f4", count=n * nlay, offset=4 * rec * n * nlay).reshape(nlay, n)
            assert np.array_equal(raw, orc.record(var)), (var, rec)


# ------------------------------------------------------------------------------------------------------------------
# The FUSED step on the SIMT emulator (tools/emu/include/simt.h): one OS thread per warp, 32 coroutine lanes, warp
# shuffles and votes, __syncthreads, mbarriers with transaction counts and bulk global->shared copies -- the kernel
# source of fused_kernel.cuh unchanged except for its six inline-PTX helpers, which build_emu.py redirects.  Shared
# memory starts filled with a byte pattern, so a read of something never written shows up as garbage.  A protocol error
# (an mbarrier wait nobody satisfies) is a hang: the workers run under a timeout.
# ------------------------------------------------------------------------------------------------------------------
FUSED_SCRIPTS = list(SCRIPTS)  # all sixteen (tidal targets: the step is handed fnud + tide term, k_tide_targets)


for _n in FUSED_SCRIPTS:
    job(_n, 16, fused=1)
for _l in (1, 3, 4):
    job("synthetic_basin", 9, kwargs=dict(n=130, mm=70, nlay=_l), fused=1)


@pytest.mark.parametrize("name", FUSED_SCRIPTS)
def test_every_reference_script_on_the_emulated_fused_step(emu_so, name):
    """Incl. the shapes that had never run on hardware when this was written: one-row and one-column tori
    (baines_ridge, upwelling_seaward_wind, morel_upwelling), five layers, a single outcropping layer, the open-boundary
    kernel after the fused step on three sides, nudged periodic duplicates."""
    run(emu_so, name, 16, fused=1, path="fused")


@pytest.mark.parametrize("nlay", [1, 3, 4])
def test_the_bench_workload_on_the_emulated_fused_step(emu_so, nlay):
    """The specialised (lean) instantiations: several x strips and y chunks, wind stress, Leith viscosity, both u/v orders,
    the gene = 0 start-up steps."""
    run(emu_so, "synthetic_basin", 9, kwargs=dict(n=130, mm=70, nlay=nlay), fused=1, path="fused")


job("option_basin", 16, kwargs=dict(tide=True), fused=1)
job("option_basin", 16, kwargs=dict(tide=True, wind=False), fused=1)
job("tide_ridge", 16, {"mcbc": "0."}, fused=1)


@pytest.mark.parametrize("opt", sorted(k for k in OPTION_MATRIX if k != "tide"))
def test_option_matrix_on_the_emulated_fused_step(emu_so, opt):
    run(emu_so, "option_basin", 16, kwargs=OPTION_MATRIX[opt], fused=1, path="fused")


def test_tidal_targets_on_the_emulated_fused_step(emu_so):
    """Tides run fused unless the Ekman term of the sponge target (wind stress and f != 0), which the reference adds
    between fnud and the tidal term, is live: then the sum would associate differently and the split path is kept."""
    run(emu_so, "option_basin", 16, kwargs=dict(tide=True, wind=False), fused=1, path="fused")
    run(emu_so, "option_basin", 16, kwargs=dict(tide=True), fused=1, path="split")
    run(emu_so, "tide_ridge", 16, {"mcbc": "0."}, fused=1, path="fused")  # + no_gradient_obc with the plain targets


SILL_FUSED = [{"mcbc": "0."}, {"bdrg": "2.e-3", "qdrg": "1."}, {"bdrg": "1.e-3", "qdrg": "0."}, {"tdrg": "1.e-3"},
              {"dt3d": "0.002", "dvis": "0.", "bvis": "30."}]  # n_3d > 1 with a constant viscosity stays on the fused step
for _e in SILL_FUSED:
    job("sill_exchange3D", 16, _e, fused=1)


@pytest.mark.parametrize("extra", SILL_FUSED)
def test_sill_options_on_the_emulated_fused_step(emu_so, extra):
    run(emu_so, "sill_exchange3D", 16, extra, fused=1, path="fused")


OBC_FUSED = [("baines_ridge", 24), ("wave_sponge", 16), ("mixed_open_bc", 16)]
for _n, _k in OBC_FUSED:
    job(_n, _k, {"mcbc": "0."}, fused=1)


@pytest.mark.parametrize("name,nsteps", OBC_FUSED)
def test_open_boundaries_after_the_emulated_fused_step(emu_so, name, nsteps):
    """no_gradient_obc reads hlay, u, v at neighbours of the segment points (pm:2635-2676).  After the fused step, which
    writes only the rows it owns, those may be stale periodic images: in baines_ridge (one row, periodic in y) EVERY
    southern neighbour is one -- the combination that failed on the B200 in round 1."""
    run(emu_so, name, nsteps, {"mcbc": "0."}, fused=1, path="fused")


@pytest.mark.parametrize("name,fused", [("lock_exchange", 1), ("sill_exchange3D", 1), ("sill_exchange3D", 0), ("stommel1948", 1)])
def test_reference_program_drives_the_library_through_the_c_abi(emu_so, name, fused):
    """The drop-in boundary, executed (tests/test_gpu_dropin.py is the same comparison on a B200): the reference's own program
    -- its read_input_data, time loop, write_outputs, translated from /root/reference by oracle/f95c -- with the edits of
    INTEGRATION.md section 2 applied to its text, linked here against the EMULATED library, writes byte-identical output
    files to the pure reference and leaves a bit-identical state."""
    import shutil
    import tempfile
    from oracle import refbuild
    from tests.test_gpu_dropin import check_pair, run_pair
    from beom_b200 import cases
    if not refbuild.reference_available():
        pytest.skip("needs the reference's sources (/root/reference)")
    gen, kw, _ = refbuild.DROPIN_CASES[name]
    blk = refbuild.named_block(cases.CASES[gen](**kw))
    libname = os.path.basename(emu_so)[3:-3]
    exe_pure = refbuild.build_case(blk, timing=True)
    exe_dropin = refbuild.build_dropin(blk, os.path.dirname(emu_so), libname)
    root = tempfile.mkdtemp(prefix="di", dir="/tmp")
    try:
        out = run_pair(name, exe_pure, exe_dropin, fused, root)
        check_pair(out)
    finally:
        shutil.rmtree(root, ignore_errors=True)


_GPU_MODULES_CMD = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), os.path.join(ROOT, "tests", "test_gpu_outputs.py"),
                    "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k", "rigid_lid or conservation_integrals or diagnostic_records or biharmonic"]
early(_GPU_MODULES_CMD, {"BEOM_TEST_EMU": "@EMU@"})


def test_gpu_test_modules_on_the_emulation(emu_so):
    """The GPU tests themselves, re-run with the emulated library swapped in (tests/conftest.py, BEOM_TEST_EMU): the ones
    whose kernels nothing above reaches -- the rigid lid (hyperplane Gauss-Seidel with __syncthreads and shuffle
    reductions), the biharmonic viscosity, the float32 diagnostic records and the conservation integrals (warp-shuffle
    reductions).  The whole of tests/test_gpu_parity.py and tests/test_gpu_outputs.py passes this way too (about five
    minutes; `BEOM_TEST_EMU=<libbeom_gpu_emu.so> pytest tests/test_gpu_parity.py tests/test_gpu_outputs.py -m gpu`)."""
    r = _sp_run(_GPU_MODULES_CMD, {"BEOM_TEST_EMU": emu_so}, timeout=1200)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


RANKS_FUSED = [("synthetic_basin", 9, 2, dict(n=60, mm=40, nlay=2), {}),
               ("synthetic_basin", 9, 4, dict(n=60, mm=90, nlay=4), {}),
               ("sill_exchange3D", 12, 2, None, {}),
               ("soliton", 12, 3, None, {}),        # periodic in x: every slab is a torus of its own
               ("conservation", 12, 3, None, {}),   # doubly periodic: deep rows across the ring
               ("unstable_jet", 12, 4, None, {}),
               ("wave_sponge", 12, 4, None, {"mcbc": "0."}),      # open-boundary segments whose neighbours are another rank's rows
               ("sill_exchange3D", 12, 3, None, {"mcbc": "0."})]


for _r in RANKS_FUSED:
    early(_ranks_cmd("@EMU@", _r[0], _r[1], _r[2], _r[4], _r[3], 1))


@pytest.mark.parametrize("name,nsteps,nranks,kwargs,extra", RANKS_FUSED,
                         ids=["%s-%d%s" % (r[0], r[2], "-obc" if r[4] else "") for r in RANKS_FUSED])
def test_y_slab_ranks_on_the_emulated_fused_step(emu_so, name, nsteps, nranks, kwargs, extra):
    """The fused step on y-slabs: one packed exchange of the 8 new fields' four boundary rows per step, the deep halo
    recomputed locally -- every rank's slab bit-identical to the oracle.  With no_gradient_obc
    (mcbc = 0) the exchange runs before the open-boundary copy and once more, for u, v, h_u, h_v, after it."""
    r = _sp_run(_ranks_cmd(emu_so, name, nsteps, nranks, extra, kwargs, 1))
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert r.returncode == 0 and all(not k["bad"] and k["path"] == "fused" for k in res["ranks"]), res
    assert all(k["exchanges"] < 3 * nsteps for k in res["ranks"])  # one exchange per fused step (two with mcbc = 0; + the start-up rebuilds)


COASTS = [dict(seed=1), dict(seed=5, lm=120, mm=21, nlay=2, sponge=True), dict(seed=6, lm=57, mm=57, nlay=5, ocrp=1.0),
          dict(seed=7, lm=29, mm=33, nlay=4, land=0.5), dict(seed=9, lm=85, mm=40, nlay=8, land=0.1)]


for _s in COASTS:
    for _f in (0, 1):
        job("random_coast", 12, kwargs=_s, fused=_f)


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("spec", COASTS, ids=["seed%d" % s["seed"] for s in COASTS])
def test_random_coastlines_on_the_emulation(emu_so, spec, fused):
    """cases.random_coast: bays, islands, one-cell channels and lakes drive every combination of the five masks through
    the masked rows of both paths (wind, quadratic drag, optional sponge and outcropping; 1 to 8 layers)."""
    run(emu_so, "random_coast", 12, kwargs=spec, fused=fused, path="fused" if fused else "split")


def test_no_out_of_bounds_access_under_addresssanitizer(emu_so, tmp_path_factory):
    """The emulated library built with -fsanitize=address: "device" memory is malloc'ed and shared memory is a vector, so
    a kernel reading or writing outside its planes, rings or staged row segments becomes an ASan report (the CPU
    counterpart of compute-sanitizer's memcheck).  The lean and the general fused step (several strips and chunks, a
    one-row torus) and the split path with tidal targets."""
    import shutil
    so = _sanitizer_builds["asan"].result(timeout=1200)
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not found")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    for args in (["synthetic_basin", "6", "1", "{}", json.dumps(dict(n=130, mm=70, nlay=3))], ["baines_ridge", "8", "1"],
                 ["sill_exchange3D", "6", "1", json.dumps({"mcbc": "0."})], ["tide_ridge", "6", "0"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_worker.py"), so] + args, capture_output=True, text=True,
                           timeout=900, cwd=ROOT, env=env)
        assert "AddressSanitizer" not in r.stderr and r.returncode == 0, (args, r.stdout[-500:], r.stderr[-3000:])
        assert json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])["bad"] == []
    shutil.rmtree(os.path.dirname(so), ignore_errors=True)


def test_no_race_between_warps_under_threadsanitizer(emu_so, tmp_path_factory):
    """The emulated library built with -fsanitize=thread, every lane a TSan fiber (tools/emu/simt.cc): two WARPS (OS
    threads) touching the same shared-memory word without an mbarrier / __syncthreads in between is a report -- the CPU
    counterpart of compute-sanitizer's racecheck for the fused step's protocol (input ring refill against its readers,
    thickness exchange between the layers of a column group, per-warp state rings).  The one race that is there by
    design (halo lanes peeking at the neighbouring column group's thickness slot; the value only feeds discarded halo
    results) goes through emu::halo_peek and is reported when it does not (seen while writing this).  Run through
    beom_run with the host single-threaded: libgomp's barriers are invisible to TSan."""
    import shutil
    from beom_b200 import cases
    so = _sanitizer_builds["tsan"].result(timeout=1200)
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        pytest.skip("libtsan.so not found")
    exe = os.path.join(ROOT, "beom_b200", "lib", "beom_run")
    env = dict(os.environ, LD_PRELOAD=tsan + " " + so, OMP_NUM_THREADS="1",
               TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4 exitcode=0")
    for k, c in enumerate((cases.synthetic_basin(n=130, mm=70, nlay=4),                 # lean, 2 strips x 4 chunks, wind
                           cases.random_coast(seed=7, lm=29, mm=33, nlay=4, land=0.5),   # general, masked rows, drag
                           cases.conservation(dl=30.0e3))):                              # torus
        d = str(tmp_path_factory.mktemp("tsan_case%d" % k))
        blk = c.write(d)
        r = subprocess.run([exe, blk, "--steps", "5", "--fused"], capture_output=True, text=True, timeout=1500, env=env, cwd=d)
        assert r.returncode == 0 and "record = 1" in r.stdout, r.stdout[-800:] + r.stderr[-2000:]
        assert "ThreadSanitizer" not in r.stderr, r.stderr[:6000]
    shutil.rmtree(os.path.dirname(so), ignore_errors=True)


_GRID_INIT_CMD = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_grid_init.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"]
_GRID_INIT_RANKS = [["synthetic_basin", "9", "3", "{}", json.dumps(dict(n=60, mm=90, nlay=4)), "1"], ["sill_exchange3D", "12", "2", "{}", "null", "1"],
                    ["conservation", "12", "3", "{}", "null", "1"],                     # ring-closed y-periodic slabs
                    ["wave_sponge", "12", "4", json.dumps({"mcbc": "0."}), "null", "1"]]  # open-boundary segments spread over the ranks
early(_GRID_INIT_CMD, {"BEOM_TEST_EMU": "@EMU@"})
for _spec in _GRID_INIT_RANKS:
    early([sys.executable, os.path.join(ROOT, "tests", "emu_ranks_worker.py"), "@EMU@"] + _spec, {"BEOM_GRID_INIT": "1"})


def test_device_side_initialisation_on_the_emulation(emu_so):
    """beom_gpu_init_grids (gridinit.cuh: index_grid_points as a prefix sum, the rest thickness incl. the Newton solve, the forcing
    files) against the host's read_input_data on every case it covers, and the refusals (tests/test_grid_init.py re-run with the
    emulated library); then every rank of a y-slab run initialised that way."""
    r = _sp_run(_GRID_INIT_CMD, {"BEOM_TEST_EMU": emu_so}, timeout=1200)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    for spec in _GRID_INIT_RANKS:
        r = _sp_run([sys.executable, os.path.join(ROOT, "tests", "emu_ranks_worker.py"), emu_so] + spec, {"BEOM_GRID_INIT": "1"})
        res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        assert r.returncode == 0 and all(not k["bad"] and k["path"] == "fused" for k in res["ranks"]), res


for _o in ("1", "0"):
    early(_ranks_cmd("@EMU@", "synthetic_basin", 9, 2, {}, dict(n=60, mm=90, nlay=2), 1), {"BEOM_OVERLAP": _o})


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_halo_overlap_variant_on_emulated_ranks(emu_so, overlap):
    """The default on several ranks: the G rows next to each neighbour first, their exchange while the rows
    in between are computed -- three launches per step instead of one; BEOM_OVERLAP=0 is the plain sequence.  Same result."""
    r = _sp_run(_ranks_cmd(emu_so, "synthetic_basin", 9, 2, {}, dict(n=60, mm=90, nlay=2), 1), {"BEOM_OVERLAP": overlap})
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert r.returncode == 0 and all(not k["bad"] and k["path"] == "fused" for k in res["ranks"]), res


# ---- whole steps as CUDA graphs on latency-bound grids (beom_gpu.cu: step_graphed).  The emulation records the launches of a
# capture with their arguments as they were when issued and replays them in order, which is what a graph launch does: what is
# checked here is the library's bookkeeping -- one graph per phase of the rotating buffers (period 12), the indices after a
# replayed step, what may and may not be captured -- not CUDA's.
GRAPH_CASES = [("lock_exchange", 45, {}, 0, 40), ("stommel1948", 45, {}, 0, 40), ("soliton", 40, {}, 0, 35),
               ("sill_exchange3D", 45, {"bdrg": "2.e-3", "qdrg": "1."}, 0, 40),   # distribute_stress in front of every step
               ("sill_exchange3D", 45, {"dt3d": "0.002"}, 0, 40),                  # n_3d = 2: upst alternates
               ("baines_ridge", 40, {"mcbc": "0."}, 0, 35),                        # periodic images + the open-boundary copy
               ("lock_exchange", 40, {}, 1, 35), ("baines_ridge", 36, {"mcbc": "0."}, 1, 31),  # the fused step
               ("upwelling_seaward_wind", 30, {}, 0, 0),                           # dt_r ramp: ramp is a kernel argument
               ("tide_ridge", 30, {}, 0, 0)]                                       # tides: so is ctim
for _n, _s, _e, _f, _g in GRAPH_CASES:
    job(_n, _s, _e, fused=_f)


@pytest.mark.parametrize("name,nsteps,extra,fused,graph_steps", GRAPH_CASES,
                         ids=["%s-%s-%d" % (n, "fused" if f else "split", i) for i, (n, _, _, f, _) in enumerate(GRAPH_CASES)])
def test_steps_replayed_from_graphs_are_the_same_steps(emu_so, name, nsteps, extra, fused, graph_steps):
    res = run(emu_so, name, nsteps, extra, fused=fused, path="fused" if fused else "split")
    assert res["graph_steps"] == graph_steps, res  # all steady steps but the first two (0: never steady / not eligible)


@pytest.mark.parametrize("env", [{"BEOM_EMU_NO_CAPTURE": "1"}, {"BEOM_GRAPH": "0"}], ids=["capture-refused", "switched-off"])
def test_step_graphs_fall_back_to_direct_launches(emu_so, env):
    """A capture that fails switches the graphs off for good; the step that was being captured runs directly, nothing runs twice."""
    cmd = _cmd(emu_so, "sill_exchange3D", 30, {"bdrg": "2.e-3", "qdrg": "1."}, None, 0, 0)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **env))
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert r.returncode == 0 and not res["bad"] and res["worst"] == 0.0 and res["graph_steps"] == 0, res


for _env in ({}, {"BEOM_GRAPH": "0"}):
    early([sys.executable, os.path.join(ROOT, "tests", "emu_reupload_worker.py"), "@EMU@", "sill_exchange3D", "24"], _env)


def test_a_second_upload_drops_the_step_graphs(emu_so):
    """beom_gpu_upload_state in the middle of a run (a restart from the downloaded state): the graphs of the first leg are dropped,
    the second leg captures its own -- same final state as with the graphs switched off."""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emu_reupload_worker.py"), emu_so, "sill_exchange3D", "24"]
    res = []
    for env in ({}, {"BEOM_GRAPH": "0"}):
        r = _sp_run(cmd, env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res.append(json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]))
    assert res[0]["graph_steps"] == [19, 19] and res[1]["graph_steps"] == [0, 0], res
    assert res[0]["sha256"] == res[1]["sha256"], res
