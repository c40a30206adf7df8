#!/usr/bin/env python
"""bench.py -- cell-layer updates/s of BEOM's per-timestep update on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

Workload at any N: the synthetic 8192 x 8192 x 4-layer closed basin of BASELINE.json / SURVEY.md 8(d)
(strong scaling: the grid is fixed, y-slabs are split over the ranks).  A "step" is one model time
step (update_h + Montgomery/vorticity/divergence + Leith viscosity + u,v momentum) in the generalized
forward-backward regime (tstp >= 4); the three start-up steps are part of the warm-up.

One JSON line on stdout (rank 0).  Keys beyond the base contract: roofline, cpu_baseline.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_UPDATE = 168.0  # 13 double reads + 8 double writes per cell-layer update


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax = max(smax, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_case(n, nlay, workdir, mm=None):
    from beom_b200 import cases
    c = cases.synthetic_basin(n=n, nlay=nlay, mm=mm)
    blk = c.write(workdir)
    return c, blk


def cpu_baseline(sample_n, nlay, steps, tmp):
    """Times the oracle's OpenMP build (the reference's loops with the reference's `!$OMP PARALLEL DO'
    placement) on a bounded sample of the same workload."""
    from beom_b200 import model
    from oracle.pyoracle import Oracle
    d = os.path.join(tmp, "cpu_sample")
    c, blk = make_case(sample_n, nlay, d)
    with open(blk) as f:
        p, idir, _, _ = model.parse_params(f.read())
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    orc = Oracle(p, idir, omp=True)
    orc.advance(1, 4)  # start-up steps + one generalized forward-backward step (warm caches)
    t0 = time.perf_counter()
    orc.advance(5, 4 + steps)
    dt = time.perf_counter() - t0
    orc.close()
    shutil.rmtree(d, ignore_errors=True)
    ups = float(sample_n) * sample_n * nlay * steps / dt
    return {"value": ups, "unit": "cell-layer updates/s", "cores": int(os.environ["OMP_NUM_THREADS"]), "kind": "port",
            "sample": "%dx%dx%d basin of the same workload, %d generalized-FB steps, oracle built -O3 -fopenmp "
                      "(the reference cannot be compiled here: no Fortran compiler)" % (sample_n, sample_n, nlay, steps),
            "seconds": dt}


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the real stdout (see _quiet_stdout)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout() -> None:
    """Everything else that writes to fd 1 (NCCL's version banner, compiler output of the in-tree build) goes to
    stderr, so that stdout carries exactly one line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=8192, help="basin is size x size cells")
    ap.add_argument("--nlay", type=int, default=4)
    ap.add_argument("--split", action="store_true", help="one kernel per reference loop instead of the fused step")
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--cpu-steps", type=int, default=120, help="steps of the CPU baseline leg on the sample (about 10-15 s of CPU work on 16 threads)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fma", action="store_true", help="opt-in: the FMA-contracted copy of the fused step (tolerance parity, not bit-exact)")
    ap.add_argument("--trace", type=int, default=0, help="diagnostic: time N further blocks of 10 steps each and log clocks/power")
    args = ap.parse_args()

    if args.fma:
        os.environ["BEOM_FMA"] = "1"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n, nlay = args.size, args.nlay
    config = {"workload": "synthetic %dx%dx%d-layer closed flat basin (SURVEY.md 8d): dl=1km, Leith dvis=0.2 every step, "
                          "generalized forward-backward, wind 0.1cos(pi y/L) Pa, seed 20261018" % (n, n, nlay),
              "grid": [n, n, nlay], "decomposition": "y-slabs x%d" % args.gpus, "l2": "inputs_exceed_L2",
              "bytes_per_update_algorithmic": ALGO_BYTES_PER_UPDATE,
              "arithmetic": "fma-contracted (tolerance parity 1e-10)" if args.fma else "strict IEEE, no contraction (bit-exact vs the oracle)"}

    from beom_b200 import build
    from oracle import build as obuild
    if rank == 0:
        build.build_all()
        obuild.build_oracle()

    # ---------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        tmp = tempfile.mkdtemp(prefix="beom_ref_")
        try:
            W = max(args.warmup, 0)
            cb = cpu_baseline(args.cpu_sample, nlay, max(args.steps, 1), tmp)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        ms = 1.0e3 * (float(n) * n * nlay) / cb["value"]  # per step of the full workload at this rate
        line = {"impl": "reference", "metric": "cell_layer_updates_per_s", "value": cb["value"], "unit": "cell-layer updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": cb, "gpu_launches": 0,
                "e2e": {"value": cb["value"], "unit": "cell-layer updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ---------------------------------------------------------------------------------- our arm
    from beom_b200 import _lib, model

    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps, warm = args.steps, max(args.warmup, 3)

    from beom_b200 import dist as bdist
    bdist.init_comm(rank, world, local_rank)
    # inputs are written once (rank 0) into a directory every rank of this node can read
    shared = [tempfile.mkdtemp(prefix="beom_bench_") if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(shared, src=0)
    tmp = shared[0]
    try:
        t0 = time.perf_counter()
        c = None
        if rank == 0:
            c, blk = make_case(n, nlay, tmp)
            log("[rank 0] inputs written: %.1f s" % (time.perf_counter() - t0))
        if world > 1:
            dist.barrier()
        blk = os.path.join(tmp, "shared_mod_block.f95")
        nd1 = (n + 1) * (n + 1) + 1
        opt = model.default_options(fused=not args.split, rank=rank, nranks=world, device=local_rank)
        gm = hl = uu = vv = None
        # read_input_data holds ~25 GB of host arrays for this grid: at most two ranks do it at a time
        for turn in range(0, world, 2):
            if turn <= rank < turn + 2:
                t0 = time.perf_counter()
                hm = model.HostModel.from_block(blk)
                gm = model.GpuModel(hm.params, hm.fields(), opt)
                first, count, own_first, own_count = gm.point_range()
                if world > 1:
                    gm.set_window(first, count)
                else:
                    first, count = 0, nd1
                hl = gm.pinned((nlay, count)); uu = gm.pinned((nlay, count)); vv = gm.pinned((nlay, count))
                hl[:] = hm.array("hlay")[:, first:first + count]
                uu[:] = hm.array("u")[:, first:first + count]
                vv[:] = hm.array("v")[:, first:first + count]
                hm.close()
                log("[rank %d] read_input_data + beom_gpu_init: %.1f s" % (rank, time.perf_counter() - t0))
            if world > 1:
                dist.barrier()
        if rank == 0:
            for f in os.listdir(tmp):
                if f.endswith(".bin"):
                    os.remove(os.path.join(tmp, f))

        def barrier():
            gm.sync()
            if world > 1:
                dist.barrier()

        updates_per_step = float(n) * n * nlay

        # ---- device-resident timing: W warm-up steps (incl. the 3 start-up steps), then K timed steps
        gm.upload_state(hl, uu, vv)
        gm.advance(1, warm)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        time.sleep(0.3)  # let nvidia-smi start sampling; every rank waits the same, then they line up again
        barrier()
        l0 = gm.launch_count()
        gm.mark(0)
        gm.advance(warm + 1, warm + steps)
        gm.mark(1)
        barrier()
        ms = gm.elapsed_ms()
        launches = gm.launch_count() - l0
        clocks = sampler.finish() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ms_per_step = ms / steps
        value = updates_per_step * steps / (ms * 1.0e-3)
        if args.trace and world == 1:  # diagnostic: how the step time evolves under sustained load (power cap)
            t = warm + steps
            for k in range(args.trace):
                gm.mark(0)
                gm.advance(t + 1, t + 10)
                gm.mark(1)
                gm.sync()
                t += 10
                q = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
                log("[trace] block %d: %.3f ms/step   sm,mem MHz, W, C = %s" % (k, gm.elapsed_ms() / 10.0, q))

        # ---- end to end through the C ABI with host buffers: upload_state + K steps + download_state
        e2e = None
        if not args.no_e2e:
            barrier()
            t0 = time.perf_counter()
            gm.upload_state(hl, uu, vv)
            gm.advance(1, steps)
            gm.download_state((hl, uu, vv))
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            nbytes = 3.0 * nlay * count * 8 * world  # every rank moves its own slab
            e2e = {"value": updates_per_step * steps / dt, "unit": "cell-layer updates/s", "h2d_bytes_per_step": nbytes / steps,
                   "d2h_bytes_per_step": nbytes / steps, "seconds": dt,
                   "what": "beom_gpu_upload_state (pinned host hlay,u,v) + %d steps + beom_gpu_download_state" % steps}
        path = gm.path
        gm.close()
    finally:
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    achieved = ALGO_BYTES_PER_UPDATE * value / 1.0e9 / world  # per GPU
    traffic = None  # DRAM bytes per launch from the committed ncu capture of this kernel on this grid (not measured live)
    try:
        with open(os.path.join(ROOT, "profiles", "r1_fused_traffic.json")) as f:
            tj = json.load(f)
        if path == "fused" and world == tj["n_gpus"] and [n, n, nlay] == tj["grid"]:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "k_fused_step" if path == "fused" else "whole step (split path: one kernel per reference loop)",
                "bytes_per_launch_algorithmic": ALGO_BYTES_PER_UPDATE * updates_per_step / world,
                "launch_ms": ms_per_step, "per_gpu": True}
    cb = None
    if not args.no_cpu:
        tmp2 = tempfile.mkdtemp(prefix="beom_cpu_")
        try:
            cb = cpu_baseline(args.cpu_sample, nlay, args.cpu_steps, tmp2)
        finally:
            shutil.rmtree(tmp2, ignore_errors=True)
    line = {"metric": "cell_layer_updates_per_s", "value": value, "unit": "cell-layer updates/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(config, path=path), "roofline": roofline, "cpu_baseline": cb,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
