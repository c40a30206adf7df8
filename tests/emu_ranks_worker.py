"""Runs in its own process: N y-slab ranks of the CPU EMULATION of the library's split path (tools/emu), each rank its own
copy of the emulated library driven by its own thread, the halo exchange handed to a Python callback that pairs the
messages like NCCL.  Compares every rank's own slab with the CPU oracle bit for bit.  Test infrastructure.

usage: emu_ranks_worker.py <libbeom_gpu_emu.so> <case> <nsteps> <nranks> ['{"param": "value"}' ['{"kwarg": value}' [fused 0|1]]]"""
import ctypes as C
import json
import os
import shutil
import sys
import tempfile
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from beom_b200 import _lib, cases, model  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from tests.conftest import SMALL  # noqa: E402

so, name, nsteps, nranks = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
extra = json.loads(sys.argv[5]) if len(sys.argv) > 5 else {}
kwargs = json.loads(sys.argv[6]) if len(sys.argv) > 6 and sys.argv[6] != "null" else SMALL.get(name, {})
fused = int(sys.argv[7]) if len(sys.argv) > 7 else 0

EXCH = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                   C.c_int, C.c_size_t)
barrier = threading.Barrier(nranks, timeout=120)
box = {}
calls = [0] * nranks


def make_exchange(rank):
    def exchange(send_lo, recv_lo, peer_lo, send_hi, recv_hi, peer_hi, count):
        try:
            calls[rank] += 1
            if peer_lo >= 0:
                box[(rank, "lo")] = np.ctypeslib.as_array(send_lo, shape=(count,)).copy()
            if peer_hi >= 0:
                box[(rank, "hi")] = np.ctypeslib.as_array(send_hi, shape=(count,)).copy()
            barrier.wait()
            if peer_lo >= 0:  # the lower neighbour's upward message
                np.ctypeslib.as_array(recv_lo, shape=(count,))[:] = box[(peer_lo, "hi")]
            if peer_hi >= 0:  # the upper neighbour's downward message
                np.ctypeslib.as_array(recv_hi, shape=(count,))[:] = box[(peer_hi, "lo")]
            barrier.wait()
            return 0
        except Exception as e:  # noqa: BLE001 (a broken barrier = the ranks disagree about the number of exchanges)
            print("exchange failed on rank %d: %r" % (rank, e), flush=True)
            return -1
    return EXCH(exchange)


GRID_INIT = bool(os.environ.get("BEOM_GRID_INIT"))  # initialise every rank on the "device" from the raw files (beom_gpu_init_grids)


class EmuRank(model.GpuModel):
    def __init__(self, lib, params, fields, options, idir=None):
        self.lib, self.params, self.opt = lib, params, options
        self.nlay, self.ndeg = params.nlay, params.ndeg
        if idir is None:
            self._ck(lib.beom_gpu_init(C.byref(params), C.byref(fields), C.byref(options)), "beom_gpu_init")
            return
        gr = _lib.Grids()
        self._keep = []
        for k in model.GpuModel.GRID_FILES:
            path = os.path.join(idir, k + ".bin")
            if os.path.exists(path):
                a = np.fromfile(path, dtype="<f4")
                self._keep.append(a)
                setattr(gr, k, a.ctypes.data_as(C.POINTER(C.c_float)))
        self._ck(lib.beom_gpu_init_grids(C.byref(params), C.byref(gr), C.byref(options)), "beom_gpu_init_grids")

    def _ck(self, rc, what):
        if rc:
            buf = C.create_string_buffer(2048)
            self.lib.beom_gpu_last_error(buf, len(buf))
            raise RuntimeError("%s: %s" % (what, buf.value.decode(errors="replace")))


c = cases.CASES[name](**kwargs)
c.params_text += "".join("%-10s = %s\n" % kv for kv in extra.items())
work = tempfile.mkdtemp(prefix="beom_emu_ranks_")
hm = model.HostModel.from_block(c.write(work))
orc = Oracle(hm.params, work)
orc.advance(1, nsteps)
results = [None] * nranks
keep_cb = []


def run_rank(rank):
    try:
        path = os.path.join(work, "rank%d.so" % rank)  # its own file = its own library instance = its own globals
        shutil.copy(so, path)
        lib = _lib.bind_gpu(C.CDLL(path, mode=C.RTLD_LOCAL))
        assert b"cpu-emulation" in lib.beom_gpu_version()
        cb = make_exchange(rank)
        keep_cb.append(cb)
        lib.emu_comm_set.argtypes = [C.c_int, C.c_int, EXCH]
        lib.emu_comm_set(rank, nranks, cb)
        opt = model.Options()
        lib.beom_gpu_default_options(C.byref(opt))
        opt.fused, opt.rank, opt.nranks, opt.device = fused, rank, nranks, 0
        gm = EmuRank(lib, hm.params, hm.fields(), opt, idir=work if GRID_INIT else None)
        first, count, own_first, own_count = gm.point_range()
        if not GRID_INIT:
            gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
        gm.advance(1, nsteps)
        state = gm.download_state()
        aux = gm.download_aux()
        sl = slice(own_first, own_first + own_count)
        keep = np.ones(own_count, dtype=bool)  # frozen periodic duplicates carry no meaningful fluxes or histories
        if hm.params.xper > 0.5 or hm.params.yper > 0.5:
            sub = hm.iarray("subc")[:, sl]
            keep = ~((sub[0] == c.lm + 1) | (sub[1] == c.mm + 1))
        bad = []
        for nm, got in zip(("hlay", "u", "v"), state):
            if not np.array_equal(got[:, sl], orc.array(nm)[:, sl]):
                bad.append(nm)
        for nm, got in zip(("h_u", "h_v", "rs_h", "dmdx", "dmdy"), aux):
            if not np.array_equal(got[:, sl][:, keep], orc.array(nm)[:, sl][:, keep]):
                bad.append(nm)
        sweeps = None
        if hm.params.rgld > 0.5:  # the rigid lid: the surface pressure of the own points and the number of sweeps of the last solve
            pi_s = gm.download_pi_s()
            if not np.array_equal(pi_s[sl], orc.array("pi_s")[0][sl]):
                bad.append("pi_s")
            sweeps = gm.pi_iterations()
        results[rank] = {"rank": rank, "points": [int(own_first), int(own_first + own_count - 1)], "bad": bad,
                         "path": lib.beom_gpu_path().decode(), "exchanges": calls[rank], "pi_sweeps": sweeps}
        lib.beom_gpu_finalize()
    except Exception as e:  # noqa: BLE001
        results[rank] = {"rank": rank, "error": repr(e)}
        barrier.abort()


threads = [threading.Thread(target=run_rank, args=(r,)) for r in range(nranks)]
for t in threads:
    t.start()
for t in threads:
    t.join()
ok = all(r is not None and not r.get("bad") and "error" not in r for r in results)
print(json.dumps({"case": name, "nranks": nranks, "ranks": results, "moved": float(np.abs(orc.array("u")).max() + np.abs(orc.array("v")).max())}))
shutil.rmtree(work, ignore_errors=True)
sys.exit(0 if ok else 1)
