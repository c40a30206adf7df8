"""Runs in its own process on the CPU EMULATION of the library (test infrastructure, like tests/emu_worker.py): a second
beom_gpu_upload_state in the middle of a run -- the state is handed back to the library and the run starts over from it, as after a
restart record -- with the step graphs on (default) or off (BEOM_GRAPH=0 in the environment).  Prints the sha256 of the final state
and the number of steps that ran as graph launches; the caller compares the two runs.

usage: emu_reupload_worker.py <libbeom_gpu_emu.so> <case> <nsteps>"""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from beom_b200 import _lib  # noqa: E402

emu = _lib.bind_gpu(C.CDLL(sys.argv[1], mode=C.RTLD_LOCAL))
assert b"cpu-emulation" in emu.beom_gpu_version()
_lib.host_lib()
_lib._gpu = emu

from beom_b200 import cases, model  # noqa: E402
from tests.conftest import SMALL  # noqa: E402

name, nsteps = sys.argv[2], int(sys.argv[3])
c = cases.CASES[name](**SMALL.get(name, {}))
with tempfile.TemporaryDirectory() as d:
    hm = model.HostModel.from_block(c.write(d))
    gm = model.GpuModel(hm.params, hm.fields(), model.default_options(fused=0))
    gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    gm.advance(1, nsteps)
    g1 = gm.graph_launch_count()
    hl, u, v = gm.download_state()
    gm.upload_state(hl, u, v)          # drops the graphs: they were captured with the stress state and buffer phases of the first leg
    gm.advance(1, nsteps)
    g2 = gm.graph_launch_count()
    state = gm.download_state()
    gm.close()
    h = hashlib.sha256()
    for a in state:
        h.update(a.tobytes())
    print(json.dumps({"case": name, "sha256": h.hexdigest(), "graph_steps": [g1, g2 - g1]}))
