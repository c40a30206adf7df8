"""The drop-in boundary, executed: THE REFERENCE'S OWN PROGRAM driving the B200 library through the C ABI.

oracle/refbuild.py builds two programs from the reference's sources (translated to C++ by oracle/f95c; no Fortran compiler
exists): the pure reference, and the same program with the edits of INTEGRATION.md section 2 applied to the text of
private_mod.f95 and main.f95 -- gpu_setup() at the end of read_input_data, distribute_stress / first_three_timesteps /
gener_forward_backward replaced one for one by beom_gpu_stress / beom_gpu_step, beom_gpu_download_state before
write_outputs, beom_gpu_finalize before quit() -- linked against beom_b200/lib/libbeom_gpu.so.  read_input_data, the time
loop, write_outputs and write_array are the reference's own code in both; the module arrays the library receives are the
reference's own, in the reference's layout.

Both run here on the same input files; the drop-in program must write BYTE-IDENTICAL output files (eta_.bin, u___.bin,
v___.bin, h_0.bin, grid.bin, time.txt) and leave a bit-identical hlay, u, v -- on both kernel paths.  The binaries are
made in the development container (oracle/_ref/pure_* and dropin_*, __graft_entry__.build()) and travel with the snapshot.
The same comparison runs in the CPU suite against the emulated library (tests/test_emulation.py)."""
import filecmp
import os
import shutil
import tempfile

import numpy as np
import pytest

from oracle import refbuild

FILES = ("eta_.bin", "u___.bin", "v___.bin", "h_0.bin", "grid.bin", "time.txt")


def run_pair(name, exe_pure, exe_dropin, fused, root):
    from beom_b200 import cases

    gen, kw, nsteps = refbuild.DROPIN_CASES[name]
    c = cases.CASES[gen](**kw)
    blk = refbuild.named_block(c)
    out = {}
    for tag, exe, env in (("pure", exe_pure, {}), ("dropin", exe_dropin, {"BEOM_DROPIN_FUSED": str(fused)})):
        d = os.path.join(root, "%s_%s_%d" % (name[:8], tag, fused)) + "/"
        shutil.rmtree(d, ignore_errors=True)
        c.write(d)
        dump, stdout = refbuild.run_with_env(exe, d, blk, nsteps, env)
        out[tag] = (d, dump, stdout)
    return out


def check_pair(out):
    for f in FILES:
        assert filecmp.cmp(out["pure"][0] + f, out["dropin"][0] + f, shallow=False), "%s differs" % f
    for k in ("hlay", "u", "v"):
        a, b = out["pure"][1][k], out["dropin"][1][k]
        assert np.array_equal(a.view(np.int64), b.view(np.int64)), k
    recs = np.fromfile(out["dropin"][0] + "eta_.bin", dtype="<f4")
    assert recs.size > 0 and np.all(np.isfinite(recs)) and float(out["dropin"][1]["ctim"]) > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [1, 0], ids=["fused", "split"])
@pytest.mark.parametrize("name", list(refbuild.DROPIN_CASES))
def test_reference_program_on_the_gpu_library_writes_identical_files(name, fused):
    exe_pure, exe_dropin = refbuild.prebuilt("pure_" + name), refbuild.prebuilt("dropin_" + name)
    if not exe_pure or not exe_dropin:
        pytest.skip("oracle/_ref/{pure,dropin}_%s were not built (they are made where /root/reference exists)" % name)
    root = tempfile.mkdtemp(prefix="di", dir="/tmp")
    try:
        check_pair(run_pair(name, exe_pure, exe_dropin, fused, root))
    finally:
        shutil.rmtree(root, ignore_errors=True)
