"""Worker of tests/test_multigpu.py: one rank per GPU under torchrun; compares this rank's slab with
the CPU oracle bit for bit."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from beom_b200 import cases, dist as bdist, model
    from oracle.pyoracle import Oracle

    rank, world, local = bdist.env_rank()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bdist.init_comm(rank, world, local)
    name, nsteps, fused = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    if name == "synthetic_basin":
        c = cases.synthetic_basin(n=300, mm=170, nlay=4)
    elif name == "rigid_lid_basin":  # the rigid lid across y-slabs: surf_pressure in rounds (beom_gpu.cu, pi_solve_slabs)
        c = cases.synthetic_basin(n=300, mm=170, nlay=2)
        c.params_text += "rgld       = 1.\nocrp       = 1.\n"
    else:
        from tests.conftest import SMALL
        c = cases.CASES[name](**SMALL.get(name, {}))
    d = tempfile.mkdtemp(prefix="beom_mg%d_" % rank)
    blk = c.write(d)
    hm = model.HostModel.from_block(blk)
    orc = Oracle(hm.params, d)
    orc.advance(1, nsteps)
    opt = model.default_options(fused=bool(fused), rank=rank, nranks=world, device=local)
    gm = model.GpuModel.from_grids(hm.params, d, opt) if len(sys.argv) > 4 and sys.argv[4] == "grids" else None  # device-side init
    if gm is None:
        gm = model.GpuModel(hm.params, hm.fields(), opt)
        gm.upload_state(hm.array("hlay"), hm.array("u"), hm.array("v"))
    first, count, own_first, own_count = gm.point_range()
    gm.advance(1, nsteps)
    hl, u, v = gm.download_state()
    aux = gm.download_aux()
    sl = slice(own_first, own_first + own_count)
    # frozen periodic duplicates carry no meaningful fluxes or histories (their state proper is compared)
    keep = np.ones(own_count, dtype=bool)
    if hm.params.xper > 0.5 or hm.params.yper > 0.5:
        sub = hm.iarray("subc")[:, sl]
        keep = ~((sub[0] == c.lm + 1) | (sub[1] == c.mm + 1))
    bad = []
    for nm, got in (("hlay", hl), ("u", u), ("v", v), ("h_u", aux[0]), ("h_v", aux[1])):
        want = orc.array(nm)
        m = keep if nm in ("h_u", "h_v") else slice(None)
        if not np.array_equal(got[:, sl][:, m], want[:, sl][:, m]):
            bad.append("%s (max abs %.3e)" % (nm, float(np.max(np.abs(got[:, sl][:, m] - want[:, sl][:, m])))))
    for nm, got in (("rs_h", aux[2]), ("dmdx", aux[3]), ("dmdy", aux[4])):
        want = orc.array(nm)
        if not np.array_equal(got[:, sl][:, keep], want[:, sl][:, keep]):
            bad.append(nm)
    if hm.params.rgld > 0.5:
        pi_s = gm.download_pi_s()
        if not np.array_equal(pi_s[sl], orc.array("pi_s")[0][sl]):
            bad.append("pi_s")
    path = gm.path
    gm.close()
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        print("rank %d/%d %s path: MISMATCH %s" % (rank, world, path, bad), flush=True)
        sys.exit(1)
    print("rank %d/%d %s path: points %d..%d bit-identical to the oracle after %d steps" % (rank, world, path, own_first, own_first + own_count - 1, nsteps), flush=True)


if __name__ == "__main__":
    main()
