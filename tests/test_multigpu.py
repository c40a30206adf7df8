"""y-slab decomposition.  CPU part: the slab arithmetic and the rendezvous plumbing over gloo
(world_size 2).  GPU part (needs >= 2 GPUs): a 2-rank run, halo rows exchanged with NCCL, each rank's
slab compared bit for bit with the CPU oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_rows_cover_the_grid():
    from beom_b200.dist import slab_rows
    for mm in (1, 2, 7, 63, 501, 8192):
        for n in (1, 2, 3, 4, 8):
            if n > mm + 1:
                continue
            rows = [slab_rows(mm, r, n) for r in range(n)]
            assert rows[0][0] == 1 and rows[-1][1] == mm + 1
            for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                assert b0 == a1 + 1 and a1 >= a0
            sizes = [b - a + 1 for a, b in rows]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from beom_b200.dist import slab_rows
        payload = [b"x" * 128 if rank == 0 else None]  # stands in for the NCCL unique id
        dist.broadcast_object_list(payload, src=0)
        rows = [None] * world
        dist.all_gather_object(rows, slab_rows(501, rank, world))
        q.put((rank, payload[0] == b"x" * 128, rows))
    finally:
        dist.destroy_process_group()


def test_rendezvous_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in out)
    assert out[0][2] == out[1][2] == [(1, 251), (252, 502)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,nsteps,fused,how", [
    ("synthetic_basin", 12, 1, ""), ("synthetic_basin", 9, 0, ""), ("sill_exchange3D", 12, 1, ""),
    ("conservation", 12, 0, ""),      # y-periodic: ring-closed exchange
    ("conservation", 12, 1, ""),      # ... and the fused step across the ring
    ("unstable_jet", 12, 1, ""),
    ("soliton", 12, 0, ""),           # x-periodic slabs
    ("soliton", 12, 1, ""),           # ... each a torus of its own: fused
    ("rigid_lid_basin", 4, 0, ""),    # surf_pressure across the slabs (18 and 55 sweeps in steps 3, 4)
    ("synthetic_basin", 12, 1, "grids"),  # every rank initialised on the device from the raw files (beom_gpu_init_grids)
    ("sill_exchange3D", 12, 1, "grids")])
def test_two_ranks_bit_exact(name, nsteps, fused, how):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 400), os.path.join(ROOT, "tests", "mgpu_worker.py"), name, str(nsteps), str(fused)] + ([how] if how else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-identical") == 2, r.stdout
