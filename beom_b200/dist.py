"""Multi-GPU plumbing: one process per GPU (torchrun), y-slab decomposition.

The data path (halo rows) uses the library's own NCCL communicator (beom_gpu_comm_init); torch.distributed
is only the rendezvous: it carries the 128-byte NCCL unique id from rank 0 to the others and provides the
barrier / max-reduce that bench.py needs."""
from __future__ import annotations

import ctypes as C
import os

from . import _lib


def slab_rows(mm: int, rank: int, nranks: int) -> tuple[int, int]:
    """Grid rows j0..j1 (inclusive) owned by `rank`: rows 1..mm+1 split evenly, remainder to the first
    ranks (the same rule as beom_gpu_init)."""
    rows = mm + 1
    base, rem = divmod(rows, nranks)
    j0 = 1 + rank * base + min(rank, rem)
    j1 = j0 + base + (1 if rank < rem else 0) - 1
    return j0, j1


def env_rank() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_comm(rank: int, world: int, local_rank: int) -> None:
    """Create the library's NCCL communicator; torch.distributed must already be initialised."""
    if world <= 1:
        return
    import torch.distributed as dist
    lib = _lib.gpu_lib()
    payload = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        if lib.beom_gpu_comm_unique_id(buf):
            raise RuntimeError(_lib.gpu_error())
        payload = [bytes(buf.raw)]
    dist.broadcast_object_list(payload, src=0)
    if lib.beom_gpu_comm_init(payload[0], rank, world, local_rank):
        raise RuntimeError(_lib.gpu_error())
