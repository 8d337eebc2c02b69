"""One-process-per-GPU partitioning of a render and the NCCL plumbing around `Renderer`.

The path shards without any exchange during rendering: every pixel-sample is a pure
function of (pixel, sample index).  Two partitions:

* sample-index range — rank r renders samples [r*S/N, (r+1)*S/N) of every pixel: only the seed
  table changes (`seeds.x += start`);
* tiles — rank r renders the 32x32 tiles t with t % N == r (`rpt_set_tile_partition`).

Either way each rank holds a full-frame accumulator and `ncclReduce(sum)` over NVLink on the
context's stream combines them on the root (`Renderer.comm_reduce_output`).  `torch.distributed`
only carries the 128-byte NCCL unique id and the barriers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def sample_range(total_samples: int, rank: int, nranks: int) -> tuple[int, int]:
    """[start, end) of the sample indices rank `rank` renders; ranges tile [0, total) exactly."""
    base, extra = divmod(total_samples, nranks)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def offset_seeds(seeds: np.ndarray, first_sample: int) -> np.ndarray:
    """Seed table for a rank whose first sample index is `first_sample` (rng.x is the sample index,
    kernels/src/rng.rs:47-49)."""
    out = np.ascontiguousarray(seeds, np.uint32).copy()
    out[:, 0] += np.uint32(first_sample)
    return out


def tile_pixels(width: int, height: int, rank: int, nranks: int) -> np.ndarray:
    n = C.c_uint32(0)
    lib = capi.lib()
    capi.check(lib.rpt_tile_partition_pixels(C.c_uint32(width), C.c_uint32(height), C.c_uint32(rank), C.c_uint32(nranks), None, C.byref(n)),
               "rpt_tile_partition_pixels")
    out = np.zeros(n.value, np.uint32)
    capi.check(lib.rpt_tile_partition_pixels(C.c_uint32(width), C.c_uint32(height), C.c_uint32(rank), C.c_uint32(nranks), capi.ptr(out), C.byref(n)),
               "rpt_tile_partition_pixels")
    return out


def init_comm(renderer, dist, rank: int, world_size: int, device=None) -> None:
    """Create the NCCL communicator of `renderer` across a torch.distributed group: rank 0 makes the
    unique id, the group broadcasts it, every rank joins."""
    import torch

    from .trace import Renderer

    backend = dist.get_backend()
    dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(Renderer.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    renderer.comm_init(uid.cpu().numpy().tobytes(), rank, world_size)
