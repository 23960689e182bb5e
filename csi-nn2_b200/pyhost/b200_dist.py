"""Host-side plumbing for batch sharding across the GPUs of one box.

The inference path has NO per-step collective: images are independent units, so a global batch is
cut into contiguous NCHW slices, one per rank, and every rank runs its own session on its own GPU.
The only communication is one broadcast of the packed weight arena (weights + per-channel tables,
one contiguous device allocation, see shl_b200_session_weight_arena in include/shl_b200.h) from
rank 0 after session_setup -- NCCL over NVLink on the GPU box, gloo in the CPU tests -- plus the
max-over-ranks reduction of the timings bench.py reports.

torch.distributed is plumbing here (process group, broadcast), not the product.
"""
from __future__ import annotations

from typing import Tuple


def shard_batch(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [start, start + count) of a batch of independent images owned by `rank`.
    Remainders go to the lowest ranks, so slices differ by at most one image and cover the batch."""
    if world <= 0 or not 0 <= rank < world or global_batch < 0:
        raise ValueError(f"bad shard request: batch {global_batch}, world {world}, rank {rank}")
    base, extra = divmod(global_batch, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def broadcast_arena(arena, src: int = 0):
    """Broadcast a flat uint8 tensor (the weight arena) from `src` in place; returns it."""
    import torch.distributed as dist

    if arena.dtype.is_floating_point or arena.dim() != 1:
        raise ValueError("the weight arena is a flat byte tensor")
    dist.broadcast(arena, src=src)
    return arena


def max_over_ranks(values, device=None):
    """Element-wise max over ranks of a list of floats (device times are reported as the max)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def device_bytes_as_tensor(ptr: int, nbytes: int, device_index: int):
    """A torch uint8 view of raw device memory owned by the b200 runtime (no copy)."""
    import torch

    class _Raw:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                    "version": 2}

    return torch.as_tensor(_Raw(), device=torch.device("cuda", device_index))
