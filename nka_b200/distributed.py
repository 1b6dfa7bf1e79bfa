"""One process per GPU: row-slab partition of every vector and the rendezvous for the
library's built-in reduction of the partial dot products.

The reference anticipates exactly this use ("works without modification when called
collectively, passing each processing element's portion of the vector", with a global dot
product supplied through the dp hook: src-F08-vector/README.md:16-22,
src-C/nonlinear_krylov_accelerator.c:61-68).  Here the hook is built in: each rank's pass A
reduces its slab, one 66-double NCCL all-reduce combines the ranks, and every rank runs the
same scalar state step on the same bits, so drop decisions are identical everywhere.

torch.distributed is used only as plumbing (to ship the 128-byte NCCL id).
"""
from __future__ import annotations

from .nka import NKA, comm_unique_id


def slab_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced, 2-element-aligned slabs: [lo, hi) of rank `rank`.

    Boundaries are even so every slab starts 16-byte aligned inside a 16-byte-aligned global
    vector (keeps the LDG.128 path)."""
    if world < 1 or not 0 <= rank < world or n < 0:
        raise ValueError("bad partition arguments")
    pairs = (n + 1) // 2
    lo = 2 * ((pairs * rank) // world)
    hi = 2 * ((pairs * (rank + 1)) // world)
    return min(lo, n), min(hi, n) if rank < world - 1 else n


def exchange_unique_id(group=None, src: int = 0) -> bytes:
    """Rank `src` creates the NCCL id, everyone receives it (any torch.distributed backend)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    box = [comm_unique_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def distributed_nka(n_global: int, mvec: int, vtol: float = 0.01, group=None, device: int = -1,
                    stream: int | None = None) -> tuple[NKA, int, int]:
    """Collective: every rank gets an accelerator over its slab of an n_global-long vector.
    Returns (accelerator, lo, hi)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = slab_bounds(n_global, world, rank)
    acc = NKA(hi - lo, mvec, vtol, device=device, stream=stream)
    if world > 1:
        acc.comm_init(world, rank, exchange_unique_id(group))
    return acc, lo, hi
