"""Multi-GPU host plumbing: one process per GPU (torchrun), torch.distributed for the rendezvous,
the exchange itself is native (kmx_dist_run_samples: grouped ncclSend/ncclRecv of bucket regions).

Sharding: stage 1 over samples (rank r parses its own n_local samples, global slot
r*n_local + i), stages 2-4 over partitions (rank g owns the contiguous block
[g*P//world, (g+1)*P//world) -- the same formula as part_first() in csrc/kmx_dist.inl).
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def part_first(P: int, world: int, g: int) -> int:
    return (g * P) // world


def owned_partitions(P: int, world: int, rank: int) -> range:
    return range(part_first(P, world, rank), part_first(P, world, rank + 1))


def owner_of(P: int, world: int, p: int) -> int:
    for g in range(world):
        if part_first(P, world, g) <= p < part_first(P, world, g + 1):
            return g
    raise ValueError(p)


def global_slot(rank: int, n_local: int, i: int) -> int:
    return rank * n_local + i


def exchange_plan(boff_end: np.ndarray, P: int, world: int, rank: int, all_boff_end: np.ndarray):
    """Record counts this rank sends to / receives from every peer for one batch.
    boff_end: this rank's bucket offsets [P+1] (record units, last = end of slab);
    all_boff_end: [world, P+1] gathered from every rank.  Mirrors dist_batch() in kmx_dist.inl."""
    send = np.array([boff_end[part_first(P, world, g + 1)] - boff_end[part_first(P, world, g)] for g in range(world)], dtype=np.int64)
    f, l = part_first(P, world, rank), part_first(P, world, rank + 1)
    recv = np.array([all_boff_end[g][l] - all_boff_end[g][f] for g in range(world)], dtype=np.int64)
    return send, recv


def init_engine(eng, nlanes: int = 2):
    """Creates the engine's NCCL communicators (one per lane).  torch.distributed must be initialised."""
    import torch.distributed as dist
    L = eng.lib
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [None]
    if rank == 0:
        buf = (C.c_uint8 * (128 * nlanes))()
        for t in range(nlanes):
            rc = L.kmx_dist_unique_id(C.byref(buf, 128 * t))
            if rc:
                raise RuntimeError("kmx_dist_unique_id failed (libnccl.so.2 not loadable?)")
        ids = [bytes(buf)]
    dist.broadcast_object_list(ids, src=0)
    raw = (C.c_uint8 * (128 * nlanes)).from_buffer_copy(ids[0])
    eng._ck(L.kmx_dist_init(eng.h, rank, world, nlanes, raw), "dist_init")
    return rank, world


def run_layout(n_samples: int, world: int, batch: int):
    """Batches of a streamed multi-GPU run (what `kmx pipeline --devices` and bench.py's other_configs do): rank r parses the
    samples [r*nl, (r+1)*nl), nl = ceil(n_samples / world), in batches of `batch`; every rank makes the same number of
    kmx_dist_run_batch calls with the same (n, slot_base, nl) -- slots >= n_samples are empty padding samples.
    Returns nl and, per batch, (slot_base, n, [per-rank list of global sample indices or None for padding])."""
    nl = (n_samples + world - 1) // world
    out = []
    for b0 in range(0, nl, batch):
        n = min(batch, nl - b0)
        out.append((b0, n, [[(r * nl + b0 + i if r * nl + b0 + i < n_samples else None) for i in range(n)] for r in range(world)]))
    return nl, out
