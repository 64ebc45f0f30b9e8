"""Host-side logic of the N>1 path on CPU: partition ownership and the exchange plan, checked
across two gloo ranks (world_size 2, 127.0.0.1) -- what rank a sends to rank b must be what b
expects from a, and every partition must have exactly one owner."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kmtricks_b200 import dist as kd


def test_ownership_covers_every_partition_once():
    for P in (1, 4, 7, 64, 512):
        for world in (1, 2, 3, 8):
            owners = [kd.owner_of(P, world, p) for p in range(P)]
            assert owners == sorted(owners)
            for g in range(world):
                assert list(kd.owned_partitions(P, world, g)) == [p for p in range(P) if owners[p] == g]
            assert sum(len(kd.owned_partitions(P, world, g)) for g in range(world)) == P


def test_global_slots_are_a_bijection():
    world, n = 4, 5
    assert sorted(kd.global_slot(r, n, i) for r in range(world) for i in range(n)) == list(range(world * n))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    caps = rng.integers(1, 1000, P)
    boff_end = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
    t = torch.from_numpy(boff_end.copy())
    allb = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allb, t)
    all_boff = np.stack([x.numpy() for x in allb])
    send, recv = kd.exchange_plan(boff_end, P, world, rank, all_boff)
    # exchange the plans and check they match pairwise
    plans = [None] * world
    dist.all_gather_object(plans, (send.tolist(), recv.tolist()))
    ok = all(plans[a][0][b] == plans[b][1][a] for a in range(world) for b in range(world))
    ok = ok and int(send.sum()) == int(boff_end[-1])
    # payload round trip with the plan's split sizes (records = int64 tags)
    payload = torch.arange(int(boff_end[-1]), dtype=torch.int64) + rank * 10**9
    outs = [torch.empty(int(n), dtype=torch.int64) for n in recv]
    ins = list(torch.split(payload, send.tolist()))
    for g in range(world):                                      # gloo: emulate all-to-all-v with send/recv pairs
        reqs = []
        if g != rank:
            reqs.append(dist.isend(ins[g].contiguous(), g)); reqs.append(dist.irecv(outs[g], g))
            for r in reqs: r.wait()
        else:
            outs[g].copy_(ins[g])
    f = kd.part_first(P, world, rank)
    for g in range(world):
        lo = int(all_boff[g][f])
        ok = ok and bool((outs[g] == torch.arange(lo, lo + len(outs[g]), dtype=torch.int64) + g * 10**9).all())
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [8, 5])
def test_exchange_plan_is_consistent_across_two_gloo_ranks(P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_streamed_run_layout_covers_every_sample_once_with_padding_at_the_end():
    for n_samples, world, batch in ((500, 8, 8), (1000, 8, 8), (5, 2, 4), (7, 4, 3), (1, 2, 8), (16, 8, 4)):
        nl, batches = kd.run_layout(n_samples, world, batch)
        assert nl * world >= n_samples and (nl - 1) * world < n_samples
        seen = []
        for b0, n, per_rank in batches:
            assert len(per_rank) == world and all(len(x) == n for x in per_rank)         # every rank: same call sequence
            for r, slots in enumerate(per_rank):
                for i, s in enumerate(slots):
                    if s is not None:
                        assert s == kd.global_slot(r, nl, b0 + i)
                        seen.append(s)
                    else:
                        assert kd.global_slot(r, nl, b0 + i) >= n_samples               # padding only beyond the last sample
        assert sorted(seen) == list(range(n_samples))
        assert sum(n for _, n, _ in batches) == nl
