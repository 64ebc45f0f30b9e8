"""Multi-GPU parity check, launched as
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Every rank parses its own samples, buckets are exchanged to the partition owners, every rank merges
its partitions; rank 0 gathers all matrix files and compares them with the CPU oracle run over all
samples.  Exit code 0 == bit-identical."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # as bench.py: one hardware queue per lane stream

import numpy as np
import torch
import torch.distributed as dist


MODES = (("hash:bf:bin", dict(bloom_size=400_000, soft_min=2, share_min=2)), ("kmer:count:bin", {}),
         ("kmer:pa:bin", dict(kmer_size=63, soft_min=2, recurrence_min=2)))


def check(rank, world, local, n_local=3, P=8, reads=4000, modes=MODES, quiet=False):
    """Runs the N-rank path on small seeded samples and compares every matrix, merge_info, counts file and .pinfo with
    the CPU oracle (rank 0).  torch.distributed (nccl) must be initialised.  Returns True on every rank iff bit-identical."""
    from kmtricks_b200 import dist as kd, engine, formats, synth
    P = max(P, world)
    ok = True
    for mode, extra in modes:
        extra = dict(extra)
        k = extra.pop("kmer_size", 31)
        cfg = engine.Config(kmer_size=k, nb_partitions=P, mode=mode, hard_min=2, **extra)
        N = world * n_local
        eng = engine.Engine(cfg, N, device=local)
        kd.init_engine(eng, nlanes=2)
        texts = [synth.make_fastq(31, kd.global_slot(rank, n_local, i), reads, L=150, G=30000, d=4e-3, e=4e-3, revcomp=True) for i in range(n_local)]
        bufs = [C.create_string_buffer(t, len(t)) for t in texts]
        ptrs = (C.c_void_p * n_local)(*[C.addressof(b) for b in bufs])
        sizes = (C.c_size_t * n_local)(*[len(t) for t in texts])
        hm = (C.c_uint32 * n_local)(*([2] * n_local))
        pin = np.zeros((n_local, P), dtype=np.uint64)
        eng._ck(eng.lib.kmx_dist_run_samples(eng.h, n_local, ptrs, sizes, 0, hm, pin.ctypes.data_as(C.POINTER(C.c_uint64))), "dist_run_samples")
        mine = {}
        for p in kd.owned_partitions(P, world, rank):
            m = eng.merge(p)
            mine[p] = (eng.matrix_file(p, m), formats.merge_info(m["stats"]), [eng.counts_file(s, p) for s in range(N)])
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, texts, pin))
        if rank == 0:
            from oracle import oracle as O
            all_texts = [None] * N
            for r, (_, tx, _) in enumerate(gathered):
                for i, t in enumerate(tx):
                    all_texts[kd.global_slot(r, n_local, i)] = [t]
            prm = O.Params(k=k, P=P, mode=mode, hard_min=2, soft_min=cfg.soft_min, recurrence_min=cfg.recurrence_min,
                           share_min=cfg.share_min, bloom_size=cfg.bloom_size)
            want = O.run_pipeline(all_texts, prm)
            seen = set()
            for r, (mats, _, pins) in enumerate(gathered):
                for i in range(n_local):
                    if list(map(int, pins[i])) != list(map(int, want["pinfo"][kd.global_slot(r, n_local, i)])):
                        print("pinfo mismatch", r, i); ok = False
                for p, (mat, mi, cfiles) in mats.items():
                    seen.add(p)
                    if mat != want["matrices"][p]:
                        print(mode, "matrix mismatch partition", p, "owner", r); ok = False
                    if mi != want["merge_info"][p]:
                        print(mode, "merge_info mismatch partition", p); ok = False
                    for s in range(N):
                        if cfiles[s] != want["counts"][(s, p)]:
                            print(mode, "counts mismatch", s, p); ok = False
            if seen != set(range(P)):
                print("partitions not covered", seen); ok = False
            if not quiet:
                print(mode, "world", world, "OK" if ok else "FAIL", flush=True)
        eng.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.broadcast(flag, src=0)
    return int(flag.item()) == 0


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = check(rank, world, local)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
