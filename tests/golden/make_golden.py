#!/usr/bin/env python
"""Generates tests/golden/*.npz -- outputs of the UNMODIFIED reference (oracle/_ref/bin, built by
oracle/build_ref.sh from /root/reference) on small inputs, plus a copy of the reference's own
tiny test DATA fixtures (tests/data: 2 FASTA files of 2x99 nt, the partition goldens that
tests/merge_test.cpp and tests/task_main.cpp check, hash.info, the fixture repartition table).

Run in the authoring container only (needs /root/reference):  python tests/golden/make_golden.py
The .npz files travel with the repo; /root/reference is never read at test time.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from kmtricks_b200 import synth          # noqa: E402
from oracle import oracle as O           # noqa: E402
from tests.conftest import EDGE_FASTA, EDGE_FASTQ_CRLF   # noqa: E402

REF = "/root/reference"
EXT = {("kmer", "count"): "count", ("kmer", "pa"): "pa", ("hash", "count"): "count_hash", ("hash", "pa"): "pa_hash", ("hash", "bf"): "cmbf"}


def run_case(name, samples, prm, fof_hard_min=None, bft=False):
    """samples: list of list of (filename, bytes)."""
    d = tempfile.mkdtemp(prefix="gold_")
    try:
        fof = os.path.join(d, "fof.txt")
        with open(fof, "w") as f:
            for i, files in enumerate(samples):
                paths = []
                for fn, b in files:
                    p = os.path.join(d, f"S{i}_{fn}")
                    open(p, "wb").write(b)
                    paths.append(p)
                line = f"S{i}: " + " ; ".join(paths)
                if fof_hard_min and i in fof_hard_min:
                    line += f" ! {fof_hard_min[i]}"
                f.write(line + "\n")
        rd = os.path.join(d, "run")
        O.run_reference(fof, rd, prm, threads=2)
        kind, what = prm.mode.split(":")[:2]
        out = {"mode": prm.mode, "k": prm.k, "m": prm.m, "P": prm.P, "hard_min": prm.hard_min, "soft_min": prm.soft_min,
               "recurrence_min": prm.recurrence_min, "share_min": prm.share_min, "bloom_size": prm.bloom_size,
               "n_samples": len(samples)}
        for i, files in enumerate(samples):
            out[f"n_files_{i}"] = len(files)
            for j, (fn, b) in enumerate(files):
                out[f"input_{i}_{j}"] = np.frombuffer(b, dtype=np.uint8)
            if fof_hard_min and i in fof_hard_min:
                out[f"fof_hard_min_{i}"] = fof_hard_min[i]
            out[f"pinfo_{i}"] = np.array([int(x) for x in open(f"{rd}/partition_infos/S{i}.pinfo").read().split()], dtype=np.uint64)
            for p in range(prm.P):
                cext = "hash" if kind == "hash" else "kmer"
                out[f"counts_{i}_{p}"] = np.fromfile(f"{rd}/counts/partition_{p}/S{i}.{cext}", dtype=np.uint8)
        for p in range(prm.P):
            out[f"matrix_{p}"] = np.fromfile(f"{rd}/matrices/matrix_{p}.{EXT[(kind, what)]}", dtype=np.uint8)
            out[f"merge_info_{p}"] = np.fromfile(f"{rd}/merge_infos/partition{p}.merge_info", dtype=np.uint8)
            if bft:   # HashMerger::write_as_bft through the harness (the CLI cannot reach it)
                W = O.window_bits(prm.bloom_size, prm.P)
                files = [f"{rd}/counts/partition_{p}/S{i}.hash" for i in range(len(samples))]
                o = os.path.join(d, f"bft_{p}")
                subprocess.run([os.path.join(ROOT, "oracle/_ref/bin/bft_harness"), "bft", o, str(W * p), str(W * (p + 1) - 1),
                                str(prm.soft_min), str(prm.recurrence_min), str(prm.share_min)] + files, check=True)
                out[f"bft_{p}"] = np.fromfile(o, dtype=np.uint8)
        out["hash_info"] = np.fromfile(f"{rd}/hash.info", dtype=np.uint8)
        out["minim_repart"] = np.frombuffer(gzip.compress(open(f"{rd}/repartition_gatb/repartition.minimRepart", "rb").read(), 9, mtime=0), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "ok", sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray)), "bytes raw")
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    assert O.have_ref(), "build oracle/_ref first (oracle/build_ref.sh)"
    # ---- the reference's own DATA fixtures (not source)
    fx = os.path.join(HERE, "ref_fixture")
    shutil.rmtree(fx, ignore_errors=True)
    os.makedirs(fx)
    for f in ("1.fasta", "2.fasta", "hash.info"):
        shutil.copy(f"{REF}/tests/data/{f}", fx)
    open(f"{fx}/repartition.minimRepart.gz", "wb").write(gzip.compress(open(f"{REF}/tests/data/repart_gatb/repartition.minimRepart", "rb").read(), 9, mtime=0))
    for kind, ext in (("kmers", "kmer"), ("hashes", "hash")):
        for p in range(4):
            for s in ("D1", "D2"):
                os.makedirs(f"{fx}/partitions/{kind}/partition_{p}", exist_ok=True)
                shutil.copy(f"{REF}/tests/data/partitions/{kind}/partition_{p}/{s}.{ext}", f"{fx}/partitions/{kind}/partition_{p}/")
    # literal golden vectors of tests/task_main.cpp (count_task tests), as data
    import json, re
    src = open(f"{REF}/tests/task_main.cpp").read()
    i2 = src.index("TEST(count_task, kmer_count_task)"); i3 = src.index("TEST(count_task, hash_count_task)")
    gold = {}
    for m in re.finditer(r'HashReader<[^>]*> kr\("\./tests_tmp/km_dir_test/counts/partition_(\d)/(D\d)\.hash"\);(.*?)\n  \}', src[i3:], re.S):
        v = [(int(a), int(b)) for a, b in re.findall(r"EXPECT_EQ\(kmer, (\d+)\); EXPECT_EQ\(count, (\d+)\)", m.group(3))]
        if v: gold[f"hash_{m.group(2)}_p{m.group(1)}"] = v
    for m in re.finditer(r'KmerReader<[^>]*> kr\("\./tests_tmp/km_dir_test/counts/partition_(\d)/(D\d)\.kmer"\);(.*?)\n  \}', src[i2:i3], re.S):
        v = [(a, int(b)) for a, b in re.findall(r'EXPECT_EQ\(kmer\.to_string\(\), "([ACGT]+)"\); EXPECT_EQ\((\d+), count\)', m.group(3))]
        if v: gold[f"kmer_{m.group(2)}_p{m.group(1)}"] = v
    json.dump(gold, open(f"{fx}/task_main_goldens.json", "w"))
    for r, _, fs in os.walk(fx):
        for f in fs:
            os.chmod(os.path.join(r, f), 0o644)

    ref1 = open(f"{REF}/tests/data/1.fasta", "rb").read(); ref2 = open(f"{REF}/tests/data/2.fasta", "rb").read()
    syn = [[("r.fastq", synth.make_fastq(11, s, 500, L=100, G=3000, d=8e-3, e=8e-3, revcomp=True))] for s in range(3)]
    edge = [[("e.fasta", EDGE_FASTA)], [("e2.fasta", EDGE_FASTA[:400] + b"\n"), ("crlf.fastq", EDGE_FASTQ_CRLF)]]
    run_case("cfg1_ref_fixture_kmer_count", [[("1.fasta", ref1)], [("2.fasta", ref2)]], O.Params(k=31, P=4, mode="kmer:count:bin", hard_min=1))
    run_case("syn_kmer_count", syn, O.Params(k=31, P=4, mode="kmer:count:bin", hard_min=2))
    run_case("syn_kmer_count_fof_hardmin", syn, O.Params(k=31, P=4, mode="kmer:count:bin", hard_min=2), fof_hard_min={1: 1})
    run_case("syn_kmer_pa_rescue", syn, O.Params(k=31, P=4, mode="kmer:pa:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=2))
    run_case("syn_hash_bf_rescue", syn, O.Params(k=31, P=4, mode="hash:bf:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=1, bloom_size=40_000), bft=True)
    run_case("syn_hash_count", syn, O.Params(k=31, P=4, mode="hash:count:bin", hard_min=2, bloom_size=40_000))
    run_case("syn_hash_pa", syn, O.Params(k=31, P=4, mode="hash:pa:bin", hard_min=2, soft_min=2, recurrence_min=2, bloom_size=40_000))
    run_case("syn_k63_kmer_pa", syn, O.Params(k=63, P=4, mode="kmer:pa:bin", hard_min=1, soft_min=2, share_min=2, recurrence_min=2))
    run_case("syn_k63_hash_bf", syn, O.Params(k=63, P=4, mode="hash:bf:bin", hard_min=2, bloom_size=30_000))
    run_case("syn_k21_m8", syn, O.Params(k=21, m=8, P=5, mode="kmer:count:bin", hard_min=1))
    run_case("edge_kmer_count", edge, O.Params(k=31, P=4, mode="kmer:count:bin", hard_min=1))
    run_case("edge_hash_bf", edge, O.Params(k=31, P=4, mode="hash:bf:bin", hard_min=1, bloom_size=20_000), bft=True)


if __name__ == "__main__":
    main()
