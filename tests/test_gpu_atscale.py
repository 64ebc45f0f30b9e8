"""GPU parity at the sizes bench.py times, against the UNMODIFIED reference binary run on the GPU box itself
(oracle/_ref travels with the snapshot; /root/reference is never read):

* BASELINE cfg2's exact per-sample shape (1 M reads x 150 nt, k=31, P=64, Bloom 2e8, hard-min 2, hash:bf:bin):
  every counts/partition_P/<id>.hash file and every matrices/matrix_P.cmbf byte-compared;
* parity-sized twins of cfg3 / cfg4 / cfg5 (8 samples x 250 k reads, P = 512 / 512 / 256, kmer:count, hash:bft,
  k=63 kmer:pa + rescue): matrices, merge_infos, counts and .pinfo byte-compared; the bft bodies against the
  reference's own HashMerger::write_as_bft driven by oracle/ref_harness/bft_harness.cpp.

Both arms read the same bytes: the text is generated on the device (kmx_synth_fastq), copied to the host, written
to /dev/shm for the reference CLI and handed to kmx_run_samples (4 lanes, the path bench.py times)."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT = {("kmer", "count"): "count", ("kmer", "pa"): "pa", ("hash", "bf"): "cmbf", ("hash", "bft"): "cmbf"}


def _device_texts(N, R, Lr, G, d, e, seed):
    """N samples of R reads each from the device generator, as host bytes."""
    from kmtricks_b200 import engine, synth
    eng = engine.Engine(engine.Config(kmer_size=31, nb_partitions=4, mode="kmer:count:bin"), 1)
    L, h = eng.lib, eng.h
    sb = R * synth.record_bytes(Lr)
    dev = C.c_void_p()
    assert L.kmx_dev_alloc(h, sb + 64, C.byref(dev)) == 0
    out = []
    for s in range(N):
        assert L.kmx_synth_fastq(h, seed, s, 0, R, Lr, G, d, e, 1, dev.value) == 0
        buf = np.empty(sb, dtype=np.uint8)
        assert L.kmx_memcpy_d2h(h, buf.ctypes.data, dev, sb) == 0
        out.append(buf.tobytes())
    eng.close()
    return out


def _reference_run(texts, prm, until=None):
    from oracle import oracle as O
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    d = tempfile.mkdtemp(prefix="kmx_scale_", dir=base)
    with open(f"{d}/fof.txt", "w") as f:
        for i, b in enumerate(texts):
            open(f"{d}/S{i}.fastq", "wb").write(b)
            f.write(f"S{i}: {d}/S{i}.fastq\n")
    O.run_reference(f"{d}/fof.txt", f"{d}/run", prm, threads=min(32, os.cpu_count() or 4), until=until)
    return d


def _compare_run_dir(got, d, prm, N, bft=False):
    kind, what = prm.mode.split(":")[:2]
    P = prm.P
    for i in range(N):
        want = [int(x) for x in open(f"{d}/run/partition_infos/S{i}.pinfo").read().split()]
        assert list(map(int, got["pinfo"][i])) == want, f"pinfo sample {i}"
    cext = "hash" if kind == "hash" else "kmer"
    for p in range(P):
        for i in range(N):
            assert got["counts"][(i, p)] == open(f"{d}/run/counts/partition_{p}/S{i}.{cext}", "rb").read(), f"counts sample {i} partition {p}"
    for p in range(P):
        assert got["merge_info"][p] == open(f"{d}/run/merge_infos/partition{p}.merge_info", "rb").read(), f"merge_info {p}"
        if not bft:
            assert got["matrices"][p] == open(f"{d}/run/matrices/matrix_{p}.{EXT[(kind, what)]}", "rb").read(), f"matrix {p}"


def test_cfg2_full_size_samples_equal_reference_binary():
    """4 samples of cfg2's exact shape: 4 x 64 .hash files and 64 .cmbf matrices, byte for byte."""
    from kmtricks_b200 import engine
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref/bin/kmtricks not shipped")
    N, R, Lr, P = 4, 1_000_000, 150, 64
    texts = _device_texts(N, R, Lr, 5_000_000, 2e-3, 2e-3, 1234)
    prm = O.Params(k=31, P=P, mode="hash:bf:bin", hard_min=2, bloom_size=200_000_000)
    d = _reference_run(texts, prm)
    try:
        cfg = engine.Config(kmer_size=31, nb_partitions=P, mode="hash:bf:bin", hard_min=2, bloom_size=200_000_000)
        got = engine.run_pipeline_lanes(texts, cfg, lanes=4)
        _compare_run_dir(got, d, prm, N)
        assert sum(int(x.sum()) for x in got["pinfo"]) == N * R * (Lr - 31 + 1)
    finally:
        shutil.rmtree(d, ignore_errors=True)


TWINS = {
    # BASELINE configs[2..4] at a size the reference CPU run finishes in seconds
    "cfg3_kmer_count_P512": dict(k=31, P=512, mode="kmer:count:bin", hard_min=3, G=2_000_000, d=1e-4, e=1e-3),
    "cfg4_hash_bft_P512": dict(k=31, P=512, mode="hash:bft:bin", hard_min=2, soft_min=2, share_min=2, bloom_size=200_000_000,
                               G=2_000_000, d=1e-4, e=1e-3),
    "cfg5_k63_kmer_pa_rescue_P256": dict(k=63, P=256, mode="kmer:pa:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=1,
                                         G=2_000_000, d=1e-4, e=2e-4),
}


@pytest.mark.parametrize("name", list(TWINS))
def test_cfg345_twins_equal_reference_binary(name):
    from kmtricks_b200 import engine
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref/bin/kmtricks not shipped")
    c = dict(TWINS[name])
    N, R, Lr = 8, 250_000, 150
    texts = _device_texts(N, R, Lr, c.pop("G"), c.pop("d"), c.pop("e"), 4321)
    bft = c["mode"].startswith("hash:bft")
    ref_mode = "hash:bf:bin" if bft else c["mode"]            # the reference CLI cannot reach write_as_bft
    prm = O.Params(k=c["k"], P=c["P"], mode=ref_mode, hard_min=c["hard_min"], soft_min=c.get("soft_min", 1),
                   recurrence_min=c.get("recurrence_min", 1), share_min=c.get("share_min", 0), bloom_size=c.get("bloom_size", 10_000_000))
    d = _reference_run(texts, prm)
    try:
        cfg = engine.Config(kmer_size=c["k"], nb_partitions=c["P"], mode=c["mode"], hard_min=c["hard_min"], soft_min=c.get("soft_min", 1),
                            recurrence_min=c.get("recurrence_min", 1), share_min=c.get("share_min", 0), bloom_size=c.get("bloom_size", 10_000_000))
        got = engine.run_pipeline_lanes(texts, cfg, lanes=4)
        _compare_run_dir(got, d, prm, N, bft=bft)
        if bft:
            W = O.window_bits(prm.bloom_size, prm.P)
            harness = os.path.join(ROOT, "oracle", "_ref", "bin", "bft_harness")
            for p in range(prm.P):
                files = [f"{d}/run/counts/partition_{p}/S{i}.hash" for i in range(N)]
                o = f"{d}/bft_{p}"
                subprocess.run([harness, "bft", o, str(W * p), str(W * (p + 1) - 1), str(prm.soft_min), str(prm.recurrence_min),
                                str(prm.share_min)] + files, check=True)
                assert got["matrices"][p] == open(o, "rb").read(), f"bft matrix {p}"
                os.unlink(o)
    finally:
        shutil.rmtree(d, ignore_errors=True)
