"""GPU parity against (1) the committed golden fixtures = outputs of the unmodified reference
binary, and (2) the reference binary itself run on the GPU box (oracle/_ref travels with the
snapshot) on a larger seeded input, plus size-independent properties at a bench-like size."""
import hashlib
import os
import shutil
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.test_oracle_golden import CASES, load_case   # noqa: E402


@pytest.mark.parametrize("name", CASES)
def test_cuda_path_equals_reference_golden(name):
    from kmtricks_b200 import engine
    z, prm, samples = load_case(name)
    cfg = engine.Config(kmer_size=prm.k, minim_size=prm.m, nb_partitions=prm.P, mode=prm.mode, hard_min=prm.hard_min,
                        soft_min=prm.soft_min, recurrence_min=prm.recurrence_min, share_min=prm.share_min, bloom_size=prm.bloom_size)
    got = engine.run_pipeline(samples, cfg, prm.sample_hard_min)
    for i in range(len(samples)):
        assert list(map(int, got["pinfo"][i])) == list(map(int, z[f"pinfo_{i}"]))
        for p in range(prm.P):
            assert got["counts"][(i, p)] == z[f"counts_{i}_{p}"].tobytes(), f"counts {i} {p}"
    for p in range(prm.P):
        assert got["matrices"][p] == z[f"matrix_{p}"].tobytes(), f"matrix {p}"
        assert got["merge_info"][p] == z[f"merge_info_{p}"].tobytes(), f"merge_info {p}"
    if "bft_0" in z.files:
        cfg.mode = "hash:bft:bin"
        got = engine.run_pipeline(samples, cfg, prm.sample_hard_min)
        for p in range(prm.P):
            assert got["matrices"][p] == z[f"bft_{p}"].tobytes(), f"bft {p}"


@pytest.mark.parametrize("mode,extra", [("kmer:count:bin", {}), ("hash:bf:bin", dict(bloom_size=3_000_000)),
                                        ("kmer:pa:bin", dict(soft_min=3, share_min=2, recurrence_min=2, k=63))])
def test_cuda_path_equals_reference_binary_on_the_box(mode, extra):
    """6 samples x 40k reads, both arms run here: kmtricks pipeline (CPU) vs the CUDA path."""
    from kmtricks_b200 import engine, synth
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref/bin/kmtricks not shipped")
    k = extra.pop("k", 31)
    samples = [synth.make_fastq(21, s, 40_000, L=150, G=200_000, d=3e-3, e=3e-3, revcomp=True) for s in range(6)]
    prm = O.Params(k=k, P=16, mode=mode, hard_min=2, **extra)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="kmx_t_", dir=base)
    try:
        with open(f"{d}/fof.txt", "w") as f:
            for i, b in enumerate(samples):
                open(f"{d}/S{i}.fastq", "wb").write(b)
                f.write(f"S{i}: {d}/S{i}.fastq\n")
        O.run_reference(f"{d}/fof.txt", f"{d}/run", prm, threads=8)
        cfg = engine.Config(kmer_size=k, nb_partitions=16, mode=mode, hard_min=2, **extra)
        got = engine.run_pipeline([[b] for b in samples], cfg)
        kind, what = mode.split(":")[:2]
        ext = {("kmer", "count"): "count", ("kmer", "pa"): "pa", ("hash", "bf"): "cmbf"}[(kind, what)]
        for p in range(16):
            assert got["matrices"][p] == open(f"{d}/run/matrices/matrix_{p}.{ext}", "rb").read(), f"matrix {p}"
            assert got["merge_info"][p] == open(f"{d}/run/merge_infos/partition{p}.merge_info", "rb").read()
            for i in range(6):
                cext = "hash" if kind == "hash" else "kmer"
                assert got["counts"][(i, p)] == open(f"{d}/run/counts/partition_{p}/S{i}.{cext}", "rb").read()
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_run_samples_lanes_equal_single_lane_and_properties():
    """Bench-shaped run (device-resident text, kmx_run_samples): results must not depend on the
    number of lanes; size-independent properties: lists ascending and distinct, counts >= hard-min,
    sum of .pinfo == number of valid k-mers, Bloom bits == union of the samples' hash lists."""
    import ctypes as C
    from kmtricks_b200 import _lib, engine, synth
    N, R, Lr, P = 6, 100_000, 150, 32
    cfg = engine.Config(kmer_size=31, nb_partitions=P, mode="hash:bf:bin", hard_min=2, bloom_size=20_000_000)
    digests = []
    for lanes in (1, 3):
        eng = engine.Engine(cfg, N); L = eng.lib; h = eng.h
        sb = R * synth.record_bytes(Lr)
        d = C.c_void_p(); assert L.kmx_dev_alloc(h, N * sb + 64, C.byref(d)) == 0
        for s in range(N):
            assert L.kmx_synth_fastq(h, 77, s, 0, R, Lr, 1_000_000, 2e-3, 2e-3, 1, d.value + s * sb) == 0
        ptrs = (C.c_void_p * N)(*[d.value + s * sb for s in range(N)])
        sizes = (C.c_size_t * N)(*([sb] * N)); hm = (C.c_uint32 * N)(*([2] * N))
        pin = np.zeros((N, P), dtype=np.uint64)
        rc = L.kmx_run_samples(h, N, ptrs, sizes, 1, None, hm, lanes, pin.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert rc == 0, L.kmx_last_error(h)
        assert (pin.sum(axis=1) == R * (Lr - 31 + 1)).all()           # no N in the synthetic reads
        hsh = hashlib.sha256()
        W = cfg.window_bits
        for p in range(P):
            union = np.zeros(W, dtype=bool)
            for s in range(N):
                keys, cnt = eng.counts(s, p)
                assert (np.diff(keys.astype(np.int64)) > 0).all() and (cnt >= 2).all()
                assert keys.min(initial=W * p) >= W * p and keys.max(initial=W * p) < W * (p + 1)
                union[(keys - np.uint64(W * p)).astype(np.int64)] = True
                hsh.update(keys.tobytes()); hsh.update(cnt.tobytes())
            m = eng.merge(p)
            rows = np.unpackbits(m["body"].reshape(W, -1), axis=1, bitorder="little")[:, :N]
            assert np.array_equal(rows.any(axis=1), union)
            hsh.update(m["body"].tobytes())
        digests.append(hsh.hexdigest())
        eng.close()
    assert digests[0] == digests[1]


def test_full_size_sample_conserves_every_kmer(monkeypatch):
    """One sample launch of BASELINE cfg2's exact shape (1M reads x 150 nt, k=31, P=64, Bloom 2e8) with hard-min 1:
    the counts of all 64 hash lists must add up to the number of k-mers (nothing lost or double counted anywhere
    between the FASTQ text and the lists), lists ascending and inside their window; the 16-bit-counter histogram
    and the 32-bit one must give the same bytes."""
    import ctypes as C
    from kmtricks_b200 import engine, synth
    R, Lr, P = 1_000_000, 150, 64
    cfg = engine.Config(kmer_size=31, nb_partitions=P, mode="hash:bf:bin", hard_min=1, bloom_size=200_000_000)
    digests = []
    for force32 in (False, True):
        if force32:
            monkeypatch.setenv("KMX_HIST32", "1")
        eng = engine.Engine(cfg, 1); L = eng.lib; h = eng.h
        sb = R * synth.record_bytes(Lr)
        d = C.c_void_p(); assert L.kmx_dev_alloc(h, sb + 64, C.byref(d)) == 0
        assert L.kmx_synth_fastq(h, 1234, 0, 0, R, Lr, 5_000_000, 2e-3, 2e-3, 1, d.value) == 0
        ptrs = (C.c_void_p * 1)(d.value); sizes = (C.c_size_t * 1)(sb); hm = (C.c_uint32 * 1)(1)
        pin = np.zeros((1, P), dtype=np.uint64)
        rc = L.kmx_run_samples(h, 1, ptrs, sizes, 1, None, hm, 1, pin.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert rc == 0, L.kmx_last_error(h)
        assert int(pin.sum()) == R * (Lr - 31 + 1)
        W = cfg.window_bits
        hsh = hashlib.sha256()
        for p in range(P):
            keys, cnt = eng.counts(0, p)
            assert int(cnt.astype(np.uint64).sum()) == int(pin[0, p]), f"partition {p}: counts do not add up to its k-mers"
            assert (np.diff(keys.astype(np.int64)) > 0).all()
            assert keys.min(initial=W * p) >= W * p and keys.max(initial=W * p) < W * (p + 1)
            hsh.update(keys.tobytes()); hsh.update(cnt.tobytes())
        digests.append(hsh.hexdigest())
        eng.close()
    assert digests[0] == digests[1]
