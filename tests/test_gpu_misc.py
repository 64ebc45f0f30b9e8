"""GPU tests of the stand-alone C-ABI entry points (transpose, synth twin, vector, interop)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(mode="hash:bf:bin", P=4, N=2, **kw):
    from kmtricks_b200 import engine
    return engine.Engine(engine.Config(kmer_size=31, nb_partitions=P, mode=mode, **kw), N)


@pytest.mark.parametrize("nrows,ncols", [(8, 8), (64, 64), (1024, 16), (25024, 8), (640, 104), (4096, 1000 // 8 * 8 + 8), (72, 40)])
def test_transpose_bits_matches_oracle_and_is_involution(nrows, ncols):
    """bit_matrix_test.cpp:60-99 only checks T(T(M)) == M; also check the definitional relation."""
    from oracle import oracle as O
    rng = np.random.default_rng(nrows * 131 + ncols)
    a = rng.integers(0, 256, size=nrows * ncols // 8, dtype=np.uint8)
    eng = _engine()
    t = eng.transpose_bits(a, nrows, ncols)
    assert np.array_equal(t, O.transpose_bits(a, nrows, ncols))
    assert np.array_equal(eng.transpose_bits(t, ncols, nrows), a)
    eng.close()


def test_synth_device_twin_equals_numpy():
    from kmtricks_b200 import synth
    eng = _engine()
    L = eng.lib
    for (R, Lr, G, rc, first) in [(1000, 150, 50000, 1, 0), (257, 60, 1000, 0, 12345)]:
        nb = R * synth.record_bytes(Lr)
        d = C.c_void_p()
        assert L.kmx_dev_alloc(eng.h, nb, C.byref(d)) == 0
        assert L.kmx_synth_fastq(eng.h, 1234, 3, first, R, Lr, G, 2e-3, 5e-3, rc, d) == 0
        out = np.empty(nb, dtype=np.uint8)
        assert L.kmx_memcpy_d2h(eng.h, out.ctypes.data, d, nb) == 0
        want = synth.make_fastq(1234, 3, R, L=Lr, G=G, d=2e-3, e=5e-3, revcomp=bool(rc), first_read=first)
        assert out.tobytes() == want
        L.kmx_dev_free(eng.h, d)
    eng.close()


def test_hash_vector_equals_bit_image_of_hash_list(synth_samples):
    """count --mode vector (count_processor.hpp:84-120) == bit image of the sample's .hash keys."""
    from oracle import oracle as O
    eng = _engine(P=4, N=1, bloom_size=200_000, hard_min=2)
    eng.superk(synth_samples[0]); eng.count(0)
    W = eng.cfg.window_bits
    for p in range(4):
        keys, cnt = eng.counts(0, p)
        assert eng.vector(0, p) == O.enc_vector(keys, W, p)
    eng.close()


def test_counts_put_interop_merges_foreign_lists(synth_samples):
    """Lists produced elsewhere (e.g. a reference counts/ dir) can be injected and merged."""
    from kmtricks_b200 import engine, formats
    from oracle import oracle as O
    prm = O.Params(k=31, P=4, mode="kmer:count:bin", hard_min=1, soft_min=2, share_min=2, recurrence_min=1)
    want = O.run_pipeline(synth_samples, prm)
    cfg = engine.Config(kmer_size=31, nb_partitions=4, mode="kmer:count:bin", hard_min=1, soft_min=2, share_min=2)
    eng = engine.Engine(cfg, len(synth_samples))
    for (s, p), (lo, hi, c) in want["lists"].items():
        eng.put_counts(s, p, lo, c)
    for p in range(4):
        m = eng.merge(p)
        assert eng.matrix_file(p, m) == want["matrices"][p]
        assert formats.merge_info(m["stats"]) == want["merge_info"][p]
    eng.close()


def test_emit_all_rows_for_plugins(synth_samples):
    """With a plugin loaded every merged row is surfaced with the default decision."""
    from kmtricks_b200 import engine
    from oracle import oracle as O
    prm = O.Params(k=31, P=4, mode="kmer:count:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=3)
    want = O.run_pipeline(synth_samples, prm)
    cfg = engine.Config(kmer_size=31, nb_partitions=4, mode="kmer:count:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=3)
    eng = engine.Engine(cfg, len(synth_samples))
    for s, b in enumerate(synth_samples):
        eng.superk(b); eng.count(s)
    N = len(synth_samples)
    for p in range(4):
        lists = [want["lists"][(s, p)] for s in range(N)]
        r = O.s3_merge(lists, 1, [3] * N, 3, 2, emit_all=True)
        m = eng.merge(p, emit_all=True)
        assert m["n_rows"] == len(r["lo"]) == m["n_union"]
        body = m["body"].reshape(m["n_rows"], m["row_bytes"])
        assert np.array_equal(body[:, :8].copy().view(np.uint64).ravel(), r["lo"])
        assert np.array_equal(body[:, 8:].copy().view(np.uint32).reshape(-1, N), r["counts"])
        assert np.array_equal(m["row_keep"], r["keep"])
    eng.close()


def test_empty_and_tiny_inputs():
    from kmtricks_b200 import engine
    from oracle import oracle as O
    samples = [[b""], [b"@r0\nACGT\n+\nIIII\n"], [b"@r0\n" + b"ACGTTGCAAGGCTTAACCGGTTAACGATCGATCGG" + b"\n+\n" + b"I" * 35 + b"\n"]]
    for mode, extra in (("kmer:count:bin", {}), ("hash:bf:bin", dict(bloom_size=1000))):
        cfg = engine.Config(kmer_size=31, nb_partitions=4, mode=mode, hard_min=1, **extra)
        prm = O.Params(k=31, P=4, mode=mode, hard_min=1, **extra)
        got = engine.run_pipeline(samples, cfg)
        want = O.run_pipeline(samples, prm)
        for p in range(4):
            assert got["matrices"][p] == want["matrices"][p]
            assert got["merge_info"][p] == want["merge_info"][p]


def test_bad_parameters_fail_loudly():
    from kmtricks_b200 import engine
    with pytest.raises(engine.KmxError):
        engine.Engine(engine.Config(kmer_size=64, nb_partitions=4), 1)
    with pytest.raises(engine.KmxError):
        engine.Engine(engine.Config(kmer_size=31, minim_size=13, nb_partitions=4), 1)
    eng = _engine(mode="kmer:count:bin")
    with pytest.raises(engine.KmxError):
        eng.merge(0, fmt="bf")          # bf rows need hash keys
    with pytest.raises(engine.KmxError):
        eng.count(0)                    # no sample finished
    eng.close()


def test_transpose_of_a_window_beyond_65535_row_blocks():
    """A Bloom window of more than 65535 x 1024 rows (large --bloom-size with few partitions): the row blocks go to gridDim.x,
    and so do the column blocks of the way back (BitMatrix::transpose is an involution, bitmatrix.hpp:209-214)."""
    from kmtricks_b200 import engine
    rng = np.random.default_rng(3)
    nrows, ncols = 68_000_000 // 8 * 8, 8
    a = rng.integers(0, 256, nrows * ncols // 8, dtype=np.uint8)
    eng = engine.Engine(engine.Config(kmer_size=31, nb_partitions=4, mode="kmer:count:bin"), 1)
    try:
        t = eng.transpose_bits(a, nrows, ncols)
        want = np.packbits(np.unpackbits(a, bitorder="little").reshape(nrows, ncols).T.copy(), bitorder="little")
        assert np.array_equal(t, want)
        back = eng.transpose_bits(t, ncols, nrows)
        assert np.array_equal(back, a)
    finally:
        eng.close()


def test_buffers_that_grow_from_sample_to_sample_over_several_lanes():
    """Samples of growing size over three lanes: every per-lane buffer (bucket slab, bin buffer, text staging, ...) is outgrown
    several times while other lanes are in flight.  Outgrown buffers are set aside, not freed on the spot (a cudaFree waits for
    the whole device; with NCCL kernels of other lanes in flight that wait closed a cycle over lanes and ranks at 8 GPUs); the
    results stay those of the oracle and the memory comes back at the next synchronisation point."""
    from kmtricks_b200 import engine, synth
    from oracle import oracle as O
    texts = [synth.make_fastq(3, s, 2000 * (s + 1), L=150, G=50000, d=4e-3, e=4e-3, revcomp=True) for s in range(7)]
    for mode, extra in (("hash:bf:bin", dict(bloom_size=400_000)), ("kmer:count:bin", {})):
        cfg = engine.Config(kmer_size=31, nb_partitions=8, mode=mode, hard_min=2, **extra)
        got = engine.run_pipeline_lanes(texts, cfg, lanes=3)
        want = O.run_pipeline([[t] for t in texts], O.Params(k=31, P=8, mode=mode, hard_min=2, **extra))
        for p in range(8):
            assert got["matrices"][p] == want["matrices"][p], f"{mode}: matrix {p}"
            assert got["merge_info"][p] == want["merge_info"][p]
