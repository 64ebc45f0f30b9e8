"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs,
file for file, byte for byte (integer / bit work => bit-exact is the bar)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    "hash_bf": dict(k=31, P=8, mode="hash:bf:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=1, bloom_size=400_000),
    "hash_bf_default": dict(k=31, P=4, mode="hash:bf:bin", hard_min=2, bloom_size=300_000),
    "hash_bft": dict(k=31, P=4, mode="hash:bft:bin", hard_min=2, soft_min=2, share_min=1, recurrence_min=2, bloom_size=200_000),
    "hash_count": dict(k=31, P=8, mode="hash:count:bin", hard_min=2, bloom_size=400_000),
    "hash_pa": dict(k=31, P=8, mode="hash:pa:bin", hard_min=2, soft_min=3, share_min=2, bloom_size=400_000),
    "kmer_count": dict(k=31, P=8, mode="kmer:count:bin", hard_min=2),
    "kmer_pa_rescue": dict(k=31, P=8, mode="kmer:pa:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=2),
    "k63_kmer_pa": dict(k=63, P=4, mode="kmer:pa:bin", hard_min=1, soft_min=2, share_min=2, recurrence_min=2),
    "k63_kmer_count": dict(k=63, P=4, mode="kmer:count:bin", hard_min=2),
    "k63_hash_bf": dict(k=63, P=4, mode="hash:bf:bin", hard_min=2, bloom_size=300_000),
    "k21_m8": dict(k=21, m=8, P=5, mode="kmer:count:bin", hard_min=1),
    # hash keys at other k: the binned pass A rolls on 32-bit halves for 28 <= k <= 32 and on 64-bit words (128-bit tail) below
    "k21_m8_hash_count": dict(k=21, m=8, P=5, mode="hash:count:bin", hard_min=1, bloom_size=300_000),
    "k27_hash_bf": dict(k=27, P=4, mode="hash:bf:bin", hard_min=2, bloom_size=300_000),
    "k28_hash_count": dict(k=28, P=4, mode="hash:count:bin", hard_min=2, bloom_size=300_000),
    "k32_hash_count": dict(k=32, P=4, mode="hash:count:bin", hard_min=1, bloom_size=300_000),
    "k40_m11": dict(k=40, m=11, P=6, mode="kmer:count:bin", hard_min=2),
    "k32": dict(k=32, P=4, mode="kmer:count:bin", hard_min=1),
    "k33": dict(k=33, P=4, mode="kmer:count:bin", hard_min=1),
}


def _run_both(samples, case, sample_hard_min=None):
    from kmtricks_b200 import engine
    from oracle import oracle as O
    c = dict(case)
    cfg = engine.Config(kmer_size=c["k"], minim_size=c.get("m", 10), nb_partitions=c["P"], mode=c["mode"],
                        hard_min=c["hard_min"], soft_min=c.get("soft_min", 1), recurrence_min=c.get("recurrence_min", 1),
                        share_min=c.get("share_min", 0), bloom_size=c.get("bloom_size", 10_000_000))
    prm = O.Params(k=c["k"], m=c.get("m", 10), P=c["P"], mode=c["mode"], hard_min=c["hard_min"],
                   soft_min=c.get("soft_min", 1), recurrence_min=c.get("recurrence_min", 1),
                   share_min=c.get("share_min", 0), bloom_size=c.get("bloom_size", 10_000_000),
                   sample_hard_min=sample_hard_min or {})
    got = engine.run_pipeline(samples, cfg, sample_hard_min)
    want = O.run_pipeline(samples, prm)
    return got, want


def _compare(got, want, P, N):
    for s in range(N):
        assert list(map(int, got["pinfo"][s])) == list(map(int, want["pinfo"][s])), f"pinfo sample {s}"
    for s in range(N):
        for p in range(P):
            assert got["counts"][(s, p)] == want["counts"][(s, p)], f"counts file sample {s} partition {p}"
    for p in range(P):
        assert got["merge_info"][p] == want["merge_info"][p], f"merge_info partition {p}"
        assert got["matrices"][p] == want["matrices"][p], f"matrix partition {p}"
    assert got["launches"] > 0


@pytest.mark.parametrize("name", list(CASES))
def test_pipeline_parity_synth(name, synth_samples):
    got, want = _run_both(synth_samples, CASES[name])
    _compare(got, want, CASES[name]["P"], len(synth_samples))


@pytest.mark.parametrize("name", ["kmer_count", "hash_bf", "k63_kmer_pa"])
def test_pipeline_parity_edge_cases(name, edge_samples):
    """single N, NN + IUPAC, lowercase, read < k, read == k, poly-A (all m-mers banned), long A
    run, multi-line FASTA, CRLF FASTQ, multi-file sample (SURVEY §9.2 edge list)."""
    got, want = _run_both(edge_samples, CASES[name])
    _compare(got, want, CASES[name]["P"], len(edge_samples))


def test_per_sample_hard_min_override(synth_samples):
    got, want = _run_both(synth_samples, CASES["kmer_count"], {1: 1, 3: 4})
    _compare(got, want, CASES["kmer_count"]["P"], len(synth_samples))


@pytest.mark.parametrize("name,case,N", [
    ("cfg3_shape", dict(k=31, P=512, mode="kmer:count:bin", hard_min=1, soft_min=2, recurrence_min=2), 130),
    ("cfg5_shape", dict(k=63, P=256, mode="kmer:pa:bin", hard_min=1, soft_min=3, share_min=2, recurrence_min=1), 70),
    ("cfg4_shape", dict(k=31, P=512, mode="hash:bft:bin", hard_min=1, soft_min=2, share_min=3, bloom_size=100_000), 130),
])
def test_many_samples_many_partitions(name, case, N):
    """The shapes of BASELINE configs 3-5 (hundreds of samples, P = 256/512, count / pa+rescue / bft)
    at a size the oracle finishes in seconds: N not a multiple of 8, most (sample, partition) lists tiny."""
    from kmtricks_b200 import synth
    samples = [[synth.make_fastq(3, s, 120, L=150, G=6000, d=1e-2, e=5e-3, revcomp=True)] for s in range(N)]
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], N)


def test_hash_keys_take_the_binned_path(synth_samples):
    """Hash keys with k <= 32 are counted by the binned shared-memory kernels (s2_bin.cu), one pass per sample; k > 32 keeps
    the L2-histogram path."""
    got, want = _run_both(synth_samples, CASES["hash_count"])
    _compare(got, want, CASES["hash_count"]["P"], len(synth_samples))
    assert got["hash_binned"] == len(synth_samples)
    got, _ = _run_both(synth_samples, CASES["k63_hash_bf"])
    assert got["hash_binned"] == 0


@pytest.mark.parametrize("flag,name", [
    ("KMX_HASH_NOBIN", "hash_bf"),         # L2-histogram path (fallback of the binned counting; what k > 32 uses)
    ("KMX_HASH_NOBIN", "hash_count"),
    ("KMX_HIST32", "hash_bf"),             # 32-bit counters from the start (binned: 16 K-slot bins)
    ("KMX_HIST_NOROLL", "hash_bf"),        # k <= 32 histogram fill by the search-and-extract kernel (the k > 32 / fallback kernel)
    ("KMX_HIST_FUSE", "hash_bf"),          # opt-in fused fill + compact persistent kernel (done counters, fences)
    ("KMX_HIST_FUSE", "hash_count"),
    ("KMX_NO_HT", "kmer_count"),           # generic path: expand -> segmented radix sort -> run-length (fallback of the hash-count)
    ("KMX_NO_HT", "k63_kmer_pa"),
    ("KMX_S1V5_OFF", "kmer_count"),        # stage 1 by the streaming thread-per-read kernel (what long sequences use)
    ("KMX_S1V5_OFF", "k63_kmer_pa"),
    ("KMX_S1V5_OFF", "hash_bf"),
])
def test_fallback_paths_stay_exact(flag, name, synth_samples, monkeypatch):
    """The alternative stage-2 paths behind the environment switches give the same bytes as the default ones."""
    monkeypatch.setenv(flag, "1")
    got, want = _run_both(synth_samples, CASES[name])
    _compare(got, want, CASES[name]["P"], len(synth_samples))


def test_stage1_streaming_kernel_edge_cases(edge_samples, monkeypatch):
    """The edge-case inputs through the streaming stage-1 kernel as well (the default for short reads is the
    position-parallel one)."""
    monkeypatch.setenv("KMX_S1V5_OFF", "1")
    got, want = _run_both(edge_samples, CASES["kmer_count"])
    _compare(got, want, CASES["kmer_count"]["P"], len(edge_samples))


@pytest.mark.parametrize("R", [16, 64, 128])
def test_stage1_position_parallel_reads_per_cta(R, synth_samples, edge_samples, monkeypatch):
    """Other CTA geometries of the position-parallel stage-1 kernel (reads per CTA)."""
    monkeypatch.setenv("KMX_S1V5_R", str(R))
    for samples in (synth_samples, edge_samples):
        got, want = _run_both(samples, CASES["kmer_count"])
        _compare(got, want, CASES["kmer_count"]["P"], len(samples))


def test_long_sequences_take_the_streaming_kernel():
    """Contigs / long reads (FASTA, multi-line, with N runs; FASTQ reads of 400-900 nt): longer than the shared arrays of
    the position-parallel kernel, so stage 1 streams them (sequences beyond 1054 nt are cut into overlapping segments
    by the host side of kmx_superk_push_reads)."""
    import random
    rnd = random.Random(11)

    def seq(n):
        s = [rnd.choice("ACGT") for _ in range(n)]
        for _ in range(n // 3000):
            p = rnd.randrange(n - 40); s[p:p + rnd.randint(1, 35)] = "N" * rnd.randint(1, 35)
        return "".join(s)

    contigs = [seq(n) for n in (12345, 5000, 1054, 1055, 2048, 2049, 31, 30, 3000)]
    fasta = "".join(">c%d\n%s\n" % (i, "\n".join(c[j:j + 70] for j in range(0, len(c), 70))) for i, c in enumerate(contigs)).encode()
    long_reads = [seq(rnd.randint(400, 900)).encode() for _ in range(60)]
    fastq = b"".join(b"@l%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(long_reads))
    samples = [[fasta], [fastq, fasta[:20000] + b"\n"], [fastq]]
    for name in ("kmer_count", "k63_kmer_pa", "hash_bf"):
        got, want = _run_both(samples, CASES[name])
        _compare(got, want, CASES[name]["P"], len(samples))


def _fq(reads, crlf=False, trailing_newline=True, comment=b""):
    nl = b"\r\n" if crlf else b"\n"
    t = b"".join(b"@q%d" % i + comment + nl + r + nl + b"+" + nl + b"I" * len(r) + nl for i, r in enumerate(reads))
    return t if trailing_newline else t[:-len(nl)]


def test_stage1_self_indexing_launch(monkeypatch):
    """After the first FASTQ block of a run, stage 1 finds its reads in the newline masks itself (no line-index
    pass).  Blocks with CRLF line ends, N / lower case, reads shorter than k, no final newline, many small
    records (several CTAs per tile) and few long ones (several tiles per CTA) must take that path and give the
    oracle's bytes; a block with a longer read than seen before falls back to the indexed path; the switch
    KMX_S1_NOFUSE gives the same bytes."""
    import random
    rnd = random.Random(5)
    genome = "".join(rnd.choice("ACGT") for _ in range(30000))

    def reads(n, lo, hi, noise=0.0):
        out = []
        for _ in range(n):
            L = rnd.randint(lo, hi); p = rnd.randrange(len(genome) - L)
            r = list(genome[p:p + L])
            for j in range(L):
                if rnd.random() < noise:
                    r[j] = rnd.choice("NnacgtRY")
            out.append("".join(r).encode())
        return out

    samples = [
        [_fq(reads(700, 250, 250))],                                   # sets the length hint (indexed path)
        [_fq(reads(900, 20, 250, noise=0.01), crlf=True)],             # self-indexed: CRLF, invalid letters, reads < k
        [_fq(reads(5000, 31, 40)), _fq(reads(300, 200, 250), trailing_newline=False)],
        [_fq(reads(64, 250, 250) + reads(1, 300, 300) + reads(64, 100, 250))],   # longer read: falls back, raises the hint
        [_fq(reads(400, 280, 300, noise=0.002), trailing_newline=False)],
        [_fq(reads(40, 310, 310))],                                     # longer again: indexed, hint 310 (about the longest the kernel's shared arrays take)
        [_fq(reads(333, 290, 310, noise=0.001), crlf=True, comment=b" " + b"x" * 90)],   # 32 records > 20 KiB: two windows of mask words per CTA
    ]
    case = dict(k=31, P=8, mode="kmer:count:bin", hard_min=1)
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], len(samples))
    assert got["s1_self_indexed"] == 5 and got["s1_indexed"] == 3, (got["s1_self_indexed"], got["s1_indexed"])
    monkeypatch.setenv("KMX_S1_NOFUSE", "1")
    got2, _ = _run_both(samples, case)
    assert got2["s1_self_indexed"] == 0
    assert got2["matrices"] == got["matrices"] and got2["counts"] == got["counts"]


def test_stage1_self_indexing_rejects_malformed_fastq():
    """A block that is not strict 4-line FASTQ gets the format status on the self-indexing path too (the host
    then parses it kseq-style), and the sample's buckets are left as they were."""
    from kmtricks_b200 import engine, _lib
    good = _fq([b"ACGTTGCA" * 10] * 50)
    bad = good.replace(b"\n+\n", b"\n-\n", 1)
    cfg = engine.Config(kmer_size=31, nb_partitions=4, mode="kmer:count:bin", hard_min=1)
    eng = engine.Engine(cfg, 2)
    try:
        L = eng.lib
        pin1 = eng.superk([good])
        assert L.kmx_superk_begin(eng.h) == 0
        assert L.kmx_superk_push_fastq(eng.h, bad, len(bad), 0) == _lib.KMX_ERR_FORMAT
        assert L.kmx_superk_push_fastq(eng.h, good, len(good), 0) == 0
        pin2 = np.zeros(4, dtype=np.uint64)
        import ctypes as C
        assert L.kmx_superk_end(eng.h, pin2.ctypes.data_as(C.POINTER(C.c_uint64))) == 0
        assert list(pin1) == list(pin2)
        assert int(L.kmx_stat(eng.h, 0)) == 1
    finally:
        eng.close()


def test_hash_mode_larger_sample_all_paths(monkeypatch):
    """One sample large enough that every partition spans several histogram tiles and sweep chunks
    (tile tickets wrap over windows, partial last chunk), default and fused kernels."""
    from kmtricks_b200 import synth
    samples = [[synth.make_fastq(7, s, 30000, L=150, G=200000, d=3e-3, e=3e-3, revcomp=True)] for s in range(2)]
    case = dict(k=31, P=6, mode="hash:bf:bin", hard_min=2, bloom_size=3_000_000)
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], len(samples))
    assert got["hash_binned"] == len(samples)
    monkeypatch.setenv("KMX_HASH_NOBIN", "1")
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], len(samples))
    assert got["hash_binned"] == 0
    monkeypatch.setenv("KMX_HIST_FUSE", "1")
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], len(samples))


def test_hash_mode_many_bins_per_window(monkeypatch):
    """A window of 3 M slots counted with 32-bit counters = 184 bins of 16 K slots per partition: the pass-A variant with the
    larger per-bin arrays (more than 128 bins per window); with 16-bit counters the same window has 92 bins."""
    from kmtricks_b200 import synth
    samples = [[synth.make_fastq(17, s, 20000, L=150, G=150000, d=3e-3, e=3e-3, revcomp=True)] for s in range(2)]
    case = dict(k=31, P=4, mode="hash:count:bin", hard_min=2, bloom_size=12_000_000)
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], len(samples))
    assert got["hash_binned"] == len(samples)
    monkeypatch.setenv("KMX_HIST32", "1")
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], len(samples))
    assert got["hash_binned"] == len(samples)


def test_hash_counter_wrap_falls_back_to_32_bit(monkeypatch):
    """The hash histogram starts with 16-bit counters; a k-mer seen more than 65535 times in one sample
    (600 poly-A reads = 72000 x A^31) must be detected by the window checksum and the sample redone with
    32-bit counters -- same bytes as the oracle, and as a run forced to 32-bit counters from the start."""
    from kmtricks_b200 import synth
    base = synth.make_fastq(11, 0, 300, L=150, G=5000, d=3e-3, e=3e-3, revcomp=True)
    polya = b"".join(b"@a%d\n%s\n+\n%s\n" % (i, b"A" * 150, b"I" * 150) for i in range(600))
    samples = [[base + polya], [synth.make_fastq(11, 1, 300, L=150, G=5000, d=3e-3, e=3e-3, revcomp=True)]]
    case = dict(k=31, P=4, mode="hash:count:bin", hard_min=1, bloom_size=200_000)
    got, want = _run_both(samples, case)
    _compare(got, want, case["P"], len(samples))
    monkeypatch.setenv("KMX_HIST32", "1")
    got32, _ = _run_both(samples, case)
    assert got32["matrices"] == got["matrices"]
