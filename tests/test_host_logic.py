"""Host-side logic of the product package (no GPU): format writers, FASTX parser, synth."""
import numpy as np

from kmtricks_b200 import engine, formats, synth
from oracle import oracle as O
from tests.conftest import EDGE_FASTA, EDGE_FASTQ_CRLF


def test_format_writers_match_oracle_encoders():
    rng = np.random.default_rng(2)
    n, N = 50, 11
    lo = np.sort(rng.integers(0, 2**62, n, dtype=np.uint64)); hi = rng.integers(0, 2**60, n, dtype=np.uint64)
    cnt = rng.integers(1, 1000, n, dtype=np.uint32)
    assert formats.kmer_file(lo, cnt, 31, 3, 2) == O.enc_kmer_file(lo, hi, cnt, 31, 3, 2)
    keys2 = np.stack([lo, hi], axis=1).reshape(-1)
    assert formats.kmer_file(keys2, cnt, 63, 1, 0) == O.enc_kmer_file(lo, hi, cnt, 63, 1, 0)
    big = np.sort(rng.integers(0, 2**40, 9000, dtype=np.uint64)); bc = rng.integers(1, 9, 9000, dtype=np.uint32)
    assert formats.hash_file(big, bc, 0, 1) == O.enc_hash_file(big, bc, 0, 1)
    counts = rng.integers(0, 3, (n, N), dtype=np.uint32)
    body = np.zeros(n, dtype=[("k", "<u8"), ("c", "<u4", (N,))]); body["k"] = lo; body["c"] = counts
    assert formats.matrix_header("count", "kmer", 31, N, 2) + body.tobytes() == O.enc_count_matrix(lo, hi, counts, 31, N)
    assert formats.matrix_header("count", "hash", 31, N, 2) + body.tobytes() == O.enc_count_hash_matrix(lo, counts, N, 2)
    pa = np.zeros(n, dtype=[("k", "<u8"), ("b", "u1", ((N + 7) // 8,))]); pa["k"] = lo; pa["b"] = O.pa_rows(counts)
    assert formats.matrix_header("pa", "kmer", 31, N, 2) + pa.tobytes() == O.enc_pa_matrix(lo, hi, counts, 31, N)
    assert formats.matrix_header("pa", "hash", 31, N, 2) + pa.tobytes() == O.enc_pa_hash_matrix(lo, counts, N, 2)
    assert formats.matrix_header("bf", "hash", 31, N, 3, 640) == O.cmbf_header(N, 640, 3)
    stats = rng.integers(0, 10**6, (6, N), dtype=np.uint64)
    assert formats.merge_info(stats) == O.enc_merge_info(stats)
    assert formats.hash_info(10**8, 4, 10) == O.enc_hash_info(10**8, 4, 10)
    t = formats.static_repart_table(8, 13)
    assert np.array_equal(t, O.repart_static(8, 13))
    assert formats.minim_repart(t, 13) == O.enc_minim_repart(t, 13)
    assert formats.read_minim_repart(formats.minim_repart(t, 13))[0] == 13
    for bloom, P in ((10**8, 4), (2 * 10**8, 64), (1000, 3), (64, 1)):
        assert formats.window_bits(bloom, P) == O.window_bits(bloom, P)


def test_host_fastx_parser_matches_oracle_parser():
    for buf in (EDGE_FASTA, EDGE_FASTQ_CRLF, b"", b">x\n", b">a\nAC\n>b\n\n>c\nGT", b"@q\nACGT\n+\nII\nII\n@r\nAA\n+\nII\n",
                synth.make_fastq(1, 0, 20, L=40, G=500)):
        assert engine.parse_fastx(buf) == O.fastx_parse(buf)


def test_synth_is_deterministic_and_well_formed():
    a = synth.make_fastq(9, 2, 100, L=80, G=5000, revcomp=True)
    b = synth.make_fastq(9, 2, 60, L=80, G=5000, revcomp=True) + synth.make_fastq(9, 2, 40, L=80, G=5000, revcomp=True, first_read=60)
    assert a == b
    lines = a.split(b"\n")
    assert len(lines) == 401 and lines[0] == b"@r00000000" and lines[2] == b"+" and set(lines[1]) <= set(b"ACGT")
    assert len(a) == 100 * synth.record_bytes(80)


def test_cli_fof_parser_kat_from_fof_test(tmp_path):
    """tests/io/fof_test.cpp: the reference's input-list fixture (tests/data/fof.txt, grammar `ID : f1 ; f2 ! hard-min`,
    include/kmtricks/io/fof.hpp:39-40,115-147) through the C++ host's parser (`kmx fof <file>`, no device needed);
    duplicate identifiers and lines without ':' are errors as in the reference."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    kmx = os.path.join(root, "kmtricks_b200", "bin", "kmx")
    if not os.path.exists(kmx):
        subprocess.run(["bash", os.path.join(root, "build.sh")], check=True)
    f = tmp_path / "fof.txt"
    f.write_text("D1 : /path/to/D1.fasta ; /path/to/D1.fasta ! 20\nD2 : /path/to/D2.fasta ; /path/to/D2.fasta ! 20\n")
    out = subprocess.run([kmx, "fof", str(f)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out == ["D1\t20\t/path/to/D1.fasta;/path/to/D1.fasta", "D2\t20\t/path/to/D2.fasta;/path/to/D2.fasta"]
    f.write_text("A: x.fa\n\n  B :y.fq.gz;z.fq   \nC : w.fa ! 3\n")
    out = subprocess.run([kmx, "fof", str(f)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out == ["A\t0\tx.fa", "B\t0\ty.fq.gz;z.fq", "C\t3\tw.fa"]
    for bad in ("A : x.fa\nA : y.fa\n", "no colon here\n", "A : \n"):
        f.write_text(bad)
        r = subprocess.run([kmx, "fof", str(f)], capture_output=True, text=True)
        assert r.returncode != 0 and r.stderr.strip()


def test_cli_streams_fastq_in_blocks_of_whole_records(tmp_path):
    """The host's streaming reader (`kmx blocks <file> <block bytes>`, no device needed): every block but the last ends on
    a record boundary (lines a multiple of 4, first byte '@'), the blocks concatenate to the file -- plain and gzipped, LF and
    CRLF, with '@' as the first quality character, a last line without a newline; a record larger than a block is an error."""
    import gzip
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    kmx = os.path.join(root, "kmtricks_b200", "bin", "kmx")
    if not os.path.exists(kmx):
        subprocess.run(["bash", os.path.join(root, "build.sh")], check=True)

    def fnv(b):
        h = 1469598103934665603
        for x in b:
            h = ((h ^ x) * 1099511628211) & (2**64 - 1)
        return h
    rng = np.random.default_rng(4)
    recs = []
    for i in range(400):
        L = int(rng.integers(20, 200))
        seq = bytes(rng.choice(list(b"ACGTN"), L).tolist())
        recs.append(b"@r%d some comment\n" % i + seq + b"\n+\n" + b"@" * L + b"\n")       # '@' opens the quality line
    text = b"".join(recs)
    cases = {"plain.fastq": text, "crlf.fastq": text.replace(b"\n", b"\r\n"), "nonl.fastq": text[:-1]}
    for name, data in cases.items():
        p = tmp_path / name
        p.write_bytes(data)
        gz = tmp_path / (name + ".gz")
        with gzip.open(gz, "wb") as g:
            g.write(data)
        for path in (p, gz):
            for block in (700, 4096, 10**7):
                out = subprocess.run([kmx, "blocks", str(path), str(block)], capture_output=True, text=True, check=True).stdout.split("\n")
                rows = [l.split("\t") for l in out if l and not l.startswith("total")]
                tot = [l.split("\t") for l in out if l.startswith("total")][0]
                assert int(tot[1]) == len(data) and int(tot[2]) == fnv(data)
                assert sum(int(r[0]) for r in rows) == len(data)
                for r in rows[:-1]:
                    assert int(r[0]) <= block and int(r[1]) % 4 == 0 and int(r[1]) > 0 and r[2] == "1"
                assert rows[-1][2] == "1"
                if block >= len(data):
                    assert len(rows) == 1
    r = subprocess.run([kmx, "blocks", str(tmp_path / "plain.fastq"), "100"], capture_output=True, text=True)
    assert r.returncode != 0 and "block" in r.stderr


def test_hash_mod_halves():
    """The hand-scheduled 32-bit formulation of XXH64(8 bytes, seed 0) and of the Barrett modulo that stage 2's pass A runs
    (hb_hash_mod, csrc/s2_bin.cu) -- restated here step for step on Python integers -- equals xxHash's XXH64 (the `xxhash`
    package when present, else the oracle's C restatement) followed by `% d`, for the window sizes in use and for edge values."""
    import random
    M32, M64 = (1 << 32) - 1, (1 << 64) - 1
    P1, P2, P3, P4, P5 = 0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5
    try:
        import xxhash
        ref = lambda w: xxhash.xxh64(int(w).to_bytes(8, "little"), seed=0).intdigest()
    except ImportError:
        ref = lambda w: int(O.xxh64_u64(np.array([w], dtype=np.uint64))[0])

    def fsl(lo, hi, s): return ((((hi << 32) | lo) << s) >> 32) & M32
    def fsr(lo, hi, s): return (((hi << 32) | lo) >> s) & M32

    def mul64(l, h, C):
        t = l * (C & M32)
        return t & M32, ((t >> 32) + l * (C >> 32) + h * (C & M32)) & M32

    def fast(w, d):
        l, h = mul64(w & M32, w >> 32, P2)
        l, h = fsl(h, l, 31), fsl(l, h, 31)
        l, h = mul64(l, h, P1)
        h0 = (P5 + 8) & M64
        l ^= h0 & M32; h ^= h0 >> 32
        l, h = fsl(h, l, 27), fsl(l, h, 27)
        t = (l * (P1 & M32) + P4) & M64
        l, h = t & M32, ((t >> 32) + l * (P1 >> 32) + h * (P1 & M32)) & M32
        l ^= h >> 1
        l, h = mul64(l, h, P2)
        l, h = l ^ fsr(l, h, 29), h ^ (h >> 29)
        l, h = mul64(l, h, P3)
        l ^= h
        assert ((h << 32) | l) == ref(w)
        m64 = M64 // d; ml, mh = m64 & M32, m64 >> 32
        s = l * mh
        t2 = h * ml + (s & M32)
        q = (h * mh + (s >> 32) + (t2 >> 32)) & M32
        r = (l - q * d) & M32
        r = min(r, (r - d) & M32)
        return min(r, (r - d) & M32)
    rnd = random.Random(7)
    for d in (3125056, 781312, 50048, 64, 2**30 - 64, 999983):
        for w in [0, 1, M64, (1 << 62) - 1, d, d - 1] + [rnd.getrandbits(rnd.choice([10, 40, 62, 64])) for _ in range(3000)]:
            assert fast(w, d) == ref(w) % d
