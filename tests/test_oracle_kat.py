"""Known-answer tests of the oracle against the reference's own test vectors and data fixtures
(copied as DATA into tests/golden/ref_fixture by make_golden.py)."""
import gzip
import os
import struct

import numpy as np
import pytest

from oracle import oracle as O

FX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fixture")


def enc(s: str) -> int:
    """k-mer string -> integer, A0 C1 T2 G3, first base most significant (kmer.hpp:165)."""
    v = 0
    for ch in s:
        v = (v << 2) | {"A": 0, "C": 1, "T": 2, "G": 3}[ch]
    return v


def test_xxh64_published_vectors():
    assert O.xxh64(b"") == 0xEF46DB3751D8E999
    assert O.xxh64(b"abc") == 0x44BC2CF5AD770999
    assert O.xxh64(b"Nobody inspects the spammish repetition") == 0xFBCEA83C8A378BF1


def test_xxh64_against_python_xxhash_all_lengths():
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(5)
    for n in list(range(0, 70)) + [100, 255, 1000]:
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        for seed in (0, 1, 0xDEADBEEF):
            assert O.xxh64(b, seed) == xxhash.xxh64_intdigest(b, seed)


def test_vectorised_hash_keys_equal_scalar():
    rng = np.random.default_rng(7)
    lo = rng.integers(0, 2**63, 2000, dtype=np.uint64); hi = rng.integers(0, 2**61, 2000, dtype=np.uint64)
    for w, W, p in ((1, 25_000_000, 3), (2, 250_048, 2), (1, 64, 0)):
        got = O.hash_keys(lo, hi, w, W, p)
        want = [O.lib().orc_hash_key(int(a), int(b), w, W, p) for a, b in zip(lo, hi)]
        assert got.tolist() == want


def test_minimizer_kat_from_kmer_test():
    """tests/kmer_test.cpp:117-154: ACGAGCAATACGA, m=4 -> minimizer AATA."""
    lut = O.minim_lut(4)
    s = "ACGAGCAATACGA"
    import ctypes as C
    v = O.lib().orc_minimizer_of(enc(s), 0, len(s), 4, lut.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert v == enc("AATA")


def test_minimizer_lut_rules():
    lut = O.minim_lut(10)
    mask = 4**10 - 1
    assert lut[enc("AAAAAAAAAA")] == mask                      # AA inside -> banned
    assert lut[enc("AACGTCGTCG")] == min(enc("AACGTCGTCG"), enc("CGACGACGTT"))   # AA only as prefix is allowed
    assert lut[enc("CGAACGTTCG")] == mask                      # reverse-palindrome with AA inside
    x = enc("ACGTCGTCGT"); assert lut[x] == min(x, enc("ACGACGACGT"))
    # strand invariance
    rng = np.random.default_rng(3)
    for v in rng.integers(0, 4**10, 500):
        rc = 0; t = int(v)
        for _ in range(10):
            rc = (rc << 2) | ((t & 3) ^ 2); t >>= 2
        assert lut[v] == lut[rc]


def test_static_repart_table_is_xxh64_mod_p():
    xxhash = pytest.importorskip("xxhash")
    t = O.repart_static(6, 7)
    for x in range(0, 4**6, 37):
        assert t[x] == xxhash.xxh64_intdigest(struct.pack("<I", x), 0) % 7


def _fixture_table():
    P, table = O.dec_minim_repart(gzip.decompress(open(os.path.join(FX, "repartition.minimRepart.gz"), "rb").read()))
    assert P == 4 and len(table) == 4**10
    return table


def test_canonical_kat_from_kmer_test():
    """tests/kmer_test.cpp:70-82: canonical(AAAAAAACCCCCCC) is itself, canonical(CGCCCCCCCCCCCT) = AGGGGGGGGGGGCG
    (k-mers compare as integers with A < C < T < G, first base most significant)."""
    t = O.repart_static(10, 4)

    def dec(v, k):
        return "".join("ACTG"[(int(v) >> (2 * (k - 1 - i))) & 3] for i in range(k))

    for seq, want in (("AAAAAAACCCCCCC", "AAAAAAACCCCCCC"), ("CGCCCCCCCCCCCT", "AGGGGGGGGGGGCG")):
        part, lo, hi = O.s1_sequences([seq.encode()], 14, 10, t)
        assert dec(lo[0], 14) == want


def test_repartition_kat_from_repartition_test():
    """tests/repartition_test.cpp:7-18: four 31-mers whose minimizer (m = 10) falls in partitions 0, 1, 2, 3 of the
    fixture table tests/data/repart_gatb/repartition.minimRepart."""
    table = _fixture_table()
    kmers = ["AATATACTATATAATATATATAGCGAGGGGG", "AAAACGACGACCGCAACACGACGCCAGCAGA",
             "AAGATATAATATATAAAATATATAGTGTCGT", "AAAAAAAAAAAAAAAAAAAACGCGGCGAAAA"]
    for want, km in enumerate(kmers):
        part, lo, hi = O.s1_sequences([km.encode()], 31, 10, table)
        assert list(part) == [want]


def test_stage1_partition_totals_from_task_main():
    """tests/task_main.cpp:59-116: k-mers per partition with the fixture repartition table."""
    table = _fixture_table()
    want = {"1.fasta": [37, 46, 12, 43], "2.fasta": [20, 21, 58, 39]}
    for f, tot in want.items():
        seqs = O.fastx_parse(open(os.path.join(FX, f), "rb").read())
        part, lo, hi = O.s1_sequences(seqs, 31, 10, table)
        # the fixture goldens count DISTINCT k-mers per partition (hard-min 1 lists)
        got = [len(np.unique(lo[part == p])) for p in range(4)]
        assert got == tot


@pytest.mark.parametrize("sample,fasta", [("D1", "1.fasta"), ("D2", "2.fasta")])
def test_stage2_kmer_and_hash_goldens_from_task_main(sample, fasta):
    """tests/task_main.cpp:118-340: ordered k-mers + counts of partition 0 (literal goldens) and of
    every partition (the tests/data/partitions/kmers fixtures); :342-507: ordered hashes of
    partition 0 with W = 25 000 000 -- the reference's only pin of XXH64 % W + W*p."""
    import json
    table = _fixture_table()
    gold = json.load(open(os.path.join(FX, "task_main_goldens.json")))
    bloom, P, W, Wbytes, m = struct.unpack("<QQQQI", open(os.path.join(FX, "hash.info"), "rb").read())
    assert (P, W, m) == (4, 25_000_000, 10)
    seqs = O.fastx_parse(open(os.path.join(FX, fasta), "rb").read())
    part, lo, hi = O.s1_sequences(seqs, 31, 10, table)
    for p in range(4):
        ref = O.dec_kmer_file(open(os.path.join(FX, f"partitions/kmers/partition_{p}/{sample}.kmer"), "rb").read())
        kl, _, kc = O.s2_count(lo[part == p], None, 1)
        assert kl.tolist() == ref["lo"].tolist() and kc.tolist() == ref["count"].tolist()
    kl, _, kc = O.s2_count(lo[part == 0], None, 1)
    assert [(enc(a), b) for a, b in gold[f"kmer_{sample}_p0"]] == list(zip(kl.tolist(), kc.tolist()))
    hk, _, hc = O.s2_count(O.hash_keys(lo[part == 0], hi[part == 0], 1, W, 0), None, 1)
    assert [tuple(x) for x in gold[f"hash_{sample}_p0"]] == list(zip(hk.tolist(), hc.tolist()))


def test_merge_row_counts_from_merge_test():
    """tests/merge_test.cpp:5-78: 57/67/70/82 merged rows, soft_min=1, r_min=1, save_if=1."""
    for kind, ext, dec in (("kmers", "kmer", O.dec_kmer_file), ("hashes", "hash", O.dec_hash_file)):
        for p, want in enumerate([57, 67, 70, 82]):
            lists = []
            for s in ("D1", "D2"):
                d = dec(open(os.path.join(FX, f"partitions/{kind}/partition_{p}/{s}.{ext}"), "rb").read())
                keys = d["lo"] if kind == "kmers" else d["keys"]
                lists.append((keys, np.zeros(len(keys), np.uint64), d["count"]))
            r = O.s3_merge(lists, 1, [1, 1], 1, 1, emit_all=True)
            assert r["n_union"] == want


def test_hard_min_semantics_from_processor_test():
    """tests/processor_test.cpp:10-75: count >= hard_min survives, below is dropped."""
    keys = np.array([5, 5, 5, 9, 9, 2], dtype=np.uint64)
    k, _, c = O.s2_count(keys, None, 2)
    assert k.tolist() == [5, 9] and c.tolist() == [3, 2]
    k, _, c = O.s2_count(keys, None, 4)
    assert k.tolist() == []


def test_transpose_is_involution_and_layout():
    """tests/bit_matrix_test.cpp:60-99 (T(T(M)) == M) + the bit-order relation of SURVEY §9.1."""
    rng = np.random.default_rng(1)
    for nrows, ncols in ((64, 64), (128, 16), (25024, 8)):
        a = rng.integers(0, 256, nrows * ncols // 8, dtype=np.uint8)
        t = O.transpose_bits(a, nrows, ncols)
        assert np.array_equal(O.transpose_bits(t, ncols, nrows), a)
        A = np.unpackbits(a.reshape(nrows, ncols // 8), axis=1, bitorder="little")
        T = np.unpackbits(t.reshape(ncols, nrows // 8), axis=1, bitorder="little")
        assert np.array_equal(A.T, T)


def test_window_bits_matches_hash_info_fixture():
    bloom, P, W, Wbytes, m = struct.unpack("<QQQQI", open(os.path.join(FX, "hash.info"), "rb").read())
    assert O.window_bits(100_000_000, 4) == W and Wbytes == W // 8 and bloom == W * P


def test_fastx_parser_edge_cases():
    recs = O.fastx_parse(b">a\nACGT\nAC\n\n>b\n\n>c desc\nGG\r\nTT\r\n@q\nACGT\n+\nIIII\n@q2\nAC\nGT\n+q2\nII\nII\n>last\nA")
    assert recs == [b"ACGTAC", b"", b"GGTT", b"ACGT", b"ACGT", b"A"]
