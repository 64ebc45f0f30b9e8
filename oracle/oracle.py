"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/kmx_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (kmtricks_b200) never does.

It wraps the C restatement with ctypes/numpy, adds the byte-level file encoders/decoders of
the reference formats (include/kmtricks/io/*.hpp), and a whole-pipeline driver
`run_pipeline` that produces, per partition, exactly the bytes the reference writes to
counts/partition_P/<id>.{kmer,hash} and matrices/matrix_P.<ext>.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkmx_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kmx_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u64p, u32p, u16p, u8p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint16), C.POINTER(C.c_uint8))
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.orc_minim_lut.argtypes = [C.c_int, u32p]
        L.orc_repart_static.argtypes = [C.c_int, C.c_uint32, u16p]
        L.orc_s1_seq.restype = C.c_size_t
        L.orc_s1_seq.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, u32p, u16p, u16p, u64p, u64p]
        L.orc_minimizer_of.restype = C.c_uint32
        L.orc_minimizer_of.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, u32p]
        L.orc_fastx_parse.restype = C.c_size_t
        L.orc_fastx_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, u64p, C.c_size_t]
        L.orc_hash_key.restype = C.c_uint64
        L.orc_hash_key.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64]
        L.orc_s2_count.restype = C.c_size_t
        L.orc_s2_count.argtypes = [u64p, u64p, C.c_size_t, C.c_uint32, u64p, u64p, u32p]
        L.orc_s3_merge.restype = C.c_size_t
        L.orc_s3_merge.argtypes = [C.c_int, C.c_uint32, u64p, u64p, u64p, u32p, u32p, C.c_uint32,
                                   C.c_uint32, C.c_int, u64p, u64p, u32p, u8p, u64p, u64p]
        L.orc_bf_slab.argtypes = [u64p, u32p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_uint64, u8p]
        L.orc_transpose_bits.argtypes = [u8p, C.c_size_t, C.c_size_t, u8p]
        L.orc_hash_vector.argtypes = [u64p, C.c_size_t, C.c_uint64, C.c_uint64, u8p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


# ---------------------------------------------------------------------------- primitives
def xxh64(data: bytes, seed: int = 0) -> int:
    return lib().orc_xxh64(data, len(data), seed)


_LUTS: dict[int, np.ndarray] = {}


def minim_lut(m: int) -> np.ndarray:
    if m not in _LUTS:
        a = np.empty(1 << (2 * m), dtype=np.uint32)
        lib().orc_minim_lut(m, _p(a, C.c_uint32))
        _LUTS[m] = a
    return _LUTS[m]


def repart_static(m: int, P: int) -> np.ndarray:
    a = np.empty(1 << (2 * m), dtype=np.uint16)
    lib().orc_repart_static(m, P, _p(a, C.c_uint16))
    return a


def window_bits(bloom_size: int, P: int) -> int:
    """include/kmtricks/hash.hpp:31-38: W = roundup64(ceil(bloom/P)) (via doubles, as there)."""
    import math
    w = int(math.ceil(float(bloom_size) / float(P)))
    return (w + 63) // 64 * 64


def fastx_parse(buf: bytes) -> list[bytes]:
    n = len(buf)
    seq = C.create_string_buffer(n + 1)
    max_rec = buf.count(b">") + buf.count(b"@") + 1
    off = np.zeros(max_rec + 1, dtype=np.uint64)
    nrec = lib().orc_fastx_parse(buf, n, seq, _p(off, C.c_uint64), max_rec)
    raw = seq.raw
    return [raw[int(off[i]):int(off[i + 1])] for i in range(nrec)]


def s1_sequences(seqs: list[bytes], k: int, m: int, table: np.ndarray):
    """-> (part u16[], lo u64[], hi u64[]) for all valid k-mers of all sequences, in order."""
    lut = minim_lut(m)
    parts, los, his = [], [], []
    for s in seqs:
        if len(s) < k:
            continue
        cap = len(s) - k + 1
        p = np.empty(cap, dtype=np.uint16)
        lo = np.empty(cap, dtype=np.uint64)
        hi = np.zeros(cap, dtype=np.uint64)
        n = lib().orc_s1_seq(s, len(s), k, m, _p(lut, C.c_uint32), _p(table, C.c_uint16),
                             _p(p, C.c_uint16), _p(lo, C.c_uint64), _p(hi, C.c_uint64))
        parts.append(p[:n]); los.append(lo[:n]); his.append(hi[:n])
    if not parts:
        z = np.zeros(0, dtype=np.uint64)
        return np.zeros(0, dtype=np.uint16), z, z.copy()
    return np.concatenate(parts), np.concatenate(los), np.concatenate(his)


def hash_keys(lo: np.ndarray, hi: np.ndarray, w: int, W: int, p: int) -> np.ndarray:
    """Vectorised closed form of XXH64 for 8- and 16-byte inputs (seed 0) % W + W*p; checked
    against orc_hash_key / orc_xxh64 in tests/test_oracle_kat.py."""
    P1, P2, P3, P4, P5 = (np.uint64(0x9E3779B185EBCA87), np.uint64(0xC2B2AE3D27D4EB4F),
                          np.uint64(0x165667B19E3779F9), np.uint64(0x85EBCA77C2B2AE63),
                          np.uint64(0x27D4EB2F165667C5))

    def rotl(x, r):
        return (x << np.uint64(r)) | (x >> np.uint64(64 - r))

    with np.errstate(over="ignore"):
        h = np.full(lo.shape, P5, dtype=np.uint64) + np.uint64(8 * w)
        for word in ([lo] if w == 1 else [lo, hi]):
            k1 = rotl(word * P2, 31) * P1
            h = rotl(h ^ k1, 27) * P1 + P4
        h ^= h >> np.uint64(33); h *= P2; h ^= h >> np.uint64(29); h *= P3; h ^= h >> np.uint64(32)
        return h % np.uint64(W) + np.uint64(W * p)


def s2_count(lo: np.ndarray, hi: np.ndarray | None, hard_min: int):
    n = len(lo)
    lo = np.ascontiguousarray(lo, dtype=np.uint64)
    olo = np.empty(n, dtype=np.uint64)
    ohi = np.zeros(n, dtype=np.uint64)
    oc = np.empty(n, dtype=np.uint32)
    if hi is not None:
        hi = np.ascontiguousarray(hi, dtype=np.uint64)
    m = lib().orc_s2_count(_p(lo, C.c_uint64), _p(hi, C.c_uint64), n, hard_min,
                           _p(olo, C.c_uint64), _p(ohi, C.c_uint64), _p(oc, C.c_uint32))
    return olo[:m].copy(), ohi[:m].copy(), oc[:m].copy()


STAT_NAMES = ["NON_SOLID", "RESCUED", "UNIQUE_WO_RESCUE", "UNIQUE_W_RESCUE",
              "TOTAL_WO_RESCUE", "TOTAL_W_RESCUE"]


def s3_merge(lists, w: int, soft_min, r_min: int, save_if: int, emit_all: bool = False):
    """lists: per sample (lo, hi, count).  -> dict(lo, hi, counts[n,N], keep, stats[6,N], n_union)"""
    N = len(lists)
    off = np.zeros(N + 1, dtype=np.uint64)
    for i, (l, _, _) in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    tot = int(off[N])
    lo = np.concatenate([x[0] for x in lists]).astype(np.uint64) if tot else np.zeros(0, np.uint64)
    hi = np.concatenate([x[1] for x in lists]).astype(np.uint64) if tot else np.zeros(0, np.uint64)
    cnt = np.concatenate([x[2] for x in lists]).astype(np.uint32) if tot else np.zeros(0, np.uint32)
    sm = np.ascontiguousarray(soft_min, dtype=np.uint32)
    stats = np.zeros((6, N), dtype=np.uint64)
    nun = C.c_uint64(0)
    rlo = np.empty(max(tot, 1), dtype=np.uint64)
    rhi = np.zeros(max(tot, 1), dtype=np.uint64)
    rc = np.empty((max(tot, 1), N), dtype=np.uint32)
    rk = np.empty(max(tot, 1), dtype=np.uint8)
    n = lib().orc_s3_merge(w, N, _p(off, C.c_uint64), _p(lo, C.c_uint64), _p(hi, C.c_uint64),
                           _p(cnt, C.c_uint32), _p(sm, C.c_uint32), r_min, save_if, int(emit_all),
                           _p(rlo, C.c_uint64), _p(rhi, C.c_uint64), _p(rc, C.c_uint32),
                           _p(rk, C.c_uint8), _p(stats, C.c_uint64), C.byref(nun))
    return dict(lo=rlo[:n].copy(), hi=rhi[:n].copy(), counts=rc[:n].copy(), keep=rk[:n].copy(),
                stats=stats, n_union=nun.value)


def transpose_bits(a: np.ndarray, nrows: int, ncols: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    out = np.empty(ncols * (nrows // 8), dtype=np.uint8)
    lib().orc_transpose_bits(_p(a, C.c_uint8), nrows, ncols, _p(out, C.c_uint8))
    return out


def pa_rows(counts: np.ndarray) -> np.ndarray:
    """[n,N] u32 -> [n, ceil(N/8)] u8, LSB-first (include/kmtricks/utils.hpp:104-116)."""
    n, N = counts.shape
    return np.packbits(counts != 0, axis=1, bitorder="little") if n else np.zeros((0, (N + 7) // 8), np.uint8)


# ---------------------------------------------------------------------------- file formats
KM_MAGIC = 0x736b636972746d6b
MAGIC = dict(kmer=0x72656d6b, hash=0x68736168, count=0x6b5f78697274616d, pa=0x6b5f74616d6170,
             count_hash=0x685f78697274616d, pa_hash=0x685f74616d6170, cmbf=0x74616d746962,
             vector=0x726f74636576, superk=0x6b7265707573)


def _kmh(compressed=0):
    return struct.pack("<QIB", KM_MAGIC, 0, compressed)


def enc_kmer_file(lo, hi, cnt, k, sample_idx, p):
    """counts/partition_P/<id>.kmer -- include/kmtricks/io/kmer_file.hpp:31-40,73-108."""
    w = (k + 31) // 32
    hdr = _kmh() + struct.pack("<QIIIII", MAGIC["kmer"], k, w, 4, sample_idx, p)
    rec = np.zeros(len(lo), dtype=[("k", "<u8", (w,)), ("c", "<u4")])
    rec["k"][:, 0] = lo
    if w == 2:
        rec["k"][:, 1] = hi
    rec["c"] = cnt
    return hdr + rec.tobytes()


def enc_hash_file(keys, cnt, sample_idx, p, block=4096):
    """counts/partition_P/<id>.hash -- include/kmtricks/io/hash_file.hpp:31-38,91-131."""
    out = [_kmh() + struct.pack("<QIII", MAGIC["hash"], 4, sample_idx, p)]
    for i in range(0, len(keys), block):
        kk = np.ascontiguousarray(keys[i:i + block], dtype="<u8")
        cc = np.ascontiguousarray(cnt[i:i + block], dtype="<u4")
        out.append(struct.pack("<Q", len(kk)) + kk.tobytes() + cc.tobytes())
    return b"".join(out)


def dec_kmer_file(b: bytes):
    km, ver, cpr, mg, k, w, cb, sid, p = struct.unpack_from("<QIBQIIIII", b, 0)
    assert km == KM_MAGIC and mg == MAGIC["kmer"] and cpr == 0
    dt = np.dtype([("k", "<u8", (w,)), ("c", {1: "u1", 2: "<u2", 4: "<u4"}[cb])])
    rec = np.frombuffer(b, dtype=dt, offset=41)
    lo = rec["k"][:, 0].copy()
    hi = rec["k"][:, 1].copy() if w == 2 else np.zeros(len(rec), np.uint64)
    return dict(k=k, w=w, count_bytes=cb, id=sid, partition=p, lo=lo, hi=hi, count=rec["c"].astype(np.uint32))


def dec_hash_file(b: bytes):
    km, ver, cpr, mg, cb, sid, p = struct.unpack_from("<QIBQIII", b, 0)
    assert km == KM_MAGIC and mg == MAGIC["hash"] and cpr == 0
    pos = 33
    ks, cs = [], []
    cdt = {1: "u1", 2: "<u2", 4: "<u4"}[cb]
    while pos < len(b):
        (n,) = struct.unpack_from("<Q", b, pos); pos += 8
        ks.append(np.frombuffer(b, "<u8", n, pos)); pos += 8 * n
        cs.append(np.frombuffer(b, cdt, n, pos).astype(np.uint32)); pos += cb * n
    z = np.zeros(0, np.uint64)
    return dict(count_bytes=cb, id=sid, partition=p, keys=np.concatenate(ks) if ks else z,
                count=np.concatenate(cs) if cs else np.zeros(0, np.uint32))


def enc_count_matrix(lo, hi, counts, k, N, partition_field=0):
    """matrices/matrix_P.count -- io/matrix_file.hpp:31-41,94-128; merge.hpp:262-272.
    count_slots is the literal 1 (F8); `partition` is uninitialised in the reference (F7),
    0 observed -> partition_field."""
    w = (k + 31) // 32
    hdr = _kmh() + struct.pack("<QIIIIII", MAGIC["count"], k, w, 1, N, 0, partition_field)
    rec = np.zeros(len(lo), dtype=[("k", "<u8", (w,)), ("c", "<u4", (N,))])
    rec["k"][:, 0] = lo
    if w == 2:
        rec["k"][:, 1] = hi
    rec["c"] = counts
    return hdr + rec.tobytes()


def enc_pa_matrix(lo, hi, counts, k, N, partition_field=0):
    """matrices/matrix_P.pa -- io/pa_matrix_file.hpp:31-41,73-106; merge.hpp:274-286."""
    w = (k + 31) // 32
    nb = (N + 7) // 8
    hdr = _kmh() + struct.pack("<QIIIIII", MAGIC["pa"], k, w, N, nb, 0, partition_field)
    rec = np.zeros(len(lo), dtype=[("k", "<u8", (w,)), ("b", "u1", (nb,))])
    rec["k"][:, 0] = lo
    if w == 2:
        rec["k"][:, 1] = hi
    rec["b"] = pa_rows(counts)
    return hdr + rec.tobytes()


def enc_count_hash_matrix(keys, counts, N, p):
    """matrices/matrix_P.count_hash (37 B header) -- merge.hpp:519-529."""
    hdr = _kmh() + struct.pack("<QIIII", MAGIC["count_hash"], 4, N, 0, p)
    rec = np.zeros(len(keys), dtype=[("k", "<u8"), ("c", "<u4", (N,))])
    rec["k"] = keys; rec["c"] = counts
    return hdr + rec.tobytes()


def enc_pa_hash_matrix(keys, counts, N, p):
    """matrices/matrix_P.pa_hash (37 B header) -- merge.hpp:546-558."""
    nb = (N + 7) // 8
    hdr = _kmh() + struct.pack("<QIIII", MAGIC["pa_hash"], N, nb, 0, p)
    rec = np.zeros(len(keys), dtype=[("k", "<u8"), ("b", "u1", (nb,))])
    rec["k"] = keys; rec["b"] = pa_rows(counts)
    return hdr + rec.tobytes()


def cmbf_header(N, W, p):
    """49 B -- io/vector_matrix_file.hpp:31-40."""
    return _kmh() + struct.pack("<QIQQII", MAGIC["cmbf"], N, W * p, W, 0, p)


def enc_cmbf(keys, counts, N, W, p):
    """matrices/matrix_P.cmbf -- merge.hpp:575-600."""
    nb = (N + 7) // 8
    slab = np.zeros((W, nb), dtype=np.uint8)
    if len(keys):
        slab[(keys - np.uint64(W * p)).astype(np.int64)] = pa_rows(counts)
    return cmbf_header(N, W, p) + slab.tobytes()


def enc_bft(keys, counts, N, W, p):
    """write_as_bft -- merge.hpp:631-644: W x (8*ceil(N/8)) bit slab transposed."""
    nb = (N + 7) // 8
    slab = np.zeros((W, nb), dtype=np.uint8)
    if len(keys):
        slab[(keys - np.uint64(W * p)).astype(np.int64)] = pa_rows(counts)
    t = transpose_bits(slab.reshape(-1), W, nb * 8)
    return cmbf_header(N, W, p) + t.tobytes()


def enc_vector(keys, W, p):
    """counts/partition_P/<id>.vector (37 B header) -- io/vector_file.hpp:26-90."""
    hdr = _kmh() + struct.pack("<QQII", MAGIC["vector"], W, 0, p)
    bits = np.zeros(W, dtype=np.uint8)
    bits[(keys - np.uint64(W * p)).astype(np.int64)] = 1
    return hdr + np.packbits(bits, bitorder="little").tobytes()


def enc_merge_info(stats: np.ndarray) -> bytes:
    """merge_infos/partitionP.merge_info -- merge.hpp:72-83."""
    out = []
    for name, row in zip(STAT_NAMES, stats):
        out.append(name + "\t" + "".join(f"{int(v)}\t" for v in row) + "\n")
    return "".join(out).encode()


def enc_hash_info(bloom_size, P, m) -> bytes:
    """hash.info -- include/kmtricks/hash.hpp:52-60."""
    W = window_bits(bloom_size, P)
    return struct.pack("<QQQQI", W * P, P, W, W // 8, m)


def enc_minim_repart(table: np.ndarray, P: int) -> bytes:
    """repartition_gatb/repartition.minimRepart -- include/kmtricks/repartition.hpp:58-92."""
    return (struct.pack("<HQH", P, len(table), 1) + np.ascontiguousarray(table, "<u2").tobytes()
            + struct.pack("<BI", 0, 0x12345678))


def dec_minim_repart(b: bytes):
    P, n, npass = struct.unpack_from("<HQH", b, 0)
    return P, np.frombuffer(b, "<u2", n, 12).copy()


# ---------------------------------------------------------------------------- whole pipeline
@dataclass
class Params:
    k: int = 31
    m: int = 10
    P: int = 4
    mode: str = "kmer:count:bin"          # kmer:count | kmer:pa | hash:count | hash:pa | hash:bf | hash:bft
    hard_min: int = 2
    soft_min: int | list = 1
    recurrence_min: int = 1
    share_min: int = 0
    bloom_size: int = 10_000_000
    table: np.ndarray | None = None        # repartition table; None = --static-repart
    sample_hard_min: dict = field(default_factory=dict)   # fof "! n" overrides by sample index


def run_pipeline(samples: list[list[bytes]], prm: Params):
    """samples[i] = list of FASTA/FASTQ byte buffers of sample i (fof order).
    Returns dict(pinfo[N][P], counts[(s,p)] -> file bytes, matrices[p] -> file bytes,
    merge_info[p] -> bytes, lists[(s,p)] -> (lo,hi,cnt))."""
    k, m, P = prm.k, prm.m, prm.P
    w = (k + 31) // 32
    N = len(samples)
    kind, what = prm.mode.split(":")[:2]
    table = prm.table if prm.table is not None else repart_static(m, P)
    W = window_bits(prm.bloom_size, P)
    soft = prm.soft_min if isinstance(prm.soft_min, (list, tuple, np.ndarray)) else [prm.soft_min] * N
    out = dict(pinfo=[], counts={}, matrices={}, merge_info={}, lists={}, W=W, table=table)
    for s, bufs in enumerate(samples):
        seqs = []
        for b in bufs:
            seqs += fastx_parse(b)
        part, lo, hi = s1_sequences(seqs, k, m, table)
        out["pinfo"].append(np.bincount(part, minlength=P).astype(np.uint64))
        hm = prm.sample_hard_min.get(s, 0) or prm.hard_min
        for p in range(P):
            sel = part == p
            l, h = lo[sel], hi[sel]
            if kind == "hash":
                keys = hash_keys(l, h, w, W, p)
                kl, kh, kc = s2_count(keys, None, hm)
                out["counts"][(s, p)] = enc_hash_file(kl, kc, s, p)
            else:
                kl, kh, kc = s2_count(l, h if w == 2 else None, hm)
                out["counts"][(s, p)] = enc_kmer_file(kl, kh, kc, k, s, p)
            out["lists"][(s, p)] = (kl, kh, kc)
    for p in range(P):
        lists = [out["lists"][(s, p)] for s in range(N)]
        r = s3_merge(lists, w if kind == "kmer" else 1, soft, prm.recurrence_min, prm.share_min)
        out["merge_info"][p] = enc_merge_info(r["stats"])
        if kind == "kmer" and what == "count":
            out["matrices"][p] = enc_count_matrix(r["lo"], r["hi"], r["counts"], k, N)
        elif kind == "kmer" and what == "pa":
            out["matrices"][p] = enc_pa_matrix(r["lo"], r["hi"], r["counts"], k, N)
        elif what == "count":
            out["matrices"][p] = enc_count_hash_matrix(r["lo"], r["counts"], N, p)
        elif what == "pa":
            out["matrices"][p] = enc_pa_hash_matrix(r["lo"], r["counts"], N, p)
        elif what == "bf":
            out["matrices"][p] = enc_cmbf(r["lo"], r["counts"], N, W, p)
        elif what == "bft":
            out["matrices"][p] = enc_bft(r["lo"], r["counts"], N, W, p)
        else:
            raise ValueError(prm.mode)
    return out


# ---------------------------------------------------------------------------- reference binary
REF_BIN = os.path.join(_HERE, "_ref", "bin", "kmtricks")
REF_BIN_V3 = os.path.join(_HERE, "_ref", "bin", "kmtricks_v3")   # -march=x86-64-v3 build of the same sources (timed baseline)


def timed_ref_bin():
    """(path, build note) of the reference binary to TIME on this host: the AVX2/BMI2/FMA build when the CPU has those
    (what the reference's -DNATIVE=ON would use here), else the portable build.  Parity always uses REF_BIN."""
    try:
        flags = set(next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split())
    except Exception:
        flags = set()
    if os.path.exists(REF_BIN_V3) and {"avx2", "bmi2", "fma", "movbe", "abm"} <= flags:
        return REF_BIN_V3, "-O3 -march=x86-64-v3"
    return REF_BIN, "-O3 (portable x86-64)"


def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def run_reference(fof_path: str, run_dir: str, prm: Params, threads: int = 4, keep_tmp=True,
                  until: str | None = None, extra: list[str] | None = None):
    """Run the UNMODIFIED reference CLI (oracle/_ref/bin/kmtricks, built by build_ref.sh)."""
    kind, what = prm.mode.split(":")[:2]
    cmd = [REF_BIN, "pipeline", "--file", fof_path, "--run-dir", run_dir, "--kmer-size", str(prm.k),
           "--mode", f"{kind}:{what}:bin", "--hard-min", str(prm.hard_min),
           "--soft-min", str(prm.soft_min), "--recurrence-min", str(prm.recurrence_min),
           "--share-min", str(prm.share_min), "--nb-partitions", str(prm.P),
           "--minimizer-size", str(prm.m), "--static-repart", "--bloom-size", str(prm.bloom_size),
           "-t", str(threads)]
    if keep_tmp:
        cmd.append("--keep-tmp")
    if until:
        cmd += ["--until", until]
    if extra:
        cmd += extra
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
