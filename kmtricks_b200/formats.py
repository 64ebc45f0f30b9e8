"""Byte-exact writers for the kmtricks on-disk formats the hot path produces (host side).

Headers follow include/kmtricks/io/*.hpp of the reference (all little-endian, packed):
KmHeader io_common.hpp:125-158; .kmer kmer_file.hpp:31-40; .hash hash_file.hpp:31-38,91-131;
.count matrix_file.hpp:31-41; .pa pa_matrix_file.hpp:31-41; .count_hash/.pa_hash
matrix_hash_file.hpp / pa_matrix_hash_file.hpp; .cmbf vector_matrix_file.hpp:31-40;
.vector vector_file.hpp:26-90; hash.info hash.hpp:52-60; minimRepart repartition.hpp:58-92;
merge_info merge.hpp:72-83.  The engine returns file *bodies*; these helpers add the headers.
"""
from __future__ import annotations

import math
import struct

import numpy as np

KM_MAGIC = 0x736b636972746d6b
MAGIC = dict(kmer=0x72656d6b, hash=0x68736168, count=0x6b5f78697274616d, pa=0x6b5f74616d6170,
             count_hash=0x685f78697274616d, pa_hash=0x685f74616d6170, cmbf=0x74616d746962,
             vector=0x726f74636576)
STAT_NAMES = ["NON_SOLID", "RESCUED", "UNIQUE_WO_RESCUE", "UNIQUE_W_RESCUE", "TOTAL_WO_RESCUE", "TOTAL_W_RESCUE"]


def km_header(compressed: int = 0) -> bytes:
    return struct.pack("<QIB", KM_MAGIC, 0, compressed)


def window_bits(bloom_size: int, P: int) -> int:
    """HashWindow (hash.hpp:31-38): roundup64(ceil(bloom/P)), computed through doubles as there."""
    return (int(math.ceil(float(bloom_size) / float(P))) + 63) // 64 * 64


def kmer_file(keys: np.ndarray, counts: np.ndarray, k: int, sample_idx: int, p: int) -> bytes:
    w = (k + 31) // 32
    rec = np.zeros(len(counts), dtype=[("k", "<u8", (w,)), ("c", "<u4")])
    rec["k"] = keys.reshape(-1, w)
    rec["c"] = counts
    return km_header() + struct.pack("<QIIIII", MAGIC["kmer"], k, w, 4, sample_idx, p) + rec.tobytes()


def hash_file(keys: np.ndarray, counts: np.ndarray, sample_idx: int, p: int, block: int = 4096) -> bytes:
    out = [km_header() + struct.pack("<QIII", MAGIC["hash"], 4, sample_idx, p)]
    for i in range(0, len(keys), block):
        kk = np.ascontiguousarray(keys[i:i + block], dtype="<u8")
        cc = np.ascontiguousarray(counts[i:i + block], dtype="<u4")
        out.append(struct.pack("<Q", len(kk)) + kk.tobytes() + cc.tobytes())
    return b"".join(out)


def matrix_header(fmt: str, key_kind: str, k: int, N: int, p: int, W: int = 0) -> bytes:
    w = (k + 31) // 32
    nb = (N + 7) // 8
    if key_kind == "kmer" and fmt == "count":   # count_slots literal 1, partition field 0
        return km_header() + struct.pack("<QIIIIII", MAGIC["count"], k, w, 1, N, 0, 0)
    if key_kind == "kmer" and fmt == "pa":
        return km_header() + struct.pack("<QIIIIII", MAGIC["pa"], k, w, N, nb, 0, 0)
    if key_kind == "hash" and fmt == "count":
        return km_header() + struct.pack("<QIIII", MAGIC["count_hash"], 4, N, 0, p)
    if key_kind == "hash" and fmt == "pa":
        return km_header() + struct.pack("<QIIII", MAGIC["pa_hash"], N, nb, 0, p)
    if key_kind == "hash" and fmt in ("bf", "bft"):
        return km_header() + struct.pack("<QIQQII", MAGIC["cmbf"], N, W * p, W, 0, p)
    raise ValueError((fmt, key_kind))


MATRIX_EXT = {("kmer", "count"): "count", ("kmer", "pa"): "pa", ("hash", "count"): "count_hash",
              ("hash", "pa"): "pa_hash", ("hash", "bf"): "cmbf", ("hash", "bft"): "cmbf"}


def vector_file(bits: bytes, W: int, p: int) -> bytes:
    return km_header() + struct.pack("<QQII", MAGIC["vector"], W, 0, p) + bits


def merge_info(stats: np.ndarray) -> bytes:
    return "".join(n + "\t" + "".join(f"{int(v)}\t" for v in row) + "\n" for n, row in zip(STAT_NAMES, stats)).encode()


def hash_info(bloom_size: int, P: int, m: int) -> bytes:
    W = window_bits(bloom_size, P)
    return struct.pack("<QQQQI", W * P, P, W, W // 8, m)


def minim_repart(table: np.ndarray, P: int) -> bytes:
    return struct.pack("<HQH", P, len(table), 1) + np.ascontiguousarray(table, "<u2").tobytes() + struct.pack("<BI", 0, 0x12345678)


def read_minim_repart(b: bytes):
    P, n, _ = struct.unpack_from("<HQH", b, 0)
    return P, np.frombuffer(b, "<u2", n, 12).copy()


def xxh64_u32(x: np.ndarray) -> np.ndarray:
    """XXH64 of little-endian uint32 inputs (len 4, seed 0), vectorised: the --static-repart
    map of repartition.hpp:45-56 is xxh64_u32(minimizer) % P."""
    P1, P2, P3, P5 = (np.uint64(0x9E3779B185EBCA87), np.uint64(0xC2B2AE3D27D4EB4F),
                      np.uint64(0x165667B19E3779F9), np.uint64(0x27D4EB2F165667C5))
    with np.errstate(over="ignore"):
        h = np.full(x.shape, P5, dtype=np.uint64) + np.uint64(4)
        h ^= x.astype(np.uint64) * P1
        h = ((h << np.uint64(23)) | (h >> np.uint64(41))) * P2 + P3
        h ^= h >> np.uint64(33); h *= P2; h ^= h >> np.uint64(29); h *= P3; h ^= h >> np.uint64(32)
    return h


def static_repart_table(m: int, P: int) -> np.ndarray:
    return (xxh64_u32(np.arange(1 << (2 * m), dtype=np.uint32)) % np.uint64(P)).astype(np.uint16)
