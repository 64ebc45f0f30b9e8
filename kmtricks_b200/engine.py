"""Host-side mirror of the reference task classes for the hot path (include/kmtricks/task.hpp):
SuperKTask -> Engine.superk, CountTask / HashCountTask -> Engine.count, KmerMergeTask /
HashMergeTask -> Engine.merge.  Everything computes in libkmx_sm100.so through the C ABI;
this module only moves bytes and adds file headers."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib, formats


class KmxError(RuntimeError):
    pass


@dataclass
class Config:
    kmer_size: int = 31
    minim_size: int = 10
    nb_partitions: int = 4
    mode: str = "kmer:count:bin"       # <kmer|hash>:<count|pa|bf|bft>:bin
    hard_min: int = 2
    soft_min: int = 1
    recurrence_min: int = 1
    share_min: int = 0
    bloom_size: int = 10_000_000
    repart_table: np.ndarray | None = None     # None = --static-repart

    @property
    def key_kind(self): return self.mode.split(":")[0]
    @property
    def fmt(self): return self.mode.split(":")[1]
    @property
    def w(self): return (self.kmer_size + 31) // 32
    @property
    def window_bits(self): return formats.window_bits(self.bloom_size, self.nb_partitions)


def parse_fastx(buf: bytes) -> list[bytes]:
    """kseq-style FASTA/FASTQ reader (host side; gatb BankFasta.cpp:391-560 semantics)."""
    out, n, pos = [], len(buf), 0
    last = 0
    while True:
        if last == 0:
            while pos < n and buf[pos] not in (62, 64):
                pos += 1
            if pos >= n:
                break
            last = buf[pos]; pos += 1
        if pos >= n:
            break
        e = buf.find(b"\n", pos)
        pos = n if e < 0 else e + 1
        seq = bytearray()
        c = -1
        while pos < n:
            c = buf[pos]; pos += 1
            if c in (62, 43, 64):
                break
            if c == 10:
                c = -1
                continue
            e = buf.find(b"\n", pos)
            e = n if e < 0 else e
            seq.append(c); seq += buf[pos:e]
            pos = min(e + 1, n)
            if len(seq) > 1 and seq[-1] == 13:
                seq.pop()
            c = -1
        if c in (62, 64):
            last = c
        if c == 43:
            e = buf.find(b"\n", pos)
            pos = n if e < 0 else e + 1
            qlen = 0
            while pos < n:
                e = buf.find(b"\n", pos)
                e2 = n if e < 0 else e
                l = e2 - pos
                qlen += l
                if qlen > 1 and l > 0 and buf[e2 - 1] == 13:
                    qlen -= 1
                pos = min(e2 + 1, n)
                if qlen >= len(seq):
                    break
            last = 0
        out.append(bytes(seq))
        if pos >= n:
            break
    return out


class Engine:
    def __init__(self, cfg: Config, nb_samples: int, device: int = 0):
        self.lib = _lib.load()
        self.cfg = cfg
        self.N = nb_samples
        self.P = cfg.nb_partitions
        table = cfg.repart_table if cfg.repart_table is not None else formats.static_repart_table(cfg.minim_size, cfg.nb_partitions)
        self.table = np.ascontiguousarray(table, dtype=np.uint16)
        prm = _lib.KmxParams(cfg.kmer_size, cfg.minim_size, cfg.nb_partitions,
                             _lib.KEY_HASH if cfg.key_kind == "hash" else _lib.KEY_KMER,
                             cfg.window_bits if cfg.key_kind == "hash" else 0,
                             self.table.ctypes.data_as(C.POINTER(C.c_uint16)), nb_samples, 0)
        h = C.c_void_p()
        rc = self.lib.kmx_create(device, C.byref(prm), C.byref(h))
        self.h = h
        if rc:
            msg = self.lib.kmx_last_error(h).decode() if h else "kmx_create failed"
            if h:
                self.lib.kmx_destroy(h)
            self.h = None
            raise KmxError(f"kmx_create: {msg} (code {rc})")

    def close(self):
        if getattr(self, "h", None):
            self.lib.kmx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise KmxError(f"{what}: {self.lib.kmx_last_error(self.h).decode()} (code {rc})")

    # ---- SuperKTask ----------------------------------------------------------------------
    def superk(self, bufs: list[bytes]) -> np.ndarray:
        """One sample = list of FASTA/FASTQ buffers.  Returns the .pinfo vector (k-mers/partition)."""
        L = self.lib
        self._ck(L.kmx_superk_begin(self.h), "superk_begin")
        for b in bufs:
            rc = L.kmx_superk_push_fastq(self.h, b, len(b), 0) if b[:1] == b"@" else _lib.KMX_ERR_FORMAT
            if rc == _lib.KMX_ERR_FORMAT:
                seqs = parse_fastx(b)
                off = np.zeros(len(seqs) + 1, dtype=np.uint64)
                off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
                cat = b"".join(seqs)
                rc = L.kmx_superk_push_reads(self.h, cat, off.ctypes.data_as(C.POINTER(C.c_uint64)), len(seqs))
            self._ck(rc, "superk_push")
        pin = np.zeros(self.P, dtype=np.uint64)
        self._ck(L.kmx_superk_end(self.h, pin.ctypes.data_as(C.POINTER(C.c_uint64))), "superk_end")
        return pin

    def superk_device(self, dev_ptr: int, nbytes: int) -> np.ndarray:
        L = self.lib
        self._ck(L.kmx_superk_begin(self.h), "superk_begin")
        self._ck(L.kmx_superk_push_fastq(self.h, dev_ptr, nbytes, 1), "superk_push(device)")
        pin = np.zeros(self.P, dtype=np.uint64)
        self._ck(L.kmx_superk_end(self.h, pin.ctypes.data_as(C.POINTER(C.c_uint64))), "superk_end")
        return pin

    # ---- CountTask / HashCountTask ------------------------------------------------------
    def count(self, sample: int, hard_min: int | None = None):
        self._ck(self.lib.kmx_count_sample(self.h, sample, self.cfg.hard_min if hard_min is None else hard_min), "count_sample")

    def counts(self, sample: int, p: int):
        n = C.c_uint64()
        self._ck(self.lib.kmx_counts_size(self.h, sample, p, C.byref(n)), "counts_size")
        w = self.cfg.w if self.cfg.key_kind == "kmer" else 1
        keys = np.empty(n.value * w, dtype=np.uint64)
        cnt = np.empty(n.value, dtype=np.uint32)
        self._ck(self.lib.kmx_counts_get(self.h, sample, p, keys.ctypes.data, cnt.ctypes.data), "counts_get")
        return keys, cnt

    def counts_file(self, sample: int, p: int) -> bytes:
        keys, cnt = self.counts(sample, p)
        if self.cfg.key_kind == "hash":
            return formats.hash_file(keys, cnt, sample, p)
        return formats.kmer_file(keys, cnt, self.cfg.kmer_size, sample, p)

    def put_counts(self, sample: int, p: int, keys: np.ndarray, cnt: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64); cnt = np.ascontiguousarray(cnt, dtype=np.uint32)
        self._ck(self.lib.kmx_counts_put(self.h, sample, p, keys.ctypes.data, cnt.ctypes.data, len(cnt)), "counts_put")

    def vector(self, sample: int, p: int) -> bytes:
        W = self.cfg.window_bits
        out = np.empty(W // 8, dtype=np.uint8)
        self._ck(self.lib.kmx_counts_vector(self.h, sample, p, out.ctypes.data), "counts_vector")
        return formats.vector_file(out.tobytes(), W, p)

    # ---- KmerMergeTask / HashMergeTask ------------------------------------------------
    def merge(self, p: int, fmt: str | None = None, soft_min=None, emit_all: bool = False, download: bool = True):
        cfg = self.cfg
        fmt = fmt or cfg.fmt
        sm = soft_min if soft_min is not None else cfg.soft_min
        sm = np.ascontiguousarray(sm if hasattr(sm, "__len__") else [sm] * self.N, dtype=np.uint32)
        mp = _lib.KmxMergeParams(sm.ctypes.data_as(C.POINTER(C.c_uint32)), cfg.recurrence_min, cfg.share_min,
                                 {"count": 0, "pa": 1, "bf": 2, "bft": 3}[fmt], int(emit_all))
        res = _lib.KmxMergeResult()
        self._ck(self.lib.kmx_merge_partition(self.h, p, C.byref(mp), C.byref(res)), "merge_partition")
        if not download:
            return res
        body = np.empty(res.n_rows * res.row_bytes, dtype=np.uint8)
        stats = np.zeros((6, self.N), dtype=np.uint64)
        keep = np.empty(res.n_rows if emit_all else 0, dtype=np.uint8)
        self._ck(self.lib.kmx_merge_get(self.h, body.ctypes.data, stats.ctypes.data, keep.ctypes.data if emit_all else None), "merge_get")
        return dict(body=body, stats=stats, row_keep=keep, n_rows=res.n_rows, row_bytes=res.row_bytes, n_union=res.n_union)

    def matrix_file(self, p: int, merged: dict, fmt: str | None = None) -> bytes:
        cfg = self.cfg
        fmt = fmt or cfg.fmt
        return formats.matrix_header(fmt, cfg.key_kind, cfg.kmer_size, self.N, p, cfg.window_bits) + merged["body"].tobytes()

    def transpose_bits(self, a: np.ndarray, nrows: int, ncols: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint8)
        out = np.empty(nrows * ncols // 8, dtype=np.uint8)
        self._ck(self.lib.kmx_transpose_bits(self.h, a.ctypes.data, nrows, ncols, out.ctypes.data), "transpose_bits")
        return out

    def launches(self) -> int:
        return self.lib.kmx_launch_count(self.h)

    def reset(self):
        self._ck(self.lib.kmx_reset(self.h), "reset")


def run_pipeline(samples: list[list[bytes]], cfg: Config, sample_hard_min: dict | None = None, device: int = 0):
    """Whole hot path for small inputs: returns dict(pinfo, counts[(s,p)] file bytes,
    matrices[p] file bytes, merge_info[p] bytes) -- the same files the reference writes."""
    N = len(samples)
    eng = Engine(cfg, N, device)
    out = dict(pinfo=[], counts={}, matrices={}, merge_info={})
    try:
        for s, bufs in enumerate(samples):
            out["pinfo"].append(eng.superk(bufs))
            hm = (sample_hard_min or {}).get(s, 0) or cfg.hard_min
            eng.count(s, hm)
            for p in range(cfg.nb_partitions):
                out["counts"][(s, p)] = eng.counts_file(s, p)
        for p in range(cfg.nb_partitions):
            m = eng.merge(p)
            out["matrices"][p] = eng.matrix_file(p, m)
            out["merge_info"][p] = formats.merge_info(m["stats"])
        out["launches"] = eng.launches()
        out["s1_self_indexed"] = int(eng.lib.kmx_stat(eng.h, 0)); out["s1_indexed"] = int(eng.lib.kmx_stat(eng.h, 1))
        out["hash_binned"] = int(eng.lib.kmx_stat(eng.h, 2))
    finally:
        eng.close()
    return out


def run_pipeline_lanes(texts: list[bytes], cfg: Config, lanes: int = 4, device: int = 0, want_counts: bool = True):
    """Same outputs as run_pipeline, but through kmx_run_samples (several samples in flight on their own
    streams: the path bench.py times).  One strict 4-line FASTQ text per sample, host buffers."""
    N = len(texts)
    eng = Engine(cfg, N, device)
    L = eng.lib
    out = dict(pinfo=[], counts={}, matrices={}, merge_info={})
    try:
        bufs = [C.create_string_buffer(t, len(t)) for t in texts]
        ptrs = (C.c_void_p * N)(*[C.addressof(b) for b in bufs])
        sizes = (C.c_size_t * N)(*[len(t) for t in texts])
        hm = (C.c_uint32 * N)(*([cfg.hard_min] * N))
        pin = np.zeros((N, cfg.nb_partitions), dtype=np.uint64)
        eng._ck(L.kmx_run_samples(eng.h, N, ptrs, sizes, 0, None, hm, lanes, pin.ctypes.data_as(C.POINTER(C.c_uint64))), "run_samples")
        out["pinfo"] = [pin[s] for s in range(N)]
        if want_counts:
            for s in range(N):
                for p in range(cfg.nb_partitions):
                    out["counts"][(s, p)] = eng.counts_file(s, p)
        for p in range(cfg.nb_partitions):
            m = eng.merge(p)
            out["matrices"][p] = eng.matrix_file(p, m)
            out["merge_info"][p] = formats.merge_info(m["stats"])
        out["launches"] = eng.launches()
        out["hash_binned"] = int(eng.lib.kmx_stat(eng.h, 2))
    finally:
        eng.close()
    return out
