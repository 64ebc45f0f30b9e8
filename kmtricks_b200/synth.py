"""Deterministic synthetic FASTQ of the BASELINE.json shape (SURVEY §8(d) "Synthetic inputs").

Counter-based (splitmix64 of (stream, index)), so this numpy generator and the CUDA twin
`kmx_synth_fastq` (csrc/synth.cu) produce identical bytes; tests/test_synth.py checks that.

Record layout (strict 4-line FASTQ, fixed width): "@r%08d\n" SEQ "\n+\n" "I"*L "\n"
=> 2L+15 bytes per read.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rnd(stream: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(stream) * np.uint64(0xD1342543DE82EF95) + idx.astype(np.uint64))


def _streams(seed: int, sample: int):
    base = (seed * 1000003) & 0x7FFFFFFFFFFF
    return dict(genome=base + 1, snp=base + 1000 + 4 * sample, start=base + 1001 + 4 * sample,
                err=base + 1002 + 4 * sample, strand=base + 1003 + 4 * sample)


def prob_thr(p: float) -> int:
    return min(int(p * 4294967296.0), 0xFFFFFFFF)


def record_bytes(L: int) -> int:
    return 2 * L + 15


def sample_bases(seed: int, sample: int, pos: np.ndarray, d: float) -> np.ndarray:
    """2-bit codes in 'ACGT' letter order for genome positions `pos` of sample `sample`."""
    st = _streams(seed, sample)
    g = (rnd(st["genome"], pos) & np.uint64(3)).astype(np.uint8)
    u = rnd(st["snp"], pos)
    mut = (u >> np.uint64(32)) < np.uint64(prob_thr(d))
    alt = ((g.astype(np.uint64) + np.uint64(1) + (u & np.uint64(0xFFFF)) % np.uint64(3)) & np.uint64(3)).astype(np.uint8)
    return np.where(mut, alt, g)


def make_fastq(seed: int, sample: int, R: int, L: int = 150, G: int = 1_000_000,
               d: float = 2e-3, e: float = 2e-3, revcomp: bool = False, first_read: int = 0) -> bytes:
    st = _streams(seed, sample)
    ridx = np.arange(first_read, first_read + R, dtype=np.uint64)
    start = rnd(st["start"], ridx) % np.uint64(G - L + 1)
    j = np.arange(L, dtype=np.uint64)[None, :]
    if revcomp:
        rc = (rnd(st["strand"], ridx) & np.uint64(1)).astype(bool)[:, None]
        pos = np.where(rc, start[:, None] + np.uint64(L - 1) - j, start[:, None] + j)
    else:
        rc = None
        pos = start[:, None] + j
    b = sample_bases(seed, sample, pos.reshape(-1), d).reshape(R, L)
    if rc is not None:
        b = np.where(rc, 3 - b, b)          # ACGT order: complement = 3 - code
    u = rnd(st["err"], (ridx[:, None] * np.uint64(L) + j).reshape(-1)).reshape(R, L)
    err = (u >> np.uint64(32)) < np.uint64(prob_thr(e))
    alt = ((b.astype(np.uint64) + np.uint64(1) + (u & np.uint64(0xFFFF)) % np.uint64(3)) & np.uint64(3)).astype(np.uint8)
    b = np.where(err, alt, b)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[b]
    rec = np.empty((R, record_bytes(L)), dtype=np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
    v = ridx.copy()
    for dpos in range(8):
        rec[:, 9 - dpos] = (v % np.uint64(10)).astype(np.uint8) + ord("0")
        v //= np.uint64(10)
    rec[:, 10] = 10
    rec[:, 11:11 + L] = letters
    rec[:, 11 + L] = 10
    rec[:, 12 + L] = ord("+")
    rec[:, 13 + L] = 10
    rec[:, 14 + L:14 + 2 * L] = ord("I")
    rec[:, 14 + 2 * L] = 10
    return rec.tobytes()
