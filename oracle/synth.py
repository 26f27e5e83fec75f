"""Counter-based synthetic data generator (numpy side).  TEST INFRASTRUCTURE ONLY.

The SAME function is implemented on the device in
``vlite_b200/csrc/synth.cuh`` (``vl_synth_fill``) so that a 1e9-row corpus never
crosses PCIe and any slice of it can be regenerated here for checking
(SURVEY.md section 8d).  All arithmetic is modulo 2**64.

    mix64(z):  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9
               z ^= z >> 27; z *= 0x94D049BB133111EB
               z ^= z >> 31
    word(seed, row, j, stride) =
        mix64((row * stride + j) * 0x9E3779B97F4A7C15
              + seed * 0xD1B54A32D192ED03 + 0x9E3779B97F4A7C15)

A corpus row of W bytes is words j = 0 .. W/8-1 with stride 16, stored
little-endian -- so the W=64 corpus is the 64-byte prefix of the W=128 one.
Distribution "D" (duplicates, mimicking the reference bench's ``chunks * N``,
tests/benchtest.py:514) replaces ``row`` by ``row % period``.
int8 document rows for the rescore step use stride 128 (1024 dims).
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_SEEDMUL = np.uint64(0xD1B54A32D192ED03)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
CORPUS_STRIDE = 16
DOC_STRIDE = 128
TAG_SALT = 0x7461675F73616C74  # "tag_salt"


def mix64(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= _M1
        z ^= z >> np.uint64(27)
        z *= _M2
        z ^= z >> np.uint64(31)
    return z


def words(seed: int, rows: np.ndarray, nwords: int, stride: int) -> np.ndarray:
    """(len(rows), nwords) uint64."""
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    j = np.arange(nwords, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        ctr = (rows * np.uint64(stride) + j) * _GOLD + np.uint64(seed) * _SEEDMUL + _GOLD
    return mix64(ctr)


def corpus_rows(seed: int, first_row: int, n: int, width: int, period: int = 0) -> np.ndarray:
    """(n, width) uint8 raw-ubinary rows [first_row, first_row + n)."""
    assert width % 8 == 0 and width // 8 <= CORPUS_STRIDE
    rows = np.arange(first_row, first_row + n, dtype=np.uint64)
    if period:
        rows = rows % np.uint64(period)
    w = words(seed, rows, width // 8, CORPUS_STRIDE)
    return np.ascontiguousarray(w).view(np.uint8).reshape(n, width)


def doc_rows_i8(seed: int, first_row: int, n: int, dims: int = 1024) -> np.ndarray:
    """(n, dims) int8 document embeddings for the rescore step."""
    assert dims % 8 == 0 and dims // 8 <= DOC_STRIDE
    rows = np.arange(first_row, first_row + n, dtype=np.uint64)
    w = words(seed, rows, dims // 8, DOC_STRIDE)
    return np.ascontiguousarray(w).view(np.int8).reshape(n, dims)


def tags_u32(seed: int, first_row: int, n: int) -> np.ndarray:
    """Per-row 32-bit tag; a row passes a filter of selectivity p iff
    tag < p * 2**32 (``filter_mask``)."""
    rows = np.arange(first_row, first_row + n, dtype=np.uint64)
    w = words(seed ^ TAG_SALT, rows, 1, 1)[:, 0]
    return (w >> np.uint64(32)).astype(np.uint32)


def selectivity_threshold(p: float) -> int:
    return min(int(round(p * 2.0 ** 32)), 2 ** 32)


def filter_valid(seed: int, first_row: int, n: int, p: float) -> np.ndarray:
    """bool (n,) -- True where the row passes the filter."""
    return tags_u32(seed, first_row, n).astype(np.uint64) < np.uint64(selectivity_threshold(p))


def pack_mask(valid: np.ndarray) -> np.ndarray:
    """bool (n,) -> uint32 bitmask words, bit (row & 31) of word (row >> 5),
    the layout ``vl_search`` takes for its filter mask."""
    v = np.asarray(valid, dtype=bool)
    n = v.shape[0]
    pad = (-n) % 32
    if pad:
        v = np.concatenate([v, np.zeros(pad, dtype=bool)])
    return np.packbits(v.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)


def clustered_queries(seed: int, corpus_seed: int, n_rows: int, nq: int, width: int,
                      flip_p: float = 0.10, period: int = 0):
    """Distribution "C": query i = corpus row r_i with each bit flipped w.p.
    flip_p.  Returns (queries (nq, width) uint8, source rows int64 (nq,))."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_rows, size=nq)
    base = np.concatenate([corpus_rows(corpus_seed, int(r), 1, width, period) for r in src])
    flips = np.packbits(rng.random((nq, width * 8)) < flip_p, axis=1)
    return base ^ flips, src.astype(np.int64)


def uniform_queries(seed: int, nq: int, width: int) -> np.ndarray:
    """Distribution "U" queries: rows of a *different* seed's corpus."""
    return corpus_rows(seed, 0, nq, width)
