"""Host-side metadata index: one row bitmap per (key, value) pair, kept up to date by
add / update / delete, so the filter mask the scan kernel reads (one bit per corpus row,
``vl_search(filter_mask=...)``) is the AND of a few word arrays -- O(N/64) per term -- instead of
the reference's per-candidate dict probes after the search (vlite/main.py:159-164) or an O(N)
walk over every row's metadata.

Matching follows the reference's predicate ``all(md.get(k) == v for k, v in metadata.items())``
(vlite/main.py:162): values are compared with ``==``, so ``1``, ``1.0`` and ``True`` select the
same rows (they are also the same dict key here).  A value that cannot be a dict key (a list, a
dict) or that is not equal to itself (NaN) is not indexed; ``mask`` then returns ``None`` for that
query and the caller falls back to checking rows one by one.  ``md.get(k) == None`` also holds for
rows that lack the key, so a per-key "has this key" bitmap is kept and a ``None`` term selects
``(k, None) | ~has(k)``.

Bits of deleted rows are cleared lazily: the device ANDs every filter mask with its own live-row
mask inside the scan, and rows are never reused, so a stale bit can never surface a result.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import numpy as np

_WORD = 64
_HAS = object()      # value slot of the per-key presence bitmaps: (key, _HAS)


def indexable(value) -> bool:
    try:
        hash(value)
    except TypeError:
        return False
    return value == value


class MetaIndex:
    def __init__(self):
        self._maps: Dict[Tuple, np.ndarray] = {}      # (key, value) -> uint64 words, bit r = row r
        self.version = 0                              # bumped by every mutation (mask cache key)
        self._cache_key = None
        self._cache_val = None

    # ---- maintenance -------------------------------------------------------------------------
    def _words(self, kv, upto_row: int) -> np.ndarray:
        need = upto_row // _WORD + 1
        m = self._maps.get(kv)
        if m is None:
            m = np.zeros(max(need, 16), dtype=np.uint64)
            self._maps[kv] = m
        elif m.shape[0] < need:
            grown = np.zeros(max(need, 2 * m.shape[0]), dtype=np.uint64)
            grown[:m.shape[0]] = m
            m = self._maps[kv] = grown
        return m

    @staticmethod
    def _terms(metadata: dict):
        """bitmaps a row with this metadata belongs to: (k, v) when v can be indexed, (k, _HAS) always"""
        for k, v in metadata.items():
            if indexable(v):
                yield (k, v)
            yield (k, _HAS)

    def set_row(self, metadata: dict, row: int):
        for kv in self._terms(metadata):
            self._words(kv, row)[row // _WORD] |= np.uint64(1 << (row % _WORD))
        self.version += 1

    def clear_row(self, metadata: dict, row: int):
        for kv in self._terms(metadata):
            m = self._maps.get(kv)
            if m is not None and row // _WORD < m.shape[0]:
                m[row // _WORD] &= np.uint64(~(1 << (row % _WORD)) & 0xFFFFFFFFFFFFFFFF)
        self.version += 1

    def set_range(self, metadata: dict, first: int, n: int):
        """rows [first, first+n) all carry ``metadata`` (bulk ingest with one shared dict)"""
        if n <= 0:
            return
        last = first + n - 1
        for kv in self._terms(metadata):
            m = self._words(kv, last)
            w0, w1 = first // _WORD, last // _WORD
            lo = (0xFFFFFFFFFFFFFFFF << (first % _WORD)) & 0xFFFFFFFFFFFFFFFF
            hi = 0xFFFFFFFFFFFFFFFF >> (_WORD - 1 - last % _WORD)
            if w0 == w1:
                m[w0] |= np.uint64(lo & hi)
            else:
                m[w0] |= np.uint64(lo)
                m[w0 + 1:w1] = np.uint64(0xFFFFFFFFFFFFFFFF)
                m[w1] |= np.uint64(hi)
        self.version += 1

    def set_rows(self, metadata: dict, rows: Iterable[int]):
        rows = np.asarray(list(rows), dtype=np.int64)
        if rows.size == 0:
            return
        for kv in self._terms(metadata):
            m = self._words(kv, int(rows.max()))
            np.bitwise_or.at(m, rows // _WORD, np.uint64(1) << (rows % _WORD).astype(np.uint64))
        self.version += 1

    def set_many(self, metadatas, first_row: int):
        """rows first_row, first_row+1, ... carry metadatas[0], metadatas[1], ...: one pass over the
        dicts, then one vectorised OR per distinct (key, value) term (bulk ingest, collection load)."""
        n = len(metadatas)
        if n == 0:
            return
        if all(md is metadatas[0] for md in metadatas):        # add(): one dict shared by every chunk
            self.set_range(metadatas[0], first_row, n)
            return
        groups = {}
        for i, md in enumerate(metadatas):
            for k, v in md.items():
                if indexable(v):
                    groups.setdefault((k, v), []).append(i)
                groups.setdefault((k, _HAS), []).append(i)
        for term, idxs in groups.items():
            rows = np.asarray(idxs, dtype=np.int64) + first_row
            m = self._words(term, int(rows[-1]))
            np.bitwise_or.at(m, rows >> 6, np.uint64(1) << (rows & 63).astype(np.uint64))
        self.version += 1

    def clear(self):
        self._maps.clear()
        self.version += 1
        self._cache_key = self._cache_val = None

    # ---- queries -----------------------------------------------------------------------------
    def mask(self, metadata: dict, n_rows: int) -> Optional[np.ndarray]:
        """uint32 words, bit r of word r//32 = row r satisfies every term; ``None`` when a term is
        not indexable.  Length (n_rows + 31) // 32 words."""
        items = list(metadata.items())
        if not all(indexable(v) for _, v in items):
            return None
        try:
            key = (frozenset(items), n_rows, self.version)
        except TypeError:
            key = None
        if key is not None and key == self._cache_key:
            return self._cache_val
        n64 = (n_rows + _WORD - 1) // _WORD
        def fitted(kv):
            m = self._maps.get(kv)
            if m is None:
                return np.zeros(n64, dtype=np.uint64)
            if m.shape[0] >= n64:
                return m[:n64].copy()
            cur = np.zeros(n64, dtype=np.uint64)
            cur[:m.shape[0]] = m
            return cur

        acc = None
        for k, v in items:
            cur = fitted((k, v))
            if v is None:                                 # md.get(k) == None: also rows without k
                cur |= ~fitted((k, _HAS))
            acc = cur if acc is None else np.bitwise_and(acc, cur, out=acc)
        if acc is None:                                   # no terms: every row passes
            acc = np.full(n64, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        if n_rows % _WORD and n64:
            acc[-1] &= np.uint64((1 << (n_rows % _WORD)) - 1)
        out = acc.view(np.uint32)[:(n_rows + 31) // 32]   # little-endian: word r//32, bit r%32
        out = np.ascontiguousarray(out)
        self._cache_key, self._cache_val = key, out
        return out

    def rows(self, metadata: dict, n_rows: int) -> Optional[np.ndarray]:
        m = self.mask(metadata, n_rows)
        if m is None:
            return None
        bits = np.unpackbits(m.view(np.uint8), bitorder="little")[:n_rows]
        return np.flatnonzero(bits).astype(np.int64)

    def __len__(self):
        return len(self._maps)
