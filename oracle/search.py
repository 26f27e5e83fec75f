"""numpy restatement of the reference's search arithmetic.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows (paths relative to
/root/reference).  Corpus rows and queries are *raw ubinary* bytes
(``np.packbits`` output, MSB-first within a byte) unless a function says
otherwise; ``to_ref_encoding`` / ``from_ref_encoding`` convert to and from the
reference's stored ``int8(u8) - 128`` form (vlite/model.py:44).

Canonical result order everywhere: ascending (score, row index).  The
reference's ``torch.topk`` leaves tie order unspecified (vlite/model.py:85), so
parity against the live torch path is "same score list, and same index list
after canonicalising tie groups"; parity against this oracle is bit-exact.
"""
from __future__ import annotations

import numpy as np

HAMMING = 0
XORSUM = 1

SENTINEL_SCORE = np.uint32(0xFFFFFFFF)
SENTINEL_ROW = np.uint64(0xFFFFFFFFFFFFFFFF)


# --------------------------------------------------------------------------
# encodings (vlite/model.py:37-46)
# --------------------------------------------------------------------------
def quantize_binary(x: np.ndarray) -> np.ndarray:
    """float embeddings (B, D) -> raw ubinary bytes (B, D/8).

    Follows vlite/model.py:41 and :44: ``(x > 0)`` then ``np.packbits(axis=-1)``
    (MSB-first).  The reference additionally slices ``[:, :512]`` (model.py:39)
    -- callers do that; this function is width-agnostic.
    """
    bits = (np.asarray(x) > 0).astype(np.uint8)
    return np.packbits(bits, axis=-1)


def to_ref_encoding(u8: np.ndarray) -> np.ndarray:
    """raw ubinary bytes -> the reference's stored values.

    vlite/model.py:44 computes ``packbits(...).astype(np.int8) - 128``.  Under the
    numpy 1.x the reference pins this value-casts to int16 with range
    [-256, -1]:  int8(u) = u (u < 128) or u - 256 (u >= 128), minus 128.
    Equivalently ``(u ^ 0x80) - 256``.  (Under numpy >= 2 the same line raises
    OverflowError; the golden vectors pin the numpy 1.x values via the notebook
    outputs quoted in SURVEY.md F3.)
    """
    u = np.asarray(u8).astype(np.uint8)
    return (u ^ np.uint8(0x80)).astype(np.int16) - np.int16(256)


def from_ref_encoding(s: np.ndarray) -> np.ndarray:
    """Inverse of :func:`to_ref_encoding`; also accepts values already in
    [0, 255] (raw ubinary kept as int/float), which pass through unchanged."""
    v = np.asarray(s)
    vi = np.rint(v).astype(np.int64)
    neg = vi < 0
    out = np.where(neg, ((vi + 256) ^ 0x80), vi)
    if ((out < 0) | (out > 255)).any():
        raise ValueError("value outside the reference encoding range [-256, 255]")
    return out.astype(np.uint8)


# --------------------------------------------------------------------------
# scores
# --------------------------------------------------------------------------
def xorsum_scores(q: np.ndarray, corpus: np.ndarray) -> np.ndarray:
    """Live reference score, vlite/model.py:82:
    ``torch.sum(q.int() ^ E.int(), dim=1)`` -- the XOR is taken on the integer
    *byte values*, so bit b of every byte carries weight 2**b.
    Works on raw ubinary bytes or on the reference encoding alike (the XOR of
    two reference-encoded values equals the XOR of their raw bytes, F3).
    q: (W,) or (Q, W); corpus: (N, W).  Returns int64 (N,) or (Q, N)."""
    qq = np.asarray(q).astype(np.int64)
    ee = np.asarray(corpus).astype(np.int64)
    if qq.ndim == 1:
        return np.sum(qq[None, :] ^ ee, axis=1)
    return np.stack([np.sum(row[None, :] ^ ee, axis=1) for row in qq])


def hamming_scores_unpacked(q: np.ndarray, corpus: np.ndarray) -> np.ndarray:
    """True Hamming distance the way the reference's dead numpy code states it,
    vlite/model.py:71-74 (``np.sum(e1 != e2, axis=1)``) and vlite/index.py:30
    (``np.count_nonzero(bv != q, axis=1)``), applied to *unpacked* 0/1 bit arrays.
    q: (W,) raw ubinary; corpus: (N, W) raw ubinary.  int64 (N,)."""
    qb = np.unpackbits(np.asarray(q, dtype=np.uint8))[None, :]
    eb = np.unpackbits(np.asarray(corpus, dtype=np.uint8), axis=1)
    return np.sum(qb != eb, axis=1).astype(np.int64)


def hamming_scores(q: np.ndarray, corpus: np.ndarray) -> np.ndarray:
    """popcount(q XOR e) -- identical to :func:`hamming_scores_unpacked`
    (asserted in tests/test_oracle.py) but fast enough for 1e6+ rows.
    q: (W,) or (Q, W).  Returns int64 (N,) or (Q, N)."""
    qq = np.asarray(q, dtype=np.uint8)
    ee = np.ascontiguousarray(corpus, dtype=np.uint8)
    w = ee.shape[1]
    if w % 8 == 0:
        ev = ee.view(np.uint64)
        def one(row):
            rv = np.ascontiguousarray(row).view(np.uint64)
            return np.bitwise_count(ev ^ rv[None, :]).sum(axis=1, dtype=np.int64)
    else:
        def one(row):
            return np.bitwise_count(ee ^ row[None, :]).sum(axis=1, dtype=np.int64)
    if qq.ndim == 1:
        return one(qq)
    return np.stack([one(r) for r in qq])


def scores(q, corpus, metric: int) -> np.ndarray:
    if metric == HAMMING:
        return hamming_scores(q, corpus)
    if metric == XORSUM:
        return xorsum_scores(q, corpus)
    raise ValueError(f"unknown metric {metric}")


# --------------------------------------------------------------------------
# top-k
# --------------------------------------------------------------------------
def canonical_topk(score_vec: np.ndarray, k: int, valid: np.ndarray | None = None):
    """k smallest by ascending (score, row index) -- the canonical order.
    ``k`` is clamped to the number of (valid) rows as vlite/model.py:84 does.
    Returns (rows int64[k'], scores int64[k'])."""
    s = np.asarray(score_vec).astype(np.int64)
    rows = np.arange(s.shape[0], dtype=np.int64)
    if valid is not None:
        keep = np.asarray(valid, dtype=bool)
        s, rows = s[keep], rows[keep]
    k = min(int(k), s.shape[0])
    order = np.argsort(s, kind="stable")[:k]     # stable => ties by row index
    return rows[order], s[order]


def search(queries: np.ndarray, corpus: np.ndarray, k: int, metric: int = HAMMING,
           valid: np.ndarray | None = None, base_row: int = 0):
    """Batched canonical search.  Returns (scores uint32 (Q,k), rows uint64 (Q,k))
    padded with SENTINEL_* where fewer than k rows are valid -- the exact
    layout ``vl_search`` writes."""
    qs = np.atleast_2d(np.asarray(queries, dtype=np.uint8))
    out_s = np.full((qs.shape[0], k), SENTINEL_SCORE, dtype=np.uint32)
    out_r = np.full((qs.shape[0], k), SENTINEL_ROW, dtype=np.uint64)
    if corpus.shape[0] == 0:
        return out_s, out_r
    for i, q in enumerate(qs):
        r, s = canonical_topk(scores(q, corpus, metric), k, valid)
        out_s[i, :len(s)] = s.astype(np.uint32)
        out_r[i, :len(r)] = r.astype(np.uint64) + np.uint64(base_row)
    return out_s, out_r


def canonicalise_ties(rows: np.ndarray, scores_: np.ndarray):
    """Sort an (index, score) result list by (score, index).  Used to compare the
    reference's torch.topk output (unspecified tie order, model.py:85) with the
    canonical order: equal score lists + equal canonicalised index lists on all
    tie groups that lie wholly inside the top-k."""
    rows = np.asarray(rows).astype(np.int64)
    scores_ = np.asarray(scores_).astype(np.int64)
    order = np.lexsort((rows, scores_))
    return rows[order], scores_[order]


def merge_topk(score_lists, row_lists, k: int):
    """Final merge of per-shard candidate lists: concatenate,
    drop sentinels, sort by (score, global row), keep k, pad with sentinels.
    score_lists/row_lists: arrays shaped (L, Q, k_in)."""
    s = np.asarray(score_lists, dtype=np.uint32)
    r = np.asarray(row_lists, dtype=np.uint64)
    L, Q, kin = s.shape
    out_s = np.full((Q, k), SENTINEL_SCORE, dtype=np.uint32)
    out_r = np.full((Q, k), SENTINEL_ROW, dtype=np.uint64)
    for q in range(Q):
        ss = s[:, q, :].reshape(-1)
        rr = r[:, q, :].reshape(-1)
        keep = rr != SENTINEL_ROW
        ss, rr = ss[keep], rr[keep]
        order = np.lexsort((rr, ss))[:k]
        out_s[q, :len(order)] = ss[order]
        out_r[q, :len(order)] = rr[order]
    return out_s, out_r


# --------------------------------------------------------------------------
# reference call sites restated
# --------------------------------------------------------------------------
def ref_search_port(query_embedding: np.ndarray, embeddings: np.ndarray, top_k: int):
    """Restatement of ``EmbeddingModel.search`` (vlite/model.py:76-90) with the
    canonical tie order: XOR-sum scores, ``top_k = min(top_k, N)``, k smallest.
    Accepts what the reference accepts: q (W,) and E (N, W) in any numeric
    dtype holding the stored integer values.  Returns (indices int64, scores
    int64) like the reference."""
    sc = xorsum_scores(np.rint(np.asarray(query_embedding)).astype(np.int64),
                       np.rint(np.asarray(embeddings)).astype(np.int64))
    rows, s = canonical_topk(sc, top_k)
    return rows, s


def rank_and_filter_port(index: dict, q: np.ndarray, top_k: int, metadata=None,
                         top_k_multiplier: int = 4, reference_bug: bool = True):
    """Restatement of ``VLite.rank_and_filter`` (vlite/main.py:130-168).

    index: the reference's ``{chunk_id: {'text','metadata','binary_vector'}}``.
    Follows the reference step by step: widen k when filtering (:134-137), keep
    only rows whose vector length equals the query's (:143-144), raise
    ValueError when none (:149), search (:150), map rows to chunk ids (:156),
    post-filter on ``all(md.get(k) == v)`` (:159-164).  With
    ``reference_bug=True`` the returned scores are ``top_k_scores[:len(ids)]``
    exactly as main.py:165 does (scores NOT realigned with the surviving ids);
    with False each id keeps its own score."""
    initial = top_k * top_k_multiplier if metadata else top_k
    q = np.asarray(q).reshape(-1)
    keys = list(index.keys())
    vecs, pos = [], []
    for i, item in enumerate(index.values()):
        if len(item["binary_vector"]) == len(q):
            vecs.append(item["binary_vector"])
            pos.append(i)
    if not vecs:
        raise ValueError("No valid binary vectors found for comparison.")
    rows, sc = ref_search_port(q, np.asarray(vecs), initial)
    ids = [keys[pos[r]] for r in rows]
    if metadata:
        kept = [(cid, s) for cid, s in zip(ids, sc)
                if all(index[cid]["metadata"].get(kk) == vv for kk, vv in metadata.items())]
        kept = kept[:top_k]
        ids2 = [c for c, _ in kept]
        sc2 = sc[:len(ids2)] if reference_bug else np.asarray([s for _, s in kept], dtype=np.int64)
        return list(zip(ids2, sc2))
    return list(zip(ids, sc))


# --------------------------------------------------------------------------
# rescore -- PARITY UNPINNED (no reference implementation)
# --------------------------------------------------------------------------
def rescore(queries: np.ndarray, docs_i8: np.ndarray, cand_rows: np.ndarray, k: int):
    """float64 dot of each query with its candidates' int8 document rows, then
    top-k by DESCENDING score, ties by ascending row.  ``cand_rows`` (Q, C)
    uint64 with SENTINEL_ROW padding.  The reference has no rescore stage; its only
    dense similarity is the unused ``utils.cos_sim`` (vlite/utils.py:205-208): ``a @ b.T``
    over the norms.  The dot products here are pinned to that function's outputs on
    seeded inputs (tests/golden/rescore_golden.npz, tests/test_oracle.py); ranking by the
    raw dot (not the cosine) and the (score desc, row asc) tie-break are north_star's
    definition, not the reference's.
    Returns (scores float64 (Q,k) padded with -inf, rows uint64 (Q,k))."""
    qs = np.atleast_2d(np.asarray(queries))
    Q, C = cand_rows.shape
    out_s = np.full((Q, k), -np.inf, dtype=np.float64)
    out_r = np.full((Q, k), SENTINEL_ROW, dtype=np.uint64)
    for i in range(Q):
        rows = cand_rows[i]
        rows = rows[rows != SENTINEL_ROW]
        if len(rows) == 0:
            continue
        d = docs_i8[rows.astype(np.int64)].astype(np.float64) @ qs[i].astype(np.float64)
        order = np.lexsort((rows, -d))[:k]
        out_s[i, :len(order)] = d[order]
        out_r[i, :len(order)] = rows[order]
    return out_s, out_r
