"""-m gpu parity tests: the CUDA scan + top-k + merge through the C ABI vs the CPU oracle.

Bar: bit-exact scores and row lists under the (score, row) tie-break.
"""
import numpy as np
import pytest

from oracle import cscan, search as osearch, synth

pytestmark = pytest.mark.gpu

HAMMING, XORSUM = 0, 1


def _oracle(corpus, queries, k, metric, mask_words=None, base_row=0):
    return cscan.search(queries, corpus, k, metric, mask_words, base_row)


def _check(c, corpus, queries, k, metric, mask_words=None, base_row=0):
    s, r = c.search(queries, k, metric, mask_words)
    es, er = _oracle(corpus, queries, k, metric, mask_words, base_row)
    np.testing.assert_array_equal(s, es)
    np.testing.assert_array_equal(r, er)


@pytest.mark.parametrize("width", [32, 64, 128])
@pytest.mark.parametrize("metric", [HAMMING, XORSUM])
def test_full_score_vector_bit_exact(gpu, width, metric):
    n = 5000
    corpus = synth.corpus_rows(3, 0, n, width)
    q = synth.uniform_queries(4, 1, width)[0]
    with gpu.Corpus(width) as c:
        c.append(corpus)
        got = c.scores(q, metric)
    np.testing.assert_array_equal(got, cscan.scores(q, corpus, metric))
    # and the C oracle agrees with the numpy restatement of the reference lines
    np.testing.assert_array_equal(got.astype(np.int64), osearch.scores(q, corpus, metric))


@pytest.mark.parametrize("width", [32, 64, 128])
@pytest.mark.parametrize("metric", [HAMMING, XORSUM])
@pytest.mark.parametrize("nq", [1, 2, 3, 5, 8, 13, 24, 40, 64, 65, 130])
def test_topk_all_query_tilings(gpu, width, metric, nq):
    n, k = 20011, 10   # ragged: not a multiple of the 256-row tile
    corpus = synth.corpus_rows(5, 0, n, width)
    queries = synth.uniform_queries(6, nq, width)
    queries[0] = corpus[n - 1]            # exact hit on the last row
    with gpu.Corpus(width) as c:
        c.append(corpus)
        _check(c, corpus, queries, k, metric)


@pytest.mark.parametrize("k", [1, 5, 10, 32, 33, 100, 128])
def test_k_values(gpu, k):
    n = 9000
    corpus = synth.corpus_rows(7, 0, n, 128)
    queries = synth.uniform_queries(8, 9, 128)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        _check(c, corpus, queries, k, HAMMING)
        _check(c, corpus, queries[:2], k, XORSUM)


@pytest.mark.parametrize("k", [129, 300, 1000])
def test_large_k_multipass(gpu, k):
    n = 6000
    corpus = synth.corpus_rows(9, 0, n, 128, period=500)   # duplicates: ties straddle pass boundaries
    queries = synth.uniform_queries(10, 5, 128)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        _check(c, corpus, queries, k, HAMMING)


@pytest.mark.parametrize("k", [33, 100, 300, 1000, 2048])
@pytest.mark.parametrize("period", [0, 700])
def test_large_k_sampled_path_and_fallback(gpu, k, period):
    """k > 32 on a corpus larger than the candidate buffer: sampled threshold + collect + select.
    period=700 makes every score value shared by ~290 rows, so buffers overflow and the exact
    multi-pass fallback must answer -- the result is the same either way."""
    n = 200_000
    corpus = synth.corpus_rows(31, 0, n, 128, period)
    queries = synth.uniform_queries(32, 6, 128)
    queries[0] = corpus[777]
    valid = synth.filter_valid(2, 0, n, 0.5)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        _check(c, corpus, queries, k, HAMMING)
        _check(c, corpus, queries[:2], k, XORSUM)
        _check(c, corpus, queries[:3], k, HAMMING, synth.pack_mask(valid))
        fast = c.search(queries, k, HAMMING)
        c.set_tuning(stages=-1)                     # force the multi-pass cursor path
        slow = c.search(queries, k, HAMMING)
        c.set_tuning()
    np.testing.assert_array_equal(fast[0], slow[0])
    np.testing.assert_array_equal(fast[1], slow[1])


def test_duplicates_tie_break(gpu):
    # distribution D: every score value is shared by n/period rows
    n, period = 30000, 97
    corpus = synth.corpus_rows(11, 0, n, 128, period)
    queries = synth.corpus_rows(11, 0, 4, 128, period)   # exact copies of rows 0..3
    with gpu.Corpus(128) as c:
        c.append(corpus)
        s, r = c.search(queries, 10, HAMMING)
        _check(c, corpus, queries, 10, HAMMING)
    assert (s == 0).all()
    np.testing.assert_array_equal(r[2], 2 + period * np.arange(10, dtype=np.uint64))


def test_clustered_queries_find_their_source(gpu):
    n = 50000
    corpus = synth.corpus_rows(13, 0, n, 128)
    queries, src = synth.clustered_queries(14, 13, n, 16, 128, flip_p=0.10)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        s, r = c.search(queries, 5, HAMMING)
        _check(c, corpus, queries, 5, HAMMING)
    np.testing.assert_array_equal(r[:, 0].astype(np.int64), src)


def test_k_larger_than_corpus_pads_with_sentinels(gpu):
    corpus = synth.corpus_rows(15, 0, 7, 64)
    queries = synth.uniform_queries(16, 3, 64)
    with gpu.Corpus(64) as c:
        c.append(corpus)
        s, r = c.search(queries, 12, XORSUM)
        _check(c, corpus, queries, 12, XORSUM)
    assert (s[:, 7:] == 0xFFFFFFFF).all() and (r[:, 7:] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()


def test_empty_corpus(gpu):
    with gpu.Corpus(128) as c:
        s, r = c.search(synth.uniform_queries(1, 2, 128), 4, HAMMING)
    assert (s == 0xFFFFFFFF).all() and (r == np.uint64(0xFFFFFFFFFFFFFFFF)).all()


@pytest.mark.parametrize("sel", [0.01, 0.10, 0.50])
@pytest.mark.parametrize("nq", [1, 20])
def test_filter_mask_in_scan(gpu, sel, nq):
    n = 40000
    corpus = synth.corpus_rows(17, 0, n, 128)
    queries = synth.uniform_queries(18, nq, 128)
    valid = synth.filter_valid(2, 0, n, sel)
    mask = synth.pack_mask(valid)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        _check(c, corpus, queries, 10, HAMMING, mask)
        s, r = c.search(queries, 10, HAMMING, mask)
    assert valid[r.astype(np.int64)].all()
    # oracle = numpy search on the filtered subset (SURVEY 8d cfg 5)
    es, er = osearch.search(queries[:1], corpus, 10, HAMMING, valid=valid)
    np.testing.assert_array_equal(r[:1], er)
    np.testing.assert_array_equal(s[:1], es)


def test_tombstones_and_append_in_place(gpu):
    n = 3000
    corpus = synth.corpus_rows(19, 0, n, 128)
    queries = synth.uniform_queries(20, 6, 128)
    with gpu.Corpus(128, capacity_rows=16) as c:      # forces several capacity doublings
        for lo in range(0, n, 700):
            assert c.append(corpus[lo:lo + 700]) == lo
        assert c.rows == n and c.live_rows == n
        np.testing.assert_array_equal(c.read(), corpus)
        _, r0 = c.search(queries, 3, HAMMING)
        dead = np.unique(r0.reshape(-1))
        c.tombstone(dead)
        c.tombstone(dead[:2])                           # double delete must not double count
        assert c.live_rows == n - len(dead)
        mask = np.ones(n, dtype=bool)
        mask[dead.astype(np.int64)] = False
        _check(c, corpus, queries, 8, HAMMING, synth.pack_mask(mask))   # same as filtering them out
        s, r = c.search(queries, 8, HAMMING)
        es, er = osearch.search(queries, corpus, 8, HAMMING, valid=mask)
        np.testing.assert_array_equal(r, er)
        np.testing.assert_array_equal(s, es)
        # appended rows after deletes keep global indices
        extra = synth.corpus_rows(21, 0, 50, 128)
        assert c.append(extra) == n
        full = np.concatenate([corpus, extra])
        mask2 = np.concatenate([mask, np.ones(50, dtype=bool)])
        s, r = c.search(queries, 8, HAMMING)
        es, er = osearch.search(queries, full, 8, HAMMING, valid=mask2)
        np.testing.assert_array_equal(r, er)


def test_base_row_offsets_results(gpu):
    n, base = 5000, 10_000_000_000
    corpus = synth.corpus_rows(23, base, n, 128)          # a shard of a bigger corpus
    queries = synth.uniform_queries(24, 3, 128)
    with gpu.Corpus(128, base_row=base) as c:
        c.fill_synthetic(23, 0, n)                         # device generator, global rows base..base+n
        np.testing.assert_array_equal(c.read(), corpus)
        _check(c, corpus, queries, 10, HAMMING, base_row=base)


def test_row_splits_do_not_change_results(gpu):
    n = 70000
    corpus = synth.corpus_rows(25, 0, n, 128, period=1000)
    queries = synth.uniform_queries(26, 70, 128)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        ref = c.search(queries, 10, HAMMING)
        for splits in (1, 3, 17):
            c.set_tuning(row_splits=splits)
            got = c.search(queries, 10, HAMMING)
            np.testing.assert_array_equal(got[0], ref[0])
            np.testing.assert_array_equal(got[1], ref[1])
        for stages in (1, 2, 5):
            c.set_tuning(row_splits=0, stages=stages)
            got = c.search(queries, 10, HAMMING)
            np.testing.assert_array_equal(got[1], ref[1])
        c.set_tuning()
    es, er = _oracle(corpus, queries, 10, HAMMING)
    np.testing.assert_array_equal(ref[1], er)


def test_device_resident_queries_and_outputs(gpu):
    torch = pytest.importorskip("torch")
    n = 12000
    corpus = synth.corpus_rows(27, 0, n, 128)
    queries = synth.uniform_queries(28, 33, 128)
    dq = torch.from_numpy(queries).cuda()
    ds = torch.empty((33, 10), dtype=torch.int32, device="cuda")
    dr = torch.empty((33, 10), dtype=torch.int64, device="cuda")
    with gpu.Corpus(128) as c:
        c.set_stream(torch.cuda.current_stream().cuda_stream)
        c.append(torch.from_numpy(corpus).cuda())
        c.search(dq, 10, HAMMING, out_scores=ds, out_rows=dr)
        torch.cuda.synchronize()
    es, er = _oracle(corpus, queries, 10, HAMMING)
    np.testing.assert_array_equal(ds.cpu().numpy().view(np.uint32), es)
    np.testing.assert_array_equal(dr.cpu().numpy().view(np.uint64), er)


def test_config2_full_size_bit_exact(gpu):
    """BASELINE configs[1] at full size: 500k x 128 B, 1024 queries, k=10, every query compared
    with the C oracle (5.1e8 pairs on the host cores), both metrics."""
    n, w, nq, k = 500_000, 128, 1024, 10
    corpus = synth.corpus_rows(0, 0, n, w)
    qc, src = synth.clustered_queries(1, 0, n, nq // 2, w)
    queries = np.ascontiguousarray(np.concatenate([qc, synth.uniform_queries(1, nq - nq // 2, w)]))
    with gpu.Corpus(w, capacity_rows=n) as c:
        c.fill_synthetic(0, 0, n)
        for metric in (HAMMING, XORSUM):
            s, r = c.search(queries, k, metric)
            es, er = cscan.search(queries, corpus, k, metric)
            np.testing.assert_array_equal(s, es)
            np.testing.assert_array_equal(r, er)
        np.testing.assert_array_equal(r[: nq // 2, 0].astype(np.int64), src)


def test_large_corpus_properties(gpu):
    """Size-independent properties on a 64M-row (8 GB) shard, where the oracle cannot follow:
    planted needles at the ends and at tile boundaries come back first, lists ascend by
    (score, row), returned scores equal a host re-score of the regenerated rows, and a second
    identical call returns identical bytes (determinism)."""
    n, w, k = 64_000_000, 128, 10
    needles = np.array([0, 255, 256, 31_999_999, 32_000_000, n - 257, n - 1])
    rng = np.random.default_rng(5)
    base = np.concatenate([synth.corpus_rows(0, int(r), 1, w) for r in needles])
    q = base ^ np.packbits(rng.random((len(needles), w * 8)) < 0.05, axis=1)
    q = np.concatenate([q, synth.uniform_queries(77, 3, w)])
    with gpu.Corpus(w, capacity_rows=n) as c:
        c.fill_synthetic(0, 0, n)
        for nq in (1, 2, len(q)):
            s, r = c.search(q[:nq], k, HAMMING)
            s2, r2 = c.search(q[:nq], k, HAMMING)
            assert np.array_equal(s, s2) and np.array_equal(r, r2)
            np.testing.assert_array_equal(r[: min(nq, len(needles)), 0].astype(np.int64), needles[:nq])
            key = s.astype(np.float64) * 2.0 ** 40 + r.astype(np.float64)
            assert (np.diff(key, axis=1) > 0).all()
            for qi in range(nq):
                regen = np.concatenate([synth.corpus_rows(0, int(x), 1, w) for x in r[qi]])
                np.testing.assert_array_equal(osearch.hamming_scores(q[qi], regen).astype(np.uint32), s[qi])
        # the k-th score is a true bound: an independent device score vector has exactly the
        # returned number of rows strictly below it
        full = c.scores(q[0], HAMMING, 0, 4_000_000)
        assert (full < s[0, -1]).sum() <= k


def test_random_differential(gpu):
    """Seeded random shapes against the oracle: corpus size, query count, k, width, metric, masks,
    tombstones and duplicates all vary; every case must match bit for bit."""
    rng = np.random.default_rng(20240607)
    for case in range(60):
        w = int(rng.choice([32, 64, 128]))
        n = int(rng.choice([1, 2, 31, 32, 33, 255, 256, 257, 511, 1000, 4097, 9999, 30000]))
        nq = int(rng.choice([1, 2, 3, 4, 7, 8, 9, 16, 17, 31, 33, 64, 65, 100]))
        k = int(rng.choice([1, 2, 5, 10, 31, 32, 33, 40, 64, 130]))
        metric = int(rng.integers(0, 2))
        period = int(rng.choice([0, 0, 0, 7, 100]))
        corpus = synth.corpus_rows(1000 + case, 0, n, w, period)
        queries = synth.uniform_queries(2000 + case, nq, w)
        if n > 3:
            queries[0] = corpus[n - 1]
        valid = np.ones(n, dtype=bool)
        mask = None
        if rng.random() < 0.5:
            valid &= rng.random(n) < rng.choice([0.02, 0.3, 0.9])
            mask = synth.pack_mask(valid)
        with gpu.Corpus(w) as c:
            c.append(corpus)
            if rng.random() < 0.4 and n > 4:
                dead = rng.choice(n, size=max(1, n // 7), replace=False)
                c.tombstone(dead)
                valid[dead] = False
            s, r = c.search(queries, k, metric, mask)
        es, er = cscan.search(queries, corpus, k, metric, synth.pack_mask(valid))
        assert np.array_equal(s, es) and np.array_equal(r, er), (case, w, n, nq, k, metric, period)


def test_very_large_query_batch(gpu):
    """17k queries in one call (past the size where the candidate histogram is skipped): still
    bit-exact, both metrics."""
    n, w, nq = 3000, 128, 17_000
    corpus = synth.corpus_rows(61, 0, n, w, period=700)
    queries = synth.uniform_queries(62, nq, w)
    with gpu.Corpus(w) as c:
        c.append(corpus)
        for metric, k in ((HAMMING, 5), (XORSUM, 32)):
            _check(c, corpus, queries, k, metric)


@pytest.mark.parametrize("period", [0, 700])
@pytest.mark.parametrize("nq,path", [(5, 0), (130, 0), (130, 1)])
def test_large_k_is_enqueue_only_graph_capturable(gpu, period, nq, path):
    """k > 32 with device pointers only never synchronises with the host: the whole search --
    sampled pass, collect pass, select, and the exact multi-pass fallback behind its device-side gate
    (opened here by period=700: mass ties overflow the candidate buffers) -- is captured into a CUDA
    graph and replayed.  A host read-back inside vl_search would make the capture fail."""
    import torch
    n, k, w = 200_000, 100, 128
    corpus = synth.corpus_rows(11, 0, n, w, period=period)
    queries = synth.uniform_queries(12, nq, w)
    es, er = cscan.search(queries, corpus, k, HAMMING)
    stream = torch.cuda.Stream()
    with gpu.Corpus(w) as c:
        c.append(corpus)
        c.set_scan_path(path)
        c.set_stream(stream.cuda_stream)
        dq = torch.from_numpy(queries).cuda()
        ds = torch.zeros((nq, k), dtype=torch.int32, device="cuda")
        dr = torch.zeros((nq, k), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        c.search(dq, k, HAMMING, None, ds, dr)            # warm-up: scratch buffers get allocated here
        stream.synchronize()
        np.testing.assert_array_equal(ds.cpu().numpy().view(np.uint32), es)
        ds.zero_()
        dr.zero_()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
            c.search(dq, k, HAMMING, None, ds, dr)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        np.testing.assert_array_equal(ds.cpu().numpy().view(np.uint32), es)
        np.testing.assert_array_equal(dr.cpu().numpy().view(np.uint64), er)
        c.set_stream(None)
        del g


@pytest.mark.parametrize("width", [32, 64, 128])
def test_search_embeddings_quantises_like_packbits(gpu, width):
    """vl_search_embeddings = sign-quantise (first 8W features, x > 0, MSB-first: vlite/model.py:39-42)
    + vl_search, for host and device float inputs, with more features than the rows hold bits."""
    import torch
    n, nq, dims = 30_000, 9, width * 8 + 64
    corpus = synth.corpus_rows(21, 0, n, width)
    x = np.random.default_rng(5).standard_normal((nq, dims)).astype(np.float32)
    x[0, : width * 8] = np.unpackbits(corpus[123]).astype(np.float32) * 2 - 1      # an exact hit
    x[1, 3] = 0.0                                                                  # 0 is not > 0
    q = np.packbits(x[:, : width * 8] > 0, axis=1)
    with gpu.Corpus(width) as c:
        c.append(corpus)
        for metric in (HAMMING, XORSUM):
            es, er = cscan.search(q, corpus, 7, metric)
            s, r = c.search_embeddings(x, 7, metric)
            np.testing.assert_array_equal(s, es)
            np.testing.assert_array_equal(r, er)
            ds = torch.empty((nq, 7), dtype=torch.int32, device="cuda")
            dr = torch.empty((nq, 7), dtype=torch.int64, device="cuda")
            c.search_embeddings(torch.from_numpy(x).cuda(), 7, metric, None, ds, dr)
            torch.cuda.synchronize()
            np.testing.assert_array_equal(ds.cpu().numpy().view(np.uint32), es)
            np.testing.assert_array_equal(dr.cpu().numpy().view(np.uint64), er)
        assert r[0, 0] == 123 and s[0, 0] == 0
        valid = synth.filter_valid(22, 0, n, 0.2)
        s, r = c.search_embeddings(x[:1], 5, HAMMING, synth.pack_mask(valid))
        es, er = cscan.search(q[:1], corpus, 5, HAMMING, synth.pack_mask(valid))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(r, er)
        with pytest.raises(gpu.VliteB200Error):
            c.search_embeddings(x[:, : width * 8 - 32], 5, HAMMING)                 # fewer features than bits


@pytest.mark.parametrize("nq", [1, 2])
@pytest.mark.parametrize("n", [300, 5000, 200_000])
def test_fused_merge_matches_separate_merge_kernels(gpu, nq, n):
    """1-2 queries on a small corpus finish inside the scan launch (last-CTA-done ticket); the result
    is identical to the scan + merge-kernel path and to the oracle, call after call (the ticket resets)."""
    corpus = synth.corpus_rows(23, 0, n, 128, period=97 if n == 5000 else 0)
    q = synth.uniform_queries(24, nq, 128)
    q[0] = corpus[n - 1]
    with gpu.Corpus(128) as c:
        c.append(corpus)
        for k in (1, 10, 32):
            es, er = cscan.search(q, corpus, k, HAMMING)
            for rep in range(3):
                s, r = c.search(q, k, HAMMING)
                np.testing.assert_array_equal(s, es)
                np.testing.assert_array_equal(r, er)
            c.set_tuning(stages=-2)                       # separate merge kernels
            s, r = c.search(q, k, HAMMING)
            np.testing.assert_array_equal(s, es)
            np.testing.assert_array_equal(r, er)
            c.set_tuning()
            for splits in (1, 7, 300):
                c.set_tuning(row_splits=splits)
                s, r = c.search(q, k, XORSUM)
                es2, er2 = cscan.search(q, corpus, k, XORSUM)
                np.testing.assert_array_equal(s, es2)
                np.testing.assert_array_equal(r, er2)
            c.set_tuning()
