"""-m gpu parity tests of the tensor-core scan (csrc/mmascan.cuh: tcgen05.mma kind::i8 on rows
expanded bit -> int8 in shared memory) through the C ABI against the CPU oracle.

Bar: bit-exact scores and row lists under the (score, row) tie-break, identical to the
integer-pipe kernel on every shape.  vl_set_scan_path forces the kernel: 2 = one CTA per query
tile, 3 = CTA pairs (cta_group::2); both metrics then run as 4-bit-float contractions (kind::mxf4, 224-row
tiles for pairs; XORSUM with its bit weights as block scales), 5 / 6 ask for the int8 form (kind::i8, 256-row tiles)."""
import numpy as np
import pytest

from oracle import cscan, synth

pytestmark = pytest.mark.gpu

HAMMING, XORSUM = 0, 1
PATHS = [2, 3, 5, 6]   # 2 / 3: e2m1 (kind::mxf4; XORSUM with block-scale weights, int8 at W = 32); 5 / 6: int8 (kind::i8)


def _check(c, corpus, queries, k, mask_words=None, base_row=0, metric=HAMMING):
    s, r = c.search(queries, k, metric, mask_words)
    es, er = cscan.search(queries, corpus, k, metric, mask_words, base_row)
    np.testing.assert_array_equal(s, es)
    np.testing.assert_array_equal(r, er)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("metric", [HAMMING, XORSUM])
@pytest.mark.parametrize("width", [128, 64, 32])
@pytest.mark.parametrize("nq", [1, 100, 128, 129, 300])
def test_topk_query_tilings(gpu, path, metric, width, nq):
    n, k = 20011, 10   # ragged: not a multiple of the 128 / 256-row tile
    corpus = synth.corpus_rows(5, 0, n, width)
    queries = synth.uniform_queries(6, nq, width)
    queries[0] = corpus[n - 1]            # exact hit on the last row
    with gpu.Corpus(width) as c:
        c.append(corpus)
        c.set_scan_path(path)
        _check(c, corpus, queries, k, metric=metric)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("width", [128, 64, 32])
def test_xorsum_extreme_scores_and_weights(gpu, path, width):
    """XORSUM on the tensor cores = a +-weight contraction with bit 7 counted as 64 + 64: rows that differ
    from the query in exactly one bit position per byte must score that bit's weight times W, the
    complement 255 W, the query itself 0."""
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, size=(1, width), dtype=np.uint8)
    rows = [q[0] ^ np.uint8(1 << b) for b in range(8)] + [q[0] ^ np.uint8(0xFF), q[0].copy(), q[0] ^ np.uint8(0x81)]
    corpus = np.concatenate([np.stack(rows), synth.corpus_rows(31, 0, 3000, width)])
    with gpu.Corpus(width) as c:
        c.append(corpus)
        c.set_scan_path(path)
        got = {int(r): int(s) for s, r in zip(*[a[0] for a in c.search(np.repeat(q, 70, 0), 16, XORSUM)])}
        assert got[9] == 0 and [got[b] for b in range(6)] == [width << b for b in range(6)]
        _check(c, corpus, np.repeat(q, 3, 0), len(corpus) if len(corpus) <= 16 else 16, metric=XORSUM)
        full = c.scores(q[0], XORSUM)
        assert full[8] == 255 * width and full[10] == 0x81 * width and full[7] == 128 * width


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("k", [1, 5, 16])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 255, 257, 5000])
def test_k_and_tiny_corpora(gpu, path, k, n):
    corpus = synth.corpus_rows(7, 0, n, 128)
    queries = synth.uniform_queries(8, 70, 128)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        c.set_scan_path(path)
        _check(c, corpus, queries, k)     # k > n: sentinel padded


@pytest.mark.parametrize("path", PATHS)
def test_duplicates_ties_break_on_row(gpu, path):
    n = 60_000
    corpus = synth.corpus_rows(9, 0, n, 128, period=500)   # every row 120 times: ties everywhere
    queries = synth.uniform_queries(10, 140, 128)
    queries[3] = corpus[17]
    with gpu.Corpus(128) as c:
        c.append(corpus)
        c.set_scan_path(path)
        _check(c, corpus, queries, 10)
        _check(c, corpus, queries, 16)
        _check(c, corpus, queries, 10, metric=XORSUM)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("sel", [0.01, 0.5])
def test_filter_mask_and_tombstones(gpu, path, sel):
    n, w = 150_003, 128
    corpus = synth.corpus_rows(11, 0, n, w)
    queries, src = synth.clustered_queries(12, 0, n, 200, w)
    valid = synth.filter_valid(13, 0, n, sel)
    dead = np.unique(np.concatenate([src[:50], np.arange(0, n, 977)])).astype(np.uint64)
    live = np.ones(n, dtype=bool)
    live[dead.astype(np.int64)] = False
    with gpu.Corpus(w) as c:
        c.append(corpus)
        c.tombstone(dead)
        c.set_scan_path(path)
        s, r = c.search(queries, 10, HAMMING, synth.pack_mask(valid))
        es, er = cscan.search(queries, corpus, 10, HAMMING, synth.pack_mask(valid & live))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(r, er)
        s, r = c.search(queries, 10, HAMMING)
        es, er = cscan.search(queries, corpus, 10, HAMMING, synth.pack_mask(live))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(r, er)
        s, r = c.search(queries, 10, XORSUM, synth.pack_mask(valid))
        es, er = cscan.search(queries, corpus, 10, XORSUM, synth.pack_mask(valid & live))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(r, er)


@pytest.mark.parametrize("path", PATHS)
def test_matches_integer_pipe_kernel_and_base_row(gpu, path):
    n, w, nq = 300_000, 128, 1024
    base = 5_000_000_000
    queries = synth.uniform_queries(15, nq, w)
    with gpu.Corpus(w, base_row=base) as c:
        c.fill_synthetic(14, 0, n)
        c.set_scan_path(1)
        s1, r1 = c.search(queries, 10, HAMMING)
        c.set_scan_path(path)
        s2, r2 = c.search(queries, 10, HAMMING)
        assert r1.min() >= base
        np.testing.assert_array_equal(s1, s2)
        np.testing.assert_array_equal(r1, r2)
        for splits in (1, 3, 40):            # result must not depend on how rows are split over CTAs
            c.set_tuning(row_splits=splits)
            s3, r3 = c.search(queries[:200], 10, HAMMING)
            np.testing.assert_array_equal(s3, s1[:200])
            np.testing.assert_array_equal(r3, r1[:200])


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("metric", [HAMMING, XORSUM])
@pytest.mark.parametrize("k", [100, 1000])
def test_large_k_collect_pass_on_tensor_cores(gpu, path, metric, k):
    n = 200_000
    corpus = synth.corpus_rows(16, 0, n, 128)
    queries = synth.uniform_queries(17, 130, 128)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        c.set_scan_path(path)
        _check(c, corpus, queries, k, metric=metric)


def test_automatic_dispatch_is_bit_identical(gpu):
    """auto mode: Q >= 24 (16 at W <= 64) with k <= 16 goes to the tensor-core kernel (both metrics),
    everything else to the integer-pipe kernel."""
    n = 50_000
    corpus = synth.corpus_rows(18, 0, n, 64)
    queries = synth.uniform_queries(19, 260, 64)
    with gpu.Corpus(64) as c:
        c.append(corpus)
        for nq in (15, 16, 63, 64, 260):
            _check(c, corpus, queries[:nq], 10)
            _check(c, corpus, queries[:nq], 10, metric=XORSUM)


@pytest.mark.parametrize("k", [17, 24, 32])
@pytest.mark.parametrize("period", [0, 700])
@pytest.mark.parametrize("width,metric", [(128, HAMMING), (64, XORSUM), (32, HAMMING)])
def test_k_17_to_32_batches_take_the_collect_path(gpu, k, period, width, metric):
    """The tensor-core kernel's register lists end at k = 16; a batch with 16 < k <= 32 runs sampled
    threshold + collect + select on the tensor cores (search_host.cuh::large_k_min) instead of the
    integer-pipe lists.  period=700 floods the candidate buffer with ties: the gated exact fallback answers.
    Same result either way, and the same as the integer-pipe kernel."""
    n = 150_000
    corpus = synth.corpus_rows(41, 0, n, width, period)
    queries = synth.uniform_queries(42, 70, width)
    queries[0] = corpus[n - 3]
    valid = synth.filter_valid(3, 0, n, 0.1)
    with gpu.Corpus(width) as c:
        c.append(corpus)
        _check(c, corpus, queries, k, metric=metric)
        _check(c, corpus, queries, k, synth.pack_mask(valid), metric=metric)
        auto = c.search(queries, k, metric)
        c.set_scan_path(1)                       # integer-pipe lists
        lists = c.search(queries, k, metric)
    np.testing.assert_array_equal(auto[0], lists[0])
    np.testing.assert_array_equal(auto[1], lists[1])


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("nq", [128, 1024])
def test_many_tiles_per_cta_repeated(gpu, path, nq):
    """Regression: 8+ tiles per CTA, launched repeatedly.  The packed-row box used to be handed back to the TMA
    producer while the expanders' shared-memory loads of it were still in flight (an mbarrier arrive does not
    wait for them): with the fast e2m1 pipeline rows of tile i then carried 16-byte chunks of tile i + 2 in
    most launches of this shape.  Every launch must be bit-exact."""
    n, w = 150_003, 128
    corpus = synth.corpus_rows(11, 0, n, w)
    queries, _ = synth.clustered_queries(12, 0, n, nq, w)
    es, er = cscan.search(queries, corpus, 10, HAMMING)
    with gpu.Corpus(w) as c:
        c.append(corpus)
        c.set_scan_path(path)
        for _ in range(12):
            s, r = c.search(queries, 10, HAMMING)
            np.testing.assert_array_equal(s, es)
            np.testing.assert_array_equal(r, er)
