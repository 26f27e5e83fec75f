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
    """auto mode: Q >= 16 with k <= 16 goes to the tensor-core kernel (both metrics; XOR-sum on big shards later),
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
@pytest.mark.parametrize("nq,metric", [(128, HAMMING), (1024, HAMMING), (1024, XORSUM)])
def test_many_tiles_per_cta_repeated(gpu, path, nq, metric):
    """Regression: 8+ tiles per CTA, launched repeatedly.  The packed-row box used to be handed back to the TMA
    producer while the expanders' shared-memory loads of it were still in flight (an mbarrier arrive does not
    wait for them): with the fast e2m1 pipeline rows of tile i then carried 16-byte chunks of tile i + 2 in
    most launches of this shape.  Every launch must be bit-exact."""
    n, w = 150_003, 128
    corpus = synth.corpus_rows(11, 0, n, w)
    queries, _ = synth.clustered_queries(12, 0, n, nq, w)
    es, er = cscan.search(queries, corpus, 10, metric)
    with gpu.Corpus(w) as c:
        c.append(corpus)
        c.set_scan_path(path)
        for _ in range(12):
            s, r = c.search(queries, 10, metric)
            np.testing.assert_array_equal(s, es)
            np.testing.assert_array_equal(r, er)


@pytest.mark.parametrize("k", [1000, 2048])
def test_large_k_batch_does_not_fall_back(gpu, k):
    """Regression: the sampled threshold of the large-k path used to be the 16th best of its sample; one query in
    ~7000 then has fewer than k rows at or below it, and ONE such query in a batch sends ALL of it to the exact
    multi-pass fallback (k = 2048 at 500k x 1024 queries: 0.8 -> 166 ms).  The tensor-core pass now takes the 64th
    best of the merged lists: three different batches must all stay on the fast path (a 20x margin on time) and
    agree with the oracle."""
    torch = pytest.importorskip("torch")
    n, w, nq = 500_000, 128, 1024
    host = synth.corpus_rows(0, 0, n, w)
    with gpu.Corpus(w, capacity_rows=n) as c:
        c.fill_synthetic(0, 0, n)
        c.set_timing(True)
        ds = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        dr = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        for seed in (1, 2, 3):
            q = synth.uniform_queries(seed, nq, w)
            dq = torch.from_numpy(q).cuda()
            for _ in range(2):
                c.search(dq, k, HAMMING, None, ds, dr)
            total_ms, _ = c.last_timing()
            assert total_ms < 20.0, f"k={k} seed={seed}: {total_ms:.1f} ms -- the exact fallback ran"
            sub = np.r_[0:4, nq - 4:nq]
            es, er = cscan.search(q[sub], host, k, HAMMING)
            np.testing.assert_array_equal(ds.cpu().numpy().view(np.uint32)[sub], es)
            np.testing.assert_array_equal(dr.cpu().numpy().view(np.uint64)[sub], er)


@pytest.mark.parametrize("nq,k", [(1024, 1000), (4096, 256)])
def test_large_k_many_tiles_per_cta_stays_on_the_fast_path(gpu, nq, k):
    """The sampled pass merges its 16-entry lists to the 64 best of their union: the bound the CTAs share must then
    stand for rank 64 as well (with the rank-16 bound, lists of CTAs that visit many sampled tiles keep only rows
    in the global top 16, the 64th of the union is junk, every buffer overflows and the batch falls back: 245 ms
    instead of 5 ms at 4M rows).  2M rows = 20+ sampled tiles per CTA; results checked by size-independent
    properties (sorted, unique rows, scores re-computed on the host from the generator)."""
    torch = pytest.importorskip("torch")
    n, w = 2_000_000, 128
    q = synth.uniform_queries(5, nq, w)
    with gpu.Corpus(w, capacity_rows=n) as c:
        c.fill_synthetic(0, 0, n)
        c.set_timing(True)
        dq = torch.from_numpy(q).cuda()
        ds = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        dr = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        for _ in range(2):
            c.search(dq, k, HAMMING, None, ds, dr)
        total_ms, _ = c.last_timing()
        assert total_ms < 60.0, f"{total_ms:.1f} ms -- the exact fallback ran"
        s = ds.cpu().numpy().view(np.uint32)
        r = dr.cpu().numpy().view(np.uint64)
    for qi in (0, nq // 2, nq - 1):
        rows = r[qi].astype(np.int64)
        assert len(np.unique(rows)) == k and rows.max() < n
        key = (s[qi].astype(np.uint64) << np.uint64(32)) | r[qi]
        assert np.all(np.diff(key.astype(np.float64)) >= 0) and np.all(key[1:] > key[:-1])
        sub = np.stack([synth.corpus_rows(0, int(x), 1, w)[0] for x in rows[:: max(1, k // 50)]])
        true = np.unpackbits(sub ^ q[qi], axis=1).sum(axis=1)
        np.testing.assert_array_equal(true, s[qi][:: max(1, k // 50)])



@pytest.mark.parametrize("nq", [2048 + 77, 4096, 5000])
@pytest.mark.parametrize("metric,k", [(HAMMING, 10), (XORSUM, 10), (HAMMING, 100)])
def test_big_batches_go_out_in_several_launches(gpu, nq, metric, k):
    """8+ query tiles: the batch is scanned in several launches of fewer tiles each (ScanParams::q_offset) so that
    tiles x row splits fills the SMs (4096 queries: 2 x (8 pairs x 9 splits) instead of 16 x 4).  Every query,
    list mode and collect mode, both metrics, ragged last launch."""
    n, w = 60_000, 128
    corpus = synth.corpus_rows(71, 0, n, w)
    queries = synth.uniform_queries(72, nq, w)
    queries[0] = corpus[n - 1]
    queries[nq - 1] = corpus[0]
    with gpu.Corpus(w) as c:
        c.append(corpus)
        _check(c, corpus, queries, k, metric=metric)
