"""-m gpu: error behaviour of the C ABI (negative code + vl_last_error, no crash, no silent fallback)."""

import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def test_argument_errors(gpu):
    corpus = synth.corpus_rows(1, 0, 100, 128)
    q = synth.uniform_queries(2, 2, 128)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        for bad_k in (0, -1, gpu.MAX_K + 1):
            with pytest.raises(gpu.VliteB200Error) as ei:
                c.search(q, bad_k)
            assert ei.value.code == gpu.EINVAL
        with pytest.raises(gpu.VliteB200Error):
            c.search(q, 5, metric=7)
        with pytest.raises(ValueError):
            c.search(synth.uniform_queries(2, 2, 64), 5)            # wrong width
        with pytest.raises(ValueError):
            c.append(synth.corpus_rows(1, 0, 3, 64))
        with pytest.raises(ValueError):
            c.search(q, 5, filter_mask=np.zeros(1, np.uint32))       # mask shorter than the corpus
        with pytest.raises(gpu.VliteB200Error):
            c.read(50, 100)
        with pytest.raises(gpu.VliteB200Error):
            c.set_row(100, corpus[0])
        with pytest.raises(gpu.VliteB200Error):
            c.last_timing()                                           # nothing timed yet
        # still healthy after all of that
        s, r = c.search(corpus[7:8], 1)
        assert int(r[0, 0]) == 7 and int(s[0, 0]) == 0
    with pytest.raises(gpu.VliteB200Error) as ei:
        gpu.Corpus(96)
    assert ei.value.code == gpu.EINVAL
    with pytest.raises(gpu.VliteB200Error):
        gpu.Corpus(128, device=99)


def test_set_row_reserve_clear_and_out_of_range_tombstones(gpu):
    corpus = synth.corpus_rows(3, 0, 600, 64)
    with gpu.Corpus(64) as c:
        c.reserve(10_000)
        rows_ptr, live_ptr = c.device_ptrs()
        c.append(corpus)
        c.reserve(5_000_000)                                          # growth maps memory: pointers do not move
        assert c.device_ptrs() == (rows_ptr, live_ptr)
        np.testing.assert_array_equal(c.read(), corpus)
        c.set_row(5, corpus[9])
        s, r = c.search(corpus[9:10], 2, gpu.HAMMING)
        assert list(r[0]) == [5, 9] and list(s[0]) == [0, 0]
        c.tombstone([5, 10_000_000, 5])                               # out-of-range and repeated rows are ignored
        assert c.live_rows == 599
        c.clear()
        assert c.rows == 0 and c.live_rows == 0
        s, r = c.search(corpus[:1], 3)
        assert (r == np.uint64(gpu.SENTINEL_ROW)).all()
        assert c.append(corpus[:10]) == 0                             # rows restart at 0 after clear
        s, r = c.search(corpus[3:4], 1)
        assert int(r[0, 0]) == 3
    c.close()
    c.close()                                                         # double close is harmless


def test_launch_counter_counts_library_kernels(gpu):
    before = gpu.launch_count()
    with gpu.Corpus(128) as c:
        c.fill_synthetic(0, 0, 1000)
        c.search(synth.uniform_queries(1, 3, 128), 4)
    assert gpu.launch_count() - before >= 4          # fill, live-mask, scan, merge


def test_out_of_memory_is_an_error_not_a_crash(gpu):
    """Asking for more rows than the device can back fails with VL_ENOMEM (one failed
    cuMemCreate, nothing half-mapped) and leaves the handle usable."""
    corpus = synth.corpus_rows(8, 0, 300, 128)
    with gpu.Corpus(128) as c:
        c.append(corpus)
        ptrs = c.device_ptrs()
        with pytest.raises(gpu.VliteB200Error) as ei:
            c.reserve(2_100_000_000)                                  # 269 GB of rows on a 180 GB device
        assert ei.value.code in (gpu.ENOMEM, gpu.EINVAL)
        assert c.device_ptrs() == ptrs and c.rows == 300
        c.append(corpus[:10])
        s, r = c.search(corpus[3:4], 2, gpu.HAMMING)
        assert list(r[0]) == [3, 303] and list(s[0]) == [0, 0]


def test_independent_handles_on_concurrent_threads(gpu):
    """Different handles are independent (include/vlite_b200.h): two threads, each with its own
    corpus, stream, k and row width, search at the same time (ctypes drops the GIL) and both stay
    exact -- kernel attributes are per device, not per handle."""
    import threading
    from oracle import search as osearch
    specs = [(128, 32, osearch.HAMMING, 41), (64, 3, osearch.XORSUM, 42), (128, 100, osearch.XORSUM, 43)]
    errors = []

    def worker(w, k, metric, seed):
        try:
            corpus = synth.corpus_rows(seed, 0, 5000, w)
            q = synth.uniform_queries(seed + 100, 9, w)
            es, er = osearch.search(q, corpus, k, metric)
            with gpu.Corpus(w) as c:
                c.append(corpus)
                for _ in range(40):
                    s, r = c.search(q, k, metric)
                    np.testing.assert_array_equal(s, es)
                    np.testing.assert_array_equal(r, er)
        except Exception as e:                                        # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=sp) for sp in specs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_rejected_f32_append_leaves_tombstones_alone(gpu):
    """A vl_corpus_append_f32 that fails validation (VL_ERANGE) appends nothing and must not bring
    deleted rows back: count, live count and the live mask stay as they were."""
    from oracle import cscan, search as osearch, synth
    n, w = 3000, 64
    rows = synth.corpus_rows(33, 0, n, w)
    q = synth.uniform_queries(34, 4, w)
    dead = np.arange(0, n, 7).astype(np.uint64)
    live = np.ones(n, dtype=bool)
    live[dead.astype(np.int64)] = False
    with gpu.Corpus(w) as c:
        c.append_f32(osearch.to_ref_encoding(rows).astype(np.float32), gpu.SRC_F32_OFFSET)
        c.tombstone(dead)
        bad = osearch.to_ref_encoding(rows[:10]).astype(np.float32)
        bad[3, 5] = 1e6
        with pytest.raises(gpu.VliteB200Error) as e:
            c.append_f32(bad, gpu.SRC_F32_OFFSET)
        assert e.value.code == gpu.ERANGE
        assert c.rows == n and c.live_rows == n - len(dead)
        s, r = c.search(q, 10, 0)
        es, er = cscan.search(q, rows, 10, 0, synth.pack_mask(live))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(r, er)
        first = c.append_f32(osearch.to_ref_encoding(rows[:10]).astype(np.float32), gpu.SRC_F32_OFFSET)
        assert first == n and c.live_rows == n - len(dead) + 10
