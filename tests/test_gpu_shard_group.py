"""-m gpu: the in-process shard group (vl_shard_group_*): several shards searched as one corpus
must return exactly what one vl_search over the concatenated corpus returns -- the same
invariance the one-process-per-GPU path is held to.  On a single-GPU box the shards share device 0
(the peer copy degenerates to a device copy); with more GPUs they are spread over all of them."""
import numpy as np
import pytest

from oracle import cscan, synth

pytestmark = pytest.mark.gpu
HAMMING, XORSUM = 0, 1


def _make_group(gpu, corpus, cuts, width, n_dev):
    shards, base = [], 0
    for i, (lo, hi) in enumerate(zip(cuts[:-1], cuts[1:])):
        c = gpu.Corpus(width, device=i % n_dev, base_row=lo)
        if hi > lo:
            c.append(corpus[lo:hi])
        shards.append(c)
    return gpu.ShardGroup(shards), shards


@pytest.mark.parametrize("width,metric", [(128, HAMMING), (64, XORSUM), (32, HAMMING)])
def test_group_equals_one_corpus(gpu, width, metric):
    n_dev = gpu.device_count()
    n = 30_011
    corpus = synth.corpus_rows(5, 0, n, width, period=977)          # duplicates: ties cross shard borders
    cuts = [0, 7_000, 7_000, 19_013, n]                              # ragged, one empty shard
    group, shards = _make_group(gpu, corpus, cuts, width, n_dev)
    try:
        assert group.rows == n and group.live_rows == n
        for nq, k in ((1, 1), (3, 10), (70, 32), (9, 100), (2, 1000)):
            q, _ = synth.clustered_queries(6, 5, n, nq, width, 0.05, period=977)
            s, r = group.search(q, k, metric)
            es, er = cscan.search(q, corpus, k, metric)
            np.testing.assert_array_equal(s, es)
            np.testing.assert_array_equal(r, er)
    finally:
        group.close()
        for c in shards:
            c.close()


def test_group_filters_tombstones_and_errors(gpu):
    n_dev = gpu.device_count()
    width, n = 128, 12_000
    corpus = synth.corpus_rows(15, 0, n, width)
    cuts = [0, 5_000, 9_000, n]
    group, shards = _make_group(gpu, corpus, cuts, width, n_dev)
    try:
        valid = synth.filter_valid(2, 0, n, 0.3)
        dead = np.array([10, 4_999, 5_000, 8_999, 11_999])
        valid_after = valid.copy()
        valid_after[dead] = False
        masks = [synth.pack_mask(valid[lo:hi]) for lo, hi in zip(cuts[:-1], cuts[1:])]
        for lo, hi, c in zip(cuts[:-1], cuts[1:], shards):
            local = [int(d - lo) for d in dead if lo <= d < hi]
            if local:
                c.tombstone(local)
        assert group.live_rows == n - len(dead)
        q, _ = synth.clustered_queries(16, 15, n, 5, width, 0.1)
        s, r = group.search(q, 10, HAMMING, masks)
        es, er = cscan.search(q, corpus, 10, HAMMING, synth.pack_mask(valid_after))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(r, er)
        masks[1] = None                                                 # a shard without a filter
        v2 = valid_after.copy()
        v2[cuts[1]:cuts[2]] = True
        v2[dead] = False
        s, r = group.search(q, 10, HAMMING, masks)
        es, er = cscan.search(q, corpus, 10, HAMMING, synth.pack_mask(v2))
        np.testing.assert_array_equal(s, es)
        np.testing.assert_array_equal(r, er)
        with pytest.raises(gpu.VliteB200Error):
            group.search(q, gpu.MAX_K + 1, HAMMING)
        with pytest.raises(ValueError):
            group.search(q[:, :64], 5, HAMMING)
        with pytest.raises(ValueError):
            group.search(q, 5, HAMMING, masks[:2])
        with pytest.raises(gpu.VliteB200Error):
            gpu.ShardGroup([shards[0], shards[0]])                      # the same handle twice
        other = gpu.Corpus(64)
        with pytest.raises(gpu.VliteB200Error):
            gpu.ShardGroup([shards[0], other])                          # mixed widths
        other.close()
    finally:
        group.close()
        for c in shards:
            c.close()


def test_group_spread_over_every_gpu(gpu):
    """With several GPUs the shards live on different devices (peer copies over NVLink)."""
    n_dev = gpu.device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    width, per = 128, 200_000
    n = per * n_dev
    corpus = synth.corpus_rows(21, 0, n, width)
    shards = []
    for d in range(n_dev):
        c = gpu.Corpus(width, device=d, base_row=d * per)
        c.append(corpus[d * per:(d + 1) * per])
        shards.append(c)
    with gpu.ShardGroup(shards) as group:
        q, _ = synth.clustered_queries(22, 21, n, 33, width, 0.08)
        for k in (10, 64):
            s, r = group.search(q, k, HAMMING)
            es, er = cscan.search(q, corpus, k, HAMMING)
            np.testing.assert_array_equal(s, es)
            np.testing.assert_array_equal(r, er)
    for c in shards:
        c.close()


def test_vlite_over_several_shards_in_one_process(gpu, tmp_path):
    """VLite(gpus=[...]) row-shards the collection inside this process; every answer must equal
    the single-GPU collection's -- same ids, same scores, same order -- through interleaved adds,
    filters, deletes, save and reload."""
    from vlite_b200 import VLite
    from oracle import search as osearch
    n_dev = gpu.device_count()
    gpus = list(range(n_dev)) if n_dev >= 2 else [0, 0, 0]
    w = 64
    corpus = synth.corpus_rows(31, 0, 2600, w, period=500)             # duplicates: ties across shards
    queries = synth.uniform_queries(32, 4, w)

    def embedder(texts):
        return np.unpackbits(queries[:len(texts)], axis=1).astype(np.float32) * 2 - 1

    def build(name, **kw):
        v = VLite(name, ctx_dir=str(tmp_path / name), embedder=embedder, width=w, metric="hamming", **kw)
        off = 0
        for b, size in enumerate((700, 1, 900, 333, 666)):               # ragged batches -> interleaved segments
            enc = osearch.to_ref_encoding(corpus[off:off + size])
            first = v._append_vectors(enc)
            for i in range(size):
                v._register(f"doc{off + i}_0", f"text {off + i}", {"batch": b, "odd": (off + i) % 2}, first + i)
            off += size
        return v

    one, many = build("one"), build("many", gpus=gpus)
    assert many.count() == one.count() == 2600

    def same(**kw):
        a = one.retrieve_batch(["q"] * 4, return_scores=True, **kw)
        b = many.retrieve_batch(["q"] * 4, return_scores=True, **kw)
        assert a == b
        return a

    same(top_k=7)
    same(top_k=40)
    same(top_k=9, metadata={"odd": 1})
    same(top_k=9, metadata={"batch": 2, "odd": 0})
    for v in (one, many):
        assert v.delete([f"doc{i}" for i in (0, 699, 700, 701, 1600, 2599)]) == 6
    got = same(top_k=12)
    assert all(cid not in ("doc700_0", "doc2599_0") for row in got for cid, *_ in row)
    np.testing.assert_array_equal(many._corpus.read(), one._corpus.read())
    many.save()
    many.close()
    again = VLite("many", ctx_dir=str(tmp_path / "many"), embedder=embedder, width=w, metric="hamming", gpus=gpus)
    assert again.count() == one.count()
    a = one.retrieve_batch(["q"] * 4, top_k=10, return_scores=True)
    b = again.retrieve_batch(["q"] * 4, top_k=10, return_scores=True)
    assert [[(c, int(s)) for c, _, _, s in row] for row in a] == [[(c, int(s)) for c, _, _, s in row] for row in b]
    again.clear()
    assert again.count() == 0
    again.close()
    one.close()
