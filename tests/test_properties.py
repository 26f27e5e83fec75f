"""Property tests (hypothesis) of the host-side pieces that need no GPU: the CTX container round
trip, the shard plan's partition of the rows, the metadata bitmap index and the oracle's own
invariants (linearity of the XOR scores, tie-break order, merge of shard top-k = global top-k)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import search as osearch
from vlite_b200.ctx import CtxFile
from vlite_b200.metaindex import MetaIndex
from vlite_b200.sharded import ShardPlan

json_scalars = st.one_of(st.none(), st.booleans(), st.integers(-2**31, 2**31), st.text(max_size=8),
                         st.floats(allow_nan=False, allow_infinity=False, width=32))
metadata = st.dictionaries(st.text(min_size=1, max_size=6), json_scalars, max_size=4)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 40), st.sampled_from([32, 64, 128]), st.sampled_from(["float32", "ubinary"]),
       st.lists(st.text(max_size=30), max_size=40), st.data())
def test_ctx_round_trip(tmp_path_factory, n, width, dtype, texts, data):
    path = str(tmp_path_factory.mktemp("ctx") / "p.ctx")
    rng = np.random.default_rng(n * 7 + width)
    rows = rng.integers(0, 256, size=(n, width), dtype=np.uint8)
    cf = CtxFile(path)
    cf.set_header("m", width, dtype, 512)
    if dtype == "ubinary":
        cf.set_embeddings_array(rows)
    else:
        cf.set_embeddings_array(osearch.to_ref_encoding(rows).astype(np.float32))
    for t in texts:
        cf.add_context(t)
    metas = {f"id{i}_0": data.draw(metadata) for i in range(min(n, 5))}
    for k, v in metas.items():
        cf.add_metadata(k, v)
    cf.save()
    back = CtxFile(path)
    back.load()
    assert back.header["embedding_size"] == width and back.n_embeddings == n
    if n:
        got = np.asarray(back.embeddings)
        want = rows if dtype == "ubinary" else osearch.to_ref_encoding(rows).astype(np.float32)
        np.testing.assert_array_equal(got, want)
    assert back.contexts == texts
    assert back.metadata == metas


@given(st.integers(0, 10_000), st.integers(1, 9))
def test_shard_plan_partitions_the_rows(total, world):
    plan = ShardPlan(total, world)
    covered = 0
    for r in range(world):
        assert plan.base_row(r) == covered or plan.shard_rows(r) == 0
        covered += plan.shard_rows(r)
    assert covered == total
    for g in {0, total // 2, max(total - 1, 0)} if total else set():
        r, local = plan.owner(g)
        assert 0 <= local < plan.shard_rows(r) and plan.base_row(r) + local == g


@settings(max_examples=30, deadline=None)
@given(st.lists(st.tuples(st.sampled_from(["a", "b", None]), st.sampled_from([1, 2.0, True, None, "x"])),
                min_size=1, max_size=120),
       st.sampled_from([{"k": "a"}, {"k": None}, {"v": 1}, {"k": "b", "v": 2}, {"v": None}, {}]))
def test_metaindex_equals_the_reference_predicate(rows, query):
    metas = []
    for k, v in rows:
        md = {}
        if k is not None or v == "x":
            md["k"] = k
        if v is not None:
            md["v"] = v
        metas.append(md)
    idx = MetaIndex()
    idx.set_many(metas, 3)
    n = len(metas) + 3
    got = np.unpackbits(idx.mask(query, n).view(np.uint8), bitorder="little")[:n].astype(bool)
    want = np.array([False] * 3 + [all(md.get(a) == b for a, b in query.items()) for md in metas])
    if not query:
        want[:3] = True                         # no terms: every row passes (unregistered rows are masked by the device)
    for a, b in query.items():
        if b is None:
            want[:3] = True                     # rows without the key satisfy md.get(k) == None ... as would an unregistered row
    np.testing.assert_array_equal(got, want)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 300), st.integers(1, 6), st.integers(1, 12), st.sampled_from([32, 64, 128]),
       st.integers(1, 5), st.integers(0, 2**31))
def test_oracle_invariants(n, nq, k, w, shards, seed):
    rng = np.random.default_rng(seed)
    corpus = rng.integers(0, 256, size=(n, w), dtype=np.uint8)
    corpus[rng.integers(0, n, size=n // 3)] = corpus[0]                  # ties
    q = rng.integers(0, 256, size=(nq, w), dtype=np.uint8)
    for metric in (osearch.HAMMING, osearch.XORSUM):
        s, r = osearch.search(q, corpus, k, metric)
        kk = min(k, n)
        keys = (s[:, :kk].astype(np.int64) << 32) | r[:, :kk].astype(np.int64)
        assert np.all(np.diff(keys, axis=1) > 0)                          # strictly ascending (score, row)
        # merging shard-local top-k lists gives the global top-k (the all-gather + merge step)
        cuts = sorted(set([0, n] + [int(x) for x in rng.integers(0, n + 1, size=shards - 1)]))
        parts = [osearch.search(q, corpus[lo:hi], k, metric, base_row=lo) for lo, hi in zip(cuts[:-1], cuts[1:])]
        ms, mr = osearch.merge_topk([p[0] for p in parts], [p[1] for p in parts], k)
        np.testing.assert_array_equal(ms, s)
        np.testing.assert_array_equal(mr, r)
    # XOR scores: d(q, e) == d(q ^ m, e ^ m) for any mask m, and d(e, e) == 0
    m = rng.integers(0, 256, size=(w,), dtype=np.uint8)
    np.testing.assert_array_equal(osearch.hamming_scores(q[0], corpus), osearch.hamming_scores(q[0] ^ m, corpus ^ m))
    assert osearch.hamming_scores(corpus[0], corpus[:1])[0] == 0
