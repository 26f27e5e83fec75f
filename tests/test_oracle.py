"""The oracle pinned against golden vectors produced by EXECUTING the reference
(tests/golden/make_golden.py; fixtures committed beside it).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import cscan, ctxfile, ref_port, search as osearch, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SG = np.load(os.path.join(GOLD, "search_golden.npz"))
CASES = ["u64", "u128", "dup128", "kbig64", "u32"]


@pytest.mark.parametrize("name", CASES)
def test_xorsum_scores_match_reference_search(name):
    """EmbeddingModel.search (vlite/model.py:76-90): same score list; same index list once ties
    are canonicalised (torch.topk leaves tie order unspecified)."""
    corpus, queries, k = SG[f"{name}_corpus_u8"], SG[f"{name}_queries_u8"], int(SG[f"{name}_k"])
    for qi, q in enumerate(queries):
        ref_idx, ref_sc = SG[f"{name}_q{qi}_idx"], SG[f"{name}_q{qi}_scores"]
        full = osearch.xorsum_scores(q, corpus)
        np.testing.assert_array_equal(full[ref_idx], ref_sc)            # scores of the rows it picked
        rows, sc = osearch.canonical_topk(full, k)
        np.testing.assert_array_equal(sc, np.sort(ref_sc, kind="stable"))   # same score multiset, ascending
        assert len(rows) == min(k, len(corpus))                         # model.py:84 clamp
        # rows strictly inside the k-th score boundary must coincide exactly
        kth = sc[-1]
        np.testing.assert_array_equal(np.sort(rows[sc < kth]), np.sort(ref_idx[ref_sc < kth]))
        # reference encoding and raw ubinary bytes score identically
        np.testing.assert_array_equal(SG[f"{name}_q{qi}_scores_raw"], ref_sc)
        enc_scores = osearch.xorsum_scores(osearch.to_ref_encoding(q), osearch.to_ref_encoding(corpus))
        np.testing.assert_array_equal(enc_scores, full)
        # the C restatement and the port used for timing agree with both
        np.testing.assert_array_equal(cscan.scores(q, corpus, osearch.XORSUM).astype(np.int64), full)
        pr, ps = osearch.ref_search_port(osearch.to_ref_encoding(q), osearch.to_ref_encoding(corpus), k)
        np.testing.assert_array_equal(ps, sc)
        ti, ts = ref_port.search_port(osearch.to_ref_encoding(q),
                                      ref_port.corpus_as_reference_f32(corpus), k)
        np.testing.assert_array_equal(np.sort(ts), sc)
        ci, cs = osearch.canonicalise_ties(ti, ts)
        np.testing.assert_array_equal(cs, sc)


@pytest.mark.parametrize("name", CASES)
def test_hamming_matches_reference_numpy(name):
    """EmbeddingModel.hamming_distance on unpacked bits (vlite/model.py:71-74)."""
    corpus, queries = SG[f"{name}_corpus_u8"], SG[f"{name}_queries_u8"]
    for qi, q in enumerate(queries):
        ref = SG[f"{name}_q{qi}_hamming"]
        np.testing.assert_array_equal(osearch.hamming_scores_unpacked(q, corpus), ref)
        np.testing.assert_array_equal(osearch.hamming_scores(q, corpus), ref)
        np.testing.assert_array_equal(cscan.scores(q, corpus, osearch.HAMMING).astype(np.int64), ref)
    np.testing.assert_array_equal(osearch.hamming_scores(queries, corpus)[1], SG[f"{name}_q1_hamming"])


def test_c_oracle_topk_equals_numpy_topk():
    rng = np.random.default_rng(0)
    for w, n, nq, k, period in [(64, 3000, 5, 7, 0), (128, 2000, 12, 33, 50), (128, 10, 2, 16, 0)]:
        corpus = synth.corpus_rows(1, 0, n, w, period)
        queries = synth.uniform_queries(2, nq, w)
        valid = rng.random(n) < 0.6
        for metric in (0, 1):
            for v in (None, valid):
                es, er = osearch.search(queries, corpus, k, metric, valid=v, base_row=1000)
                mask = None if v is None else synth.pack_mask(v)
                for rowpar in (False, True):
                    cs, cr = cscan.search(queries, corpus, k, metric, mask, 1000, rowpar=rowpar)
                    np.testing.assert_array_equal(cs, es)
                    np.testing.assert_array_equal(cr, er)


def test_merge_topk_is_shard_count_invariant():
    corpus = synth.corpus_rows(3, 0, 4000, 128, period=300)
    queries = synth.uniform_queries(4, 6, 128)
    es, er = osearch.search(queries, corpus, 10, 0)
    for shards in (2, 3, 8):
        per = -(-4000 // shards)
        parts = [osearch.search(queries, corpus[g * per:(g + 1) * per], 10, 0, base_row=g * per) for g in range(shards)]
        ms, mr = osearch.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 10)
        np.testing.assert_array_equal(ms, es)
        np.testing.assert_array_equal(mr, er)


def test_encoding_round_trip_and_notebook_values():
    u = np.arange(256, dtype=np.uint8)
    s = osearch.to_ref_encoding(u)
    assert s.min() == -256 and s.max() == -1
    np.testing.assert_array_equal(osearch.from_ref_encoding(s), u)
    np.testing.assert_array_equal(osearch.from_ref_encoding(u.astype(np.float32)), u)   # raw bytes pass through
    # stored -145, -214, -256 <-> sentence-transformers ubinary 111, 42, 0 ... minus the 0x80 flip
    np.testing.assert_array_equal(osearch.to_ref_encoding(np.array([239, 170, 128], np.uint8)), [-145, -214, -256])
    with pytest.raises(ValueError):
        osearch.from_ref_encoding(np.array([300]))
    x = np.random.default_rng(1).standard_normal((3, 64)).astype(np.float32)
    np.testing.assert_array_equal(osearch.quantize_binary(x), np.packbits(x > 0, axis=-1))


def test_rank_and_filter_port_matches_reference():
    g = json.load(open(os.path.join(GOLD, "collection_golden.json")))
    w = g["width"]
    corpus = np.frombuffer(bytes.fromhex(g["corpus_u8_hex"]), dtype=np.uint8).reshape(-1, w)
    queries = np.frombuffer(bytes.fromhex(g["queries_u8_hex"]), dtype=np.uint8).reshape(-1, w)
    enc = osearch.to_ref_encoding(corpus)
    index = {}
    for i, key in enumerate(g["keys"]):
        vec = enc[i].tolist() if key != "odd_0" else [0] * 32
        index[key] = {"text": f"text {i}", "metadata": g["metadata"][key], "binary_vector": vec}
    for case in g["rank_and_filter"]:
        got = osearch.rank_and_filter_port(index, osearch.to_ref_encoding(queries[case["q"]]), case["top_k"],
                                           case["metadata"], case["mult"], reference_bug=True)
        assert [c for c, _ in got] == case["ids"]
        assert [int(s) for _, s in got] == case["scores"]
    with pytest.raises(ValueError, match=g["empty_error"]):
        osearch.rank_and_filter_port({}, osearch.to_ref_encoding(queries[0]), 3)


def test_ctx_reader_writer_match_reference_bytes():
    path = os.path.join(GOLD, "ref_written.ctx")
    facts = json.load(open(path + ".json"))
    header, emb, contexts, metadata, sections = ctxfile.read_ctx(path)
    assert header == facts["header"] and contexts == facts["contexts"] and metadata == facts["metadata"]
    np.testing.assert_array_equal(emb, np.asarray(facts["embeddings"], dtype=np.float32))
    u8 = ctxfile.embeddings_to_ubinary(emb, "offset")
    assert u8.tobytes().hex() == facts["expected_u8_hex"]
    # our writer reproduces the reference's file byte for byte
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        p2 = os.path.join(td, "again.ctx")
        ctxfile.write_ctx(p2, header, [list(map(float, r)) for r in emb], contexts, metadata)
        assert open(p2, "rb").read() == open(path, "rb").read()
        for name, blob, msg in [("m", b"XXXX\x01\0\0\0", facts["errors"]["bad_magic"]),
                                ("v", b"CTXF\x07\0\0\0", facts["errors"]["bad_version"]),
                                ("s", b"CTXF\x01\0\0\0\x09\0\0\0\0\0\0\0", facts["errors"]["bad_section"])]:
            pp = os.path.join(td, name)
            open(pp, "wb").write(blob)
            with pytest.raises(ValueError) as ei:
                ctxfile.read_ctx(pp)
            assert str(ei.value) == msg


def test_legacy_fixture_facts_are_self_consistent():
    f = json.load(open(os.path.join(GOLD, "fixture_facts.json")))
    assert f["sections"][1][2] == f["legacy_rows"] * 512 * 4        # 398 rows x 512 f32
    assert f["loader_rows_x64"] == f["legacy_rows"] * 8             # the current loader's misreading
    assert f["legacy_values"] == [0.0, 1.0]
    first8 = np.frombuffer(bytes.fromhex(f["legacy_first8_packed_hex"]), dtype=np.uint8).reshape(8, 64)
    bits = np.unpackbits(first8, axis=1).astype(np.float32)
    np.testing.assert_array_equal(ctxfile.embeddings_to_ubinary(bits, "bits"), first8)


def test_synth_generator_properties():
    a = synth.corpus_rows(0, 0, 1000, 128)
    b = synth.corpus_rows(0, 500, 10, 128)
    np.testing.assert_array_equal(a[500:510], b)                        # any slice regenerates
    np.testing.assert_array_equal(synth.corpus_rows(0, 0, 50, 64), a[:50, :64])   # W=64 is the prefix
    d = synth.corpus_rows(0, 0, 300, 128, period=100)
    np.testing.assert_array_equal(d[:100], d[200:300])
    assert abs(np.unpackbits(a).mean() - 0.5) < 0.01
    v = synth.filter_valid(2, 0, 200000, 0.1)
    assert abs(v.mean() - 0.1) < 0.005
    m = synth.pack_mask(v[:70])
    assert m.dtype == np.uint32 and len(m) == 3
    assert all(((m[i >> 5] >> (i & 31)) & 1) == int(v[i]) for i in range(70))
    # mix64 known answer (splitmix64 finaliser of 1)
    assert int(synth.mix64(np.array([1], dtype=np.uint64))[0]) == 0x5692161D100B05E5


def test_rescore_dot_matches_reference_cos_sim():
    """The reference's only dense similarity is cos_sim (vlite/utils.py:205-208): a @ b.T over the norms.
    Its outputs on seeded queries / int8 documents (rescore_golden.npz, produced by executing it) pin the
    oracle's rescore: same dot products once divided by the same norms, and the ranking they induce
    (descending dot, ties by ascending row)."""
    g = np.load(os.path.join(GOLD, "rescore_golden.npz"))
    docs = g["docs_i8"]
    n = docs.shape[0]
    doc_norm = np.linalg.norm(docs.astype(np.float64), axis=1)
    rows = np.arange(n, dtype=np.uint64)[None, :]
    for name in ("f32", "i8"):
        qs = g[f"queries_{name}"]
        for qi, q in enumerate(qs):
            s, r = osearch.rescore(q[None, :], docs, rows, n)           # every document, ranked
            dots = np.empty(n)
            dots[r[0].astype(np.int64)] = s[0]
            cos = dots / (np.linalg.norm(q.astype(np.float64)) * doc_norm)
            np.testing.assert_allclose(cos, g[f"{name}_q{qi}_cos_f64"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(cos, g[f"{name}_q{qi}_cos_f32"].astype(np.float64), rtol=0, atol=2e-6)
            # the ranking the reference's numbers induce once the document norm is multiplied back
            ref_dot = g[f"{name}_q{qi}_cos_f64"] * doc_norm
            assert np.all(np.diff(ref_dot[r[0].astype(np.int64)]) <= 1e-9 * np.abs(ref_dot).max())
            pos5, pos17 = int(np.nonzero(r[0] == 5)[0][0]), int(np.nonzero(r[0] == 17)[0][0])
            assert pos17 == pos5 + 1                                    # identical documents: lower row first
