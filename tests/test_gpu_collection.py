"""-m gpu: the VLite / EmbeddingModel / CTX mirrors, driven like the reference's own call sites and
checked against outputs the reference produced (tests/golden/collection_golden.json)."""
import json
import os

import numpy as np
import pytest

from oracle import ctxfile as octx, search as osearch, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def bits_embedder(queries_u8):
    """A stand-in for the HF model whose sign pattern IS the given packed vectors."""
    def f(texts):
        n = len(texts)
        return np.unpackbits(queries_u8[:n], axis=1).astype(np.float32) * 2 - 1
    return f


@pytest.fixture()
def golden():
    g = json.load(open(os.path.join(GOLD, "collection_golden.json")))
    w = g["width"]
    g["corpus"] = np.frombuffer(bytes.fromhex(g["corpus_u8_hex"]), dtype=np.uint8).reshape(-1, w)
    g["queries"] = np.frombuffer(bytes.fromhex(g["queries_u8_hex"]), dtype=np.uint8).reshape(-1, w)
    return g


def _build(gpu, golden, tmp_path, **kw):
    from vlite_b200 import VLite
    v = VLite("golden", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(golden["queries"]), **kw)
    enc = osearch.to_ref_encoding(golden["corpus"])
    keys = [k for k in golden["keys"] if k != "odd_0"]     # the reference skips the odd-width row (main.py:143)
    first = v._append_vectors(enc)
    for i, key in enumerate(keys):
        v._register(key, f"text {i}", dict(golden["metadata"][key]), first + i)
    return v


def test_rank_and_filter_matches_reference_outputs(gpu, golden, tmp_path):
    v = _build(gpu, golden, tmp_path, filter_mode="reference")
    for case in golden["rank_and_filter"]:
        q = osearch.to_ref_encoding(golden["queries"][case["q"]])
        got = v.rank_and_filter(q, case["top_k"], case["metadata"], case["mult"])
        assert [c for c, _ in got] == case["ids"], case
        assert [int(s) for _, s in got] == case["scores"], case      # includes main.py:165's mis-pairing
    v.close()


def test_prefilter_is_true_filtered_topk(gpu, golden, tmp_path):
    v = _build(gpu, golden, tmp_path)                                  # default: filter inside the scan
    keys = [k for k in golden["keys"] if k != "odd_0"]
    for md in ({"tag": "b"}, {"tag": "c", "parity": 1}, {"tag": "zzz"}):
        valid = np.array([all(golden["metadata"][k].get(a) == b for a, b in md.items()) for k in keys])
        for qi in range(2):
            got = v.rank_and_filter(osearch.to_ref_encoding(golden["queries"][qi]), 5, md)
            es, er = osearch.search(golden["queries"][qi:qi + 1], golden["corpus"], 5, osearch.XORSUM, valid=valid)
            keep = er[0] != osearch.SENTINEL_ROW
            assert [c for c, _ in got] == [keys[int(r)] for r in er[0][keep]]
            assert [int(s) for _, s in got] == [int(s) for s in es[0][keep]]
    v.close()


def test_filter_masks_follow_updates_deletes_and_none_terms(gpu, tmp_path):
    """The (key, value) bitmap index behind the in-scan filter stays exact through set_batch,
    update (chunks sharing one metadata dict), delete, terms with value None (md.get(k) == None
    also matches rows without k, vlite/main.py:162) and values the index cannot hold."""
    from vlite_b200 import VLite
    rng = np.random.default_rng(21)
    n, w = 700, 64
    corpus = synth.corpus_rows(77, 0, n, w)
    queries = synth.uniform_queries(78, 3, w)
    v = VLite("filt", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(queries), width=w)
    metas = []
    for i in range(n):
        md = {"lang": ["en", "fr", None][int(rng.integers(3))]} if rng.random() < 0.8 else {}
        if rng.random() < 0.5:
            md["year"] = int(rng.integers(2020, 2023))
        if i % 50 == 0:
            md["tags"] = ["x", i % 100]
        metas.append(md)
    v.set_batch([f"t{i}" for i in range(n)], osearch.to_ref_encoding(corpus), [dict(m) for m in metas])
    keys = list(v.dump().keys())

    def check(filters):
        live = [k in v.dump() for k in keys]
        for md in filters:
            valid = np.array([live[i] and all(v.dump()[k]["metadata"].get(a) == b for a, b in md.items())
                              if live[i] else False for i, k in enumerate(keys)])
            got = v.retrieve_batch(["q"] * 3, top_k=7, metadata=md, return_scores=True)
            es, er = osearch.search(queries, corpus, 7, osearch.XORSUM, valid=valid)
            for qi in range(3):
                keep = er[qi] != osearch.SENTINEL_ROW
                assert [g[0] for g in got[qi]] == [keys[int(r)] for r in er[qi][keep]], md
                assert [int(g[3]) for g in got[qi]] == [int(x) for x in es[qi][keep]], md

    filters = [{"lang": "en"}, {"lang": None}, {"lang": "fr", "year": 2021}, {"year": 2022.0},
               {"tags": ["x", 0]}, {"nope": 1}]
    check(filters)
    for k in keys[10:200:7]:                                         # update: value replaced in place
        assert v.update(k.rsplit("_", 1)[0], metadata={"lang": "de", "year": 2021})
    for k in keys[300:400:9]:
        assert v.delete(k.rsplit("_", 1)[0]) == 1
    check(filters + [{"lang": "de"}, {"lang": "de", "year": 2021}])
    # chunks of one add() share a metadata dict; update must move every chunk's bits
    v.add(["a", "b", "c"], metadata={"lang": "xx"}, item_id="shared")
    assert len(v.retrieve_batch(["q"], top_k=10, metadata={"lang": "xx"})[0]) == 3
    v.update("shared", metadata={"lang": "yy"})
    assert v.retrieve_batch(["q"], top_k=10, metadata={"lang": "xx"})[0] == []
    assert len(v.retrieve_batch(["q"], top_k=10, metadata={"lang": "yy"})[0]) == 3
    v.close()


def test_concurrent_callers_take_turns(gpu, tmp_path):
    """One VLite shared by several threads (the reference's server does exactly that,
    vlite/server.py:14): filtered retrieves stay exact while another thread keeps appending."""
    import threading
    from vlite_b200 import VLite
    n, w = 2000, 128
    corpus = synth.corpus_rows(91, 0, n, w)
    queries = synth.uniform_queries(92, 4, w)
    v = VLite("mt", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(queries), width=w, metric="hamming")
    v.set_batch([f"t{i}" for i in range(n)], osearch.to_ref_encoding(corpus), [{"g": "old"} for _ in range(n)])
    keys = list(v.dump().keys())
    es, er = osearch.search(queries, corpus, 6, osearch.HAMMING)
    want = [[(keys[int(r)], int(s)) for s, r in zip(es[q], er[q])] for q in range(4)]
    errors, stop = [], threading.Event()

    def reader():
        try:
            for _ in range(25):
                got = v.retrieve_batch(["q"] * 4, top_k=6, metadata={"g": "old"}, return_scores=True)
                assert [[(g[0], int(g[3])) for g in row] for row in got] == want
        except Exception as e:                                            # noqa: BLE001
            errors.append(e)

    def writer():
        i = 0
        try:
            while not stop.is_set() and i < 40:
                extra = synth.corpus_rows(1000 + i, 0, 300, w)
                v.set_batch([f"n{i}_{j}" for j in range(300)], osearch.to_ref_encoding(extra),
                            [{"g": "new"} for _ in range(300)])
                i += 1
        except Exception as e:                                            # noqa: BLE001
            errors.append(e)

    readers = [threading.Thread(target=reader) for _ in range(4)]
    wr = threading.Thread(target=writer)
    for t in readers + [wr]:
        t.start()
    for t in readers:
        t.join()
    stop.set()
    wr.join()
    assert not errors, errors
    assert v.count() > n
    v.close()


def test_retrieve_merges_all_queries_like_reference(gpu, golden, tmp_path):
    v = _build(gpu, golden, tmp_path)
    r = v.retrieve(["a", "b"], top_k=4, return_scores=True)          # two query rows -> ONE list
    assert [[a, b, c, int(d)] for a, b, c, d in r] == golden["retrieve_scores"]
    assert [[a, b, c] for a, b, c in v.retrieve(["a", "b"], top_k=3)] == golden["retrieve_plain"]
    assert v.retrieve(None) is golden["retrieve_none"] and v.retrieve("") is None
    per_query = v.retrieve_batch(["a", "b"], top_k=2, return_scores=True)
    assert len(per_query) == 2 and per_query[1][0][0] == "doc5_2" and int(per_query[1][0][3]) == 64
    v.close()


def test_delete_prefix_semantics_and_persistence(gpu, golden, tmp_path):
    v = _build(gpu, golden, tmp_path)
    enc = osearch.to_ref_encoding(golden["corpus"])
    v.set_batch(["x"], enc[:1])
    extra_key = [k for k in v.index.keys()][-1]
    assert extra_key.endswith("_0")
    v.delete(extra_key[:-2])
    for cid, row in (("doc1extra_0", 0), ("doc1_extra_0", 1)):
        first = v._append_vectors(enc[row:row + 1])
        v._register(cid, cid, {}, first)
    before = v.count()
    assert before == golden["delete"]["before"]     # (the golden run dropped its odd-width row before deleting)
    assert v.delete("doc1") == golden["delete"]["deleted"]
    assert ("doc1extra_0" in v.index) == golden["delete"]["doc1extra_0_survives"]
    assert ("doc1_extra_0" in v.index) == golden["delete"]["doc1_extra_0_survives"]
    assert v.delete(["nope", "nada"]) == golden["delete_missing"]
    # deleted rows never come back from a search
    got = v.rank_and_filter(osearch.to_ref_encoding(golden["corpus"][3]), 3)     # doc1_0 was row 3
    assert all(not c.startswith("doc1_") for c, _ in got)
    # delete() saved a CLEAN file: reload gives exactly the live collection
    from vlite_b200 import VLite
    v2 = VLite("golden", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(golden["queries"]))
    assert v2.count() == v.count() and list(v2.index.keys()) == list(v.index.keys())
    q = osearch.to_ref_encoding(golden["queries"][0])
    assert v2.rank_and_filter(q, 5) == v.rank_and_filter(q, 5)
    # and that file is what the reference's loader would accept (oracle restatement of ctx.py:85-139)
    header, emb, contexts, metadata, _ = octx.read_ctx(str(tmp_path / "contexts" / "golden.ctx"))
    assert header["embedding_size"] == 64 and len(emb) == len(contexts) == len(metadata) == v.count()
    v.close()
    v2.close()


def test_errors_and_set_batch_like_reference(gpu, golden, tmp_path):
    from vlite_b200 import VLite
    v = VLite("empty", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(golden["queries"]))
    with pytest.raises(ValueError, match=golden["empty_error"]):
        v.rank_and_filter(osearch.to_ref_encoding(golden["queries"][0]), 3)
    with pytest.raises(ValueError, match=golden["empty_error"]):        # the same with a metadata filter
        v.rank_and_filter(osearch.to_ref_encoding(golden["queries"][0]), 3, {"i": 0})
    enc = osearch.to_ref_encoding(golden["corpus"])
    with pytest.raises(ValueError) as e1:
        v.set_batch(["a", "b"], enc[:1])
    with pytest.raises(ValueError) as e2:
        v.set_batch(["a"], enc[:1], [{}, {}])
    assert [str(e1.value), str(e2.value)] == golden["set_batch_errors"]
    v.set_batch(["t0", "t1", "t2"], enc[:3], [{"i": 0}, {"i": 1}, {"i": 2}])
    assert v.count() == golden["set_batch_count"]
    assert sorted({k.rsplit("_", 1)[1] for k in v.index}) == golden["set_batch_key_suffixes"]
    assert os.path.exists(tmp_path / "contexts" / "empty.ctx")           # set_batch saves (main.py:271)
    v.clear()
    assert v.count() == 0 and not os.path.exists(tmp_path / "contexts" / "empty.ctx")
    v.close()


def test_add_get_update_dump(gpu, golden, tmp_path):
    from vlite_b200 import VLite
    qs = synth.uniform_queries(40, 8, 64)
    v = VLite("api", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(qs))
    res = v.add(["alpha", {"text": "beta", "metadata": {"lang": "el"}}], metadata={"src": "t"}, item_id="it")
    assert len(res) == 1 and res[0][0] == "it" and res[0][1].shape == (2, 64) and res[0][2] == {"lang": "el", "src": "t"}
    assert v.count() == 2 and list(v.index.keys()) == ["it_0", "it_1"]
    np.testing.assert_array_equal(res[0][1], osearch.to_ref_encoding(qs[:2]))      # reference encoding out
    assert res[0][1].min() >= -256 and res[0][1].max() <= -1
    long = v.add("z" * 5000, need_chunks=True, item_id="long")                     # fast char chunks
    assert v.count() == 2 + 3 and long[0][1].shape[0] == 3
    assert v.get("it") == [("it", "alpha beta", {"src": "t", "lang": "el"})]
    assert v.get(where={"lang": "el"}) == [("it", "beta", {"lang": "el", "src": "t"})]
    assert v.update("it", metadata={"seen": 1}) is True and v.update("nope") is False
    assert v.get("it")[0][2]["seen"] == 1
    d = v.dump()
    e = d["it_1"]
    assert e["text"] == "beta" and e["metadata"]["lang"] == "el"
    np.testing.assert_array_equal(e["binary_vector"], osearch.to_ref_encoding(qs[1]))
    v.set("it", text="gamma")
    assert v.get("it")[0][1] == "gamma gamma"
    v.set("fresh", text="delta", metadata={"a": 1})
    assert "fresh_0" in v.index
    hits = v.retrieve("alpha", top_k=1, metadata={"a": 1})
    assert hits[0][0] == "fresh_0"
    v.close()


def test_embedding_model_search_is_the_reference_seam(gpu):
    from vlite_b200 import EmbeddingModel
    g = np.load(os.path.join(GOLD, "search_golden.npz"))
    m = EmbeddingModel(embedder=lambda t: None)
    for name in ("u64", "u128", "dup128", "kbig64", "u32"):
        corpus, queries, k = g[f"{name}_corpus_u8"], g[f"{name}_queries_u8"], int(g[f"{name}_k"])
        e_f32 = osearch.to_ref_encoding(corpus).astype(np.float32)     # what main.py:146 hands to search
        for qi, q in enumerate(queries):
            idx, sc = m.search(osearch.to_ref_encoding(q), e_f32, k)
            assert idx.dtype == np.int64 and sc.dtype == np.int64 and len(idx) == min(k, len(corpus))
            ref_idx, ref_sc = g[f"{name}_q{qi}_idx"], g[f"{name}_q{qi}_scores"]
            np.testing.assert_array_equal(sc, np.sort(ref_sc, kind="stable"))       # same scores
            ci, cs = osearch.canonicalise_ties(ref_idx, ref_sc)
            kth = sc[-1]
            np.testing.assert_array_equal(idx[sc < kth], ci[cs < kth])               # same rows off the tie boundary
            er, es = osearch.ref_search_port(osearch.to_ref_encoding(q), e_f32, k)   # canonical order: bit exact
            np.testing.assert_array_equal(idx, er)
            idx2, sc2 = m.search(q, corpus, k)                                       # raw bytes score the same
            np.testing.assert_array_equal(sc2, sc)
        bits_q = np.unpackbits(queries[0])[None, :]
        bits_c = np.unpackbits(corpus, axis=1)
        np.testing.assert_array_equal(m.hamming_distance(bits_q, bits_c), g[f"{name}_q0_hamming"])


def test_quantize_binary_matches_packbits(gpu):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((37, 512)).astype(np.float32)
    x[0, :8] = 0.0                                                       # (x > 0): zero is a 0 bit
    np.testing.assert_array_equal(gpu.quantize_binary(x), osearch.quantize_binary(x))
    x2 = rng.standard_normal((5, 1024)).astype(np.float32)
    np.testing.assert_array_equal(gpu.quantize_binary(x2), np.packbits(x2 > 0, axis=-1))


def test_fused_quantisation_tail_on_device(gpu):
    """normalize -> sign -> pack (+ int8 / float32 copies) in one launch on device tensors
    (vlite/model.py:35-44); normalisation is compared with torch's own on the CPU in float64."""
    import torch
    from vlite_b200.model import EmbeddingModel
    rng = np.random.default_rng(11)
    for n, dims, bdims in ((37, 1024, 512), (5, 1024, 1024), (3, 96, 64), (1, 32, 32), (2, 4096, 1024)):
        x = (rng.standard_normal((n, dims)) * rng.uniform(0.01, 30.0, size=(n, 1))).astype(np.float32)
        x[0, :4] = 0.0
        if n > 2:
            x[2] = 0.0                                                   # zero row: eps guard, all-zero outputs
        xd = torch.from_numpy(x).cuda()
        u8 = torch.empty((n, bdims // 8), dtype=torch.uint8, device="cuda")
        i8 = torch.empty((n, dims), dtype=torch.int8, device="cuda")
        f32 = torch.empty((n, dims), dtype=torch.float32, device="cuda")
        scale = 400.0
        gpu.quantize_embeddings(xd, bdims, u8, i8, f32, i8_scale=scale)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(u8.cpu().numpy(), osearch.quantize_binary(x[:, :bdims]))
        want = torch.nn.functional.normalize(torch.from_numpy(x).double(), p=2, dim=1).numpy()
        np.testing.assert_allclose(f32.cpu().numpy(), want, rtol=2e-6, atol=1e-12)
        t = np.clip(want * scale, -127, 127)
        got = i8.cpu().numpy().astype(np.int64)
        off = got != np.rint(t)
        assert np.all(np.abs(np.abs(t[off] - np.floor(t[off])) - 0.5) < 1e-3)    # only at rounding ties
        assert off.mean() < 1e-3
        only_bits = torch.empty_like(u8)
        gpu.quantize_embeddings(xd, bdims, only_bits)                       # optional outputs omitted
        assert torch.equal(only_bits, u8)
    # the model mirror: a CUDA-tensor embedder goes through the fused tail
    emb = torch.from_numpy(rng.standard_normal((6, 1024)).astype(np.float32)).cuda()
    m = EmbeddingModel(embedder=lambda texts: emb[:len(texts)])
    out = m.embed_device(["t"] * 6, int8=True, float32=True)
    np.testing.assert_array_equal(out["ubinary"].cpu().numpy(), osearch.quantize_binary(emb.cpu().numpy()[:, :512]))
    assert out["int8"].shape == (6, 1024) and out["float32"].shape == (6, 1024)
    np.testing.assert_array_equal(m.embed(["t"] * 6), osearch.to_ref_encoding(out["ubinary"].cpu().numpy()))
    with pytest.raises(gpu.VliteB200Error):
        gpu.quantize_embeddings(emb, 2048, out["ubinary"])                 # binary_dims > dims
    with pytest.raises(gpu.VliteB200Error):
        gpu.quantize_embeddings(emb.cpu(), 512, out["ubinary"])            # host pointer


def test_ctx_embeddings_stream_to_hbm(gpu, tmp_path):
    w = 64
    rows = synth.corpus_rows(50, 0, 3000, w)
    p = str(tmp_path / "big.ctx")
    octx.write_ctx(p, {"embedding_model": "m", "embedding_size": 64, "embedding_dtype": "float32", "context_length": 512},
                   osearch.to_ref_encoding(rows).astype(np.float32), [f"c{i}" for i in range(3000)],
                   {f"id{i}_0": {"i": i} for i in range(2990)})          # fewer keys than rows: surplus ignored
    from vlite_b200 import VLite
    v = VLite("big", ctx_dir=str(tmp_path))
    assert v.count() == 2990
    np.testing.assert_array_equal(v._corpus.read(), rows[:2990])
    got = v.rank_and_filter(osearch.to_ref_encoding(rows[1234]), 1)
    assert got == [("id1234_0", np.int64(0))]
    v.close()
    # legacy layout (tests/contexts/vlite-bench.ctx): 512 floats of 0.0/1.0 per row
    f = json.load(open(os.path.join(GOLD, "fixture_facts.json")))
    first8 = np.frombuffer(bytes.fromhex(f["legacy_first8_packed_hex"]), dtype=np.uint8).reshape(8, 64)
    p2 = str(tmp_path / "legacy.ctx")
    import struct
    with open(p2, "wb") as fh:
        hj = json.dumps(f["header"]).encode()
        fh.write(b"CTXF" + struct.pack("<I", 1) + struct.pack("<II", 0, len(hj)) + hj)
        data = np.unpackbits(first8, axis=1).astype("<f4").tobytes()
        fh.write(struct.pack("<II", 1, len(data)) + data)
        fh.write(struct.pack("<II", 2, 0))
        mj = json.dumps({f"k{i}_0": {} for i in range(8)}).encode()
        fh.write(struct.pack("<II", 3, len(mj)) + mj)
    v = VLite("legacy", ctx_dir=str(tmp_path))
    assert v.count() == 8 and v._corpus.width == 64
    np.testing.assert_array_equal(v._corpus.read(), first8)
    v.close()
    # values outside the encoding are refused, loudly
    with gpu.Corpus(64) as c:
        bad = np.full((2, 64), 300.0, dtype=np.float32)
        with pytest.raises(gpu.VliteB200Error) as ei:
            c.append_f32(bad, gpu.SRC_F32_OFFSET)
        assert ei.value.code == gpu.ERANGE and c.rows == 0
        c.append_f32(osearch.to_ref_encoding(rows[:5]).astype(np.float32), gpu.SRC_F32_OFFSET)
        c.append_f32(rows[5:9].astype(np.float32), gpu.SRC_F32_OFFSET)            # raw 0..255 also accepted
        np.testing.assert_array_equal(c.read(), rows[:9])


def test_large_collection_splits_into_v1_part_files(gpu, golden, tmp_path):
    """CTXF v1's u32 section length caps a file at 4 GiB of rows; bigger collections are written as
    several complete v1 files.  Exercised here with a tiny per-file cap."""
    from vlite_b200 import VLite
    v = _build(gpu, golden, tmp_path)
    v.max_rows_per_file = 50
    v.save()
    d = tmp_path / "contexts"
    names = sorted(p.name for p in d.iterdir())
    assert names == ["golden.ctx", "golden.part1.ctx", "golden.part2.ctx"]
    header, emb, contexts, metadata, _ = octx.read_ctx(str(d / "golden.part1.ctx"))   # each part is plain CTXF v1
    assert header["parts"] == 3 and header["part"] == 1 and len(emb) == len(contexts) == len(metadata) == 50
    v2 = VLite("golden", ctx_dir=str(d), embedder=bits_embedder(golden["queries"]))
    assert v2.count() == v.count() == 120 and list(v2.index.keys()) == list(v.index.keys())
    q = osearch.to_ref_encoding(golden["queries"][1])
    assert v2.rank_and_filter(q, 7) == v.rank_and_filter(q, 7)
    np.testing.assert_array_equal(v2._corpus.read(), v._corpus.read())
    v2.max_rows_per_file = 100
    v2.save()                                               # fewer parts now: the stale part file goes away
    assert sorted(p.name for p in d.iterdir()) == ["golden.ctx", "golden.part1.ctx"]
    v2.clear()
    assert list(d.iterdir()) == []
    v.close()
    v2.close()


def test_compact_ubinary_ctx_round_trip(gpu, golden, tmp_path):
    """Opt-in CTX extension: embedding_dtype "ubinary" stores packed bytes (a quarter of the float32
    layout); the default stays the reference's float32 layout."""
    from vlite_b200 import VLite
    v = _build(gpu, golden, tmp_path)
    v.save()
    f32_size = (tmp_path / "contexts" / "golden.ctx").stat().st_size
    v.ctx_dtype = "ubinary"
    v.save()
    path = tmp_path / "contexts" / "golden.ctx"
    assert path.stat().st_size < f32_size - 120 * 64 * 3 + 64
    v2 = VLite("golden", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(golden["queries"]))
    assert v2.count() == 120
    np.testing.assert_array_equal(v2._corpus.read(), v._corpus.read())
    q = osearch.to_ref_encoding(golden["queries"][0])
    assert v2.rank_and_filter(q, 5) == v.rank_and_filter(q, 5)
    v.close()
    v2.close()


@pytest.mark.parametrize("width", [32, 128])
def test_other_row_widths_round_trip_through_ctx(gpu, tmp_path, width):
    """256-bit and 1024-bit collections: add -> retrieve -> save -> reload; the CTX header carries
    the row width (the reference always assumes 64)."""
    from vlite_b200 import VLite
    rows = synth.corpus_rows(70, 0, 500, width)
    qs = rows[[3, 499]].copy()
    emb = lambda texts: np.unpackbits(qs[: len(texts)], axis=1).astype(np.float32) * 2 - 1
    v = VLite("wide", ctx_dir=str(tmp_path), embedder=emb, width=width, metric="hamming")
    v.set_batch([f"t{i}" for i in range(500)], osearch.to_ref_encoding(rows))
    got = v.retrieve(["a", "b"], top_k=2, return_scores=True)
    assert [(g[1], int(g[3])) for g in got] == [("t3", 0), ("t499", 0)]
    header, emb_f32, _, _, _ = octx.read_ctx(str(tmp_path / "wide.ctx"), row_floats=width)
    assert header["embedding_size"] == width and emb_f32.shape == (500, width)
    v2 = VLite("wide", ctx_dir=str(tmp_path), embedder=emb, width=width, metric="hamming")
    assert v2.count() == 500 and v2._corpus.width == width
    np.testing.assert_array_equal(v2._corpus.read(), rows)
    es, er = osearch.search(qs, rows, 5, osearch.HAMMING)
    got = v2.retrieve_batch(["a", "b"], top_k=5, return_scores=True)
    assert [[int(g[3]) for g in one] for one in got] == [list(map(int, x)) for x in es]
    v.close()
    v2.close()


def test_shared_metadata_dict_follows_update_like_reference(gpu, tmp_path):
    """set_batch without metadatas stores ONE dict for every row (vlite/main.py:249); update() of one
    id mutates it in place (vlite/main.py:178), so every row matches the new term -- in the
    in-scan filter as well as in reference filter mode -- until a save/reload separates the dicts."""
    from vlite_b200 import VLite
    rows = synth.corpus_rows(90, 0, 40, 64)
    emb = lambda texts: np.unpackbits(rows[: len(texts)], axis=1).astype(np.float32) * 2 - 1
    for mode in ("prefilter", "reference"):
        d = tmp_path / mode
        v = VLite("shared", ctx_dir=str(d), embedder=emb, filter_mode=mode)
        v.set_batch([f"t{i}" for i in range(40)], osearch.to_ref_encoding(rows))
        one = list(v.index.keys())[7][:-2]
        assert v.update(one, metadata={"tag": "x"})
        assert all(md == {"tag": "x"} for md in v._metas.values())
        hits = v.retrieve("q", top_k=40, metadata={"tag": "x"}, top_k_multiplier=40)
        assert len(hits) == 40                                       # every row carries the term
        assert len(v.get(where={"tag": "x"})) == 40
        v.close()
        v2 = VLite("shared", ctx_dir=str(d), embedder=emb, filter_mode=mode)   # reload: separate dicts,
        assert len(v2.retrieve("q", top_k=40, metadata={"tag": "x"}, top_k_multiplier=40)) == 40  # same answer
        assert v2.update(one, metadata={"tag": "y"})
        assert len(v2.retrieve("q", top_k=40, metadata={"tag": "y"}, top_k_multiplier=40)) == 1
        v2.close()
    # a caller-owned dict reused across add() items behaves the same way
    v = VLite("reuse", ctx_dir=str(tmp_path / "reuse"), embedder=emb)
    md = {"src": "a"}
    v.add([{"text": "one", "metadata": md}], item_id="i1")
    v.add([{"text": "two", "metadata": md}], item_id="i2")
    v.update("i1", metadata={"src": "b"})
    assert {h[0] for h in v.retrieve("q", top_k=5, metadata={"src": "b"})} == {"i1_0", "i2_0"}
    v.close()


@pytest.mark.parametrize("width", [32, 128])
def test_reopen_without_width_adopts_the_files_width(gpu, tmp_path, width):
    """A width-32/128 collection reopened WITHOUT width=: the row width comes from the CTX header
    and queries / later add() calls are embedded at that width."""
    from vlite_b200 import VLite
    rows = synth.corpus_rows(91, 0, 300, width)
    full = lambda texts: np.unpackbits(rows[: len(texts)], axis=1).astype(np.float32) * 2 - 1
    v = VLite("w", ctx_dir=str(tmp_path), embedder=full, width=width, metric="hamming")
    v.set_batch([f"t{i}" for i in range(300)], osearch.to_ref_encoding(rows))
    v.close()
    v2 = VLite("w", ctx_dir=str(tmp_path), embedder=full, metric="hamming")      # no width=
    assert v2._corpus.width == width and v2.model.binary_dims == width * 8
    got = v2.retrieve(["a", "b", "c"], top_k=3, return_scores=True)
    assert [(g[1], int(g[3])) for g in got] == [("t0", 0), ("t1", 0), ("t2", 0)]
    v2.add("extra", item_id="new")                                                # same width: accepted
    assert v2.count() == 301
    v2.close()


def test_save_writes_a_zero_row_for_keys_without_a_vector(gpu, golden, tmp_path):
    """Keys past the stored rows are kept with np.zeros (vlite/main.py:50) and written back as a
    zero row by the reference's save (vlite/main.py:289-293); so here, instead of refusing."""
    from vlite_b200 import VLite
    v = _build(gpu, golden, tmp_path)
    v._register("ghost_0", "no vector", {"tag": "ghost"}, None)
    n = v.count()
    v.save()
    header, emb_f32, contexts, metadata, _ = octx.read_ctx(str(tmp_path / "contexts" / "golden.ctx"))
    assert emb_f32.shape[0] == n == len(metadata) == len(contexts)
    pos = list(metadata.keys()).index("ghost_0")
    assert not emb_f32[pos].any() and emb_f32[pos - 1].any()
    v2 = VLite("golden", ctx_dir=str(tmp_path / "contexts"), embedder=bits_embedder(golden["queries"]))
    assert v2.count() == n and "ghost_0" in v2.index
    v.close()
    v2.close()


def test_journal_mode_persists_without_rewriting_the_collection(gpu, tmp_path):
    """journal=True: set_batch / delete / update append O(change) records (new rows go to a small side
    file) instead of reading the corpus back and rewriting the CTX file; a reopen replays them; save()
    folds everything into clean v1 files a reference reader loads, and removes the journal."""
    from vlite_b200 import VLite
    rows = synth.corpus_rows(95, 0, 900, 64)
    emb = lambda texts: np.unpackbits(rows[[int(t) for t in texts]], axis=1).astype(np.float32) * 2 - 1
    d = str(tmp_path)
    v = VLite("jr", ctx_dir=d, embedder=emb, journal=True)
    v.set_batch([f"t{i}" for i in range(600)], osearch.to_ref_encoding(rows[:600]), [{"g": i % 3} for i in range(600)])
    assert not os.path.exists(os.path.join(d, "jr.ctx"))                # nothing rewritten: the rows are in jr.j0.ctx
    assert os.path.exists(os.path.join(d, "jr.j0.ctx")) and os.path.exists(os.path.join(d, "jr.journal"))
    v.save()                                                             # a full save: clean file, journal gone
    assert os.path.exists(os.path.join(d, "jr.ctx")) and not os.path.exists(os.path.join(d, "jr.journal"))
    assert not os.path.exists(os.path.join(d, "jr.j0.ctx"))
    stamp = os.stat(os.path.join(d, "jr.ctx")).st_mtime_ns
    keys = list(v.index.keys())
    v.set_batch([f"u{i}" for i in range(300)], osearch.to_ref_encoding(rows[600:]), [{"g": 7} for _ in range(300)])
    assert v.delete([keys[5][:-2], keys[17][:-2]]) == 2
    assert v.update(keys[40][:-2], text="changed", metadata={"g": 9})
    assert os.stat(os.path.join(d, "jr.ctx")).st_mtime_ns == stamp      # the collection file was not touched
    want = [v.retrieve(str(q), top_k=4, return_scores=True) for q in (5, 17, 40, 777)]
    want_f = v.retrieve("777", top_k=3, metadata={"g": 7}, return_scores=True)
    v.close()
    v2 = VLite("jr", ctx_dir=d, embedder=emb, journal=True)              # last full save + journal replay
    assert v2.count() == 898 and keys[5] not in v2.index and keys[17] not in v2.index
    assert v2.index[keys[40]]["text"] == "changed" and v2.index[keys[40]]["metadata"]["g"] == 9
    assert [v2.retrieve(str(q), top_k=4, return_scores=True) for q in (5, 17, 40, 777)] == want
    assert v2.retrieve("777", top_k=3, metadata={"g": 7}, return_scores=True) == want_f
    v2.save()
    assert not os.path.exists(os.path.join(d, "jr.journal")) and not os.path.exists(os.path.join(d, "jr.j0.ctx"))
    header, emb_f32, contexts, metadata, _ = octx.read_ctx(os.path.join(d, "jr.ctx"))     # reference-readable
    assert emb_f32.shape == (898, 64) and len(contexts) == 898 and len(metadata) == 898
    v2.close()
    v3 = VLite("jr", ctx_dir=d, embedder=emb)                             # plain mode reads the same collection
    assert [v3.retrieve(str(q), top_k=4, return_scores=True) for q in (5, 17, 40, 777)] == want
    v3.close()
