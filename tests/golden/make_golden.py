#!/usr/bin/env python
"""Generate the golden vectors in this directory by EXECUTING THE REFERENCE.

Run in the build container only (needs /root/reference; the GPU box has no
copy):   python tests/golden/make_golden.py

It imports the reference's own modules -- vlite/model.py by file path, and
vlite.main / vlite.ctx with inert stub modules standing in for the reference's
network/ingest dependencies (posthog, PyPDF2, docx2txt, bs4, tiktoken ...), as
SURVEY.md section 8c describes -- runs them on seeded inputs, and stores inputs
and outputs.  Nothing from the reference's sources is copied; only the data it
produced.  Files written:

  search_golden.npz        EmbeddingModel.search + hamming_distance (vlite/model.py:71-90)
  collection_golden.json   VLite.rank_and_filter / retrieve / delete / set_batch (vlite/main.py)
  ref_written.ctx(+.json)  a CTXF v1 file written by CtxFile.save, and what CtxFile.load returns
  fixture_facts.json       facts about tests/contexts/vlite-bench.ctx (the reference's fixture)
  rescore_golden.npz       cos_sim (vlite/utils.py:205-208), the only dense similarity the reference holds:
                           queries x int8 documents, float32 and float64 -- pins the rescore's dot product
                           (``--only rescore`` regenerates this file alone)
"""
import hashlib
import importlib.util
import json
import os
import struct
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import synth  # noqa: E402  (inputs only; outputs come from the reference)
from oracle.search import to_ref_encoding  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    class Posthog:
        def __init__(self, *a, **k):
            pass

        def capture(self, *a, **k):
            pass

    _stub("posthog", Posthog=Posthog)
    _stub("PyPDF2")
    _stub("docx2txt")
    _stub("bs4", BeautifulSoup=object)
    for opt in ("tiktoken", "surya"):
        if opt not in sys.modules:
            try:
                __import__(opt)
            except Exception:
                _stub(opt)
    spec = importlib.util.spec_from_file_location("ref_model", os.path.join(REF, "vlite/model.py"))
    ref_model = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_model)
    sys.path.insert(0, REF)
    import vlite.main as ref_main
    import vlite.ctx as ref_ctx
    return ref_model, ref_main, ref_ctx


def golden_search(ref_model):
    self_ = types.SimpleNamespace(device="cpu")
    out = {}
    cases = [
        ("u64", 64, 700, 5, 0, 11),
        ("u128", 128, 900, 10, 0, 12),
        ("dup128", 128, 600, 8, 37, 13),   # heavy ties (distribution D)
        ("kbig64", 64, 50, 80, 0, 14),     # top_k > N -> clamp (model.py:84)
        ("u32", 32, 500, 7, 0, 15),        # 256-bit rows: the reference's search is width-agnostic
    ]
    for name, w, n, k, period, seed in cases:
        corpus = synth.corpus_rows(seed, 0, n, w, period)
        qs = synth.uniform_queries(seed + 100, 3, w)
        qs[1] = corpus[n // 2]                 # exact hit
        out[f"{name}_corpus_u8"] = corpus
        out[f"{name}_queries_u8"] = qs
        out[f"{name}_k"] = np.int64(k)
        e_ref = to_ref_encoding(corpus).astype(np.float32)          # what main.py:146 builds
        for qi, q in enumerate(qs):
            q_ref = to_ref_encoding(q)                              # int16, model.py:44
            idx, sc = ref_model.EmbeddingModel.search(self_, q_ref, e_ref, k)
            out[f"{name}_q{qi}_idx"] = np.asarray(idx, dtype=np.int64)
            out[f"{name}_q{qi}_scores"] = np.asarray(sc, dtype=np.int64)
            # same call on RAW ubinary bytes: scores must be identical
            idx2, sc2 = ref_model.EmbeddingModel.search(self_, q.astype(np.int16), corpus.astype(np.float32), k)
            out[f"{name}_q{qi}_scores_raw"] = np.asarray(sc2, dtype=np.int64)
            # dead-code true Hamming on unpacked bits, model.py:71-74
            hb = ref_model.EmbeddingModel.hamming_distance(
                self_, np.unpackbits(q)[None, :], np.unpackbits(corpus, axis=1))
            out[f"{name}_q{qi}_hamming"] = np.asarray(hb, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "search_golden.npz"), **out)
    print("search_golden.npz:", len(out), "arrays")


def _make_instance(ref_main, ref_model, ctx_dir, collection):
    v = ref_main.VLite.__new__(ref_main.VLite)          # skip HF download in __init__
    v.collection = collection
    v.device = "cpu"
    v.index = {}
    v.ctx = ref_main.Ctx(ctx_dir)
    model = types.SimpleNamespace(device="cpu", embedding_dtype="float32", context_length=512, dimension=1024)
    model.search = lambda q, e, k: ref_model.EmbeddingModel.search(model, q, e, k)
    v.model = model
    return v


def golden_collection(ref_model, ref_main):
    res = {}
    w, n = 64, 120
    corpus = synth.corpus_rows(21, 0, n, w)
    enc = to_ref_encoding(corpus)
    tags = ["a", "b", "c"]
    with tempfile.TemporaryDirectory() as td:
        v = _make_instance(ref_main, ref_model, td, "golden")
        for i in range(n):
            v.index[f"doc{i // 3}_{i % 3}"] = {
                "text": f"text {i}",
                "metadata": {"tag": tags[i % 3], "parity": i % 2},
                "binary_vector": enc[i].tolist(),
            }
        # one row of a different width: must be skipped (main.py:143)
        v.index["odd_0"] = {"text": "odd", "metadata": {}, "binary_vector": [0] * 32}
        qs = synth.uniform_queries(22, 2, w)
        qs[1] = corpus[17] ^ np.uint8(1)
        res["width"] = w
        res["corpus_u8_hex"] = corpus.tobytes().hex()
        res["queries_u8_hex"] = qs.tobytes().hex()
        res["keys"] = list(v.index.keys())
        res["metadata"] = {k: it["metadata"] for k, it in v.index.items()}
        raf = []
        for qi, q in enumerate(qs):
            q_ref = to_ref_encoding(q)
            for md, k, mult in [(None, 5, 4), ({"tag": "b"}, 5, 4), ({"tag": "c", "parity": 1}, 4, 2), ({"tag": "zzz"}, 3, 4)]:
                out = v.rank_and_filter(q_ref, k, md, mult)
                raf.append({"q": qi, "metadata": md, "top_k": k, "mult": mult,
                            "ids": [o[0] for o in out], "scores": [int(o[1]) for o in out]})
        res["rank_and_filter"] = raf
        # retrieve(): Q query rows -> ONE merged list (main.py:116-121)
        q_both = to_ref_encoding(qs)
        v.model.embed = lambda text, precision="binary": q_both
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            r = v.retrieve("ignored", top_k=4, return_scores=True)
            r_none = v.retrieve(None)
            r_plain = v.retrieve("ignored", top_k=3)
        res["retrieve_scores"] = [[a, b, c, int(d)] for a, b, c, d in r]
        res["retrieve_none"] = r_none
        res["retrieve_plain"] = [[a, b, c] for a, b, c in r_plain]
        # delete(): prefix semantics + count (main.py:190-205).  The short row must go first:
        # the reference's save() cannot pack a row with fewer than 64 values (ctx.py:70).
        del v.index["odd_0"]
        v.index["doc1extra_0"] = {"text": "x", "metadata": {}, "binary_vector": enc[0].tolist()}
        v.index["doc1_extra_0"] = {"text": "y", "metadata": {}, "binary_vector": enc[1].tolist()}
        before = len(v.index)
        deleted = v.delete("doc1")
        res["delete"] = {"before": before, "deleted": int(deleted), "after": len(v.index),
                         "doc1extra_0_survives": "doc1extra_0" in v.index,
                         "doc1_extra_0_survives": "doc1_extra_0" in v.index}
        res["delete_missing"] = int(v.delete(["nope", "nada"]))
        # empty collection -> ValueError (main.py:149)
        v2 = _make_instance(ref_main, ref_model, td, "empty")
        try:
            v2.rank_and_filter(to_ref_encoding(qs[0]), 3)
            res["empty_error"] = None
        except ValueError as e:
            res["empty_error"] = str(e)
        # set_batch length checks (main.py:253-258)
        errs = []
        for texts, embs, mds in [(["a", "b"], enc[:1], None), (["a"], enc[:1], [{}, {}])]:
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    v2.set_batch(texts, embs, mds)
                errs.append(None)
            except ValueError as e:
                errs.append(str(e))
        res["set_batch_errors"] = errs
        with contextlib.redirect_stdout(io.StringIO()):
            v2.set_batch(["t0", "t1", "t2"], enc[:3], [{"i": 0}, {"i": 1}, {"i": 2}])
        res["set_batch_count"] = v2.count()
        res["set_batch_key_suffixes"] = sorted({k.rsplit("_", 1)[1] for k in v2.index})
    with open(os.path.join(HERE, "collection_golden.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("collection_golden.json written")


def golden_ctx(ref_ctx):
    w = 64
    rows = to_ref_encoding(synth.corpus_rows(31, 0, 6, w))
    contexts = ["alpha", "béta ünïcode", "", "delta " * 40, "ε", "last"]
    path = os.path.join(HERE, "ref_written.ctx")
    if os.path.exists(path):
        os.remove(path)
    cf = ref_ctx.CtxFile(path)
    cf.set_header(embedding_model="mixedbread-ai/mxbai-embed-large-v1", embedding_size=64,
                  embedding_dtype="float32", context_length=512)
    for i in range(6):
        cf.add_embedding(rows[i].tolist() + [999.0] * (i % 2))   # >64 values: truncated (ctx.py:70)
        cf.add_context(contexts[i])
        cf.add_metadata(f"id{i}_0", {"n": i, "s": contexts[i][:3]})
    cf.save()
    back = ref_ctx.CtxFile(path)
    back.load()
    facts = {
        "sha256": hashlib.sha256(open(path, "rb").read()).hexdigest(),
        "header": back.header,
        "embeddings": [[float(x) for x in e] for e in back.embeddings],
        "contexts": back.contexts,
        "metadata": back.metadata,
        "expected_u8_hex": synth.corpus_rows(31, 0, 6, w).tobytes().hex(),
    }
    # save-on-exit of an existing file APPENDS again (ctx.py:157-162)
    with tempfile.TemporaryDirectory() as td:
        p2 = os.path.join(td, "twice.ctx")
        for _ in range(2):
            with ref_ctx.CtxFile(p2) as c2:
                c2.set_header("m", 64, "float32", 512)
                for i in range(2):
                    c2.add_embedding(rows[i].tolist())
                    c2.add_context(contexts[i])
                    c2.add_metadata(f"k{i}_0", {})
        c3 = ref_ctx.CtxFile(p2)
        c3.load()
        facts["double_save_counts"] = [len(c3.embeddings), len(c3.contexts), len(c3.metadata)]
    # error strings
    errs = {}
    with tempfile.TemporaryDirectory() as td:
        for name, blob in [("bad_magic", b"XXXX" + struct.pack("<I", 1)),
                           ("bad_version", b"CTXF" + struct.pack("<I", 7)),
                           ("bad_section", b"CTXF" + struct.pack("<I", 1) + struct.pack("<II", 9, 0))]:
            p = os.path.join(td, name)
            open(p, "wb").write(blob)
            try:
                ref_ctx.CtxFile(p).load()
                errs[name] = None
            except ValueError as e:
                errs[name] = str(e)
        c4 = ref_ctx.CtxFile(os.path.join(td, "missing.ctx"))
        c4.load()                                  # swallowed (ctx.py:138-139)
        errs["missing_ok"] = (c4.embeddings == [] and c4.contexts == [])
    facts["errors"] = errs
    with open(path + ".json", "w") as f:
        json.dump(facts, f, indent=1)
    print("ref_written.ctx:", os.path.getsize(path), "bytes")


def fixture_facts(ref_ctx):
    path = os.path.join(REF, "tests/contexts/vlite-bench.ctx")
    blob = open(path, "rb").read()
    facts = {"size": len(blob), "sha256": hashlib.sha256(blob).hexdigest()}
    off, sections = 8, []
    while off + 8 <= len(blob):
        typ, ln = struct.unpack_from("<II", blob, off)
        sections.append([typ, off + 8, ln])
        off += 8 + ln
    facts["sections"] = sections
    cf = ref_ctx.CtxFile(path)
    cf.load()
    facts["header"] = cf.header
    facts["loader_rows_x64"] = len(cf.embeddings)        # current loader misreads 512-wide rows as 64-wide
    facts["n_contexts"] = len(cf.contexts)
    facts["n_metadata"] = len(cf.metadata)
    facts["first_metadata_keys"] = list(cf.metadata.keys())[:3]
    facts["first_context_sha256"] = hashlib.sha256(cf.contexts[0].encode("utf-8")).hexdigest()
    emb = np.asarray(cf.embeddings, dtype=np.float32).reshape(-1, 512)   # legacy: 512 floats of 0/1 per row
    facts["legacy_rows"] = int(emb.shape[0])
    facts["legacy_values"] = sorted({float(x) for x in np.unique(emb)})
    facts["legacy_bit_density"] = float(emb.mean())
    packed = np.packbits((emb > 0.5).astype(np.uint8), axis=1)
    facts["legacy_unique_rows"] = int(len(np.unique(packed, axis=0)))
    facts["legacy_first8_packed_hex"] = packed[:8].tobytes().hex()
    facts["legacy_packed_sha256"] = hashlib.sha256(packed.tobytes()).hexdigest()
    with open(os.path.join(HERE, "fixture_facts.json"), "w") as f:
        json.dump(facts, f, indent=1)
    print("fixture_facts.json:", facts["legacy_rows"], "rows,", facts["n_contexts"], "contexts")


def golden_rescore():
    """The reference has no rescore stage; its only dense similarity is ``cos_sim`` (vlite/utils.py:205-208):
    ``a @ b.T`` divided by the norms.  The rescore here is that numerator over int8 document rows, so
    cos_sim's outputs on seeded queries / documents pin the oracle's dot products (divided by the same
    norms) and the ranking they induce."""
    import vlite.utils as ref_utils                    # importable once import_reference() has stubbed its deps
    rng = np.random.default_rng(77)
    dims, n_docs = 1024, 300
    docs_i8 = rng.integers(-128, 128, size=(n_docs, dims), dtype=np.int8)
    docs_i8[17] = docs_i8[5]                           # an exact tie: the rank test must break it by row
    q_f32 = rng.standard_normal((4, dims)).astype(np.float32)
    q_i8 = rng.integers(-128, 128, size=(2, dims), dtype=np.int8)
    out = {"docs_i8": docs_i8, "queries_f32": q_f32, "queries_i8": q_i8}
    for name, qs in (("f32", q_f32), ("i8", q_i8)):
        for qi, q in enumerate(qs):
            out[f"{name}_q{qi}_cos_f32"] = np.asarray(ref_utils.cos_sim(q.astype(np.float32), docs_i8.astype(np.float32)))
            out[f"{name}_q{qi}_cos_f64"] = np.asarray(ref_utils.cos_sim(q.astype(np.float64), docs_i8.astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, "rescore_golden.npz"), **out)
    print("rescore_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    ref_model, ref_main, ref_ctx = import_reference()
    if sys.argv[1:3] == ["--only", "rescore"]:
        golden_rescore()
        sys.exit(0)
    # the reference drops a telemetry id file under ./contexts on every save/load
    # (ctx.py:58, :89): run from a scratch cwd so nothing lands in the repo
    _scratch = tempfile.mkdtemp()
    os.makedirs(os.path.join(_scratch, "contexts"))
    os.chdir(_scratch)
    golden_search(ref_model)
    golden_collection(ref_model, ref_main)
    golden_ctx(ref_ctx)
    fixture_facts(ref_ctx)
    golden_rescore()
