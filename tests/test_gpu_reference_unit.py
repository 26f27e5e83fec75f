"""-m gpu: the scenarios of the reference's own ``tests/unit.py`` (add / retrieve / delete / update /
get / set / count / clear), driven through the mirror with a deterministic stand-in for the Hugging
Face model, and asserting what unit.py asserts -- plus the ranks and scores unit.py never checks
(every retrieve is compared with the oracle's search over the texts' vectors)."""
import hashlib

import numpy as np
import pytest

from oracle import search as osearch

pytestmark = pytest.mark.gpu


def text_embedder(texts):
    """text -> 1024 floats from a hash of the text: equal texts embed equally, nothing else does"""
    out = np.empty((len(texts), 1024), dtype=np.float32)
    for i, t in enumerate(texts):
        seed = int.from_bytes(hashlib.sha256(t.encode("utf-8")).digest()[:8], "little")
        out[i] = np.random.default_rng(seed).standard_normal(1024)
    return out


def vec(text):
    return osearch.quantize_binary(text_embedder([text])[:, :512])[0]          # model.py:39-44


@pytest.fixture()
def vlite(gpu, tmp_path):
    from vlite_b200 import VLite
    v = VLite("vlite-unit", ctx_dir=str(tmp_path / "contexts"), embedder=text_embedder)
    yield v
    v.close()


def test_add_text_and_texts(vlite):
    metadata = {"source": "test", "author": "John Doe", "timestamp": "2023-06-08"}
    result = vlite.add("This is a test text. " * 86, metadata=metadata, item_id="test_text_1", need_chunks=False)
    assert vlite.count() == 1 and len(result) == 1
    assert result[0][0] == "test_text_1" and result[0][2] == metadata
    assert result[0][1].shape == (1, 64) and result[0][1].min() >= -256 and result[0][1].max() <= -1
    r2 = vlite.add("some other words entirely", metadata={"source": "b"}, item_id="test_text_2")
    r3 = vlite.add(["a list", "of three", "texts"], item_id="test_text_3")
    assert vlite.count() == 5
    assert r2[0][0] == "test_text_2" and r3[0][0] == "test_text_3" and len(r3) == 1
    assert {"test_text_3_0", "test_text_3_1", "test_text_3_2"} <= set(vlite.dump().keys())


def test_retrieve(vlite):
    docs = [f"document number {i} about topic {i % 7}" for i in range(40)]
    for i, d in enumerate(docs):
        vlite.add(d, metadata={"topic": i % 7}, item_id=f"doc{i}")
    corpus = np.stack([vec(d) for d in docs])
    for query in ("document number 3 about topic 3", "an unrelated question?", docs[17]):
        results = vlite.retrieve(query, top_k=3, return_scores=True)
        assert len(results) == 3
        assert isinstance(results[0][1], str) and isinstance(results[0][2], dict)
        es, er = osearch.search(vec(query)[None], corpus, 3, osearch.XORSUM)      # the reference's live score
        assert [r[0] for r in results] == [f"doc{int(i)}_0" for i in er[0]]
        assert [int(r[3]) for r in results] == [int(s) for s in es[0]]
    exact = vlite.retrieve(docs[17], top_k=1, return_scores=True)[0]
    assert exact[0] == "doc17_0" and int(exact[3]) == 0 and exact[1] == docs[17]
    only3 = vlite.retrieve(docs[17], top_k=5, metadata={"topic": 3})
    assert only3 and all(md["topic"] == 3 for _, _, md in only3)
    assert vlite.retrieve("") is None                                             # main.py:109-110


def test_delete_update_get_set_count_clear(vlite):
    vlite.add("This is a test text.", item_id="test_delete_1")
    vlite.add("Another test text.", item_id="test_delete_2")
    assert vlite.delete(["test_delete_1", "test_delete_2"]) == 2
    assert vlite.count() == 0 and vlite.delete("test_delete_1") == 0

    vlite.add("This is a test text.", item_id="test_update_1")
    assert vlite.update("test_update_1", text="This is an updated text.", metadata={"updated": True})
    assert not vlite.update("no_such_item", text="x")
    assert vlite.count() == 1

    vlite.add("Text 1", item_id="test_get_1", metadata={"category": "A"})
    vlite.add("Text 2", item_id="test_get_2", metadata={"category": "B"})
    vlite.add("Text 3", item_id="test_get_3", metadata={"category": "A"})
    items = vlite.get(ids=["test_get_1", "test_get_3"])
    assert len(items) == 2 and items[0][1] == "Text 1" and items[1][1] == "Text 3"
    items = vlite.get(where={"category": "A"})
    assert len(items) == 2 and all(it[2]["category"] == "A" for it in items)

    vlite.add("Original text", item_id="test_set_1", metadata={"original": True})
    vlite.set("test_set_1", text="Updated text", metadata={"updated": True})
    items = vlite.get(ids=["test_set_1"])
    assert len(items) == 1 and items[0][1] == "Updated text" and items[0][2]["updated"] is True
    vlite.set("test_set_2", text="Created by set", metadata={"fresh": 1})           # unknown id: set adds
    assert vlite.get(ids=["test_set_2"])[0][1] == "Created by set"

    n = vlite.count()
    assert n == 6
    vlite.clear()
    assert vlite.count() == 0
