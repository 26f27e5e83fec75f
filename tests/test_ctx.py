"""CPU-only tests of the CTXF v1 reader/writer mirror against files the reference wrote."""
import json
import os
import struct

import numpy as np
import pytest

from vlite_b200.ctx import Ctx, CtxFile, CtxSectionType, row_floats_for

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reads_reference_written_file():
    facts = json.load(open(os.path.join(GOLD, "ref_written.ctx.json")))
    cf = CtxFile(os.path.join(GOLD, "ref_written.ctx"))
    cf.load()
    assert cf.header == facts["header"]
    assert cf.contexts == facts["contexts"]
    assert cf.metadata == facts["metadata"]
    np.testing.assert_array_equal(np.asarray(cf.embeddings), np.asarray(facts["embeddings"], dtype=np.float32))
    off, ln = cf.embeddings_section
    assert ln == 6 * 64 * 4 and cf.n_embeddings == 6
    lazy = CtxFile(cf.file_path)
    lazy.load(materialise=False)
    assert lazy.embeddings_section == (off, ln) and lazy.n_embeddings == 6 and len(lazy.embeddings) == 0


def test_writer_is_byte_identical_to_reference(tmp_path):
    src = os.path.join(GOLD, "ref_written.ctx")
    cf = CtxFile(src)
    cf.load()
    out = CtxFile(str(tmp_path / "copy.ctx"))
    out.set_header(**cf.header)
    for row, ctx_ in zip(cf.embeddings, cf.contexts):
        out.add_embedding([float(x) for x in row] + [123.0])      # >64 values are truncated (ctx.py:70)
        out.add_context(ctx_)
    for k, v in cf.metadata.items():
        out.add_metadata(k, v)
    out.save()
    assert open(out.file_path, "rb").read() == open(src, "rb").read()
    bulk = CtxFile(str(tmp_path / "bulk.ctx"))
    bulk.set_header(**cf.header)
    bulk.set_embeddings_array(np.asarray(cf.embeddings))
    bulk.contexts, bulk.metadata = list(cf.contexts), dict(cf.metadata)
    bulk.save()
    assert open(bulk.file_path, "rb").read() == open(src, "rb").read()


def test_context_manager_loads_then_appends_like_reference(tmp_path):
    facts = json.load(open(os.path.join(GOLD, "ref_written.ctx.json")))
    p = str(tmp_path / "twice.ctx")
    for _ in range(2):
        with CtxFile(p) as c:
            c.set_header("m", 64, "float32", 512)
            for i in range(2):
                c.add_embedding([float(-1 - i)] * 64)
                c.add_context(f"c{i}")
                c.add_metadata(f"k{i}_0", {})
    c = CtxFile(p)
    c.load()
    assert [len(c.embeddings), len(c.contexts), len(c.metadata)] == facts["double_save_counts"]


def test_errors_match_reference_strings(tmp_path):
    facts = json.load(open(os.path.join(GOLD, "ref_written.ctx.json")))["errors"]
    for name, blob, key in [("m", b"XXXX" + struct.pack("<I", 1), "bad_magic"),
                            ("v", b"CTXF" + struct.pack("<I", 7), "bad_version"),
                            ("s", b"CTXF" + struct.pack("<I", 1) + struct.pack("<II", 9, 0), "bad_section")]:
        p = tmp_path / name
        p.write_bytes(blob)
        with pytest.raises(ValueError) as ei:
            CtxFile(str(p)).load()
        assert str(ei.value) == facts[key]
    missing = CtxFile(str(tmp_path / "missing.ctx"))
    missing.load()                                   # swallowed like vlite/ctx.py:138-139
    assert missing.embeddings == [] and missing.contexts == []
    short = CtxFile(str(tmp_path / "short.ctx"))
    short.add_embedding([0.0] * 32)
    with pytest.raises(struct.error):                # the reference cannot pack short rows either
        short.save()


def test_empty_collection_omits_embeddings_section(tmp_path):
    c = CtxFile(str(tmp_path / "e.ctx"))
    c.save()
    blob = open(c.file_path, "rb").read()
    types, off = [], 8
    while off < len(blob):
        t, ln = struct.unpack_from("<II", blob, off)
        types.append(t)
        off += 8 + ln
    assert types == [CtxSectionType.HEADER.value, CtxSectionType.CONTEXTS.value, CtxSectionType.METADATA.value]


def test_width_extension_and_dir_helper(tmp_path):
    assert row_floats_for({"embedding_size": 64}) == 64
    assert row_floats_for({"embedding_size": 128}) == 128
    assert row_floats_for({"embedding_size": 512}) == 512          # legacy fixture: one float per bit
    assert row_floats_for({"embedding_size": 0}) == 64 and row_floats_for({}) == 64
    ctx = Ctx(str(tmp_path / "contexts"))
    assert os.path.isdir(ctx.directory)
    f = ctx.create("u")
    assert f.file_path == os.path.join(ctx.directory, "u.ctx") and ctx.read("u").file_path == f.file_path
    f.set_header("m", 128, "float32", 512)
    f.set_embeddings_array(np.full((3, 128), -7.0, dtype=np.float32))
    f.save()
    g = ctx.read("u")
    g.load()
    assert np.asarray(g.embeddings).shape == (3, 128)
    ctx.delete("u")
    assert not os.path.exists(f.file_path)
    ctx.delete("u")                                   # deleting a missing file is a no-op
