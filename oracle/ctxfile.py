"""Restatement of the CTXF v1 container.  TEST INFRASTRUCTURE ONLY.

Follows vlite/ctx.py: section ids (:16-20), writer (:55-82), reader (:85-139).
Layout: ``b"CTXF"``, ``<I`` version (=1), then sections ``<II`` (type, length)
followed by ``length`` payload bytes.  HEADER and METADATA payloads are JSON;
EMBEDDINGS is rows of 64 little-endian float32 (fixed, ctx.py:70 and :114);
CONTEXTS is a run of (``<I`` byte length, utf-8 bytes).
"""
from __future__ import annotations

import json
import struct

import numpy as np

MAGIC = b"CTXF"
VERSION = 1
SEC_HEADER, SEC_EMBEDDINGS, SEC_CONTEXTS, SEC_METADATA = 0, 1, 2, 3
ROW_FLOATS = 64  # ctx.py:70, :114


def write_ctx(path, header: dict, embeddings, contexts, metadata: dict) -> None:
    """What ``CtxFile.save`` writes (ctx.py:55-82): each embedding truncated to
    its first 64 values and packed as float32; the EMBEDDINGS section is
    omitted when there are no embeddings (ctx.py:68)."""
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", VERSION))
        hj = json.dumps(header).encode("utf-8")
        f.write(struct.pack("<II", SEC_HEADER, len(hj)))
        f.write(hj)
        if len(embeddings):
            rows = [np.asarray(e[:ROW_FLOATS], dtype="<f4") for e in embeddings]
            data = b"".join(r.tobytes() for r in rows)
            f.write(struct.pack("<II", SEC_EMBEDDINGS, len(data)))
            f.write(data)
        cd = b"".join(struct.pack("<I", len(c.encode("utf-8"))) + c.encode("utf-8") for c in contexts)
        f.write(struct.pack("<II", SEC_CONTEXTS, len(cd)))
        f.write(cd)
        mj = json.dumps(metadata).encode("utf-8")
        f.write(struct.pack("<II", SEC_METADATA, len(mj)))
        f.write(mj)


def read_ctx(path, row_floats: int = ROW_FLOATS):
    """What ``CtxFile.load`` reads (ctx.py:85-139).  Returns
    (header, embeddings float32 (n, row_floats), contexts list[str], metadata
    dict, sections list[(type, offset, length)]).  Raises ValueError on bad
    magic / version / unknown section like ctx.py:95-99, :137."""
    header, emb, contexts, metadata, sections = {}, np.zeros((0, row_floats), np.float32), [], {}, []
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:4] != MAGIC:
        raise ValueError(f"Invalid magic number: {blob[:4]}")
    (ver,) = struct.unpack_from("<I", blob, 4)
    if ver != VERSION:
        raise ValueError(f"Unsupported version: {ver}")
    off = 8
    while off + 8 <= len(blob):
        typ, ln = struct.unpack_from("<II", blob, off)
        off += 8
        payload = blob[off:off + ln]
        sections.append((typ, off, ln))
        if typ == SEC_HEADER:
            header = json.loads(payload.decode("utf-8"))
        elif typ == SEC_EMBEDDINGS:
            n = len(payload) // (row_floats * 4)        # ctx.py:115 floor division
            emb = np.frombuffer(payload[:n * row_floats * 4], dtype="<f4").reshape(n, row_floats).copy()
        elif typ == SEC_CONTEXTS:
            contexts, o = [], 0
            while o < len(payload):
                (cl,) = struct.unpack_from("<I", payload, o)
                o += 4
                try:
                    contexts.append(payload[o:o + cl].decode("utf-8"))
                except UnicodeDecodeError:
                    pass                                 # ctx.py:130-131 logs and skips
                o += cl
        elif typ == SEC_METADATA:
            metadata = json.loads(payload.decode("utf-8"))
        else:
            raise ValueError(f"Unknown section type: {typ}")
        off += ln
    return header, emb, contexts, metadata, sections


def embeddings_to_ubinary(emb_f32: np.ndarray, kind: str) -> np.ndarray:
    """Stored float rows -> raw ubinary bytes.
    kind "offset": values are the reference encoding in [-256, -1] (or raw
    [0, 255]) one per byte -> (n, cols) uint8.
    kind "bits": values are 0.0/1.0, one per BIT (the legacy fixture
    tests/contexts/vlite-bench.ctx, 512 floats per row) -> packbits MSB-first."""
    from .search import from_ref_encoding
    if kind == "offset":
        return from_ref_encoding(emb_f32)
    if kind == "bits":
        return np.packbits((np.asarray(emb_f32) > 0.5).astype(np.uint8), axis=1)
    raise ValueError(kind)
