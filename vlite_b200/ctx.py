"""CTXF v1 container -- host-side mirror of the reference's ``vlite/ctx.py`` (same names,
argument meaning and error behaviour), with the EMBEDDINGS payload kept as bytes so it can be
streamed to HBM (``vl_corpus_load_file``) instead of being unpacked row by row into Python
lists (vlite/ctx.py:116-119).

Byte layout (vlite/ctx.py:55-82): ``b"CTXF"``, ``<I`` version = 1, then sections
``<II`` (type, length) + payload: HEADER json, EMBEDDINGS rows of float32, CONTEXTS
(``<I`` len + utf-8)*, METADATA json.  A file this module writes with 64-byte rows is byte for
byte what the reference writes for the same collection (tests/test_ctx.py checks it against a
file the reference itself wrote).

Differences from the reference, all deliberate:
  * no telemetry (the reference posts to PostHog and drops a ``.telm`` id file on every load/save,
    vlite/ctx.py:38, :58-59, :89-90);
  * the row width follows ``header["embedding_size"]`` when it is 32, 64 or 128 (the reference writes
    that field, vlite/main.py:285, but always reads and writes 64 floats, vlite/ctx.py:70, :114);
    512 is accepted as the legacy one-float-per-bit layout of tests/contexts/vlite-bench.ctx;
  * ``header["embedding_dtype"] == "ubinary"`` (opt-in extension, not readable by the reference):
    rows are stored as raw packed bytes, ``embedding_size`` bytes per row -- a quarter of the
    float32 layout, and what the device holds, so loading is a straight copy;
  * a ``with`` block still loads on enter and saves on exit (vlite/ctx.py:157-162); ``VLite.save``
    writes a fresh ``CtxFile`` without entering it, so a collection is not appended to itself.
"""
from __future__ import annotations

import json
import os
import struct
from enum import Enum
from typing import Dict, List, Union

import numpy as np

LEGACY_ROW_FLOATS = 64   # vlite/ctx.py:70, :114


class CtxSectionType(Enum):       # vlite/ctx.py:16-20
    HEADER = 0
    EMBEDDINGS = 1
    CONTEXTS = 2
    METADATA = 3


def is_ubinary(header: dict) -> bool:
    return isinstance(header, dict) and header.get("embedding_dtype") == "ubinary"


def row_floats_for(header: dict) -> int:
    """How many float32 one stored row holds."""
    size = header.get("embedding_size", 0) if isinstance(header, dict) else 0
    return size if size in (32, 64, 128, 512) else LEGACY_ROW_FLOATS


class CtxFile:
    MAGIC_NUMBER = b"CTXF"
    VERSION = 1

    def __init__(self, file_path):
        self.file_path = file_path
        self.header = {                      # vlite/ctx.py:27-32
            "embedding_model": "default",
            "embedding_size": 0,
            "embedding_dtype": "float32",
            "context_length": 0,
        }
        self.embeddings: List = []           # rows (lists / arrays of numbers), like the reference
        self.contexts: List[str] = []
        self.metadata: Dict = {}
        # where the EMBEDDINGS payload of the loaded file lives: (offset, length) or None
        self.embeddings_section = None
        self._emb_array = None               # float32 (n, row_floats) view of the loaded payload

    # ---- builders (vlite/ctx.py:40-53) ----
    def set_header(self, embedding_model: str, embedding_size: int, embedding_dtype: str, context_length: int):
        self.header["embedding_model"] = embedding_model
        self.header["embedding_size"] = embedding_size
        self.header["embedding_dtype"] = embedding_dtype
        self.header["context_length"] = context_length

    def add_embedding(self, embedding: List[float]):
        if self._emb_array is not None:      # materialise loaded rows before mixing in new ones
            self.embeddings = [r for r in self._emb_array]
            self._emb_array = None
        self.embeddings.append(embedding)

    def set_embeddings_array(self, rows_f32: np.ndarray):
        """Bulk alternative to add_embedding: an (n, row_floats) float32 array written as is."""
        self.embeddings = rows_f32
        self._emb_array = None

    def add_context(self, context: str):
        self.contexts.append(context)

    def add_metadata(self, key: str, value: Union[int, float, str]):
        self.metadata[key] = value

    # ---- writer: byte layout of vlite/ctx.py:55-82 ----
    def _embedding_payload(self) -> bytes:
        """Rows as one contiguous block; every row is cut to the stored width like ``emb[:64]``
        (vlite/ctx.py:70) and a shorter row is the same struct.error the reference raises."""
        width = row_floats_for(self.header)
        width = LEGACY_ROW_FLOATS if width == 512 else width
        elem = np.dtype("u1") if is_ubinary(self.header) else np.dtype("<f4")
        if isinstance(self.embeddings, np.ndarray):
            block = np.ascontiguousarray(self.embeddings[:, :width], dtype=elem)
        else:
            block = np.empty((len(self.embeddings), width), dtype=elem)
            for i, emb in enumerate(self.embeddings):
                row = np.asarray(emb, dtype=elem).reshape(-1)[:width]
                if row.shape[0] != width:
                    raise struct.error(f"pack expected {width} items for packing (got {row.shape[0]})")
                block[i] = row
        if block.nbytes >= 1 << 32:
            raise ValueError("EMBEDDINGS section exceeds the u32 length field of CTXF v1 "
                             "(at most 16.7M rows of 64 floats); split the collection")
        return block.tobytes()

    def save(self):
        def utf8(text):
            return text.encode("utf-8")

        sections = [(CtxSectionType.HEADER, utf8(json.dumps(self.header)))]
        if len(self.embeddings):                              # section omitted when empty (ctx.py:68)
            sections.append((CtxSectionType.EMBEDDINGS, self._embedding_payload()))
        packed_contexts = bytearray()
        for text in self.contexts:
            raw = utf8(text)
            packed_contexts += struct.pack("<I", len(raw)) + raw
        sections.append((CtxSectionType.CONTEXTS, bytes(packed_contexts)))
        sections.append((CtxSectionType.METADATA, utf8(json.dumps(self.metadata))))
        with open(self.file_path, "wb") as out:
            out.write(self.MAGIC_NUMBER + struct.pack("<I", self.VERSION))
            for kind, payload in sections:
                out.write(struct.pack("<II", kind.value, len(payload)))
                out.write(payload)

    # ---- reader: accepts exactly what vlite/ctx.py:85-139 accepts ----
    def load(self, materialise: bool = True):
        """Parse the file.  ``materialise=False`` leaves the EMBEDDINGS payload on disk (only its
        offset/length are recorded in ``embeddings_section``) for streaming to the device.
        A missing file is not an error (vlite/ctx.py:138-139); bad magic, version or section type
        raise the reference's ValueErrors (vlite/ctx.py:95-99, :137)."""
        if not os.path.exists(self.file_path):
            return
        size = os.path.getsize(self.file_path)
        with open(self.file_path, "rb") as src:
            preamble = src.read(8)
            if preamble[:4] != self.MAGIC_NUMBER:
                raise ValueError(f"Invalid magic number: {preamble[:4]}")
            (version,) = struct.unpack("<I", preamble[4:8])
            if version != self.VERSION:
                raise ValueError(f"Unsupported version: {version}")
            pos = 8
            while pos + 8 <= size:
                src.seek(pos)
                kind, length = struct.unpack("<II", src.read(8))
                body = pos + 8
                pos = body + length
                if kind == CtxSectionType.EMBEDDINGS.value:
                    self.embeddings_section = (body, length)
                    if materialise and length:
                        self._materialise(src.read(length))
                elif kind == CtxSectionType.HEADER.value:
                    self.header = json.loads(src.read(length).decode("utf-8"))
                elif kind == CtxSectionType.CONTEXTS.value:
                    self.contexts = self._split_contexts(src.read(length))
                elif kind == CtxSectionType.METADATA.value:
                    self.metadata = json.loads(src.read(length).decode("utf-8"))
                else:
                    raise ValueError(f"Unknown section type: {kind}")

    def _materialise(self, payload: bytes):
        width = row_floats_for(self.header)
        elem = np.dtype("u1") if is_ubinary(self.header) else np.dtype("<f4")
        n = len(payload) // (width * elem.itemsize)               # whole rows only (ctx.py:115)
        self._emb_array = np.frombuffer(payload, dtype=elem, count=n * width).reshape(n, width)
        self.embeddings = self._emb_array

    @staticmethod
    def _split_contexts(payload: bytes) -> List[str]:
        """(u32 length, utf-8 bytes)*; an undecodable string is dropped (ctx.py:130-131)."""
        out, pos = [], 0
        while pos < len(payload):
            (n,) = struct.unpack_from("<I", payload, pos)
            pos += 4
            try:
                out.append(payload[pos:pos + n].decode("utf-8"))
            except UnicodeDecodeError:
                pass
            pos += n
        return out

    @property
    def n_embeddings(self) -> int:
        if self.embeddings_section is not None and self._emb_array is None and not len(self.embeddings):
            return self.embeddings_section[1] // (row_floats_for(self.header) * (1 if is_ubinary(self.header) else 4))
        return len(self.embeddings)

    def __repr__(self):
        out = "CtxFile:\n\nHeader:\n"
        for k, v in self.header.items():
            out += f"  {k}: {v}\n"
        out += f"\nEmbeddings: {self.n_embeddings} rows\n\nContexts: {len(self.contexts)}\n\nMetadata: {len(self.metadata)} keys\n"
        return out

    def __enter__(self):                                     # vlite/ctx.py:157-159
        self.load()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):           # vlite/ctx.py:161-162
        self.save()


class Ctx:                                                   # vlite/ctx.py:164-184
    def __init__(self, directory="contexts"):
        self.directory = directory
        os.makedirs(directory, exist_ok=True)      # several ranks may create it at once

    def get(self, user):
        return os.path.join(self.directory, f"{user}.ctx")

    def create(self, user: str) -> CtxFile:
        return CtxFile(self.get(user))

    def read(self, user_id: str) -> CtxFile:
        return CtxFile(self.get(user_id))

    def delete(self, user_id: str):
        path = self.get(user_id)
        if os.path.exists(path):
            os.remove(path)
