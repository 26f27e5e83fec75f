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
  * the row width follows ``header["embedding_size"]`` when it is 64 or 128 (the reference writes
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
    return size if size in (64, 128, 512) else LEGACY_ROW_FLOATS


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

    # ---- writer (vlite/ctx.py:55-82) ----
    def save(self):
        width = row_floats_for(self.header)
        if width == 512:
            width = LEGACY_ROW_FLOATS
        elem = np.dtype("u1") if is_ubinary(self.header) else np.dtype("<f4")
        with open(self.file_path, "wb") as f:
            f.write(self.MAGIC_NUMBER)
            f.write(struct.pack("<I", self.VERSION))
            header_json = json.dumps(self.header).encode("utf-8")
            f.write(struct.pack("<II", CtxSectionType.HEADER.value, len(header_json)))
            f.write(header_json)
            n_emb = len(self.embeddings)
            if n_emb:                                        # section omitted when empty (ctx.py:68)
                if isinstance(self.embeddings, np.ndarray):
                    arr = np.ascontiguousarray(self.embeddings[:, :width], dtype=elem)
                else:
                    rows = []
                    for emb in self.embeddings:
                        r = np.asarray(emb, dtype=elem).reshape(-1)[:width]   # emb[:64] (ctx.py:70)
                        if r.shape[0] != width:
                            raise struct.error(f"pack expected {width} items for packing (got {r.shape[0]})")
                        rows.append(r)
                    arr = np.stack(rows)
                data_len = arr.shape[0] * width * elem.itemsize
                if data_len >= 1 << 32:
                    raise ValueError("EMBEDDINGS section exceeds the u32 length field of CTXF v1 "
                                     "(at most 16.7M rows of 64 floats); split the collection")
                f.write(struct.pack("<II", CtxSectionType.EMBEDDINGS.value, data_len))
                f.write(arr.tobytes())
            contexts_data = b"".join(
                struct.pack("<I", len(c.encode("utf-8"))) + c.encode("utf-8") for c in self.contexts)
            f.write(struct.pack("<II", CtxSectionType.CONTEXTS.value, len(contexts_data)))
            f.write(contexts_data)
            metadata_json = json.dumps(self.metadata).encode("utf-8")
            f.write(struct.pack("<II", CtxSectionType.METADATA.value, len(metadata_json)))
            f.write(metadata_json)

    # ---- reader (vlite/ctx.py:85-139) ----
    def load(self, materialise: bool = True):
        """Parse the file.  ``materialise=False`` leaves the EMBEDDINGS payload on disk (only its
        offset/length are recorded in ``embeddings_section``) for streaming to the device."""
        try:
            f = open(self.file_path, "rb")
        except FileNotFoundError:
            return                                           # swallowed, like vlite/ctx.py:138-139
        with f:
            magic = f.read(len(self.MAGIC_NUMBER))
            if magic != self.MAGIC_NUMBER:
                raise ValueError(f"Invalid magic number: {magic}")
            version = struct.unpack("<I", f.read(4))[0]
            if version != self.VERSION:
                raise ValueError(f"Unsupported version: {version}")
            while True:
                section_header = f.read(8)
                if not section_header:
                    break
                section_type, section_length = struct.unpack("<II", section_header)
                if section_type == CtxSectionType.HEADER.value:
                    self.header = json.loads(f.read(section_length).decode("utf-8"))
                elif section_type == CtxSectionType.EMBEDDINGS.value:
                    self.embeddings_section = (f.tell(), section_length)
                    if materialise and section_length:
                        width = row_floats_for(self.header)
                        elem = np.dtype("u1") if is_ubinary(self.header) else np.dtype("<f4")
                        data = f.read(section_length)
                        n = len(data) // (width * elem.itemsize)         # floor (ctx.py:115)
                        self._emb_array = np.frombuffer(data, dtype=elem, count=n * width).reshape(n, width)
                        self.embeddings = self._emb_array
                    else:
                        f.seek(section_length, os.SEEK_CUR)
                elif section_type == CtxSectionType.CONTEXTS.value:
                    data = f.read(section_length)
                    self.contexts = []
                    off = 0
                    while off < len(data):
                        (clen,) = struct.unpack_from("<I", data, off)
                        off += 4
                        try:
                            self.contexts.append(data[off:off + clen].decode("utf-8"))
                        except UnicodeDecodeError:
                            pass                             # logged and skipped (ctx.py:130-131)
                        off += clen
                elif section_type == CtxSectionType.METADATA.value:
                    self.metadata = json.loads(f.read(section_length).decode("utf-8"))
                else:
                    raise ValueError(f"Unknown section type: {section_type}")

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
        if not os.path.exists(directory):
            os.makedirs(directory)

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
