"""``VLite`` -- host-side mirror of the reference's collection API (``vlite/main.py``) over a
corpus that lives packed in HBM.

Same method names, argument meaning, return shapes and error strings as the reference class
(vlite/main.py:19-311); what changes is where the work happens:

  reference                                                  here
  ---------------------------------------------------------  -----------------------------------
  index = dict of Python lists, rebuilt into an (N, W)        rows appended in place to a device
  float32 array on EVERY query (main.py:143-146)              corpus (vl_corpus_append*); host keeps
                                                              chunk_id <-> row, texts, metadata
  EmbeddingModel.search per query (main.py:150)               one batched vl_search for all queries
  metadata post-filter over top_k*4 candidates, scores        filter bitmask applied INSIDE the scan
  mis-paired afterwards (main.py:159-165)                     (true filtered top-k); the reference's
                                                              behaviour is kept as filter_mode="reference"
  delete = dict delete + whole-file rewrite (main.py:190)     tombstone bit + whole-file rewrite
  CTX load: struct.unpack per row into lists (ctx.py:116)     EMBEDDINGS payload streamed to HBM and
                                                              repacked on the device (vl_corpus_load_file)

Embedding generation (the Hugging Face forward pass) stays on the reference path; pass
``embedder=`` to plug in any ``texts -> float array`` callable.  No telemetry is sent
(the reference posts to PostHog on init/load/save, main.py:17-24).
"""
from __future__ import annotations

import datetime
import functools
import logging
import os
import threading
from uuid import uuid4

import numpy as np

from . import _capi
from .ctx import Ctx, is_ubinary, row_floats_for
from .metaindex import MetaIndex
from .model import EmbeddingModel, from_reference_encoding, to_reference_encoding

logger = logging.getLogger(__name__)

METRICS = {"xorsum": _capi.XORSUM, "hamming": _capi.HAMMING}


def chop_and_chunk(text, max_seq_length=512, fast=False):
    """vlite/utils.py:25-48.  fast=True: slices of max_seq_length*4 characters.  The token-based
    path needs tiktoken (ingest is out of scope for this build; it is used if importable)."""
    if isinstance(text, str):
        text = [text]
    chunks = []
    for t in text:
        if fast:
            size = max_seq_length * 4
            chunks.extend([t[i:i + size] for i in range(0, len(t), size)])
        else:
            import tiktoken
            enc = tiktoken.get_encoding("cl100k_base")
            ids = enc.encode(t, disallowed_special=())
            if len(ids) <= max_seq_length:
                chunks.append(t)
            else:
                chunks.extend(enc.decode(ids[i:i + max_seq_length]) for i in range(0, len(ids), max_seq_length))
    return chunks


def _locked(method):
    """Serialise a public method on the instance lock: the reference shares ONE VLite between all
    server requests with no locking (vlite/server.py:14); here a handle owns a CUDA stream and
    grow-only workspaces, so concurrent callers take turns (re-entrant: set -> update -> save)."""
    @functools.wraps(method)
    def wrapper(self, *args, **kw):
        with self._lock:
            return method(self, *args, **kw)
    return wrapper


class _IndexView(dict):
    """What ``VLite.index`` / ``dump()`` expose: chunk_id -> {'text', 'metadata', 'binary_vector'}
    like the reference's dict (main.py:41), except that the vector is read back from the device
    when asked for instead of being held as a Python list."""

    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def __getitem__(self, chunk_id):
        return self._owner._entry(chunk_id)

    def get(self, chunk_id, default=None):
        return self._owner._entry(chunk_id) if chunk_id in self._owner._row_of else default

    def __contains__(self, chunk_id):
        return chunk_id in self._owner._row_of

    def __iter__(self):
        return iter(self._owner._row_of)

    def __len__(self):
        return len(self._owner._row_of)

    def keys(self):
        return self._owner._row_of.keys()

    def values(self):
        return (self._owner._entry(k) for k in self._owner._row_of)

    def items(self):
        return ((k, self._owner._entry(k)) for k in self._owner._row_of)

    def __repr__(self):
        return f"<VLite index: {len(self)} chunks on device>"


class VLite:
    def __init__(self, collection=None, device=None, model_name="mixedbread-ai/mxbai-embed-large-v1", *,
                 embedder=None, metric="xorsum", width=None, filter_mode="prefilter", ctx_dir="contexts",
                 gpu_index=0, gpus=None, journal=False):
        """collection / device / model_name: as vlite/main.py:20.  Keyword-only additions:
        embedder (texts -> float array), metric ("xorsum" = the reference's live score,
        "hamming" = true Hamming distance), width (bytes per row, default from the data: 64),
        filter_mode ("prefilter" | "reference"), ctx_dir, gpu_index, gpus (a list of device
        indices: the collection is row-sharded over them inside this process; results are the
        same as on one GPU), journal (False = the reference's persistence: update / delete / set_batch
        rewrite the whole collection file, vlite/main.py:181, :201, :271; True = they append an O(change)
        record to ``<collection>.journal`` instead -- no device read-back, no rewrite -- and ``save()``
        folds everything into clean, reference-readable CTX v1 files)."""
        if metric not in METRICS:
            raise ValueError(f"metric must be one of {sorted(METRICS)}")
        if filter_mode not in ("prefilter", "reference"):
            raise ValueError("filter_mode must be 'prefilter' or 'reference'")
        self._lock = threading.RLock()
        self.device = device or "cuda"
        if collection is None:
            collection = f"vlite_{datetime.datetime.now().strftime('%Y%m%d_%H%M%S')}"
        self.collection = f"{collection}"
        self.metric = METRICS[metric]
        self.filter_mode = filter_mode
        self.gpus = list(gpus) if gpus else None
        if self.gpus:
            gpu_index = self.gpus[0]
        self.gpu_index = gpu_index
        self.model = EmbeddingModel(model_name, device=self.device, embedder=embedder, gpu_index=gpu_index)
        if width:
            self.model.binary_dims = width * 8      # e.g. width=128: keep all 1024 dims (no MRL cut)
        self.ctx = Ctx(ctx_dir)
        self.max_rows_per_file = 1 << 26       # further capped by the u32 section length of CTXF v1
        # "float32" = the reference's layout (default); "ubinary" = packed bytes, 4x smaller, opt-in
        self.ctx_dtype = "float32"
        self._width = width
        self.journal = bool(journal)
        self._journal_seq = 0      # number of side files (<collection>.jN.ctx) the journal refers to
        self._replaying = False
        self._corpus = None
        self._row_of = {}          # chunk_id -> row (None: no stored vector)
        self._ids = []             # row -> chunk_id (None once deleted)
        self._texts = {}           # chunk_id -> text
        self._metas = {}           # chunk_id -> metadata dict
        self._extra = {}           # chunk_id -> {'vector': ...} written by update() (main.py:180)
        self._meta_index = MetaIndex()   # (key, value) -> row bitmap, for filter masks
        self._by_prefix = {}       # "{id}" -> {chunk_id: None} for every chunk_id that starts with "{id}_"
        self.index = _IndexView(self)
        self._load_collection()

    # ------------------------------------------------------------------ internals
    def _new_ids(self, n: int):
        """fresh item ids (uuid4 like vlite/main.py:76, :261)"""
        return [str(uuid4()) for _ in range(n)]

    def _adopt_width(self, width: int):
        """The collection's row width is fixed by its first vectors (or its file): queries and later
        add() calls must be embedded at the same width (the reference hard-codes 512 bits,
        vlite/model.py:39; here width=32/128 collections reopen without repeating ``width=``)."""
        self._width = self._width or width
        if getattr(self.model, "binary_dims", None) != self._width * 8:
            self.model.binary_dims = self._width * 8

    def _ensure_corpus(self, width: int):
        if self._corpus is None:
            self._adopt_width(width)
            if self.gpus and len(self.gpus) > 1:
                from .sharded import LocalShardedCorpus
                self._corpus = LocalShardedCorpus(self._width, self.gpus)
            else:
                self._corpus = _capi.Corpus(self._width, device=self.gpu_index)
        return self._corpus

    def _entry(self, chunk_id):
        row = self._row_of[chunk_id]
        if row is None:
            vec = np.zeros(self.model.dimension)                         # main.py:50
        else:
            vec = to_reference_encoding(self._corpus.read(row, 1)[0])
        e = {"text": self._texts[chunk_id], "metadata": self._metas[chunk_id], "binary_vector": vec}
        e.update(self._extra.get(chunk_id, {}))
        return e

    # The reference finds an item's chunks with ``chunk_id.startswith(f"{id}_")`` over the whole
    # index (vlite/main.py:173, :195, :215, :235): O(N) per id.  Here every chunk id is filed under
    # each of its underscore-terminated prefixes, in registration order, so the same match is a lookup.
    def _file_prefixes(self, chunk_id):
        p = chunk_id.find("_")
        while p != -1:
            self._by_prefix.setdefault(chunk_id[:p], {})[chunk_id] = None
            p = chunk_id.find("_", p + 1)

    def _unfile_prefixes(self, chunk_id):
        p = chunk_id.find("_")
        while p != -1:
            bucket = self._by_prefix.get(chunk_id[:p])
            if bucket is not None:
                bucket.pop(chunk_id, None)
                if not bucket:
                    del self._by_prefix[chunk_id[:p]]
            p = chunk_id.find("_", p + 1)

    def _chunks_of(self, id):
        """chunk ids that start with ``f"{id}_"``, in index order"""
        return list(self._by_prefix.get(f"{id}", ()))

    def _post(self, chunk_id, row):
        self._meta_index.set_row(self._metas[chunk_id], row)

    def _register(self, chunk_id, text, metadata, row):
        old = self._row_of.get(chunk_id)
        if old is not None:                                              # dict overwrite (main.py:91)
            self._corpus.tombstone([old])
            self._ids[old] = None
        if chunk_id not in self._row_of:
            self._file_prefixes(chunk_id)
        self._row_of[chunk_id] = row
        self._texts[chunk_id] = text
        self._metas[chunk_id] = metadata
        self._extra.pop(chunk_id, None)
        if row is not None:
            while len(self._ids) <= row:
                self._ids.append(None)
            self._ids[row] = chunk_id
            self._post(chunk_id, row)

    def _register_many(self, chunk_ids, texts, metadatas, first_row):
        """Bulk form of ``_register`` for rows first_row.. that all have stored vectors: dict and
        bitmap updates are batched (9.5 -> ~1 us per row).  Ids that already exist, or repeat inside
        the batch, take the one-by-one path so the overwrite rules stay in one place."""
        n = len(chunk_ids)
        if n == 0:
            return
        if len(set(chunk_ids)) != n or any(cid in self._row_of for cid in chunk_ids):
            for i in range(n):
                self._register(chunk_ids[i], texts[i], metadatas[i], first_row + i)
            return
        if len(self._ids) < first_row:
            self._ids.extend([None] * (first_row - len(self._ids)))
        if len(self._ids) == first_row:
            self._ids.extend(chunk_ids)
        else:                                                            # rows already reserved
            for i, cid in enumerate(chunk_ids):
                row = first_row + i
                if row < len(self._ids):
                    self._ids[row] = cid
                else:
                    self._ids.append(cid)
        for cid in chunk_ids:
            self._file_prefixes(cid)
        self._row_of.update(zip(chunk_ids, range(first_row, first_row + n)))
        self._texts.update(zip(chunk_ids, texts))
        self._metas.update(zip(chunk_ids, metadatas))
        self._meta_index.set_many(metadatas, first_row)

    def _append_vectors(self, vectors) -> int:
        """vectors: (n, W) in the reference encoding or raw bytes.  Returns the first row."""
        u8 = from_reference_encoding(np.asarray(vectors))
        u8 = np.ascontiguousarray(np.atleast_2d(u8))
        return self._ensure_corpus(u8.shape[1]).append(u8)

    def _load_collection(self):
        """vlite/main.py:42-53: index keyed by metadata.keys() order; embeddings[idx] and
        contexts[idx] are positional; surplus rows in the file are ignored.  A collection that was
        split over several files by ``save`` (header key "parts") is read back part by part."""
        parts, i = 1, 0
        while i < parts:
            cf = self.ctx.read(self._part_name(i))
            cf.load(materialise=False)
            if i == 0:
                parts = int(cf.header.get("parts", 1)) if isinstance(cf.header, dict) else 1
            self._load_part(cf)
            i += 1
        self._replay_journal()

    # ------------------------------------------------------------------ journal (deferred persistence)
    # The reference persists by rewriting the whole collection on every update / delete / set_batch.
    # With ``journal=True`` those calls append one JSON line each to ``<collection>.journal``:
    #   {"op": "del", "ids": [chunk ids]}
    #   {"op": "upd", "id": item id, "text": ..., "metadata": ..., "vector": ...}
    #   {"op": "add", "file": "<collection>.jN", "n": rows}   rows of a set_batch, written as a small
    #                                                         standalone CTX v1 file of just those rows
    # A reopen loads the last full save and replays the journal; ``save()`` writes clean v1 files (what a
    # reference reader loads) and removes the journal and its side files.
    def _journal_path(self) -> str:
        return self.ctx.get(self.collection)[: -len(".ctx")] + ".journal"

    def _journal_append(self, record: dict):
        import json
        os.makedirs(os.path.dirname(self._journal_path()) or ".", exist_ok=True)
        with open(self._journal_path(), "a", encoding="utf-8") as f:
            f.write(json.dumps(record, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)) + "\n")
            f.flush()
            os.fsync(f.fileno())

    def _persist(self, record_fn):
        """what the reference does with ``self.save()`` after a mutation"""
        if getattr(self, "_replaying", False):
            return
        if getattr(self, "journal", False):
            self._journal_append(record_fn())
        else:
            self.save()

    def _replay_journal(self):
        import json
        path = self._journal_path()
        if not os.path.exists(path):
            return
        self._replaying = True
        try:
            with open(path, "r", encoding="utf-8") as f:
                for line in f:
                    line = line.strip()
                    if not line:
                        continue
                    rec = json.loads(line)
                    if rec["op"] == "del":
                        self._delete_chunks(rec["ids"])
                    elif rec["op"] == "upd":
                        self.update(rec["id"], rec.get("text"), rec.get("metadata"), rec.get("vector"))
                    elif rec["op"] == "add":
                        cf = self.ctx.read(rec["file"])
                        cf.load(materialise=False)
                        self._load_part(cf)
                        self._journal_seq = max(self._journal_seq, int(rec["file"].rsplit(".j", 1)[1]) + 1)
        finally:
            self._replaying = False

    def _drop_journal(self):
        if os.path.exists(self._journal_path()):
            os.remove(self._journal_path())
        i = 0
        while True:
            name = f"{self.collection}.j{i}"
            if not os.path.exists(self.ctx.get(name)):
                if i >= self._journal_seq:
                    break
            else:
                self.ctx.delete(name)
            i += 1
        self._journal_seq = 0

    def _part_name(self, i: int) -> str:
        return self.collection if i == 0 else f"{self.collection}.part{i}"

    def _load_part(self, cf):
        keys = list(cf.metadata.keys())
        if not keys:
            return
        floats = row_floats_for(cf.header)
        n_emb = cf.n_embeddings
        n_take = min(len(keys), n_emb)
        if n_take:
            if is_ubinary(cf.header):
                kind = _capi.SRC_U8
            else:
                kind = _capi.SRC_F32_BITS if floats == 512 else _capi.SRC_F32_OFFSET
            width = floats // 8 if floats == 512 else floats
            off, length = cf.embeddings_section
            c = self._ensure_corpus(width)
            first, n = c.load_file(cf.file_path, off, length, kind, max_rows=n_take)
            assert n == n_take
        else:
            first = 0
        texts = [cf.contexts[idx] if idx < len(cf.contexts) else "" for idx in range(len(keys))]
        self._register_many(keys[:n_take], texts[:n_take], [cf.metadata.get(k, {}) for k in keys[:n_take]], first)
        for idx in range(n_take, len(keys)):                             # keys past the stored rows (main.py:50)
            self._register(keys[idx], texts[idx], cf.metadata.get(keys[idx], {}), None)

    def _filter_rows(self, metadata: dict):
        """Rows whose metadata satisfies all(md.get(k) == v) (main.py:162), found by walking the
        rows: only for terms the bitmap index cannot hold (unhashable values, NaN)."""
        items = list(metadata.items())
        rows = [r for cid, r in self._row_of.items() if r is not None
                and all(self._metas[cid].get(k) == v for k, v in items)]
        return np.unique(np.asarray(rows, dtype=np.int64))

    def _filter_mask(self, metadata: dict):
        """One bit per corpus row, uint32 words: the AND of the per-(key, value) bitmaps."""
        n = self._corpus.rows
        words = self._meta_index.mask(metadata, n)
        if words is None:
            bits = np.zeros((n + 31) // 32 * 32, dtype=bool)
            bits[self._filter_rows(metadata)] = True
            words = np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
        return words

    def _search(self, queries_u8, k, mask=None, floats=None):
        live = self._corpus.live_rows if self._corpus is not None else 0
        if live == 0 or (floats is None and queries_u8.shape[1] != self._corpus.width):
            raise ValueError("No valid binary vectors found for comparison.")     # main.py:149
        k = max(1, min(int(k), int(live)))                            # vlite/model.py:84
        if k > _capi.MAX_K:
            raise ValueError(f"top_k={k} exceeds the device top-k limit VL_MAX_K={_capi.MAX_K}")
        if floats is not None:
            return self._corpus.search_embeddings(floats, k, self.metric, mask)
        return self._corpus.search(queries_u8, k, self.metric, mask)

    # ------------------------------------------------------------------ API
    @_locked
    def add(self, data, metadata=None, item_id=None, need_chunks=False, fast=True):
        """vlite/main.py:61-104 -- including its quirks: one item_id is generated and reused for
        every item of the call, chunk ids run ``{item_id}_{n}`` over all chunks, the return value is
        a one-element list ``[(item_id, encoded_vectors, last_metadata)]``, and nothing is saved."""
        data = [data] if not isinstance(data, list) else data
        all_chunks, all_metadata = [], []
        item_metadata = {}
        for item in data:
            if isinstance(item, dict):
                text_content = item["text"]
                item_metadata = item.get("metadata", {})
            else:
                text_content = item
                item_metadata = {}
            if item_id is None:
                item_id = self._new_ids(1)[0]
            item_metadata.update(metadata or {})
            chunks = chop_and_chunk(text_content, fast=fast) if need_chunks else [text_content]
            all_chunks.extend(chunks)
            all_metadata.extend([item_metadata] * len(chunks))
        encoded = self.model.embed(all_chunks, precision="binary")
        if len(all_chunks):
            first = self._append_vectors(encoded)
            self._register_many([f"{item_id}_{idx}" for idx in range(len(all_chunks))], all_chunks, all_metadata, first)
        return [(item_id, encoded, all_metadata[-1] if all_metadata else item_metadata)]

    @_locked
    def retrieve(self, text=None, top_k=5, metadata=None, return_scores=False, top_k_multiplier=4):
        """vlite/main.py:108-128: every query row is ranked, ALL results are merged, sorted by
        score and cut to top_k -- a list of Q texts yields ONE list.  Falsy text returns None."""
        if not text:
            return None
        results = []
        for chunk_results in self._rank_texts(text, top_k, metadata, top_k_multiplier):
            results.extend(chunk_results)
        results.sort(key=lambda x: x[1])                 # stable, by score only (main.py:120)
        results = results[:top_k]
        if return_scores:
            return [(cid, self._texts[cid], self._metas[cid], score) for cid, score in results]
        return [(cid, self._texts[cid], self._metas[cid]) for cid, _ in results]

    @_locked
    def retrieve_batch(self, texts, top_k=5, metadata=None, return_scores=False, top_k_multiplier=4):
        """One result list PER query (the reference merges them, main.py:116-121).  One batched scan."""
        out = []
        for chunk_results in self._rank_texts(texts, top_k, metadata, top_k_multiplier):
            if return_scores:
                out.append([(cid, self._texts[cid], self._metas[cid], s) for cid, s in chunk_results])
            else:
                out.append([(cid, self._texts[cid], self._metas[cid]) for cid, _ in chunk_results])
        return out

    def _rank_texts(self, text, top_k, metadata, top_k_multiplier):
        """embed + rank.  On a single-GPU collection the float embeddings go straight to
        ``vl_search_embeddings``: quantisation (vlite/model.py:39-42) and scan are ONE native call
        with one copy in and one copy out, instead of a device round trip for the packed query."""
        c = self._corpus
        if isinstance(c, _capi.Corpus) and c.live_rows and self.model.binary_dims == c.width * 8:
            x = self.model.embed_float(text)
            if x.shape[1] >= c.width * 8:
                return self._rank_batch(None, top_k, metadata, top_k_multiplier, floats=x)
        queries = self.model.embed(text, precision="binary")
        return self._rank_batch(np.atleast_2d(queries), top_k, metadata, top_k_multiplier)

    def _rank_batch(self, queries, top_k, metadata, top_k_multiplier, floats=None):
        if self._corpus is None or self._corpus.live_rows == 0:
            raise ValueError("No valid binary vectors found for comparison.")     # main.py:149
        q = from_reference_encoding(queries) if floats is None else None
        if metadata and self.filter_mode == "reference":
            scores, rows = self._search(q, top_k * top_k_multiplier, floats=floats)   # main.py:134-137
            out = []
            for s_row, r_row in zip(scores, rows):
                ids = [self._ids[int(r)] for r in r_row if r != _capi.SENTINEL_ROW]
                kept = [cid for cid in ids
                        if all(self._metas[cid].get(k) == v for k, v in metadata.items())][:top_k]
                # main.py:165 pairs the surviving ids with the FIRST len(ids) scores
                out.append(list(zip(kept, [np.int64(s) for s in s_row[:len(kept)]])))
            return out
        mask = self._filter_mask(metadata) if metadata else None
        scores, rows = self._search(q, top_k, mask, floats=floats)
        out = []
        for s_row, r_row in zip(scores, rows):
            keep = r_row != _capi.SENTINEL_ROW
            out.append([(self._ids[int(r)], np.int64(s)) for s, r in zip(s_row[keep], r_row[keep])])
        return out

    @_locked
    def rank_and_filter(self, query_binary_vector, top_k, metadata=None, top_k_multiplier=4):
        """vlite/main.py:130-168 for ONE query vector: list of (chunk_id, score)."""
        q = np.array(query_binary_vector).reshape(1, -1)
        return self._rank_batch(q, top_k, metadata, top_k_multiplier)[0]

    @_locked
    def update(self, id, text=None, metadata=None, vector=None):
        """vlite/main.py:170-188 (the vector is stored under 'vector' and not searched, as there)."""
        chunk_ids = self._chunks_of(id)
        if not chunk_ids:
            return False
        if metadata is not None:
            # Metadata dicts can be SHARED: by the chunks of one add() (vlite/main.py:83), by every
            # row of a set_batch without metadatas (``[{}] * n``, vlite/main.py:249) or by a caller
            # that reuses one dict.  The reference mutates the shared object in place
            # (vlite/main.py:178), so every chunk holding it changes: the bitmap index must follow all
            # of them, not only this id's rows.  Read every old value before changing any.
            touched = [self._metas[cid] for cid in chunk_ids]      # (`id` is a parameter here)
            sharers = [cid for cid, md in self._metas.items() if any(md is t for t in touched)]
            for cid in sharers:
                row = self._row_of[cid]
                if row is not None:
                    replaced = {k: self._metas[cid][k] for k in metadata if k in self._metas[cid]}
                    self._meta_index.clear_row(replaced, row)
                    self._meta_index.set_row(metadata, row)
        for cid in chunk_ids:
            if text is not None:
                self._texts[cid] = text
            if metadata is not None:
                self._metas[cid].update(metadata)
            if vector is not None:
                self._extra.setdefault(cid, {})["vector"] = vector
        self._persist(lambda: {"op": "upd", "id": id, "text": text, "metadata": metadata, "vector": vector})
        return True

    @_locked
    def delete(self, ids):
        """vlite/main.py:190-205: prefix match on ``{id}_``; returns the number of chunks removed."""
        if isinstance(ids, str):
            ids = [ids]
        gone = []
        for id in ids:
            gone.extend(self._chunks_of(id))
        gone = list(dict.fromkeys(gone))
        deleted = self._delete_chunks(gone)
        if deleted > 0:
            self._persist(lambda: {"op": "del", "ids": gone})
        return deleted

    def _delete_chunks(self, chunk_ids):
        deleted, dead_rows = 0, []
        for cid in chunk_ids:
            if cid not in self._row_of:
                continue
            row = self._row_of.pop(cid)
            self._unfile_prefixes(cid)
            self._texts.pop(cid, None)
            self._metas.pop(cid, None)
            self._extra.pop(cid, None)
            if row is not None:
                dead_rows.append(row)
                self._ids[row] = None
            deleted += 1
        if dead_rows:
            self._corpus.tombstone(dead_rows)
        return deleted

    @_locked
    def get(self, ids=None, where=None):
        """vlite/main.py:207-231."""
        items = []
        if ids is not None:
            if isinstance(ids, str):
                ids = [ids]
            for id in ids:
                chunks, md = [], {}
                for cid in self._chunks_of(id):
                    chunks.append(self._texts[cid])
                    md.update(self._metas[cid])
                if chunks:
                    items.append((id, " ".join(chunks), md))
        else:
            for cid in self._row_of:
                items.append((cid.split("_")[0], self._texts[cid], self._metas[cid]))
        if where is not None:
            items = [it for it in items if all(it[2].get(k) == v for k, v in where.items())]
        return items

    @_locked
    def set(self, id, text=None, metadata=None, vector=None):
        """vlite/main.py:233-240."""
        if self._chunks_of(id):
            self.update(id, text, metadata, vector)
        else:
            self.add(text, metadata=metadata, item_id=id)

    @_locked
    def set_batch(self, texts, embeddings, metadatas=None):
        """vlite/main.py:243-274: precomputed embeddings, one fresh uuid per row, then save()."""
        if not isinstance(texts, list):
            texts = [texts]
        if metadatas is None:
            metadatas = [{}] * len(texts)      # ONE shared dict, like vlite/main.py:249 (see update())
        elif not isinstance(metadatas, list):
            metadatas = [metadatas]
        if len(texts) != len(embeddings):
            raise ValueError("The number of texts and embeddings must be the same.")
        if len(texts) != len(metadatas):
            raise ValueError("The number of texts and metadatas must be the same.")
        keys = []
        if len(texts):
            first = self._append_vectors(np.asarray(embeddings))
            keys = [f"{uid}_0" for uid in self._new_ids(len(texts))]
            self._register_many(keys, texts, metadatas, first)
        if self.journal and not self._replaying:
            if keys:
                # only the new rows go to disk: a standalone CTX v1 file named by the journal
                name = f"{self.collection}.j{self._journal_seq}"
                self._journal_seq += 1
                self._write_rows_file(name, keys, np.asarray(embeddings))
                self._journal_append({"op": "add", "file": name, "n": len(keys)})
        else:
            self.save()

    def _write_rows_file(self, name, keys, embeddings):
        """a complete CTX v1 file holding just these rows (given in the reference encoding or raw bytes)"""
        width = self._width or 64
        compact = self.ctx_dtype == "ubinary"
        cf = self.ctx.create(name)
        cf.set_header(embedding_model="mixedbread-ai/mxbai-embed-large-v1", embedding_size=width,
                      embedding_dtype="ubinary" if compact else self.model.embedding_dtype,
                      context_length=self.model.context_length)
        u8 = np.ascontiguousarray(np.atleast_2d(from_reference_encoding(np.asarray(embeddings))))
        cf.set_embeddings_array(u8 if compact else to_reference_encoding(u8).astype(np.float32))
        for k in keys:
            cf.add_context(self._texts[k])
            cf.add_metadata(k, self._metas[k])
        cf.save()

    @_locked
    def count(self):
        return len(self._row_of)

    @_locked
    def save(self, _all_rows=None):
        """vlite/main.py:280-294, written CLEAN: one row per live key in key order (the reference's
        context manager loads the old file first and appends the whole collection to it again,
        vlite/ctx.py:157-162; a reference reader loads a clean file identically).

        CTXF v1 frames a section with a u32 length, i.e. at most 4 GiB of float32 rows per file
        (16.7M rows of 64 floats).  A larger collection is written as several v1 files --
        ``<name>.ctx`` plus ``<name>.partN.ctx`` -- each a complete, reference-readable CTX file
        for its slice of the keys; the first header carries ``"parts": n``."""
        width = self._width or 64
        keys = list(self._row_of.keys())
        rows = [self._row_of[k] for k in keys]
        # a key without a stored vector (the file held fewer rows than keys) is kept by the reference
        # as np.zeros(dimension) (vlite/main.py:50) and saved as such (:289-293): write a zero row
        no_vec = [r is None for r in rows]
        compact = self.ctx_dtype == "ubinary"
        per_file = max(1, min(self.max_rows_per_file, ((1 << 32) - 1) // (width * (1 if compact else 4))))
        n_parts = max(1, -(-len(keys) // per_file))
        if _all_rows is None and keys and self._corpus is not None and self._corpus.rows:
            _all_rows = self._corpus.read()                                  # one D2H copy
        all_rows = _all_rows
        for part in range(n_parts):
            cf = self.ctx.create(self._part_name(part))
            cf.set_header(embedding_model="mixedbread-ai/mxbai-embed-large-v1", embedding_size=width,
                          embedding_dtype="ubinary" if compact else self.model.embedding_dtype,
                          context_length=self.model.context_length)
            if n_parts > 1:
                cf.header["parts"] = n_parts
                cf.header["part"] = part
            lo, hi = part * per_file, min(len(keys), (part + 1) * per_file)
            if hi > lo:
                missing = np.asarray(no_vec[lo:hi])
                sel = np.asarray([0 if r is None else r for r in rows[lo:hi]], dtype=np.int64)
                if all_rows is None:                                         # no row has a vector
                    block = np.zeros((hi - lo, width), dtype=np.uint8)
                else:
                    block = all_rows[sel]
                if compact:
                    block = block.copy() if missing.any() else block
                    block[missing] = 0
                    cf.set_embeddings_array(block)
                else:
                    vals = to_reference_encoding(block).astype(np.float32)
                    vals[missing] = 0.0
                    cf.set_embeddings_array(vals)
            for k in keys[lo:hi]:
                cf.add_context(self._texts[k])
                cf.add_metadata(k, self._metas[k])
            cf.save()
        stale = n_parts                                                      # parts left by a bigger past save
        while os.path.exists(self.ctx.get(self._part_name(stale))):
            self.ctx.delete(self._part_name(stale))
            stale += 1
        self._drop_journal()                                                 # everything is in the v1 files now

    @_locked
    def clear(self):
        """vlite/main.py:296-300."""
        if self._corpus is not None:
            self._corpus.clear()
        self._row_of.clear()
        self._ids.clear()
        self._texts.clear()
        self._metas.clear()
        self._extra.clear()
        self._meta_index.clear()
        self._by_prefix.clear()
        i = 0
        while i == 0 or os.path.exists(self.ctx.get(self._part_name(i))):
            self.ctx.delete(self._part_name(i))
            i += 1
        self._drop_journal()

    def info(self):
        print("[VLite.info] Collection Information:")
        print(f"[VLite.info] Items: {self.count()}")
        print(f"[VLite.info] Collection file: {self.collection}")
        print(f"[VLite.info] Embedding model: {self.model}")

    def __repr__(self):
        return f"VLite(collection={self.collection}, device={self.device}, model={self.model})"

    def dump(self):
        return self.index

    @_locked
    def close(self):
        """Release the device memory of the collection (not part of the reference API; the files on
        disk are untouched)."""
        if self._corpus is not None:
            self._corpus.close()
            self._corpus = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class ShardedVLite(VLite):
    """The same collection API over a corpus row-sharded across the GPUs of one box (one process
    per GPU, ``torch.distributed`` initialised by the launcher).  Driven SPMD: every rank makes the
    same calls with the same arguments; host-side ids, texts and metadata are replicated, vectors
    live on exactly one GPU each (``ShardedCorpus``: least-full appends, shard-local tombstones, one
    all-gather of the k candidates per search).  Every rank returns the same results; rank 0 alone
    writes the CTX file.  The reference has no multi-GPU counterpart."""

    def __init__(self, *args, group=None, backend_factory=None, **kw):
        self._group = group
        self._backend_factory = backend_factory
        super().__init__(*args, **kw)

    def _new_ids(self, n: int):
        """rank 0 draws the ids, every rank uses them: the replicated tables must agree"""
        import torch.distributed as dist
        ids = super()._new_ids(n)
        if dist.is_initialized() and dist.get_world_size(self._group) > 1:
            box = [ids]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self._group, 0) if self._group else 0,
                                       group=self._group)
            ids = box[0]
        return ids

    def _ensure_corpus(self, width: int):
        if self._corpus is None:
            from .sharded import ShardedCorpus
            self._adopt_width(width)
            backend = self._backend_factory(self._width) if self._backend_factory else None
            self._corpus = ShardedCorpus(self._width, device=self.gpu_index, group=self._group, backend=backend)
        return self._corpus

    def save(self, _all_rows=None):
        import torch.distributed as dist
        rows = self._corpus.read() if (self._corpus is not None and self._row_of) else None   # collective
        rank = dist.get_rank(self._group) if dist.is_initialized() else 0
        self._rank0_then_all(lambda: VLite.save(self, _all_rows=rows), rank)

    def _journal_append(self, record):
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_rank(self._group) == 0:      # rank 0 alone writes files
            super()._journal_append(record)

    def _write_rows_file(self, name, keys, embeddings):
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_rank(self._group) == 0:
            super()._write_rows_file(name, keys, embeddings)

    def _rank0_then_all(self, fn, rank):
        """Run ``fn`` on rank 0 only; its outcome is broadcast so that a failure (I/O error, bad
        value) raises on EVERY rank instead of leaving the others waiting in a barrier."""
        import torch.distributed as dist
        err = None
        if rank == 0:
            try:
                fn()
            except Exception as e:            # noqa: BLE001 -- re-raised below on every rank
                err = e
        if dist.is_initialized() and dist.get_world_size(self._group) > 1:
            box = [None if err is None else f"{type(err).__name__}: {err}"]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self._group, 0) if self._group else 0,
                                       group=self._group)
            if box[0] is not None and err is None:
                err = RuntimeError(f"rank 0 failed: {box[0]}")
        if err is not None:
            raise err

    def clear(self):
        import torch.distributed as dist
        rank = dist.get_rank(self._group) if dist.is_initialized() else 0
        if rank != 0:
            if self._corpus is not None:
                self._corpus.clear()
            for d in (self._row_of, self._texts, self._metas, self._extra, self._meta_index, self._by_prefix):
                d.clear()
            self._ids.clear()
        self._rank0_then_all(lambda: VLite.clear(self), rank)
