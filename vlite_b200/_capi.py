"""ctypes binding of libvlite_b200.so (the C ABI declared in include/vlite_b200.h).

There is no CPU fallback: if the library is missing or no sm_100 device is
usable, compute calls raise.  ``load()`` only dlopens the library and checks
that every symbol the header declares is exported, so it works on a CPU-only
box (that is what the ``-m "not gpu"`` tests exercise).
"""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# VLITE_B200_LIB: load an experiment build instead of the shipped library (tuning only)
LIB_PATH = os.environ.get("VLITE_B200_LIB") or os.path.join(_PKG, "lib", "libvlite_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "vlite_b200.h")

HAMMING, XORSUM = 0, 1
SRC_U8, SRC_F32_OFFSET, SRC_F32_BITS = 0, 1, 2
QTYPE_F32, QTYPE_I8 = 0, 1
SENTINEL_SCORE = 0xFFFFFFFF
SENTINEL_ROW = 0xFFFFFFFFFFFFFFFF
MAX_K = 2048
EINVAL, ECUDA, ENOMEM, ENODEV, EIO, ERANGE = -1, -2, -3, -4, -5, -6

_lib = None


class VliteB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vlite_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


def declared_symbols() -> list[str]:
    """Every function name include/vlite_b200.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vl_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen the in-tree library (building it with nvcc if absent or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("VLITE_B200_LIB") or LIB_PATH      # e.g. the -DVL_DEBUG_BOUNDS build
    if path == LIB_PATH and not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    L = ctypes.CDLL(path)
    c_void_p, c_int, c_u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    P = ctypes.POINTER
    sig = {
        "vl_version": (ctypes.c_char_p, []),
        "vl_last_error": (ctypes.c_char_p, []),
        "vl_device_count": (c_int, []),
        "vl_launch_count": (c_u64, []),
        "vl_corpus_create": (c_int, [c_int, c_int, c_u64, P(c_void_p)]),
        "vl_corpus_destroy": (c_int, [c_void_p]),
        "vl_corpus_set_stream": (c_int, [c_void_p, c_void_p]),
        "vl_corpus_reserve": (c_int, [c_void_p, c_u64]),
        "vl_corpus_append": (c_int, [c_void_p, c_void_p, c_u64, P(c_u64)]),
        "vl_corpus_append_f32": (c_int, [c_void_p, c_void_p, c_u64, c_int, P(c_u64)]),
        "vl_corpus_load_file": (c_int, [c_void_p, ctypes.c_char_p, c_u64, c_u64, c_int, c_u64, P(c_u64), P(c_u64)]),
        "vl_corpus_append_sign_i8": (c_int, [c_void_p, c_void_p, c_u64, P(c_u64)]),
        "vl_corpus_set_row": (c_int, [c_void_p, c_u64, c_void_p]),
        "vl_corpus_tombstone": (c_int, [c_void_p, c_void_p, c_u64]),
        "vl_corpus_clear": (c_int, [c_void_p]),
        "vl_corpus_rows": (c_int, [c_void_p, P(c_u64), P(c_u64)]),
        "vl_corpus_width": (c_int, [c_void_p]),
        "vl_corpus_set_base_row": (c_int, [c_void_p, c_u64]),
        "vl_corpus_read": (c_int, [c_void_p, c_u64, c_u64, c_void_p]),
        "vl_corpus_device_ptrs": (c_int, [c_void_p, P(c_void_p), P(c_void_p)]),
        "vl_corpus_fill_synthetic": (c_int, [c_void_p, c_u64, c_u64, c_u64, c_u64]),
        "vl_synth_filter_mask": (c_int, [c_void_p, c_u64, c_u64, c_void_p]),
        "vl_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
        "vl_scores": (c_int, [c_void_p, c_void_p, c_int, c_u64, c_u64, c_void_p]),
        "vl_merge_topk": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_u64, c_u64,
                                  c_void_p, c_void_p]),
        "vl_map_rows": (c_int, [c_int, c_void_p, c_void_p, c_u64, c_void_p, c_void_p, c_int]),
        "vl_rescore": (c_int, [c_void_p, c_void_p, c_u64, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                               c_void_p, c_void_p]),
        "vl_shard_group_create": (c_int, [c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
        "vl_shard_group_destroy": (c_int, [c_void_p]),
        "vl_shard_group_set_segments": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int]),
        "vl_shard_group_rows": (c_int, [c_void_p, ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]),
        "vl_shard_group_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p),
                                          c_void_p, c_void_p]),
        "vl_quantize_binary": (c_int, [c_int, c_void_p, c_void_p, c_u64, c_int, c_void_p]),
        "vl_quantize_embeddings": (c_int, [c_int, c_void_p, c_void_p, c_u64, c_int, c_int, ctypes.c_float,
                                           c_void_p, c_void_p, c_void_p]),
        "vl_synth_docs_i8": (c_int, [c_int, c_void_p, c_u64, c_u64, c_u64, c_int, c_void_p]),
        "vl_microbench": (c_int, [c_int, c_int, c_int, P(ctypes.c_double), P(ctypes.c_double)]),
        "vl_last_timing": (c_int, [c_void_p, P(ctypes.c_float), P(ctypes.c_float)]),
        "vl_set_timing": (c_int, [c_void_p, c_int]),
        "vl_set_tuning": (c_int, [c_void_p, c_int, c_int]),
        "vl_set_scan_path": (c_int, [c_void_p, c_int]),
        "vl_debug_checks": (c_int, [c_int, P(ctypes.c_uint32), P(ctypes.c_uint32)]),
        "vl_exchange_create": (c_int, [c_int, c_int, c_int, c_u64, P(c_void_p)]),
        "vl_exchange_handle": (c_int, [c_void_p, c_void_p]),
        "vl_exchange_attach": (c_int, [c_void_p, c_int, c_void_p]),
        "vl_exchange_gather_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_u64, c_int, c_int, c_void_p, c_void_p]),
        "vl_exchange_destroy": (c_int, [c_void_p]),
        "vl_search_embeddings": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise VliteB200Error(rc, load().vl_last_error().decode("utf-8", "replace"))


def ptr(x) -> ctypes.c_void_p:
    """Address of a numpy array (host), a torch tensor (host or device), an int, or None."""
    if x is None:
        return ctypes.c_void_p(None)
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch.Tensor
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError(f"cannot take the address of {type(x)}")


class Corpus:
    """Thin object wrapper over a ``vl_corpus*`` handle (one GPU-resident shard)."""

    def __init__(self, width_bytes: int = 128, device: int = 0, capacity_rows: int = 0, base_row: int = 0):
        self._L = load()
        self._h = ctypes.c_void_p()
        check(self._L.vl_corpus_create(device, width_bytes, capacity_rows, ctypes.byref(self._h)))
        self.width = width_bytes
        self.device = device
        if base_row:
            self.set_base_row(base_row)

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.vl_corpus_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state --
    @property
    def rows(self) -> int:
        n, live = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._L.vl_corpus_rows(self._h, ctypes.byref(n), ctypes.byref(live)))
        return n.value

    @property
    def live_rows(self) -> int:
        n, live = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._L.vl_corpus_rows(self._h, ctypes.byref(n), ctypes.byref(live)))
        return live.value

    def set_stream(self, cuda_stream: int | None):
        check(self._L.vl_corpus_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def set_base_row(self, base_row: int):
        check(self._L.vl_corpus_set_base_row(self._h, base_row))
        self.base_row = base_row

    def reserve(self, capacity_rows: int):
        check(self._L.vl_corpus_reserve(self._h, capacity_rows))

    def device_ptrs(self):
        r, m = ctypes.c_void_p(), ctypes.c_void_p()
        check(self._L.vl_corpus_device_ptrs(self._h, ctypes.byref(r), ctypes.byref(m)))
        return r.value, m.value

    # -- ingest --
    def append(self, rows) -> int:
        """rows: (n, W) uint8 raw ubinary (numpy or torch, host or device). Returns first row."""
        n = int(rows.shape[0])
        if n and (rows.ndim != 2 or int(rows.shape[1]) != self.width):
            raise ValueError(f"rows must be (n, {self.width}) uint8")
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=np.uint8)
        first = ctypes.c_uint64()
        check(self._L.vl_corpus_append(self._h, ptr(rows) if n else None, n, ctypes.byref(first)))
        return first.value

    def append_f32(self, vals, kind: int) -> int:
        cols = self.width if kind == SRC_F32_OFFSET else self.width * 8
        n = int(vals.shape[0])
        if n and (vals.ndim != 2 or int(vals.shape[1]) != cols):
            raise ValueError(f"values must be (n, {cols}) float32")
        if isinstance(vals, np.ndarray):
            vals = np.ascontiguousarray(vals, dtype=np.float32)
        first = ctypes.c_uint64()
        check(self._L.vl_corpus_append_f32(self._h, ptr(vals) if n else None, n, kind, ctypes.byref(first)))
        return first.value

    def append_sign_i8(self, docs_i8_dev, n: int) -> int:
        """Append the sign bits of n device-resident int8 document rows (n x 8*W int8)."""
        first = ctypes.c_uint64()
        check(self._L.vl_corpus_append_sign_i8(self._h, ptr(docs_i8_dev), n, ctypes.byref(first)))
        return first.value

    def load_file(self, path: str, byte_off: int, byte_len: int, kind: int, max_rows: int = 0):
        first, n = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._L.vl_corpus_load_file(self._h, os.fsencode(path), byte_off, byte_len, kind, max_rows,
                                          ctypes.byref(first), ctypes.byref(n)))
        return first.value, n.value

    def set_row(self, row: int, data):
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        if data.shape[0] != self.width:
            raise ValueError("row has the wrong width")
        check(self._L.vl_corpus_set_row(self._h, row, ptr(data)))

    def tombstone(self, rows) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.uint64).reshape(-1)
        check(self._L.vl_corpus_tombstone(self._h, ptr(rows) if len(rows) else None, len(rows)))

    def clear(self):
        check(self._L.vl_corpus_clear(self._h))

    def read(self, first: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = self.rows - first
        out = np.empty((n, self.width), dtype=np.uint8)
        check(self._L.vl_corpus_read(self._h, first, n, ptr(out) if n else None))
        return out

    def fill_synthetic(self, seed: int, first: int, n: int, period: int = 0):
        check(self._L.vl_corpus_fill_synthetic(self._h, seed, first, n, period))

    def synth_filter_mask(self, seed: int, threshold: int, mask_dev):
        check(self._L.vl_synth_filter_mask(self._h, seed, threshold, ptr(mask_dev)))

    # -- query --
    def search(self, queries, k: int, metric: int = HAMMING, filter_mask=None, out_scores=None, out_rows=None):
        """queries (Q, W) uint8.  With numpy inputs returns (scores uint32 (Q,k), rows uint64 (Q,k))
        host arrays; device tensors may be passed for queries / filter_mask / outputs, in which
        case the call only enqueues work on the handle's stream."""
        if isinstance(queries, np.ndarray):
            queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.uint8)
        nq = int(queries.shape[0])
        if int(queries.shape[1]) != self.width:
            raise ValueError(f"queries must be (Q, {self.width}) uint8")
        if isinstance(filter_mask, np.ndarray):
            filter_mask = np.ascontiguousarray(filter_mask, dtype=np.uint32)
            if filter_mask.shape[0] * 32 < self.rows:
                raise ValueError("filter mask is shorter than the corpus")
        if k < 1:
            raise VliteB200Error(EINVAL, f"k must be in [1, {MAX_K}]")
        if out_scores is None:
            out_scores = np.empty((nq, k), dtype=np.uint32)
        if out_rows is None:
            out_rows = np.empty((nq, k), dtype=np.uint64)
        check(self._L.vl_search(self._h, ptr(queries), nq, k, metric, ptr(filter_mask), ptr(out_scores), ptr(out_rows)))
        return out_scores, out_rows

    def search_embeddings(self, x, k: int, metric: int = HAMMING, filter_mask=None, out_scores=None, out_rows=None):
        """x (Q, D) float32 embeddings (numpy or a CUDA tensor), D >= 8 * width: sign-quantised on
        the device (first 8 * width features, MSB-first like np.packbits) and searched in one call."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
        nq, dims = int(x.shape[0]), int(x.shape[1])
        if isinstance(filter_mask, np.ndarray):
            filter_mask = np.ascontiguousarray(filter_mask, dtype=np.uint32)
            if filter_mask.shape[0] * 32 < self.rows:
                raise ValueError("filter mask is shorter than the corpus")
        if k < 1:
            raise VliteB200Error(EINVAL, f"k must be in [1, {MAX_K}]")
        if out_scores is None:
            out_scores = np.empty((nq, k), dtype=np.uint32)
        if out_rows is None:
            out_rows = np.empty((nq, k), dtype=np.uint64)
        check(self._L.vl_search_embeddings(self._h, ptr(x), nq, dims, k, metric, ptr(filter_mask), ptr(out_scores),
                                           ptr(out_rows)))
        return out_scores, out_rows

    def scores(self, query, metric: int = HAMMING, first: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = self.rows - first
        query = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1)
        out = np.empty(n, dtype=np.uint32)
        check(self._L.vl_scores(self._h, ptr(query), metric, first, n, ptr(out) if n else None))
        return out

    def rescore(self, docs_i8_dev, n_docs: int, dims: int, queries, cand_rows, k: int, out_scores=None, out_rows=None,
                gather_probe: bool = False):
        """docs_i8_dev: device tensor / address of (n_docs, dims) int8.  queries (Q, dims) float32
        or int8; cand_rows (Q, C) uint64 global rows.  gather_probe (int8 queries): measurement only --
        the same gathers without the dot products (VL_QTYPE_GATHER_PROBE)."""
        if isinstance(queries, np.ndarray):
            queries = np.ascontiguousarray(queries)
            qtype = QTYPE_I8 if queries.dtype == np.int8 else QTYPE_F32
            if qtype == QTYPE_F32:
                queries = np.ascontiguousarray(queries, dtype=np.float32)
        else:
            qtype = QTYPE_I8 if "int8" in str(queries.dtype) else QTYPE_F32
        if isinstance(cand_rows, np.ndarray):
            cand_rows = np.ascontiguousarray(cand_rows, dtype=np.uint64)
        if gather_probe:
            qtype = 2
        nq, nc = int(cand_rows.shape[0]), int(cand_rows.shape[1])
        if out_scores is None:
            out_scores = np.empty((nq, k), dtype=np.float32)
        if out_rows is None:
            out_rows = np.empty((nq, k), dtype=np.uint64)
        check(self._L.vl_rescore(self._h, ptr(docs_i8_dev), n_docs, dims, ptr(queries), qtype, ptr(cand_rows),
                                 nq, nc, k, ptr(out_scores), ptr(out_rows)))
        return out_scores, out_rows

    # -- measurement --
    def set_timing(self, enabled: bool):
        check(self._L.vl_set_timing(self._h, int(enabled)))

    def last_timing(self):
        t, s = ctypes.c_float(), ctypes.c_float()
        check(self._L.vl_last_timing(self._h, ctypes.byref(t), ctypes.byref(s)))
        return t.value, s.value

    def set_tuning(self, row_splits: int = 0, stages: int = 0):
        check(self._L.vl_set_tuning(self._h, row_splits, stages))

    def set_scan_path(self, path: int = 0):
        """0 auto, 1 integer-pipe kernel only, 2 / 3 tensor-core kernel (1 / 2 CTAs per query tile)."""
        check(self._L.vl_set_scan_path(self._h, path))


class ShardGroup:
    """Several ``Corpus`` shards (usually one per GPU) searched as one corpus from this process
    (``vl_shard_group_*``).  The group borrows the shards; their base rows must be disjoint."""

    def __init__(self, shards):
        self._L = load()
        self.shards = list(shards)
        arr = (ctypes.c_void_p * len(self.shards))(*[c._h.value for c in self.shards])
        self._h = ctypes.c_void_p()
        check(self._L.vl_shard_group_create(len(self.shards), arr, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.vl_shard_group_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def rows(self) -> int:
        n, live = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._L.vl_shard_group_rows(self._h, ctypes.byref(n), ctypes.byref(live)))
        return n.value

    @property
    def live_rows(self) -> int:
        n, live = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._L.vl_shard_group_rows(self._h, ctypes.byref(n), ctypes.byref(live)))
        return live.value

    def set_segments(self, shard: int, seg_local, seg_global):
        sl = np.ascontiguousarray(seg_local, dtype=np.uint64)
        sg = np.ascontiguousarray(seg_global, dtype=np.uint64)
        if sl.shape != sg.shape:
            raise ValueError("segment tables differ in length")
        check(self._L.vl_shard_group_set_segments(self._h, shard, ptr(sl), ptr(sg), int(sl.shape[0])))

    def search(self, queries, k: int, metric: int = HAMMING, filter_masks=None):
        """queries (Q, W) uint8 host array; filter_masks: None or one uint32 mask (or None) per
        shard in that shard's local row numbering.  Returns host (scores (Q,k), rows (Q,k))."""
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.uint8)
        if q.shape[1] != self.shards[0].width:
            raise ValueError("query width does not match the corpus")
        if k < 1:
            raise VliteB200Error(EINVAL, "k must be >= 1")
        nq = q.shape[0]
        scores = np.empty((nq, k), dtype=np.uint32)
        rows = np.empty((nq, k), dtype=np.uint64)
        masks = None
        keep = []
        if filter_masks is not None:
            if len(filter_masks) != len(self.shards):
                raise ValueError("one filter mask (or None) per shard")
            masks = (ctypes.c_void_p * len(self.shards))()
            for i, (m, c) in enumerate(zip(filter_masks, self.shards)):
                if m is None:
                    masks[i] = None
                    continue
                m = np.ascontiguousarray(m, dtype=np.uint32)
                if m.shape[0] * 32 < c.rows:
                    raise ValueError("filter mask shorter than its shard")
                keep.append(m)
                masks[i] = m.ctypes.data
        check(self._L.vl_shard_group_search(self._h, ptr(q), nq, int(k), int(metric), masks, ptr(scores), ptr(rows)))
        return scores, rows


class Exchange:
    """``vl_exchange_*``: the peer-store exchange of the one-process-per-GPU path.  ``handle()`` is
    published to every rank, ``attach(rank, handle)`` maps a peer's buffer (CUDA IPC)."""

    def __init__(self, device: int, world: int, rank: int, slot_bytes: int):
        self._L = load()
        self._h = ctypes.c_void_p()
        self.world, self.rank, self.slot_bytes = world, rank, slot_bytes
        check(self._L.vl_exchange_create(device, world, rank, slot_bytes, ctypes.byref(self._h)))

    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        check(self._L.vl_exchange_handle(self._h, buf))
        return buf.raw

    def attach(self, peer_rank: int, handle: bytes):
        check(self._L.vl_exchange_attach(self._h, peer_rank, ctypes.c_char_p(handle)))

    def gather_merge(self, stream, local_block_dev, block_bytes: int, n_queries: int, k: int, out_scores, out_rows):
        check(self._L.vl_exchange_gather_merge(self._h, ctypes.c_void_p(stream or 0), ptr(local_block_dev), block_bytes,
                                               n_queries, k, ptr(out_scores), ptr(out_rows)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.vl_exchange_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_topk(scores_dev, rows_dev, n_lists: int, n_queries: int, k_in: int, k_out: int,
               out_scores_dev, out_rows_dev, device: int = 0, stream: int | None = None,
               score_list_stride: int = 0, row_list_stride: int = 0):
    """Merge n_lists device-resident candidate lists; strides in elements (0 = dense)."""
    check(load().vl_merge_topk(device, ctypes.c_void_p(stream or 0), ptr(scores_dev), ptr(rows_dev), n_lists,
                               n_queries, k_in, k_out, score_list_stride, row_list_stride,
                               ptr(out_scores_dev), ptr(out_rows_dev)))


def map_rows(rows_dev, n: int, seg_local_dev, seg_global_dev, n_segments: int, device: int = 0,
             stream: int | None = None):
    """Rewrite local rows into global rows in place (device tensors); see vl_map_rows."""
    check(load().vl_map_rows(device, ctypes.c_void_p(stream or 0), ptr(rows_dev), n, ptr(seg_local_dev),
                             ptr(seg_global_dev), n_segments))


def quantize_binary(x, out=None, device: int = 0, stream: int | None = None):
    """(n, dims) float32 -> (n, dims/8) uint8 raw ubinary, sign quantisation on the device."""
    if isinstance(x, np.ndarray):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    n, dims = int(x.shape[0]), int(x.shape[1])
    if out is None:
        out = np.empty((n, dims // 8), dtype=np.uint8)
    check(load().vl_quantize_binary(device, ctypes.c_void_p(stream or 0), ptr(x), n, dims, ptr(out)))
    return out


def quantize_embeddings(x_dev, binary_dims: int, out_u8_dev, out_i8_dev=None, out_f32_dev=None,
                        i8_scale: float = 127.0, device: int = 0, stream: int | None = None):
    """Fused normalise + sign-pack (+ int8 / float32 copies) on device tensors; see
    ``vl_quantize_embeddings``.  x_dev: (n, dims) float32 CUDA tensor."""
    n, dims = int(x_dev.shape[0]), int(x_dev.shape[1])
    nul = ctypes.c_void_p(0)
    check(load().vl_quantize_embeddings(device, ctypes.c_void_p(stream or 0), ptr(x_dev), n, dims, int(binary_dims),
                                        float(i8_scale), ptr(out_u8_dev),
                                        ptr(out_i8_dev) if out_i8_dev is not None else nul,
                                        ptr(out_f32_dev) if out_f32_dev is not None else nul))
    return out_u8_dev


def synth_docs_i8(seed: int, first_row: int, n: int, dims: int, out_dev, device: int = 0, stream: int | None = None):
    check(load().vl_synth_docs_i8(device, ctypes.c_void_p(stream or 0), seed, first_row, n, dims, ptr(out_dev)))


def microbench(kind: int, iters: int = 4096, device: int = 0):
    ops, ms = ctypes.c_double(), ctypes.c_double()
    check(load().vl_microbench(device, kind, iters, ctypes.byref(ops), ctypes.byref(ms)))
    return ops.value, ms.value


def debug_checks(device: int = 0):
    """(failures, id of the first) from the -DVL_DEBUG_BOUNDS build; None in the shipped build."""
    n, first = ctypes.c_uint32(), ctypes.c_uint32()
    if load().vl_debug_checks(device, ctypes.byref(n), ctypes.byref(first)) != 0:
        return None
    return n.value, first.value


def device_count() -> int:
    return load().vl_device_count()


def launch_count() -> int:
    return load().vl_launch_count()


def version() -> str:
    return load().vl_version().decode()
