"""ctypes wrapper over oracle_scan.c (the fast CPU checker).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "oracle_scan.c")
_LIB = os.path.join(_HERE, "_build", "liboracle_scan.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/_build/ (git-ignored, travels via gpurun)."""
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    # no -march=native: the .so is built in one container and run on another box
    cmd = ["gcc", "-O3", "-mpopcnt", "-fopenmp", "-shared", "-fPIC", _SRC, "-o", _LIB]
    subprocess.run(cmd, check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        L.oracle_scores.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int, u8p, ctypes.c_int, u32p]
        L.oracle_scores.restype = None
        for fn in (L.oracle_search, L.oracle_search_rowpar):
            fn.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int,
                           ctypes.c_int, u32p, ctypes.c_uint64, u32p, u64p]
            fn.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def scores(q: np.ndarray, corpus: np.ndarray, metric: int) -> np.ndarray:
    corpus = np.ascontiguousarray(corpus, dtype=np.uint8)
    q = np.ascontiguousarray(q, dtype=np.uint8)
    n, w = corpus.shape
    out = np.empty(n, dtype=np.uint32)
    lib().oracle_scores(_p(corpus, ctypes.c_uint8), n, w, _p(q, ctypes.c_uint8), metric,
                        _p(out, ctypes.c_uint32))
    return out


def search(queries: np.ndarray, corpus: np.ndarray, k: int, metric: int = 0,
           mask_words: np.ndarray | None = None, base_row: int = 0, rowpar: bool | None = None):
    """(scores uint32 (Q,k), rows uint64 (Q,k)), sentinel padded -- same contract
    as oracle.search.search / vl_search."""
    corpus = np.ascontiguousarray(corpus, dtype=np.uint8)
    queries = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.uint8)
    n, w = corpus.shape
    nq = queries.shape[0]
    assert queries.shape[1] == w
    out_s = np.empty((nq, k), dtype=np.uint32)
    out_r = np.empty((nq, k), dtype=np.uint64)
    mp = None
    if mask_words is not None:
        mask_words = np.ascontiguousarray(mask_words, dtype=np.uint32)
        assert mask_words.shape[0] * 32 >= n
        mp = _p(mask_words, ctypes.c_uint32)
    L = lib()
    if rowpar is None:
        rowpar = nq < L.oracle_num_threads()
    fn = L.oracle_search_rowpar if rowpar else L.oracle_search
    fn(_p(corpus, ctypes.c_uint8), n, w, _p(queries, ctypes.c_uint8), nq, k, metric, mp,
       base_row, _p(out_s, ctypes.c_uint32), _p(out_r, ctypes.c_uint64))
    return out_s, out_r


def num_threads() -> int:
    return lib().oracle_num_threads()
