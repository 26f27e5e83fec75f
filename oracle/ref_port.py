"""Torch-op port of the reference's live CPU search, for TIMING.  TEST INFRASTRUCTURE ONLY.

``bench.py --impl reference`` and the ``cpu_baseline`` leg time this on the GPU
box's host cores (the reference itself is a Python package under
/root/reference and cannot travel to the box).  It performs the same tensor
operations in the same order as ``EmbeddingModel.search``
(vlite/model.py:76-90): float32 corpus tensor -> ``.int()`` -> XOR with the
query -> ``sum(dim=1)`` -> ``topk(largest=False)`` -> gather of the scores,
with torch's intra-op thread pool at its default (all host cores).  Validated
against the real reference function in tests/test_oracle.py via the golden
vectors (score lists equal; index lists equal after tie canonicalisation).
"""
from __future__ import annotations

import numpy as np
import torch


def search_port(query_embedding: np.ndarray, embeddings: np.ndarray, top_k: int, device: str = "cpu"):
    """One reference ``search`` call: q (W,), E (N, W) float32 -> (idx, scores)."""
    e = torch.from_numpy(embeddings).to(device).to(dtype=torch.float32)          # model.py:78
    q = torch.from_numpy(query_embedding).unsqueeze(0).to(device).to(dtype=torch.float32)  # :79
    d = torch.sum(q.int() ^ e.int(), dim=1)                                       # :82
    k = min(top_k, len(embeddings))                                               # :84
    idx = torch.topk(d, k, largest=False).indices.cpu().numpy()                   # :85
    sc = d[idx].cpu().numpy()                                                     # :88
    return idx, sc


def corpus_as_reference_f32(u8_rows: np.ndarray) -> np.ndarray:
    """raw ubinary rows -> the (N, W) float32 array ``rank_and_filter`` hands to
    ``search`` (vlite/main.py:146), holding the reference encoding
    ``int8(u8) - 128`` (vlite/model.py:44)."""
    from .search import to_ref_encoding
    return to_ref_encoding(u8_rows).astype(np.float32)


def dict_rebuild_port(index: dict, qlen: int) -> np.ndarray:
    """The per-query corpus rebuild of vlite/main.py:143-146 (two dict walks and
    a list-of-lists -> float32 conversion), reported separately by bench.py."""
    vecs = [item["binary_vector"] for item in index.values() if len(item["binary_vector"]) == qlen]
    _ = [i for i, item in enumerate(index.values()) if len(item["binary_vector"]) == qlen]
    return np.asarray(vecs, dtype=np.float32)
