"""Row-sharded search across the GPUs of one box: one process per GPU, one exchange step.

SURVEY.md section 8e: shard g owns the contiguous global rows
[g * ceil(N/G), (g+1) * ceil(N/G)); queries are replicated; every rank scans its
shard, keeps its local top-k, ONE all-gather of the k candidates per query
moves 12 * Q * k bytes per rank over NVLink, and the same deterministic
(score, global row) merge runs on every rank -- so the result does not depend
on the shard count.  The reference has no multi-GPU path at all (its whole
search is vlite/model.py:76-90 on one device).

``torch.distributed`` is plumbing only (NCCL on GPUs; the host logic is
backend-agnostic so the world_size-2 gloo tests can drive it on CPU with the
local search and merge swapped for the oracle's).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous row ranges; global row = base_row(rank) + local row."""
    total_rows: int
    world_size: int

    @property
    def rows_per_shard(self) -> int:
        return -(-self.total_rows // self.world_size) if self.world_size else 0

    def base_row(self, rank: int) -> int:
        return min(rank * self.rows_per_shard, self.total_rows)

    def shard_rows(self, rank: int) -> int:
        return max(0, min(self.total_rows, (rank + 1) * self.rows_per_shard) - self.base_row(rank))

    def owner(self, global_row: int) -> tuple[int, int]:
        """(rank, local row) of a global row."""
        r = min(global_row // self.rows_per_shard, self.world_size - 1)
        return r, global_row - self.base_row(r)


def _cuda_local_search(corpus, queries, k, metric, filter_mask, out_scores, out_rows):
    corpus.search(queries, k, metric, filter_mask, out_scores, out_rows)


def _cuda_merge(block, n_lists, n_queries, k, out_scores, out_rows, device, stream):
    """block: uint8 device tensor (n_lists, 12*Q*k): per rank [Q*k uint32 scores | Q*k uint64 rows]."""
    from . import _capi
    nb = n_queries * k
    base = block.data_ptr()
    _capi.merge_topk(base, base + nb * 4, n_lists, n_queries, k, k, out_scores, out_rows,
                     device=device, stream=stream,
                     score_list_stride=nb * 3, row_list_stride=(nb * 12) // 8)


class ShardedSearcher:
    """Search over a corpus row-sharded across the ranks of a process group."""

    def __init__(self, corpus, group=None, local_search=None, merge=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.corpus = corpus
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_search = local_search or _cuda_local_search
        self._merge = merge or _cuda_merge
        self.device = device if device is not None else (
            torch.device("cuda", corpus.device) if local_search is None else torch.device("cpu"))
        self._bufs = {}

    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            t = self.torch
            nb = nq * k
            if (nb * 4) % 8:
                raise ValueError("Q*k must be even so the row block stays 8-byte aligned")
            send = t.empty(nb * 12, dtype=t.uint8, device=self.device)
            recv = t.empty((self.world, nb * 12), dtype=t.uint8, device=self.device)
            out_s = t.empty((nq, k), dtype=t.int32, device=self.device)
            out_r = t.empty((nq, k), dtype=t.int64, device=self.device)
            self._bufs[key] = (send, recv, out_s, out_r)
        return self._bufs[key]

    def search(self, queries, k: int, metric: int = 0, filter_mask=None):
        """queries: (Q, W) uint8 tensor on self.device, identical on every rank.  Returns
        (scores int32 (Q,k), rows int64 (Q,k)) tensors holding the uint32/uint64 bit patterns --
        the same on every rank."""
        t = self.torch
        nq = int(queries.shape[0])
        nb = nq * k
        send, recv, out_s, out_r = self._buffers(nq, k)
        ls = send[: nb * 4].view(t.int32).view(nq, k)
        lr = send[nb * 4:].view(t.int64).view(nq, k)
        self._local_search(self.corpus, queries, k, metric, filter_mask, ls, lr)
        if self.world == 1:
            return ls, lr
        # the ONE collective of the path: k candidates per query per rank
        self.dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        stream = t.cuda.current_stream().cuda_stream if self.device.type == "cuda" else None
        dev_index = self.device.index if self.device.type == "cuda" else 0
        self._merge(recv, self.world, nq, k, out_s, out_r, dev_index, stream)
        return out_s, out_r
