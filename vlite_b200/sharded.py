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

Three layers live here: ``ShardedSearcher`` (search over fixed shards: local scan, the one
all-gather, merge), ``ShardedCorpus`` (an append-able collection over the ranks of a process
group, driven SPMD) and ``LocalShardedCorpus`` (the same collection rules inside ONE process,
over the C library's shard group: peer copies instead of a collective).  All three return
identical results for identical content.
"""
from __future__ import annotations

import os
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
        if device is None:
            device = torch.device("cuda", corpus.device) if local_search is None else torch.device("cpu")
        self.device = device
        self._bufs = {}
        # the exchange: peer stores over CUDA IPC mappings (vl_exchange_*) on GPUs; the all-gather +
        # merge of the host-logic tests (gloo) and wherever IPC cannot be set up
        self._xchg = None
        self._xchg_ok = ((merge is None or merge is _cuda_merge) and self.world > 1 and device.type == "cuda"
                         and self.world <= 16 and os.environ.get("VLITE_B200_EXCHANGE", "ipc") == "ipc")

    def _exchange(self, block_bytes: int):
        """collective (every rank calls it with the same size): (re)build the peer mappings"""
        from . import _capi
        if self._xchg is not None and self._xchg.slot_bytes >= block_bytes:
            return self._xchg
        t, dist = self.torch, self.dist
        if self._xchg is not None:
            t.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            self._xchg.close()
            self._xchg = None
        ok, x, handle = 1, None, b""
        try:
            x = _capi.Exchange(self.device.index, self.world, self.rank, max(block_bytes, 1 << 20))
            handle = x.handle()
        except Exception:          # noqa: BLE001 -- decided collectively below
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, handle), group=self.group)
        if all(h[0] for h in handles):
            try:
                for r, (_, h) in enumerate(handles):
                    if r != self.rank:
                        x.attach(r, h)
            except Exception:      # noqa: BLE001
                ok = 0
        flag = t.tensor([ok if all(h[0] for h in handles) else 0], dtype=t.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            if x is not None:
                x.close()
            self._xchg_ok = False
            return None
        self._xchg = x
        return x

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
        stream = t.cuda.current_stream().cuda_stream if self.device.type == "cuda" else None
        if self._xchg_ok:
            x = self._exchange(nb * 12)
            if x is not None:
                # the ONE exchange of the path: every rank stores its k candidates per query into every
                # peer's buffer over NVLink and raises a flag; the merge kernel acquires all flags
                x.gather_merge(stream, send, nb * 12, nq, k, out_s, out_r)
                return out_s, out_r
        # the ONE collective of the path: k candidates per query per rank
        self.dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        dev_index = self.device.index if self.device.type == "cuda" else 0
        self._merge(recv, self.world, nq, k, out_s, out_r, dev_index, stream)
        return out_s, out_r


# ==========================================================================================
# A row-sharded, append-able collection (second half)
# ==========================================================================================
class CudaShardBackend:
    """The product backend of ShardedCorpus: a local ``Corpus`` per rank + the CUDA merge."""

    def __init__(self, width: int, device_index: int):
        import torch
        from . import _capi
        self.torch, self._capi = torch, _capi
        self.device = torch.device("cuda", device_index)
        self.local = _capi.Corpus(width, device=device_index)
        self.local.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._seg_dev = None

    def append(self, rows_u8):
        return self.local.append(rows_u8)

    def load_file(self, path, off, length, kind, max_rows):
        return self.local.load_file(path, off, length, kind, max_rows)

    def tombstone(self, local_rows):
        self.local.tombstone(local_rows)

    def read(self, first, n):
        return self.local.read(first, n)

    def clear(self):
        self.local.clear()

    def live_rows(self):
        return self.local.live_rows

    def set_segments(self, seg_local, seg_global):
        t = self.torch
        self._seg_dev = (t.tensor(seg_local, dtype=t.int64, device=self.device),
                         t.tensor(seg_global, dtype=t.int64, device=self.device)) if seg_local else None

    def local_search(self, _corpus, queries, k, metric, filter_mask, out_scores, out_rows):
        """local top-k with LOCAL rows, rewritten to GLOBAL rows on the device before the exchange"""
        self.local.search(queries, k, metric, filter_mask, out_scores, out_rows)
        if self._seg_dev is not None:
            self._capi.map_rows(out_rows, out_rows.numel(), self._seg_dev[0], self._seg_dev[1],
                                int(self._seg_dev[0].numel()), device=self.device.index,
                                stream=self.torch.cuda.current_stream(self.device).cuda_stream)

    merge = staticmethod(_cuda_merge)

    def to_device(self, host_array):
        return self.torch.from_numpy(np.ascontiguousarray(host_array)).to(self.device)

    def close(self):
        self.local.close()


class SegmentTable:
    """The placement bookkeeping both sharded collections share: which shard owns which GLOBAL rows.

    A batch goes to the least-full shard (ties: lowest index) and takes the next global rows, so
    global rows are insertion order whatever the shard count.  Segments are
    ``(global_base, owner, local_start, n)``, ascending in ``global_base``; an append that continues
    the previous segment (same owner -- its local range is then contiguous too) extends it instead
    of adding one, the base array for ``locate``'s bisect and the per-owner tables the device needs
    (``local_start -> global_base``, for ``vl_map_rows``) are kept incrementally, and a shard whose
    table changed is only marked dirty: the upload happens once, before the next search, so a stream
    of small ``add()`` calls costs O(1) each on the host."""

    def __init__(self, n_shards: int):
        self.n_shards = n_shards
        self.clear()

    def clear(self):
        self.counts = [0] * self.n_shards        # local rows per shard
        self.segments = []                       # (global_base, owner, local_start, n)
        self._bases = []                         # [s[0] for s in segments], kept for bisect
        self._by_owner = [([], []) for _ in range(self.n_shards)]   # (local starts, global bases) per shard
        self.dirty = set(range(self.n_shards))   # shards whose device table is stale
        self.total = 0

    def least_full(self) -> int:
        return min(range(self.n_shards), key=lambda r: (self.counts[r], r))

    def add(self, owner: int, n: int) -> int:
        """place n rows on `owner`; returns their first GLOBAL row"""
        base, lstart = self.total, self.counts[owner]
        last = self.segments[-1] if self.segments else None
        if last is not None and last[1] == owner and last[0] + last[3] == base and last[2] + last[3] == lstart:
            self.segments[-1] = (last[0], owner, last[2], last[3] + n)      # same mapping: nothing to upload
        else:
            self.segments.append((base, owner, lstart, n))
            self._bases.append(base)
            self._by_owner[owner][0].append(lstart)
            self._by_owner[owner][1].append(base)
            self.dirty.add(owner)
        self.counts[owner] += n
        self.total += n
        return base

    def locate(self, global_row: int):
        """(owner, local row) of a global row"""
        import bisect
        i = bisect.bisect_right(self._bases, global_row) - 1
        if i < 0 or global_row >= self.segments[i][0] + self.segments[i][3]:
            raise IndexError(f"global row {global_row} out of range")
        base, owner, lstart, _ = self.segments[i]
        return owner, lstart + (global_row - base)

    def table_of(self, owner: int):
        """(local starts, global bases) of one shard, ascending"""
        return self._by_owner[owner]

    def local_masks(self, mask_words_global, owners=None):
        """global bit-per-row mask -> one local mask per shard (host-side bit gather); None for an
        empty shard or one not in `owners`"""
        bits = np.unpackbits(np.ascontiguousarray(mask_words_global, dtype=np.uint32).view(np.uint8),
                             bitorder="little")
        want = range(self.n_shards) if owners is None else owners
        local = {r: np.zeros((self.counts[r] + 31) // 32 * 32, dtype=np.uint8) for r in want}
        for base, owner, lstart, n in self.segments:
            if owner in local:
                local[owner][lstart:lstart + n] = bits[base:base + n]
        return [np.packbits(local[r].reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
                if r in local and local[r].size else None for r in range(self.n_shards)]

    def overlapping(self, first: int, n: int):
        """(offset in [first, first+n), owner, local start, count) pieces of a global row range"""
        for base, owner, lstart, cnt in self.segments:
            lo, hi = max(base, first), min(base + cnt, first + n)
            if lo < hi:
                yield lo - first, owner, lstart + (lo - base), hi - lo


class ShardedCorpus:
    """A collection row-sharded over the ranks of a process group, driven SPMD: every rank
    constructs it and makes the SAME calls with the SAME host arguments.

      * append: the whole batch goes to the least-full shard (ties: lowest rank); its GLOBAL base
        row is fixed at append time = the running global count, so global rows are insertion order
        and results do not depend on the shard count or on placement;
      * tombstones are shard-local (the owner clears its live bit);
      * search: local scan -> local rows rewritten to global rows on the device (vl_map_rows) ->
        ONE all-gather of the k candidates -> the same merge on every rank;
      * load_file: each rank reads its own byte range of the EMBEDDINGS payload.

    Offers the slice of the ``Corpus`` interface that ``VLite`` uses, with global row numbers, so
    ``ShardedVLite`` is the single-GPU class with a different corpus behind it."""

    def __init__(self, width: int, device: int | None = None, group=None, backend=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.width = width
        self.backend = backend or CudaShardBackend(width, device if device is not None else 0)
        self.table = SegmentTable(self.world)     # replicated bookkeeping: every rank makes the same calls
        self._searcher = ShardedSearcher(None, group=group, local_search=self.backend.local_search,
                                         merge=self.backend.merge, device=getattr(self.backend, "device", None))

    # ---- bookkeeping (SegmentTable) ----
    counts = property(lambda self: self.table.counts)
    segments = property(lambda self: self.table.segments)
    total = property(lambda self: self.table.total)

    def _add_segment(self, owner: int, n: int) -> int:
        return self.table.add(owner, n)

    def _sync_segments(self):
        if self.rank in self.table.dirty:
            self.backend.set_segments(*self.table.table_of(self.rank))
        self.table.dirty.clear()

    def locate(self, global_row: int):
        """(owner rank, local row) of a global row."""
        return self.table.locate(global_row)

    @property
    def rows(self) -> int:
        return self.table.total

    @property
    def live_rows(self) -> int:
        t = self.torch.tensor([self.backend.live_rows()], dtype=self.torch.int64,
                              device=self._searcher.device)
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    # ---- ingest ----
    def append(self, rows_u8) -> int:
        rows_u8 = np.ascontiguousarray(rows_u8, dtype=np.uint8)
        n = int(rows_u8.shape[0])
        if n == 0:
            return self.total
        owner = self.table.least_full()
        if owner == self.rank:
            first_local = self.backend.append(rows_u8)
            assert first_local == self.counts[owner]
        return self._add_segment(owner, n)

    def load_file(self, path: str, byte_off: int, byte_len: int, kind: int, max_rows: int = 0):
        """Every rank streams ITS contiguous slice of the file's rows to its own GPU."""
        from . import _capi
        row_bytes = {_capi.SRC_U8: self.width, _capi.SRC_F32_OFFSET: self.width * 4,
                     _capi.SRC_F32_BITS: self.width * 32}[kind]
        n = byte_len // row_bytes
        if max_rows:
            n = min(n, max_rows)
        plan = ShardPlan(n, self.world)
        first = self.total
        lo, cnt = plan.base_row(self.rank), plan.shard_rows(self.rank)
        if cnt:
            self.backend.load_file(path, byte_off + lo * row_bytes, cnt * row_bytes, kind, cnt)
        for r in range(self.world):
            if plan.shard_rows(r):
                self._add_segment(r, plan.shard_rows(r))
        return first, n

    def tombstone(self, global_rows) -> None:
        mine = []
        for g in np.asarray(global_rows, dtype=np.int64).reshape(-1):
            try:
                owner, local = self.locate(int(g))
            except IndexError:
                continue
            if owner == self.rank:
                mine.append(local)
        if mine:
            self.backend.tombstone(mine)

    def clear(self):
        self.backend.clear()
        self.table.clear()
        self.backend.set_segments([], [])

    # ---- query ----
    def _local_mask(self, mask_words_global):
        """global bit-per-row mask -> this shard's local mask (host-side bit gather)"""
        return self.table.local_masks(mask_words_global, owners=[self.rank])[self.rank]

    def search(self, queries, k: int, metric: int = 0, filter_mask=None):
        """queries (Q, W) uint8 host array, identical on every rank; filter_mask: optional GLOBAL
        bit-per-row mask.  Returns host arrays (scores uint32 (Q,k), rows uint64 (Q,k)) with global
        rows, the same on every rank."""
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.uint8)
        nq = q.shape[0]
        pad = (nq * k) % 2                       # the exchange block wants an even Q*k
        if pad:
            q = np.concatenate([q, q[:1]])
        mask = None
        if filter_mask is not None and self.counts[self.rank]:
            mask = self._local_mask(filter_mask)
        self._sync_segments()
        dq = self.backend.to_device(q)
        s, r = self._searcher.search(dq, k, metric, mask)
        # copies: the searcher hands out views of its reusable exchange buffers
        s = s.cpu().numpy().view(np.uint32)[:nq].copy()
        r = r.cpu().numpy().view(np.uint64)[:nq].copy()
        return s, r

    def read(self, first: int = 0, n: int | None = None) -> np.ndarray:
        """rows [first, first+n) by GLOBAL number, assembled on every rank (collective)."""
        if n is None:
            n = self.total - first
        out = np.zeros((n, self.width), dtype=np.uint8)
        pieces = [(off, self.backend.read(lstart, cnt)) for off, owner, lstart, cnt in self.table.overlapping(first, n)
                  if owner == self.rank]
        if self.world > 1:
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, pieces, group=self.group)
            pieces = [p for part in gathered for p in part]
        for off, block in pieces:
            out[off:off + len(block)] = block
        return out

    def close(self):
        self.backend.close()


class LocalShardedCorpus:
    """The single-process twin of ``ShardedCorpus``: one ``Corpus`` per listed GPU, searched as one
    collection through ``vl_shard_group_*`` (per-device scans run concurrently, candidates travel to
    the first GPU by peer copy, one merge there).  Placement and numbering follow the same rules --
    a batch goes to the least-full shard, its GLOBAL base row is the running global count, tombstones
    are shard-local -- so results equal the single-GPU collection's and the torchrun path's.

    Offers the slice of the ``Corpus`` interface ``VLite`` uses, with global row numbers."""

    def __init__(self, width: int, devices):
        from . import _capi
        self._capi = _capi
        self.width = width
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("need at least one GPU")
        self.shards = [_capi.Corpus(width, device=d) for d in self.devices]
        self.group = _capi.ShardGroup(self.shards)
        self.table = SegmentTable(len(self.shards))

    # ---- bookkeeping (the same SegmentTable rules as ShardedCorpus) ----
    counts = property(lambda self: self.table.counts)
    segments = property(lambda self: self.table.segments)
    total = property(lambda self: self.table.total)

    def _add_segment(self, owner: int, n: int) -> int:
        return self.table.add(owner, n)

    def _sync_segments(self):
        for owner in sorted(self.table.dirty):
            self.group.set_segments(owner, *self.table.table_of(owner))
        self.table.dirty.clear()

    def locate(self, global_row: int):
        return self.table.locate(global_row)

    @property
    def rows(self) -> int:
        return self.table.total

    @property
    def live_rows(self) -> int:
        return self.group.live_rows

    # ---- ingest ----
    def append(self, rows_u8) -> int:
        rows_u8 = np.ascontiguousarray(rows_u8, dtype=np.uint8)
        n = int(rows_u8.shape[0])
        if n == 0:
            return self.total
        owner = self.table.least_full()
        first_local = self.shards[owner].append(rows_u8)
        assert first_local == self.counts[owner]
        return self._add_segment(owner, n)

    def load_file(self, path: str, byte_off: int, byte_len: int, kind: int, max_rows: int = 0):
        """The file's rows are cut into one contiguous slice per GPU, each streamed to its device."""
        _capi = self._capi
        row_bytes = {_capi.SRC_U8: self.width, _capi.SRC_F32_OFFSET: self.width * 4,
                     _capi.SRC_F32_BITS: self.width * 32}[kind]
        n = byte_len // row_bytes
        if max_rows:
            n = min(n, max_rows)
        plan = ShardPlan(n, len(self.shards))
        first = self.total
        for r, c in enumerate(self.shards):
            lo, cnt = plan.base_row(r), plan.shard_rows(r)
            if cnt:
                c.load_file(path, byte_off + lo * row_bytes, cnt * row_bytes, kind, max_rows=cnt)
                self._add_segment(r, cnt)
        return first, n

    def tombstone(self, global_rows) -> None:
        per = {}
        for g in np.asarray(global_rows, dtype=np.int64).reshape(-1):
            try:
                owner, local = self.locate(int(g))
            except IndexError:
                continue
            per.setdefault(owner, []).append(local)
        for owner, rows in per.items():
            self.shards[owner].tombstone(rows)

    def clear(self):
        for i, c in enumerate(self.shards):
            c.clear()
            self.group.set_segments(i, [], [])
        self.table.clear()

    # ---- query ----
    def search(self, queries, k: int, metric: int = 0, filter_mask=None):
        self._sync_segments()
        masks = self.table.local_masks(filter_mask) if filter_mask is not None else None
        return self.group.search(queries, k, metric, masks)

    def read(self, first: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = self.total - first
        out = np.zeros((n, self.width), dtype=np.uint8)
        for off, owner, lstart, cnt in self.table.overlapping(first, n):
            out[off:off + cnt] = self.shards[owner].read(lstart, cnt)
        return out

    def close(self):
        if self.group is not None:
            self.group.close()
            self.group = None
        for c in self.shards:
            c.close()
        self.shards = []
