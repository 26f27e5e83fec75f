"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: shard plan, the single
all-gather exchange, and shard-count invariance of the merged result.  The per-shard search and
the merge are the CPU oracle's here (there is no GPU); on the GPU box the same ShardedSearcher
runs with the CUDA kernels (tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cscan, search as osearch, synth
from vlite_b200.sharded import ShardedSearcher, ShardPlan


def oracle_hooks():
    def local_search(corpus, queries, k, metric, filter_mask, out_scores, out_rows):
        rows_u8, base_row = corpus
        mask = None if filter_mask is None else np.asarray(filter_mask)
        s, r = cscan.search(queries.numpy(), rows_u8, k, metric, mask, base_row)
        out_scores.copy_(torch.from_numpy(s.view(np.int32)))
        out_rows.copy_(torch.from_numpy(r.view(np.int64)))

    def merge(block, n_lists, n_queries, k, out_scores, out_rows, device, stream):
        nb = n_queries * k
        raw = block.numpy()
        s = np.stack([raw[l, : nb * 4].view(np.uint32).reshape(n_queries, k) for l in range(n_lists)])
        r = np.stack([raw[l, nb * 4:].view(np.uint64).reshape(n_queries, k) for l in range(n_lists)])
        ms, mr = osearch.merge_topk(s, r, k)
        out_scores.copy_(torch.from_numpy(ms.view(np.int32)))
        out_rows.copy_(torch.from_numpy(mr.view(np.int64)))

    return local_search, merge


def test_shard_plan():
    p = ShardPlan(10, 4)
    assert [p.base_row(r) for r in range(4)] == [0, 3, 6, 9]
    assert [p.shard_rows(r) for r in range(4)] == [3, 3, 3, 1]
    assert sum(p.shard_rows(r) for r in range(4)) == 10
    assert p.owner(7) == (2, 1) and p.owner(9) == (3, 0)
    p = ShardPlan(1_000_000_000, 8)
    assert p.rows_per_shard == 125_000_000 and p.base_row(7) == 875_000_000
    p = ShardPlan(3, 8)          # more ranks than rows: trailing shards are empty
    assert [p.shard_rows(r) for r in range(8)] == [1, 1, 1, 0, 0, 0, 0, 0]


def _worker(rank, world, port, n, w, nq, k, period, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(n, world)
        base, rows = plan.base_row(rank), plan.shard_rows(rank)
        shard = synth.corpus_rows(5, base, rows, w, period) if rows else np.zeros((0, w), np.uint8)
        queries = torch.from_numpy(synth.uniform_queries(6, nq, w))
        ls, mg = oracle_hooks()
        s = ShardedSearcher((shard, base), local_search=ls, merge=mg)
        assert s.world == world and s.rank == rank
        out = {}
        for metric in (0, 1):
            sc, rw = s.search(queries, k, metric)
            out[metric] = (sc.numpy().view(np.uint32).copy(), rw.numpy().view(np.uint64).copy())
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,period", [(2, 5000, 0), (2, 3001, 200), (3, 1000, 0)])
def test_sharded_search_equals_single_shard(world, n, period):
    w, nq, k = 128, 6, 10
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, w, nq, k, period, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = synth.corpus_rows(5, 0, n, w, period)
    queries = synth.uniform_queries(6, nq, w)
    for metric in (0, 1):
        es, er = osearch.search(queries, full, k, metric)
        for rank in range(world):                      # every rank holds the same merged answer
            np.testing.assert_array_equal(results[rank][metric][0], es)
            np.testing.assert_array_equal(results[rank][metric][1], er)
