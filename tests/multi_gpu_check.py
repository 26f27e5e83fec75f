"""Real-GPU check of the sharded collection (CUDA backend + NCCL), run under torchrun on >= 2 GPUs:

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank compares what the sharded classes return with the CPU oracle on the whole collection."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from oracle import cscan, search as osearch, synth  # noqa: E402
from vlite_b200 import _capi  # noqa: E402
from vlite_b200.main import ShardedVLite  # noqa: E402
from vlite_b200.sharded import ShardedCorpus  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    w = 128
    sizes = [5000, 300, 20000, 7, 12000, 4096, 1]
    batches = [synth.corpus_rows(50 + b, 0, n, w, period=997) for b, n in enumerate(sizes)]
    full = np.concatenate(batches)
    c = ShardedCorpus(w, device=lr)
    firsts = [c.append(b) for b in batches]
    assert firsts == list(np.cumsum([0] + sizes[:-1]))
    assert c.rows == len(full) == c.live_rows and sum(c.counts) == len(full)
    assert max(c.counts) <= sum(c.counts) / world + max(sizes)            # least-full placement
    np.testing.assert_array_equal(c.read(4990, 400), full[4990:5390])
    queries = np.concatenate([synth.uniform_queries(60, 30, w), full[[0, 5299, 41402]]])
    for metric in (_capi.HAMMING, _capi.XORSUM):
        for k in (1, 10, 100):
            s, r = c.search(queries, k, metric)
            es, er = cscan.search(queries, full, k, metric)
            assert np.array_equal(s, es) and np.array_equal(r, er), (metric, k)
    dead = np.unique(c.search(queries, 3, 0)[1].reshape(-1)).astype(np.int64)
    c.tombstone(dead)
    valid = np.ones(len(full), bool)
    valid[dead] = False
    assert c.live_rows == len(full) - len(dead)
    filt = synth.filter_valid(4, 0, len(full), 0.1)
    s, r = c.search(queries[:5], 7, 0, synth.pack_mask(filt))             # odd Q*k, global filter + tombstones
    es, er = cscan.search(queries[:5], full, 7, 0, synth.pack_mask(valid & filt))
    assert np.array_equal(s, es) and np.array_equal(r, er)
    c.close()

    # ---- the collection class, including CTX save by rank 0 and a sharded reload ----
    qv = synth.uniform_queries(61, 4, 64)
    emb = lambda texts: np.unpackbits(qv[: len(texts)], axis=1).astype(np.float32) * 2 - 1
    tmp = os.path.join(tempfile.gettempdir(), "vlite_b200_multi_gpu_check")
    if rank == 0:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    dist.barrier()
    rows64 = synth.corpus_rows(62, 0, 3000, 64)
    enc = osearch.to_ref_encoding(rows64)
    v = ShardedVLite("multi", ctx_dir=tmp, embedder=emb, width=64, gpu_index=lr)
    v.set_batch([f"t{i}" for i in range(2000)], enc[:2000], [{"tag": i % 4} for i in range(2000)])
    v.set_batch([f"u{i}" for i in range(1000)], enc[2000:], [{"tag": 9}] * 1000)
    got = v.retrieve(["q0", "q1"], top_k=6, return_scores=True)
    s, r = osearch.search(qv[:2], rows64, 6, osearch.XORSUM)
    want = sorted(((int(a), int(b)) for a, b in zip(s.reshape(-1), r.reshape(-1))), key=lambda x: x[0])[:6]
    assert [int(g[3]) for g in got] == [x[0] for x in want]
    got_f = v.retrieve("q0", top_k=5, metadata={"tag": 9}, return_scores=True)
    sf, rf = osearch.search(qv[:1], rows64, 5, osearch.XORSUM, valid=np.arange(3000) >= 2000)
    assert [g[1] for g in got_f] == [f"u{int(x) - 2000}" for x in rf[0]] and [int(g[3]) for g in got_f] == list(sf[0])
    key = got[0][0]
    assert v.delete(key[:-2]) == 1 and v.count() == 2999
    assert key not in [g[0] for g in v.retrieve(["q0", "q1"], top_k=6)]
    keys_before = list(v.index.keys())
    v.close()
    v2 = ShardedVLite("multi", ctx_dir=tmp, embedder=emb, width=64, gpu_index=lr)     # every rank loads its byte range
    assert v2.count() == 2999 and list(v2.index.keys()) == keys_before, (rank, v2.count(), len(keys_before),
                                                                       os.listdir(tmp), v2._corpus.counts)
    assert sorted(v2._corpus.counts, reverse=True)[0] <= -(-2999 // world)
    again = v2.retrieve(["q0", "q1"], top_k=6, return_scores=True)
    assert key not in [g[0] for g in again]
    v2.close()
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok on {world} GPUs", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
