"""Stateful randomized run of the VLite mirror against a plain-Python model of the reference's
collection semantics (vlite/main.py): set_batch / add / update / delete / save+reload interleaved
with filtered and unfiltered retrieve_batch calls, every answer checked against the oracle's search
over the model's live rows.  usage: fuzz_collection.py [steps] [seed] [gpus]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import search as osearch, synth  # noqa: E402
from vlite_b200 import VLite  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
gpus = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
rng = np.random.default_rng(seed)
W = 64
POOL = synth.corpus_rows(seed, 0, 60_000, W, period=4000)          # duplicates -> ties
QUERIES = synth.uniform_queries(seed + 1, 6, W)
LANGS, YEARS = ["en", "fr", "de", None], [2020, 2021, 2022]


def embedder(texts):
    rows = np.stack([POOL[int(t.split(":")[1]) % len(POOL)] if t.startswith("v:") else QUERIES[int(t.split(":")[1])] for t in texts])
    return np.unpackbits(rows, axis=1).astype(np.float32) * 2 - 1


def rand_meta():
    md = {}
    if rng.random() < 0.8:
        md["lang"] = LANGS[int(rng.integers(len(LANGS)))]
    if rng.random() < 0.6:
        md["year"] = YEARS[int(rng.integers(len(YEARS)))]
    if rng.random() < 0.05:
        md["tags"] = ["t", int(rng.integers(3))]
    return md


world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:                                    # under torchrun: the SPMD ShardedVLite, one rank per GPU
    import torch
    import torch.distributed as dist
    from vlite_b200.main import ShardedVLite
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [tempfile.mkdtemp() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    kw = dict(ctx_dir=tmp, embedder=embedder, width=W, metric="hamming", gpu_index=local)
    VLite = ShardedVLite                          # noqa: F811
else:
    tmp = tempfile.mkdtemp()
    kw = dict(ctx_dir=tmp, embedder=embedder, width=W, metric="hamming")
    if gpus:
        kw["gpus"] = gpus
v = VLite("fuzz", **kw)
model = []          # live entries in row order: [chunk_id, pool index, metadata dict]
next_vec, next_item, checks, stored = 0, 0, 0, 0   # stored: rows the device must hold (tombstones included)


def check(md, k):
    global checks
    assert (v._corpus.rows if v._corpus is not None else 0) == stored and v.count() == len(model)
    valid = np.array([all(e[2].get(a) == b for a, b in md.items()) for e in model], dtype=bool) if md else \
        np.ones(len(model), dtype=bool)
    corpus = POOL[np.asarray([e[1] for e in model]) % len(POOL)] if model else np.zeros((0, W), np.uint8)
    texts = [f"q:{i}" for i in range(len(QUERIES))]
    if not model:                                   # vlite/main.py:149: an empty collection cannot be ranked
        try:
            v.retrieve_batch(texts, top_k=k, metadata=md or None)
        except ValueError:
            return
        raise AssertionError("empty collection must raise ValueError")
    got = v.retrieve_batch(texts, top_k=k, metadata=md or None, return_scores=True)
    es, er = osearch.search(QUERIES, corpus, k, osearch.HAMMING, valid=valid)
    for qi in range(len(QUERIES)):
        keep = er[qi] != osearch.SENTINEL_ROW
        want = [(model[int(r)][0], int(s)) for s, r in zip(es[qi][keep], er[qi][keep])]
        have = [(g[0], int(g[3])) for g in got[qi]]
        assert have == want, (md, k, qi, have[:5], want[:5])
    checks += 1


for step in range(steps):
    op = rng.choice(["set_batch", "add", "delete", "update", "query", "query", "reload"],
                    p=[0.2, 0.15, 0.12, 0.1, 0.2, 0.18, 0.05])
    if op == "set_batch":
        n = int(rng.choice([1, 3, 50, 700]))
        idx = list(range(next_vec, next_vec + n))
        next_vec += n
        stored += n
        metas = [rand_meta() for _ in idx]
        before = set(v.dump().keys())
        v.set_batch([f"v:{i}" for i in idx], osearch.to_ref_encoding(POOL[np.asarray(idx) % len(POOL)]), [dict(m) for m in metas])
        new_keys = [key for key in v.dump().keys() if key not in before]
        assert len(new_keys) == n
        model += [[key, i, dict(m)] for key, i, m in zip(new_keys, idx, metas)]
    elif op == "add":
        n = int(rng.integers(1, 6))
        idx = list(range(next_vec, next_vec + n))
        next_vec += n
        stored += n
        md = rand_meta()
        item = f"item{next_item}"
        next_item += 1
        v.add([f"v:{i}" for i in idx], metadata=dict(md), item_id=item)
        model += [[f"{item}_{j}", i, dict(md)] for j, i in enumerate(idx)]
    elif op == "delete" and model:
        victims = {model[int(i)][0].rsplit("_", 1)[0] for i in rng.integers(0, len(model), size=int(rng.integers(1, 5)))}
        want = sum(1 for e in model if e[0].rsplit("_", 1)[0] in victims)
        assert v.delete(sorted(victims)) == want
        model = [e for e in model if e[0].rsplit("_", 1)[0] not in victims]
    elif op == "update" and model:
        target = model[int(rng.integers(len(model)))][0].rsplit("_", 1)[0]
        new = {"lang": LANGS[int(rng.integers(len(LANGS)))]}
        assert v.update(target, metadata=dict(new))
        for e in model:
            if e[0].rsplit("_", 1)[0] == target:
                e[2].update(new)
    elif op == "query":
        md = [{}, {"lang": "en"}, {"lang": None}, {"lang": "fr", "year": 2021}, {"year": 2022}, {"tags": ["t", 1]},
              {"lang": "xx"}][int(rng.integers(7))]
        check(md, int(rng.choice([1, 5, 10, 40])))
    elif op == "reload" and model:
        v.save()
        v.close()
        v = VLite("fuzz", **kw)
        stored = len(model)                       # a reload compacts: tombstoned rows are gone
        assert v.count() == len(model)
        assert list(v.dump().keys()) == [e[0] for e in model]
check({}, 10)
check({"lang": "en"}, 10)
print(f"fuzz_collection: {steps} steps, {checks} checked query batches, {len(model)} live rows, {v._corpus.rows} rows stored"
      f" (seed {seed}, gpus {gpus})")
v.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
