"""Long randomized differential run of vl_search against the C oracle (a superset of
tests/test_gpu_search.py::test_random_differential: free-form sizes, large k, base rows, appends in
several batches, shard groups).  usage: fuzz_search.py [cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import cscan, synth  # noqa: E402
from vlite_b200 import _capi  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
bad = 0
for case in range(cases):
    w = int(rng.choice([32, 64, 128]))
    n = int(rng.choice([rng.integers(1, 600), rng.integers(600, 9000), rng.integers(9000, 120000)]))
    nq = int(rng.choice([rng.integers(1, 10), rng.integers(10, 80), rng.integers(80, 400)]))
    k = int(rng.choice([rng.integers(1, 33), rng.integers(33, 200), rng.integers(200, 2049)]))
    metric = int(rng.integers(0, 2))
    period = int(rng.choice([0, 0, 0, 3, 50, 1000]))
    base = int(rng.choice([0, 0, 12345, 1 << 33]))
    corpus = synth.corpus_rows(5000 + case, base, n, w, period)
    queries = synth.uniform_queries(9000 + case, nq, w)
    if n > 3:
        queries[0] = corpus[n - 1]
        queries[-1] = corpus[0] ^ np.uint8(1)
    valid = np.ones(n, dtype=bool)
    mask = None
    if rng.random() < 0.5:
        valid &= rng.random(n) < rng.choice([0.001, 0.02, 0.3, 0.9])
        mask = synth.pack_mask(valid)
    use_group = rng.random() < 0.25 and n > 10
    if use_group:
        cuts = sorted(set([0, n] + [int(x) for x in rng.integers(0, n, size=int(rng.integers(1, 4)))]))
        shards = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            c = _capi.Corpus(w, base_row=base + lo)
            c.append(corpus[lo:hi])
            shards.append(c)
        masks = None if mask is None else [synth.pack_mask(valid[lo:hi]) for lo, hi in zip(cuts[:-1], cuts[1:])]
        with _capi.ShardGroup(shards) as g:
            s, r = g.search(queries, k, metric, masks)
        for c in shards:
            c.close()
    else:
        with _capi.Corpus(w, base_row=base) as c:
            cut = int(rng.integers(0, n + 1))
            c.append(corpus[:cut])
            c.append(corpus[cut:])
            if rng.random() < 0.4 and n > 4:
                dead = rng.choice(n, size=max(1, n // int(rng.integers(2, 9))), replace=False)
                c.tombstone(dead)
                valid[dead] = False
            s, r = c.search(queries, k, metric, mask)
    es, er = cscan.search(queries, corpus, k, metric, synth.pack_mask(valid), base)
    ok = np.array_equal(s, es) and np.array_equal(r, er)
    if not ok:
        bad += 1
        print("MISMATCH", dict(case=case, w=w, n=n, nq=nq, k=k, metric=metric, period=period, base=base,
                               group=use_group, masked=mask is not None), flush=True)
print(f"fuzz: {cases} cases, {bad} mismatches (seed {seed})")
sys.exit(1 if bad else 0)
