"""In-process shard group over every visible GPU: rows_per_gpu synthetic rows each, wall-clock
time of vl_shard_group_search (host queries in, host results out).
usage: group_bench.py rows_per_gpu,nq,k[,metric] ..."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import synth  # noqa: E402  (query generation only)
from vlite_b200 import _capi  # noqa: E402

W = 128
n_dev = _capi.device_count()
reps = int(os.environ.get("REPS", "10"))
shards, per_now = [], 0
for spec in sys.argv[1:]:
    f = [int(x) for x in spec.split(",")]
    per, nq, k = f[:3]
    metric = f[3] if len(f) > 3 else 0
    if per != per_now:
        for c in shards:
            c.close()
        shards = []
        for d in range(n_dev):
            c = _capi.Corpus(W, device=d, capacity_rows=per, base_row=d * per)
            c.fill_synthetic(0, 0, per)        # device-side generator: global rows [d*per, (d+1)*per)
            shards.append(c)
        per_now = per
    group = _capi.ShardGroup(shards)
    q = synth.uniform_queries(1, nq, W)
    for _ in range(3):
        group.search(q, k, metric)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        s, r = group.search(q, k, metric)
        ts.append(time.perf_counter() - t0)
    ms = float(np.median(ts)) * 1e3
    total = per * n_dev
    print(json.dumps({"gpus": n_dev, "rows_per_gpu": per, "total_rows": total, "nq": nq, "k": k, "metric": metric,
                      "ms": ms, "min_ms": float(min(ts)) * 1e3, "qps": nq / (ms * 1e-3),
                      "aggregate_GBps": total * W / (ms * 1e-3) / 1e9,
                      "digest": int(np.bitwise_xor.reduce(r.reshape(-1) & np.uint64(0xFFFFFFFF)) ^ int(s.sum()) & 0xFFFFFFFF)}),
          flush=True)
    group.close()
