"""Time vl_search at given shapes (device-resident queries/outputs, CUDA events inside the library).
usage: scan_bench.py n,w,nq,k,metric[,splits[,stages]] ...   -> prints one JSON line per shape."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import synth  # noqa: E402
from vlite_b200 import _capi  # noqa: E402

POPC_PEAK = 4.62e12
HBM_PEAK = 6554.9e9
reps = int(os.environ.get("REPS", "5"))
cache = {}
for spec in sys.argv[1:]:
    f = [int(x) for x in spec.split(",")]
    n, w, nq, k, metric = f[:5]
    splits = f[5] if len(f) > 5 else 0
    stages = f[6] if len(f) > 6 else 0
    key = (n, w)
    if key not in cache:
        for c in cache.values():
            c.close()
        cache.clear()
        c = _capi.Corpus(w, capacity_rows=n)
        c.fill_synthetic(0, 0, n)
        cache[key] = c
    c = cache[key]
    c.set_tuning(splits, stages)
    c.set_scan_path(int(os.environ.get("SCAN_PATH", "0")))
    q = synth.uniform_queries(1, nq, w)
    dq = torch.from_numpy(q).cuda()
    ds = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    dr = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    c.set_timing(True)
    ts, ss = [], []
    for i in range(reps + 2):
        c.search(dq, k, metric, None, ds, dr)
        t, s = c.last_timing()
        if i >= 2:
            ts.append(t)
            ss.append(s)
    t, s = float(np.median(ts)), float(np.median(ss))
    pairs = n * nq
    t_roof = max(n * w / HBM_PEAK, pairs * (w / 4) / POPC_PEAK)
    print(json.dumps({"n": n, "w": w, "nq": nq, "k": k, "metric": metric, "splits": splits, "stages": stages, "path": int(os.environ.get("SCAN_PATH", "0")),
                      "total_ms": round(t, 4), "scan_ms": round(s, 4), "qps": round(nq / (t * 1e-3)),
                      "GBps": round(n * w / (s * 1e-3) / 1e9, 1),
                      "Tpopc": round(pairs * (w / 4) / (s * 1e-3) / 1e12, 3),
                      "roof_frac": round(t_roof / (s * 1e-3), 3)}), flush=True)
