"""First-contact GPU probe: integer-pipe micro-benchmarks + scan timings at a few shapes.
Writes gpurun_out/probe.json.  Not the bench; numbers here guide tuning."""
import json
import os
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlite_b200 import _capi  # noqa: E402

out = {"version": _capi.version(), "micro": {}, "scan": []}
names = {0: "popc", 1: "lop3", 2: "iadd", 3: "idp4a", 4: "xor+popc+add"}
for kind, name in names.items():
    ops, ms = _capi.microbench(kind, 8192)
    out["micro"][name] = {"ops_per_s": ops, "ms": ms}
    print(name, f"{ops/1e12:.3f} T lane-ops/s", f"{ms:.3f} ms", flush=True)


def time_search(c, queries, k, metric, reps=5, mask=None):
    import torch
    dq = torch.from_numpy(queries).cuda()
    ds = torch.empty((queries.shape[0], k), dtype=torch.int32, device="cuda")
    dr = torch.empty((queries.shape[0], k), dtype=torch.int64, device="cuda")
    c.set_timing(True)
    best_t, best_s = 1e9, 1e9
    for i in range(reps + 2):
        c.search(dq, k, metric, mask, ds, dr)
        t, s = c.last_timing()
        if i >= 2:
            best_t, best_s = min(best_t, t), min(best_s, s)
    return best_t, best_s


from oracle import synth  # noqa: E402

for (n, w, nq, k, metric) in [
    (500_000, 128, 1024, 10, 0), (500_000, 128, 1024, 10, 1), (500_000, 128, 64, 10, 0),
    (500_000, 64, 1024, 10, 0),
    (16_000_000, 128, 1, 10, 0), (16_000_000, 128, 2, 10, 0), (16_000_000, 128, 4, 10, 0),
    (16_000_000, 128, 8, 10, 0), (16_000_000, 128, 16, 10, 0), (16_000_000, 128, 64, 10, 0),
    (16_000_000, 128, 1, 10, 1), (16_000_000, 128, 8, 10, 1),
    (32_000_000, 64, 1, 10, 0),
]:
    with _capi.Corpus(w, capacity_rows=n) as c:
        c.fill_synthetic(0, 0, n)
        q = synth.uniform_queries(1, nq, w)
        t, s = time_search(c, q, k, metric)
        rec = {"n": n, "w": w, "nq": nq, "k": k, "metric": metric, "total_ms": t, "scan_ms": s,
               "qps": nq / (t * 1e-3), "GBps": n * w / (s * 1e-3) / 1e9,
               "Tpopc_per_s": n * nq * (w / 4) / (s * 1e-3) / 1e12}
        out["scan"].append(rec)
        print(rec, flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
