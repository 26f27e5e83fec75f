"""HBM-bound leg for ncu: Q=1 and Q=2 scans of a 4 GB shard (32M x 128 B)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from oracle import synth
from vlite_b200 import _capi
n = 32_000_000
c = _capi.Corpus(128, capacity_rows=n); c.fill_synthetic(0, 0, n); c.set_timing(True)
for nq in (1, 2):
    dq = torch.from_numpy(synth.uniform_queries(1, nq, 128)).cuda()
    ds = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); dr = torch.empty((nq, 10), dtype=torch.int64, device="cuda")
    for i in range(4):
        c.search(dq, 10, 0, None, ds, dr); t, s = c.last_timing()
    print(f"Q={nq}: scan {s:.4f} ms = {n*128/s/1e6:.1f} GB/s, total {t:.4f} ms", flush=True)
c.close()
