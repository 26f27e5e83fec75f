"""Gather-rate of the rescore kernel: Q queries x C scattered int8 rows of 1 KB."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from vlite_b200 import _capi
n, dims, nq, C, k = 20_000_000, 1024, 4096, 1000, 10
dev = torch.device("cuda", 0)
docs = torch.empty((n, dims), dtype=torch.int8, device=dev); _capi.synth_docs_i8(7, 0, n, dims, docs)
rng = np.random.default_rng(0)
cand = torch.from_numpy(rng.integers(0, n, size=(nq, C)).astype(np.int64)).to(dev)
c = _capi.Corpus(128); c.set_timing(True)
qi8 = torch.randint(-128, 128, (nq, dims), dtype=torch.int8, device=dev)
for name, q, probe in (("f32 (exact fp64 accumulate)", torch.randn((nq, dims), device=dev), False), ("i8 (DP4A)", qi8, False),
                       ("gather probe (no dot products: the HBM gather floor)", qi8, True)):
    os_ = torch.empty((nq, k), dtype=torch.float32, device=dev); or_ = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ts = []
    for i in range(5):
        c.rescore(docs, n, dims, q, cand, k, os_, or_, gather_probe=probe); ts.append(c.last_timing()[0])
    t = float(np.median(ts[1:]))
    print(f"{name}: {t:.3f} ms  gather {nq*C*dims/t/1e6:.0f} GB/s  {nq/t*1e3:.0f} q/s", flush=True)
