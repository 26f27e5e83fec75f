"""Where one CTA of the tensor-core scan spends its time (globaltimer stamps, vl_scan_trace)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from oracle import synth
from vlite_b200 import _capi
L = _capi.load()
L.vl_scan_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for n, nq, k in [(500_000, 1024, 10), (500_000, 1024, 1), (100_000, 1024, 10), (500_000, 128, 10)]:
    c = _capi.Corpus(128, capacity_rows=n); c.fill_synthetic(0, 0, n)
    dq = torch.from_numpy(synth.uniform_queries(1, nq, 128)).cuda()
    ds = torch.empty((nq, k), dtype=torch.int32, device="cuda"); dr = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    L.vl_scan_trace(c._h, None)
    c.set_timing(True)
    for _ in range(5): c.search(dq, k, 0, None, ds, dr)
    out = np.zeros(16, dtype=np.uint64)
    L.vl_scan_trace(c._h, out.ctypes.data)
    t = out.astype(np.int64)
    rel = [int(x - t[0]) if x else None for x in t]
    print(f"n={n} nq={nq} k={k}: scan_ms={c.last_timing()[1]:.4f}  prologue {rel[1]} ns, last epilogue at {rel[2]} ns, exit {rel[3]} ns")
    print(f"   accumulator 0 ready {rel[4]}, tile 0 chunks loaded at {rel[5:13]}, accumulator 1 ready {rel[13]}, 2 ready {rel[14]}")
    c.close()
