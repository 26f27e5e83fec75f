"""Per-tile timeline of CTA (0,0) of the tensor-core scan (vl_scan_trace_tiles): when the epilogue of each
tile finished, how many 32-row chunks took the rare (insert) path and how many insert-loop iterations ran."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from oracle import synth
from vlite_b200 import _capi
L = _capi.load()
L.vl_scan_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.vl_scan_trace_tiles.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.vl_scan_trace_ctas.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
shapes = [(500_000, 1024, 10), (500_000, 1024, 1), (4_000_000, 1024, 10)]
for n, nq, k in shapes:
    c = _capi.Corpus(128, capacity_rows=n); c.fill_synthetic(0, 0, n)
    dq = torch.from_numpy(synth.uniform_queries(1, nq, 128)).cuda()
    ds = torch.empty((nq, k), dtype=torch.int32, device="cuda"); dr = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    L.vl_scan_trace(c._h, None)
    c.set_timing(True)
    for _ in range(5): c.search(dq, k, 0, None, ds, dr)
    head = np.zeros(16, dtype=np.uint64); tiles = np.zeros(512, dtype=np.uint64)
    L.vl_scan_trace(c._h, head.ctypes.data); L.vl_scan_trace_tiles(c._h, tiles.ctypes.data)
    t0 = int(head[0]); done = tiles[:256].astype(np.int64); cnt = tiles[256:].astype(np.int64)
    nt = int((done > 0).sum())
    rel = (done[:nt] - t0) / 1000.0
    d = np.diff(np.concatenate([[0.0], rel]))
    print(f"n={n} nq={nq} k={k}: scan_ms={c.last_timing()[1]:.4f}, tiles traced {nt}, exit at {(int(head[3]) - t0) / 1000:.1f} us")
    print("  first 24 tiles, us each:", " ".join(f"{x:.1f}" for x in d[:24]), "| inserts:", " ".join(str(int(x & 0xFFFF)) for x in cnt[:24]))
    for lo in range(0, nt, 16):
        hi = min(nt, lo + 16)
        print(f"  tiles {lo:3d}-{hi - 1:3d}: end {rel[hi - 1]:8.1f} us, us/tile {d[lo:hi].mean():6.2f}, rare chunks/tile {(cnt[lo:hi] >> 16).mean():5.2f} of 8, insert iterations/tile {(cnt[lo:hi] & 0xFFFF).mean():6.2f}")
    ct = np.zeros(592, dtype=np.uint64); L.vl_scan_trace_ctas(c._h, ct.ctypes.data)
    ent = ct[:296].astype(np.int64); ex = ct[296:].astype(np.int64); m = ent > 0
    e0 = ent[m].min()
    print(f"  {int(m.sum())} CTAs: entry spread {(ent[m].max() - e0) / 1000:.1f} us, exits at {(ex[m].min() - e0) / 1000:.1f} .. {(ex[m].max() - e0) / 1000:.1f} us after the first entry (median {(np.median(ex[m]) - e0) / 1000:.1f})")
    c.close()
