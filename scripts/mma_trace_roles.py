"""Hand-off timeline of ONE steady-state tile (tile 64 of CTA (0,0)) of the tensor-core scan across its warp roles
(vl_scan_trace_roles): when each role starts waiting, is released, and finishes -- the critical path of a tile."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from oracle import synth
from vlite_b200 import _capi
L = _capi.load()
L.vl_scan_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.vl_scan_trace_roles.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
NAMES = {0: "TMA: waits for a free box", 1: "TMA: box free, load issued", 8: "MMA: waits for the accumulator buffer", 9: "MMA: buffer free",
         10: "MMA: slab 0 full", 11: "MMA: slab 1 full", 12: "MMA: slab 2 full", 13: "MMA: slab 3 full", 14: "MMA: tile committed",
         16: "EXP: waits for the box", 17: "EXP: box landed", 18: "EXP: slab 0 ring slot free", 19: "EXP: slab 0 written",
         20: "EXP: slab 1 slot free", 21: "EXP: slab 1 written", 22: "EXP: slab 2 slot free", 23: "EXP: slab 2 written",
         24: "EXP: slab 3 slot free", 25: "EXP: slab 3 written",
         32: "EPI0: waits for the accumulator", 33: "EPI0: accumulator ready", 34: "EPI0: chunk 0 loaded", 35: "EPI0: chunk 2 loaded",
         36: "EPI0: chunk 4 loaded", 37: "EPI0: chunk 6 loaded", 40: "EPI0: tile done",
         48: "EPI1: waits for the accumulator", 49: "EPI1: accumulator ready", 50: "EPI1: chunk 1 loaded", 51: "EPI1: chunk 3 loaded",
         52: "EPI1: chunk 5 loaded", 56: "EPI1: tile done"}
for n, w, nq, metric in [(4_000_000, 128, 1024, 0), (4_000_000, 32, 1024, 0), (4_000_000, 128, 1024, 1)]:
    c = _capi.Corpus(w, capacity_rows=n); c.fill_synthetic(0, 0, n)
    dq = torch.from_numpy(synth.uniform_queries(1, nq, w)).cuda()
    ds = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); dr = torch.empty((nq, 10), dtype=torch.int64, device="cuda")
    L.vl_scan_trace(c._h, None)
    c.set_timing(True)
    for _ in range(4): c.search(dq, 10, metric, None, ds, dr)
    r = np.zeros(64, dtype=np.uint64); L.vl_scan_trace_roles(c._h, r.ctypes.data)
    r = r.astype(np.int64); t0 = r[r > 0].min()
    print(f"n={n} W={w} nq={nq} metric={metric}: scan {c.last_timing()[1]:.3f} ms; tile 64 of CTA (0,0), ns after the earliest stamp:")
    for slot, ts in sorted(((s, t) for s, t in enumerate(r) if t > 0), key=lambda x: x[1]):
        print(f"   {ts - t0:7d}  {NAMES.get(slot, 'slot %d' % slot)}")
    c.close()
