"""Where a sharded step spends its time after the scan (torchrun, one rank per GPU): CUDA events around the
local search (scan + local merge) and around the exchange (push + flags + final merge), per rank, plus the
spread of the ranks' scan completion.  usage: torchrun --nproc-per-node N scripts/exchange_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from oracle import synth
from vlite_b200 import _capi
from vlite_b200.sharded import ShardedSearcher

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, w, nq, k = 500_000, 128, 1024, 10
c = _capi.Corpus(w, device=lr, capacity_rows=n, base_row=rank * n)
c.set_stream(torch.cuda.current_stream().cuda_stream)
c.fill_synthetic(0, 0, n)
dq = torch.from_numpy(synth.uniform_queries(1, nq, w)).cuda()
s = ShardedSearcher(c)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
one = torch.zeros(1, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
orig_local = s._local_search
def timed_local(*a, **kw):
    ev[0].record(); orig_local(*a, **kw); ev[1].record()
s._local_search = timed_local
c.set_timing(True)
loc, exch, scan = [], [], []
for it in range(25):
    flush.zero_(); dist.all_reduce(one); torch.cuda.synchronize()
    s.search(dq, k, _capi.HAMMING); ev[2].record(); torch.cuda.synchronize()
    if it >= 5:
        loc.append(ev[0].elapsed_time(ev[1])); exch.append(ev[1].elapsed_time(ev[2])); scan.append(c.last_timing()[1])
res = torch.tensor([np.mean(scan), np.mean(loc), np.mean(exch)], device="cuda", dtype=torch.float64)
allr = [torch.zeros_like(res) for _ in range(world)]
dist.all_gather(allr, res)
if rank == 0:
    a = torch.stack(allr).cpu().numpy() * 1e3
    print(f"world={world}: per rank [scan kernel | scan + local merge | exchange (push + wait + final merge)] in us")
    for r in range(world): print(f"  rank {r}: {a[r, 0]:7.1f} {a[r, 1]:7.1f} {a[r, 2]:7.1f}")
    print(f"  mean: scan {a[:, 0].mean():.1f}, local {a[:, 1].mean():.1f} (merge + launches {a[:, 1].mean() - a[:, 0].mean():.1f}), exchange {a[:, 2].mean():.1f}; slowest rank's scan - mean {a[:, 0].max() - a[:, 0].mean():.1f}")
dist.destroy_process_group()
