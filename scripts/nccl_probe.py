"""2+ GPU probe: where does a sharded step spend its time (scan / all-gather / merge)?"""
import os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from vlite_b200 import _capi
from vlite_b200.sharded import ShardedSearcher
from oracle import synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows, w, nq, k = 500_000, 128, 1024, 10
c = _capi.Corpus(w, device=lr, capacity_rows=rows, base_row=rank * rows)
c.set_stream(torch.cuda.current_stream().cuda_stream)
c.fill_synthetic(0, 0, rows)
dq = torch.from_numpy(synth.uniform_queries(1, nq, w)).to(dev)
s = ShardedSearcher(c)
send, recv, out_s, out_r = s._buffers(nq, k)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(8):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev[0].record()
    nb = nq * k
    ls = send[: nb * 4].view(torch.int32).view(nq, k); lrr = send[nb * 4:].view(torch.int64).view(nq, k)
    c.search(dq, k, 0, None, ls, lrr)
    ev[1].record()
    dist.all_gather_into_tensor(recv.view(-1), send)
    ev[2].record()
    s._merge(recv, world, nq, k, out_s, out_r, lr, torch.cuda.current_stream().cuda_stream)
    ev[3].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if rank == 0:
        print(f"it{it}: scan {ev[0].elapsed_time(ev[1]):.3f} ms, allgather {ev[1].elapsed_time(ev[2]):.3f} ms, "
              f"merge {ev[2].elapsed_time(ev[3]):.3f} ms, wall {1e3*(t1-t0):.3f} ms", flush=True)
dist.destroy_process_group()
