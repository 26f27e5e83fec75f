#!/usr/bin/env python
"""bench.py -- the retrieval hot path on N B200s of one node (contract in the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--legs a,b,...]

Headline step (BASELINE.json configs[1]): a 500k x 128 B synthetic ubinary corpus per GPU shard, a
batch of 1024 queries, top_k = 10, HAMMING.  A "step" = one vl_search over the shard (N > 1: + the
all-gather of the per-shard top-k and the merge).  Row-sharded WEAK scaling: every rank holds its
own 500k-row shard of an N x 500k corpus, all ranks scan all queries.

``value``   queries/s in 500k-row-corpus equivalents (Q x total_rows / 500k per step / time; plain
            queries/s at N = 1), inputs resident in HBM, CUDA events, max over ranks.  The reference
            arm reports the SAME unit, so the driver's ratio compares like with like at every N;
            ``qps_over_whole_corpus`` is the plain queries/s over the N x 500k corpus.
``e2e``     the same metric through the C ABI with HOST buffers (pinned queries in, results out).
``roofline`` of the dominant kernel.  Batches of >= 16 queries run on the tensor-core scan
            (HAMMING: tcgen05.mma kind::mxf4 on rows expanded bit -> +-1.0 e2m1 in shared memory, exact
            fp32 sums): bound "tensor", achieved = 2 N Q 8W ops / kernel time, peak = 4 x the measured
            bf16 cuBLAS rate of MEASURED_PEAKS.json (e2m1 issues at four times the bf16 rate on B200;
            the int8 form of the same kernel is timed beside it under ``extra``).  The north star's own
            roofline (popcount work at the INT-pipe peak, corpus bytes at HBM bandwidth) is kept
            beside it as ``north_star_view``; the HBM-bound regime (1-2 queries) is ``roofline_hbm``.
``cpu_baseline`` the oracle's torch-op port of the reference's EmbeddingModel.search
            (vlite/model.py:76-90) timed on ALL of this box's host cores on a bounded sample.
``legs``    every other BASELINE config as a parity-gated measurement (bench_legs.py): N = 1:
            cfg1 (VLite.retrieve latency), hbm (1-2 queries over 4 GB), cfg3 (100M rows, 4096 queries,
            top-1000 -> int8 rescore), cfg5 (filters, streaming add, 4 GiB CTX load); N > 1: hbm (16 GB
            per GPU: the 1B-vector corpus at 8 GPUs), strong (500k rows split N ways), group (the
            in-process shard group over the N GPUs, driven by rank 0), ctxfan (one 4 GiB CTX file
            loaded by the N ranks, each its own byte range).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_legs as legs_mod  # noqa: E402

ROWS_PER_GPU = 500_000
WIDTH = 128
N_QUERIES = 1024
TOP_K = 10
HBM_FALLBACK_GBS = 6650.0
BF16_FALLBACK_TFLOPS = 1590.0
VALUE_DEFINITION = ("Q x total_rows / 500k per step / time: queries/s in 500k-row-corpus equivalents "
                    "(plain queries/s at N = 1); both arms use this unit")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": float(d["hbm_gbs"]), "bf16": float(d["bf16_tflops"]),
                "bf16_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": HBM_FALLBACK_GBS, "bf16": BF16_FALLBACK_TFLOPS, "bf16_sustained": 1400.0,
            "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 10 ms from a
    thread (nvidia-smi -lms is too slow to start for a sub-second region; it is the fallback)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index, self.samples, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self._smi = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
            try:
                self._smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                              "-lms", "50", "-i", str(self.index)],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                threading.Thread(target=self._read_smi, daemon=True).start()
            except OSError:
                self._smi = None

    def _read_smi(self):
        for line in self._smi.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                self.mask |= int(f[2], 16)
            except (ValueError, IndexError):
                pass

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1)
        if self._smi:
            self._smi.terminate()
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples)}


def make_queries(n_rows_total: int, nq: int, width: int):
    """Half clustered (true neighbours exist), half uniform (tie-heavy).  Same on every rank."""
    from oracle import synth
    qc, _ = synth.clustered_queries(1, 0, n_rows_total, nq // 2, width)
    qu = synth.uniform_queries(1, nq - nq // 2, width)
    return np.ascontiguousarray(np.concatenate([qc, qu]))


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's torch-op port of EmbeddingModel.search on host cores
# ------------------------------------------------------------------------------------------
def time_reference(steps, warmup: int, budget_s: float = 150.0, target_s: float = 15.0):
    """steps=None: as many one-query calls as fit ~target_s seconds of CPU work (6..1024)."""
    from oracle import ref_port, synth
    cores = legs_mod.use_all_host_cores()        # torchrun sets OMP_NUM_THREADS=1: undo it for the CPU arm
    rows = ROWS_PER_GPU
    queries = make_queries(rows, 8, WIDTH)
    # probe on 50k rows to size the per-step sample so that steps+warmup fit the budget
    probe_n = 50_000
    e = ref_port.corpus_as_reference_f32(synth.corpus_rows(0, 0, probe_n, WIDTH))
    from oracle.search import to_ref_encoding
    q_ref = to_ref_encoding(queries)
    ref_port.search_port(q_ref[0], e, TOP_K)
    t0 = time.perf_counter()
    ref_port.search_port(q_ref[0], e, TOP_K)
    per_row = (time.perf_counter() - t0) / probe_n
    n_sample = rows
    if steps is None:
        steps = int(min(1024, max(6, target_s / max(per_row * rows, 1e-9))))
    est = per_row * rows * (steps + warmup)
    if est > budget_s:
        n_sample = max(probe_n, int(rows * budget_s / est) // 1000 * 1000)
    e = ref_port.corpus_as_reference_f32(synth.corpus_rows(0, 0, n_sample, WIDTH))
    for i in range(warmup):
        ref_port.search_port(q_ref[i % 8], e, TOP_K)
    t0 = time.perf_counter()
    for i in range(steps):
        ref_port.search_port(q_ref[i % 8], e, TOP_K)     # one query per call, like main.py:116-117
    dt = (time.perf_counter() - t0) / steps
    qps = (1.0 / dt) * (n_sample / rows)                  # search cost is linear in N
    sample = (f"{steps} sequential EmbeddingModel.search calls (1 query, k={TOP_K}) over "
              f"{n_sample} of the {rows} x {WIDTH} B rows held as float32"
              + ("" if n_sample == rows else "; q/s scaled linearly to the full corpus")
              + "; q/s per 500k rows (= 500k-row-corpus equivalents at any shard count: the search is linear in rows)")
    return qps, dt * 1e3, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    qps, ms, cores, sample = time_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "queries_per_sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 (via float32 storage)",
        "data": "synthetic",
        "config": workload_config(args.gpus, reference=True),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "value_definition": VALUE_DEFINITION,
        "qps_over_whole_corpus": qps / max(1, args.gpus),
    }
    if not args.no_legs:
        # BASELINE configs[0] as the reference pays for it: dict rebuild + search, reported separately
        line["cfg1_reference_port"] = legs_mod.cfg1_reference_port()
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, reference: bool = False):
    return {
        "workload": (f"BASELINE configs[1]: {ROWS_PER_GPU} x {WIDTH} B ubinary rows per GPU shard "
                     f"({n_gpus} shard(s), row-sharded), batch of {N_QUERIES} queries, top_k={TOP_K}, HAMMING"),
        "rows_per_gpu": ROWS_PER_GPU, "width_bytes": WIDTH, "queries": N_QUERIES, "top_k": TOP_K,
        # the reference's live score is the XOR-sum of byte values (vlite/model.py:82); ours is timed on
        # true Hamming, the north-star metric (same memory traffic; our XOR-sum time is under extra)
        "metric_fn": "xorsum" if reference else "hamming", "parallelism": f"row-shard x{n_gpus}",
        "l2": ("corpus (64 MB) fits L2: a 256 MB buffer is overwritten between timed iterations (L2 flush); at N > 1 an "
               "untimed one-element all-reduce re-aligns the GPUs after each flush, before the step's first event"),
    }


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import cscan, synth
    from vlite_b200 import _capi
    from vlite_b200.sharded import ShardedSearcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION (set in some images) prints a banner
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        # ... and NCCL can still print its version line from a config file: stdout carries ONE JSON line,
        # so fd 1 points at /dev/null while the communicators come up
        sys.stdout.flush()
        saved_fd, null_fd = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null_fd, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
            cpu_group = dist.new_group(backend="gloo")  # host-side waits that put no kernel on the GPUs
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
            os.close(null_fd)
    _capi.load()
    if _capi.device_count() < 1:
        raise SystemExit("no sm_100 device: vlite_b200 has no CPU fallback")
    peaks = load_peaks()
    hbm_peak = peaks["hbm"]

    rows, w, nq, k = ROWS_PER_GPU, WIDTH, N_QUERIES, TOP_K
    total_rows = rows * world
    corpus = _capi.Corpus(w, device=local_rank, capacity_rows=rows, base_row=rank * rows)
    corpus.set_stream(torch.cuda.current_stream().cuda_stream)
    corpus.fill_synthetic(0, 0, rows)
    queries = make_queries(total_rows, nq, w)
    dq = torch.from_numpy(queries).to(dev)
    searcher = ShardedSearcher(corpus)

    # ---- parity gate before any timing: device result vs the CPU oracle ----
    out_s, out_r = searcher.search(dq, k, _capi.HAMMING)
    torch.cuda.synchronize()
    got_s = out_s.cpu().numpy().view(np.uint32)
    got_r = out_r.cpu().numpy().view(np.uint64)
    if rank == 0:
        sub = np.r_[0:16, nq - 16:nq]
        host_rows = np.concatenate([synth.corpus_rows(0, g * rows, rows, w) for g in range(world)])
        es, er = cscan.search(queries[sub], host_rows, k, _capi.HAMMING)
        if not (np.array_equal(got_s[sub], es) and np.array_equal(got_r[sub], er)):
            print(json.dumps({"error": "parity gate failed: device top-k differs from the oracle"}), flush=True)
            raise SystemExit(1)
        del host_rows

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    corpus.set_timing(True)
    for _ in range(args.warmup):
        flush.zero_()
        searcher.search(dq, k, _capi.HAMMING)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _capi.launch_count()
    scan_ms = []
    barrier()
    tick = torch.zeros(1, dtype=torch.int32, device=dev)

    def align():
        """the 256 MB flush finishes at slightly different times on every GPU: an (untimed) one-element
        all-reduce lines the GPUs up again before the timed events, so a step's time is scan + exchange,
        not scan + exchange + how far the flush let the ranks drift apart"""
        if world > 1:
            dist.all_reduce(tick)

    for i in range(args.steps):
        flush.zero_()                              # L2 flush, outside the timed events
        align()
        ev0[i].record()
        searcher.search(dq, k, _capi.HAMMING)
        ev1[i].record()
        ev1[i].synchronize()
        scan_ms.append(corpus.last_timing()[1])    # the scan kernel alone (library's own events)
    barrier()
    launches = _capi.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    t_total = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_total.item()) / args.steps
    value = nq * world / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI (pinned queries in, host results out) ----
    hq = torch.from_numpy(queries).pin_memory()
    hs = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    hr = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    corpus.set_timing(False)

    def e2e_step():
        if world == 1:
            corpus.search(hq, k, _capi.HAMMING, None, hs, hr)        # H2D + scan + merge + D2H inside
        else:
            dq2 = hq.to(dev, non_blocking=True)
            s, r = searcher.search(dq2, k, _capi.HAMMING)
            hs.copy_(s, non_blocking=True)
            hr.copy_(r, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        flush.zero_()
        e2e_step()
    barrier()
    e2e_t = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()
        e2e_t += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    e2e_total = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_total.item()) / args.steps * 1e3
    e2e_value = nq * world / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (the tensor-core scan at this batch size) ----
    scan_avg_ms = float(np.mean(scan_ms))
    popc_peak, _ = _capi.microbench(0, 8192, device=local_rank)           # lane-POPC/s, measured live
    alg_bytes = rows * w + rows / 8 + nq * w + nq * k * 12
    alg_popc = rows * nq * (w / 4)
    alg_mac_ops = 2.0 * rows * nq * (w * 8)
    t_hbm = alg_bytes / (hbm_peak * 1e9)
    t_int = alg_popc / popc_peak
    tensor_peak = 4.0 * peaks["bf16"]                                        # e2m1 dense: 4 x the bf16 rate (TOP/s)
    traffic, traffic_src, pipes = None, None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_cfg2.json")
    if os.path.exists(tpath):       # a STATIC number: read from the committed ncu --set full capture of this kernel + workload
        tj = json.load(open(tpath))
        traffic = tj.get("dram_bytes_per_launch")
        traffic_src = f"not measured in this run: committed ncu capture profiles/{tj.get('source', '?')} of the same kernel and workload"
        pipes = {kk: tj[kk] for kk in ("tensor_pipe_pct", "alu_pipe_pct", "xu_pipe_pct", "fma_pipe_pct", "issue_slots_pct") if kk in tj}
    achieved_tops = alg_mac_ops / (scan_avg_ms * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "achieved": achieved_tops, "peak": tensor_peak, "unit": "TFLOP/s",
        "frac": achieved_tops / tensor_peak, "traffic": traffic, "traffic_source": traffic_src,
        "pipe_utilisation_ncu": pipes,
        "kernel": "mma_scan_kernel<W=128, CG=2, lists, e2m1> (tcgen05.mma.cta_group::2.kind::mxf4.block_scale, unit scales, fp32 TMEM accumulators)",
        "kernel_ms": scan_avg_ms,
        "algorithmic": {"mac_ops_per_launch": alg_mac_ops, "bytes_per_launch": alg_bytes, "popc_per_launch": alg_popc,
                        "per_unit": "per (query,row) pair at W=128: 1024 +-1 MAC = 2048 ops (tensor view), "
                                    "32 32-bit POPC (north-star INT view), W = 128 corpus bytes per row once per batch"},
        "peak_source": ("e2m1 (kind::mxf4) dense = 4 x bf16: 4 x " + f"{peaks['bf16']:.1f}" + " TFLOP/s cuBLAS bf16 burst, " + peaks["src"] +
                        " (nominal 9000; scripts/umma_probe.cu issues 8920 from fixed operands, profiles/r2_umma_probe.log); the unit is "
                        "TOP/s of exact +-1 multiply-adds (fp32 accumulation of integers < 2^24)"),
        "north_star_view": {"bound": "int_pipe" if t_int >= t_hbm else "hbm",
                            "achieved": alg_popc / (scan_avg_ms * 1e-3) / 1e12, "peak": popc_peak / 1e12, "unit": "TPOPC/s",
                            "frac": max(t_hbm, t_int) / (scan_avg_ms * 1e-3),
                            "note": "north_star's roofline = slower of corpus bytes at HBM bandwidth and popcount work at "
                                    "the INT-pipe peak (POPC peak: vl_microbench live in this run).  frac >> 1 because the "
                                    "popcount work runs on the tensor pipe as an exact +-1 contraction instead"},
        "hbm_view": {"achieved": rows * w / (scan_avg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": rows * w / (scan_avg_ms * 1e-3) / 1e9 / hbm_peak},
    }

    ctx = {"torch": torch, "dist": dist, "capi": _capi, "world": world, "rank": rank, "local_rank": local_rank,
           "dev": dev, "hbm_peak": hbm_peak, "barrier": barrier, "flush": flush, "no_cpu": args.no_cpu, "align": align}
    default_legs = ["extra", "hbm", "cfg1", "cfg3", "cfg5"] if world == 1 else ["hbm", "strong", "group", "ctxfan"]
    want = default_legs if args.legs == "default" else [x for x in args.legs.split(",") if x]
    if args.no_legs:
        want = []
    legs, failed = {}, []

    def run_leg(name, fn, collective=True):
        """collective legs run on every rank; a gate failure anywhere fails the leg everywhere"""
        t0 = time.perf_counter()
        err = None
        try:
            res = fn()
        except legs_mod.GateError as e:
            res, err = None, f"GATE FAILED: {e}"
        except Exception as e:            # noqa: BLE001 -- a broken leg must not take the headline down with it
            res, err = None, f"{type(e).__name__}: {e}"
        if world > 1 and collective:
            flag = torch.tensor([1 if err else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag.item()) and not err:
                err = "failed on another rank"
        if err:
            failed.append(name)
            legs[name] = {"error": err}
        else:
            legs[name] = res
        legs[name]["leg_seconds"] = time.perf_counter() - t0
        torch.cuda.empty_cache()

    extra = {}
    if "extra" in want and world == 1:
        # the reference's LIVE score (XOR-sum of byte values, vlite/model.py:82) on the same workload, and
        # the integer-pipe kernel (LOP3 + POPC, the north-star-literal scan) on the same workload
        ds = torch.empty((nq, k), dtype=torch.int32, device=dev)
        dr = torch.empty((nq, k), dtype=torch.int64, device=dev)
        corpus.set_timing(True)

        def timed(metric, path, flush_l2=True):
            corpus.set_scan_path(path)
            tt, ts = [], []
            for i in range(8):
                if flush_l2:
                    flush.zero_()
                corpus.search(dq, k, metric, None, ds, dr)
                if i >= 3:
                    tt.append(corpus.last_timing()[0])
                    ts.append(corpus.last_timing()[1])
            corpus.set_scan_path(0)
            return float(np.mean(tt)), float(np.mean(ts))
        t, s_ = timed(_capi.XORSUM, 0)
        extra["xorsum_same_workload"] = {"ms_per_step": t, "scan_ms": s_, "qps": nq / (t * 1e-3),
                                         "kernel": "mma_scan_kernel<128,2,lists,XORSUM,e2m1> (kind::mxf4, plane-major K, bit weights 2^(7-p) as the query operand's UE8M0 block scales)"}
        t, s_ = timed(_capi.XORSUM, 1)
        extra["xorsum_integer_pipe_kernel"] = {"ms_per_step": t, "scan_ms": s_, "qps": nq / (t * 1e-3),
                                               "kernel": "scan_topk_kernel<128,XORSUM,8,8> (LOP3 + IDP4A)"}
        t, s_ = timed(_capi.HAMMING, 6)
        extra["int8_tensor_kernel_same_workload"] = {
            "ms_per_step": t, "scan_ms": s_, "qps": nq / (t * 1e-3),
            "kernel": "mma_scan_kernel<128,2,lists,HAMMING,int8> (tcgen05.mma kind::i8, 256-row tiles): round 2's first tensor-core form"}
        t, s_ = timed(_capi.HAMMING, 1)
        extra["integer_pipe_kernel_same_workload"] = {
            "ms_per_step": t, "scan_ms": s_, "qps": nq / (t * 1e-3), "kernel": "scan_topk_kernel<128,HAMMING,8,8> (LOP3 carry-save + POPC)",
            "north_star_frac": max(t_hbm, t_int) / (s_ * 1e-3)}
        # the 64 MB corpus fits L2 -- the headline is timed with L2 flushed between
        # iterations; this is the same step with the corpus left resident
        t, s_ = timed(_capi.HAMMING, 0, flush_l2=False)
        extra["l2_resident_same_workload"] = {"ms_per_step": t, "qps": nq / (t * 1e-3)}
        corpus.set_timing(False)
    if "hbm" in want:
        run_leg("hbm", lambda: legs_mod.leg_hbm(ctx, 32_000_000 if world == 1 else 125_000_000))
    if "strong" in want and world > 1:
        run_leg("strong", lambda: legs_mod.leg_strong(ctx, queries))
    if "group" in want and world > 1:
        # rank 0 drives every GPU from one process; the other ranks wait on the HOST (gloo), so no
        # collective kernel spins on the GPUs meanwhile
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            run_leg("group", lambda: legs_mod.leg_group(ctx, queries), collective=False)
        dist.barrier(group=cpu_group)
    if "ctxfan" in want and world > 1:
        run_leg("ctxfan", lambda: legs_mod.leg_ctx_fanout(ctx))
    if "cfg1" in want and rank == 0:
        run_leg("cfg1", lambda: legs_mod.leg_cfg1(ctx), collective=False)
    if "cfg5" in want and world == 1:
        run_leg("cfg5", lambda: legs_mod.leg_cfg5(ctx))
    if "cfg3" in want and world == 1:
        corpus.close()
        corpus = None
        del flush
        ctx["flush"] = None
        torch.cuda.empty_cache()
        run_leg("cfg3", lambda: legs_mod.leg_cfg3(ctx))

    cpu_baseline = None
    if rank == 0 and not args.no_cpu:
        # N = 1: ~15 s of CPU work; N > 1: a shorter sample so that the scaling runs stay short
        qps, ms, cores, sample = time_reference(None, 2, budget_s=40.0 if world == 1 else 12.0,
                                                target_s=15.0 if world == 1 else 5.0)
        cpu_baseline = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        roofline_hbm = None
        if "hbm" in legs and "q1" in legs["hbm"]:
            roofline_hbm = dict(legs["hbm"]["q1"], traffic=None, gate=legs["hbm"].get("gate"),
                                rows_total=legs["hbm"]["rows_total"], gpus=world, kernel=legs["hbm"]["kernel"],
                                q2=legs["hbm"].get("q2"))
        line = {
            "metric": "queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "e2m1 x e2m1 -> f32 (packed bits expanded to +-1.0 4-bit floats in shared memory; every sum an exact integer)",
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": nq * w,
                    "d2h_bytes_per_step": nq * k * 12, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline, "clocks": clocks,
            "scan_GBps_aggregate": rows * w * world / (scan_avg_ms * 1e-3) / 1e9,
            "qps_over_whole_corpus": nq / (ms_per_step * 1e-3),
            "value_definition": VALUE_DEFINITION,
            "parity_gate": "32 of the 1024 queries bit-exact (scores and rows) vs the C oracle over the whole corpus, before timing",
            "legs": legs, "legs_failed": failed, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if corpus is not None:
        corpus.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-legs", action="store_true", help="headline step only (profiling runs)")
    ap.add_argument("--legs", default="default",
                    help="comma list of extra,hbm,cfg1,cfg3,cfg5 (N=1) / hbm,strong,group,ctxfan (N>1); default = all of them")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
