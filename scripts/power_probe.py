"""Power, SM clock and throttle reasons while the tensor-core scan runs back to back (NVML, 20 ms polls):
is the steady-state tile rate clock-bound by the power limit?  usage: power_probe.py [seconds]"""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import pynvml
from oracle import synth
from vlite_b200 import _capi

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
limit = pynvml.nvmlDeviceGetPowerManagementLimit(h) / 1000.0
samples, stop = [], threading.Event()


def poll():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)


for name, n, nq, path in (("tensor-core scan, 4M x 1024 queries", 4_000_000, 1024, 0), ("integer-pipe scan, 4M x 1024 queries", 4_000_000, 1024, 1),
                          ("HBM-bound scan, 32M rows x 1 query", 32_000_000, 1, 0)):
    c = _capi.Corpus(128, capacity_rows=n); c.fill_synthetic(0, 0, n)
    c.set_scan_path(path)
    dq = torch.from_numpy(synth.uniform_queries(1, nq, 128)).cuda()
    ds = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); dr = torch.empty((nq, 10), dtype=torch.int64, device="cuda")
    for _ in range(3): c.search(dq, 10, 0, None, ds, dr)
    torch.cuda.synchronize()
    samples.clear(); stop.clear()
    th = threading.Thread(target=poll, daemon=True); th.start()
    c.set_timing(True)
    t0 = time.perf_counter(); first = last = None; it = 0
    while time.perf_counter() - t0 < secs:
        for _ in range(8): c.search(dq, 10, 0, None, ds, dr)
        torch.cuda.synchronize()
        ms = c.last_timing()[1]
        first = ms if first is None else first
        last = ms; it += 8
    stop.set(); th.join()
    t_mid = [s for s in samples if s[0] - t0 > 0.3 * secs]
    pw = np.array([s[1] for s in t_mid]); ck = np.array([s[2] for s in t_mid]); rs = 0
    for s in t_mid: rs |= s[3]
    names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal", 0x40: "hw_thermal", 0x80: "hw_power_brake"}
    print(f"{name}: scan {first:.3f} ms (first) -> {last:.3f} ms (after {secs:.0f} s, {it} launches); power median {np.median(pw):.0f} W, max {pw.max():.0f} W of {limit:.0f} W limit; "
          f"SM clock median {np.median(ck):.0f} MHz, min {ck.min():.0f}; reasons {[v for k, v in names.items() if rs & k]}", flush=True)
    c.close()
