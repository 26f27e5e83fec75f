"""profiles/r1_sweep.md from the JSON lines scripts/scan_bench.py printed for the sweep shapes.
usage: sweep_table.py <sweep.jsonl> <out.md>   (shapes: see SHAPES below)"""
import json
import sys


def shapes():
    specs = []
    for w in (128, 64):
        for n in (500000, 32000000):
            for metric in (0, 1):
                for nq in (1, 2, 4, 8, 16, 32, 64, 256, 1024):
                    specs.append(f"{n},{w},{nq},10,{metric}")
    n_main = len(specs)
    ks = (1, 5, 10, 32, 33, 64, 100, 256, 1000, 2048)
    specs += [f"500000,128,1024,{k},0" for k in ks]
    specs += [f"32000000,128,1,{k},0" for k in (1, 10, 100, 1000)]
    return specs, n_main, len(ks)


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "shapes":
        print(" ".join(shapes()[0]))
        sys.exit(0)
    recs = [json.loads(line) for line in open(sys.argv[1]) if line.strip()]
    _, n_main, n_k = shapes()
    main, ksw, k1 = recs[:n_main], recs[n_main:n_main + n_k], recs[n_main + n_k:]
    out = ["# Scan sweep, one B200 (final kernels of the round that wrote this file)", "",
           "`scripts/scan_bench.py $(python scripts/sweep_table.py shapes)`: device-resident queries and outputs, CUDA events",
           "inside the library around scan + merge, median of 5 after 2 warm-ups, k = 10 unless stated. The 500k-row corpora",
           "(32-64 MB) stay in L2 between repetitions; the 32M-row ones (2-4 GB) do not. `roof` = max(N*W / 6554.9 GB/s,",
           "N*Q*(W/4) / 4.62 T POPC/s) / scan time: the north-star roofline (slower of corpus bytes at HBM bandwidth and popcount",
           "work at INT-pipe peak). It exceeds 1 where the carry-save adders (HAMMING) or IDP4A (XOR-sum) beat the plain-POPC",
           "pipe bound.", "",
           "| rows | W | metric | Q | scan ms | total ms | queries/s | corpus GB/s | T POPC/s | roof |",
           "|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|"]
    for r in main:
        out.append(f"| {r['n']:,} | {r['w']} | {'HAMMING' if r['metric'] == 0 else 'XORSUM'} | {r['nq']} | {r['scan_ms']:.4f} "
                   f"| {r['total_ms']:.4f} | {r['qps']:,} | {r['GBps']:,} | {r['Tpopc']:.2f} | {r['roof_frac']:.2f} |")
    out += ["", "## k sweep at config 2 (500k x 128 B, 1024 queries, HAMMING)", "",
            "k <= 16: per-thread register lists of the tensor-core kernel; larger k: sampled threshold + collect + select.", "",
            "| k | scan ms | total ms | queries/s | roof |", "|---:|---:|---:|---:|---:|"]
    for r in ksw:
        out.append(f"| {r['k']} | {r['scan_ms']:.4f} | {r['total_ms']:.4f} | {r['qps']:,} | {r['roof_frac']:.2f} |")
    out += ["", "## k sweep, HBM-bound leg (32M x 128 B = 4.096 GB, 1 query, HAMMING)", "",
            "| k | scan ms | total ms | corpus GB/s | roof |", "|---:|---:|---:|---:|---:|"]
    for r in k1:
        out.append(f"| {r['k']} | {r['scan_ms']:.4f} | {r['total_ms']:.4f} | {r['GBps']:,} | {r['roof_frac']:.2f} |")
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
