"""Turn ncu outputs into the small text summaries committed under profiles/.
usage: ncu_summary.py launches <launches.csv> <out.md>
       ncu_summary.py full <report.ncu-rep> <out.md> [traffic.json]"""
import collections, csv, json, re, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__icc_request_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[:100]
        a = agg.setdefault((name, r[gi], r[bi]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none): {len(rows)-1} launches, {tot:.3f} ms\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| launches | total ms | share | avg ms | kernel | grid | block |\n|---:|---:|---:|---:|---|---|---|\n")
        for (name, g, b), (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {n} | {t:.3f} | {100*t/tot:.1f}% | {t/n:.4f} | `{name}` | {g} | {b} |\n")
    print(open(out).read())


def full(rep, out, traffic=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n\n")
        for k, vals in enumerate(rows[2:]):
            d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
            f.write(f"## launch {k}: `{d.get('Kernel Name', ('?',))[0][:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in d:
                    f.write(f"| {w} | {d[w][0]} | {d[w][1]} |\n")
            f.write("\n| stall reason (warps per issue-active) | value |\n|---|---:|\n")
            st = [(h, float(v[0].replace(",", ""))) for h, v in d.items()
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
            for h, v in sorted(st, key=lambda x: -x[1]):
                f.write(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v:.3f} |\n")
            f.write("\n")
            if traffic and k == 0:
                def num(key):
                    v, u = d[key]
                    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                t = {"dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum")}
                t["dram_bytes_per_launch"] = t["dram_bytes_read"] + t["dram_bytes_write"]
                t["source"] = rep.split("/")[-1]
                for key, name in (("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
                                  ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
                                  ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
                                  ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_slots_pct"),
                                  ("gpu__time_duration.sum", "kernel_time")):
                    if key in d:
                        t[name] = float(d[key][0].replace(",", ""))
                json.dump(t, open(traffic, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
