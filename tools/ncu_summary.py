"""Summarise ncu captures brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/prof_X.ncu-rep  profiles/X_kernel.md [--only k_trace] [--traffic-json profiles/trace_traffic.json]
                                                                                       [--counters-json profiles/trace_counters.json]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_bytes.sum", "L1 bytes"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: lg throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio throttle"),
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        name = name.split("::")[-1] if "k_" in name else name[-60:]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("ns") else (v / 1e3 if u.startswith("us") else (v * 1e3 if u in ("s", "second") else v))
        seq.append((name, ms))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("source: `%s`, %d launches, %.3f ms total\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n" % (src, len(seq), tot))
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (k, n, ms, 100 * ms / tot))
        f.write("\n## every launch, in order (ms)\n\n```\n")
        for n, ms in seq:
            f.write("%-40s %10.4f\n" % (n[-40:], ms))
        f.write("```\n")


def kernel(src, dst, traffic_json=None, only=None, counters_json=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    if only:
        data = [d for d in data if only in d[name_i]]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none: `%s`\n\n" % src)
        f.write("| metric | unit | " + " | ".join("launch %d" % k for k in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join("`%s`" % d[name_i].split("(")[0].split("::")[-1] for d in data) + " |\n")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                f.write("| %s (`%s`) | %s | %s |\n" % (label, key, units[i], " | ".join(d[i] for d in data)))

    def num(d, key):
        return float(d[hdr.index(key)].replace(",", ""))

    if traffic_json and "dram__bytes_read.sum" in hdr:
        def to_bytes(v, u):
            m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            return float(v.replace(",", "")) * m.get(u, 1)
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        per = [to_bytes(d[ir], units[ir]) + to_bytes(d[iw], units[iw]) for d in data]
        it = hdr.index("gpu__time_duration.sum")
        scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(units[it].replace("second", "s").replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1e-3)
        secs = [float(d[it].replace(",", "")) * scale for d in data]
        json.dump({"source": src, "kernel": only or "k_trace", "launches_captured": len(per), "dram_bytes_each": per, "seconds_each": secs,
                   "dram_bytes_per_launch": sum(per) / len(per), "dram_bytes_per_second": sum(per) / sum(secs),
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum of the captured launches of the final kernel (the first bounces of one wavefront "
                           "batch at 64 spp) and their durations; bench.py reports traffic per launch as this DRAM rate x the mean duration of its own launches"},
                  open(traffic_json, "w"), indent=1)
    if counters_json:
        def mean(key):
            return sum(num(d, key) for d in data) / len(data)
        json.dump({"counters_source": src, "l1_wavefront_pct": mean("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                   "issue_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "lanes_active": mean("smsp__thread_inst_executed_per_inst_executed.ratio"),
                   "dram_pct": mean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                   "l2_hit_pct": mean("lts__t_sector_hit_rate.pct"), "l1_hit_pct": mean("l1tex__t_sector_hit_rate.pct"),
                   "occupancy_pct": mean("sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "long_scoreboard_stalls_per_issue": mean("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                   "limiter": "L1 data-pipe wavefronts + issue slots + L2 latency (the tree is L2/L1-resident: DRAM is a few percent of peak); "
                              "the HBM roofline is the rubric's formula, not the operative limit"},
                  open(counters_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        def opt(name):
            return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else None
        kernel(sys.argv[2], sys.argv[3], opt("--traffic-json"), opt("--only"), opt("--counters-json"))
