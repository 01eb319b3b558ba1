#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

  python profiles/make_summary.py <tag> <launches.csv> [<kernel>=<report.ncu-rep> ...]

* launch list  (ncu --metrics gpu__time_duration.sum --clock-control none --csv): per-kernel totals and SHARE of the step
* full capture (ncu --set full --import-source on): the handful of metrics DESIGN.md / bench.py quote
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
       "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
       "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio"]


def launch_table(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(list)
    for r in rows[1:]:
        if len(r) > vi and r[mi] == "gpu__time_duration.sum":
            agg[r[ki]].append(float(r[vi].replace(",", "")))
    tot_ours = sum(sum(v) for k, v in agg.items() if "cbk::" in k or k.startswith("k_") or "k_dr" in k or "k_ac" in k or "k_token" in k)
    out = ["| kernel | launches | total us | mean us | share of our kernels |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        ours = ("cbk::" in k) or ("k_dr" in k) or ("k_ac" in k) or ("k_token" in k) or ("k_edit" in k)
        name = k.split("(")[0][-70:]
        out.append("| %s%s | %d | %.1f | %.1f | %s |" % ("" if ours else "(torch) ", name, len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3,
                                                         "%.3f" % (sum(v) / tot_ours) if ours and tot_ours else "-"))
    return "\n".join(out)


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = ["| metric | value | unit |", "|---|---|---|"]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    for m in RAW:
        if m in hdr:
            i = hdr.index(m)
            out.append("| %s | %s | %s |" % (m, vals[i], units[i]))
    return name, "\n".join(out)


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    md = ["# ncu summary %s" % tag, "",
          "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline`",
          "(per-launch times are cold-cache and serialised: compare SHARES, not absolutes; bench numbers come from CUDA events).", "",
          launch_table(launches), ""]
    dst = os.path.join(HERE, "%s_launches.csv" % tag)
    with open(dst, "w") as fh:
        fh.writelines(l for l in open(launches) if not l.startswith("=="))
    for arg in sys.argv[3:]:
        k, rep = arg.split("=", 1)
        name, table = raw_metrics(rep)
        md += ["## `ncu --set full --clock-control none --import-source on -k regex:%s`" % k, "", "`%s`" % name[:160], "", table, ""]
    with open(os.path.join(HERE, "%s_summary.md" % tag), "w") as fh:
        fh.write("\n".join(md) + "\n")
    print("wrote", os.path.join(HERE, "%s_summary.md" % tag))


if __name__ == "__main__":
    main()
