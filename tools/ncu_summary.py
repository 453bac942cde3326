#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run here, no GPU needed).

  tools/ncu_summary.py launches <launches.csv>          per-kernel totals and shares of a
                                                        `--metrics gpu__time_duration.sum` launch list
  tools/ncu_summary.py full <report.ncu-rep>            key metrics per profiled launch of a `--set full` report
"""
import collections
import csv
import re
import subprocess
import sys

OURS = ("k_prefilter", "k_seed", "k_filter", "k_literal", "k_finalize", "k_fq_", "k_part")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "")


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3, "s": 1e3}[row["Metric Unit"]]
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for k, v in agg.items() if any(o in k for o in OURS))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = v[1] / tot if any(o in k for o in OURS) else float("nan")
        print(f"{k:42s} launches={v[0]:4d} total_ms={v[1]:10.3f} share_of_path={share:.4f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    idx = [head.index(w) for w in WANT if w in head]
    kn = head.index("Kernel Name")
    for row in rows[2:]:
        print(f"\nKernel {row[kn][:100]}")
        for i in idx:
            print(f"  {head[i]:72s} {units[i]:16s} {row[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
