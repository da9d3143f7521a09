#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

usage: python profiles/summarize.py <round-tag>        e.g. r01a
  gpurun_out/launches.csv            (ncu --metrics gpu__time_duration.sum --csv)  -> profiles/<tag>_launches.csv + _launch_shares.txt
  gpurun_out/prof_<kernel>.ncu-rep   (ncu --set full)                              -> profiles/<tag>_<kernel>_metrics.csv
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
           "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "sm__cycles_elapsed.max", "sm__inst_executed.sum", "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "derived__smsp__sass_thread_inst_executed_op_ffma_pred_on_x2"]


def main(tag):
    src = os.path.join(OUT, "launches.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
        with open(src) as f:
            lines = [l for l in f if not l.startswith("==")]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for row in csv.DictReader(lines):
            v = float(row["Metric Value"].replace(",", ""))
            u = row["Metric Unit"]
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
            k = row["Kernel Name"].split("(")[0]
            agg[k][0] += 1
            agg[k][1] += v
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(ROOT, "profiles", f"{tag}_launch_shares.txt"), "w") as o:
            o.write("# per-kernel share of device time, ncu gpu__time_duration.sum (cold-cache, serialised: compare shares, not absolutes)\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                o.write(f"{k[:70]:72s} launches={v[0]:5d} total_ms={v[1] / 1e3:10.3f} avg_us={v[1] / v[0]:10.1f} share={v[1] / tot:6.1%}\n")
        print(open(os.path.join(ROOT, "profiles", f"{tag}_launch_shares.txt")).read())
    for fn in sorted(os.listdir(OUT)):
        if not fn.endswith(".ncu-rep"):
            continue
        raw = subprocess.run(["ncu", "-i", os.path.join(OUT, fn), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr = rows[0]
        name = fn[: -len(".ncu-rep")].replace("prof_", "")
        with open(os.path.join(ROOT, "profiles", f"{tag}_{name}_metrics.csv"), "w") as o:
            w = csv.writer(o)
            w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
            kn = hdr.index("Kernel Name")
            w.writerow(["kernel", ""] + [r[kn][:60] for r in rows[2:]])
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    w.writerow([m, rows[1][i]] + [r[i] for r in rows[2:]])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
