#!/usr/bin/env python
"""Top stall sites from an `ncu --page source --csv` dump: python profiles/top_stalls.py <ncu-rep> [n]"""
import csv
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]] or 0)
    except ValueError:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
keys = ["stall_long_sb", "stall_wait", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_lg", "stall_mio",
        "stall_no_inst", "stall_dispatch", "stall_tex"]
print("kernel:", rows[0][1][:80] if rows and len(rows[0]) > 1 else "")
print("samples", tot, {k: round(sum(f(r, k) for r in data) / max(tot, 1), 3) for k in keys})
ti = sum(f(r, "Thread Instructions Executed") for r in data)
wi = sum(f(r, "Instructions Executed") for r in data)
print("warp instructions", wi, "avg active threads", ti / max(wi, 1))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    print(r[ix["# Samples"]].rjust(6), ("%.1f" % f(r, "Avg. Threads Executed")).rjust(5), "lsb", r[ix["stall_long_sb"]].rjust(5), "wait", r[ix["stall_wait"]].rjust(5),
          r[ix["Address"]][-5:], r[ix["Source"]][:110])
