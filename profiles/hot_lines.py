#!/usr/bin/env python
"""Per-source-line hot spots of one kernel: joins `ncu --page source` (per SASS instruction samples / executed counts) with
`nvdisasm -g` line info of the same build.   usage: python profiles/hot_lines.py <ncu-rep> <kernel-substring> [n]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = sys.argv[4] if len(sys.argv) > 4 else "0"      # which launch of the report (ncu --launch-skip)
base = re.sub(r"I[Lib0-9E]+$", "", kern)             # ncu filters on the base name; the mangled suffix picks the instantiation here
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + base, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == "Address" or r[0] in seen:
        continue
    seen.add(r[0])
    data.append(r)
mangled = rows[0][1] if rows and len(rows[0]) > 1 else ""
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "raytracer-795_b200", "lib", "librt795.so")], cwd=td, capture_output=True)
    cub = sorted(glob.glob(os.path.join(td, "core*.cubin")))[0]
    dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.splitlines()
# locate the function whose demangled name contains `kern` and whose instruction count matches
funcs, cur, line = {}, None, None
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        cur = m.group(1); funcs[cur] = []; line = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        funcs[cur].append(line)
cands = [k for k in funcs if kern in k and 0 <= len(funcs[k]) - len(data) <= 8]      # nvdisasm also lists the trailing padding NOPs
if not cands:
    print("no function with", len(data), "instructions matching", kern, {k: len(v) for k, v in funcs.items() if kern in k})
    sys.exit(1)
lines = funcs[cands[0]]


def f(r, k):
    try:
        return float(r[ix[k]] or 0)
    except ValueError:
        return 0.0


agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r, ln in zip(data, lines):
    a = agg[ln]
    a[0] += f(r, "# Samples"); a[1] += f(r, "Instructions Executed"); a[2] += f(r, "Thread Instructions Executed")
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{cands[0][:70]}  samples={ts:.0f} warp_instr={ti:.0f} avg_active={sum(a[2] for a in agg.values()) / max(ti, 1):.1f}")
src_cache = {}
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    text = ""
    if ln:
        p = os.path.join(ROOT, "raytracer-795_b200", "csrc", ln[0])
        if os.path.exists(p):
            src_cache.setdefault(p, open(p).read().splitlines())
            text = src_cache[p][ln[1] - 1].strip()[:90] if ln[1] - 1 < len(src_cache[p]) else ""
    print(f"{a[0] / max(ts, 1):6.1%} samples {a[1] / max(ti, 1):6.1%} instr  act {a[2] / max(a[1], 1):4.1f}  {ln}  {text}")
