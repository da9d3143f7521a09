#!/usr/bin/env python
"""Step 2 of the full-size parity check (runs where /root/reference was compiled into oracle/_ref): renders the SAME
full-size scene XMLs with the UNMODIFIED reference (8 threads, its own mt19937 streams) and compares the GPU
framebuffers brought back in gpurun_out/full_<cfg>.f32.  Writes profiles/r02_fullsize_parity.json (RT795_PARITY_OUT overrides the name).
usage: python profiles/fullsize_parity.py c3 c2 c4"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rt = ge.load_package()
out_path = os.path.join(ROOT, "profiles", os.environ.get("RT795_PARITY_OUT", "r02_fullsize_parity.json"))
res = json.load(open(out_path)) if os.path.exists(out_path) else {}


def bm(x, b):
    h, w, _ = x.shape
    return x[: h // b * b, : w // b * b].reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3))


for cfg in sys.argv[1:]:
    xml = bench.make_scene(rt, cfg, bench.WORKLOADS[cfg][0], "full")
    hs = rt.HostScene(xml)
    gpu = rt.read_f32(os.path.join(ROOT, "gpurun_out", f"full_{cfg}.f32"))
    with tempfile.TemporaryDirectory() as td:
        t0 = time.time()
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_raytracer"), xml], cwd=td, capture_output=True, text=True, check=True)
        wall = time.time() - t0
        ref = rt.read_f32(os.path.join(td, hs.image_name(0) + ".f32"))
    secs = float(r.stderr.split("RT795_RENDER_SECONDS")[1].split()[0])
    rel = np.abs(gpu - ref) / np.maximum(np.abs(ref), 1.0)
    entry = {"scene": rt.scenes.CONFIGS[cfg]["label"], "reference_render_seconds_8_threads_here": secs, "reference_wall_seconds": wall,
             "mean_radiance": {"gpu": float(gpu.mean()), "reference": float(ref.mean()), "rel_diff": float(abs(gpu.mean() - ref.mean()) / ref.mean())},
             "per_pixel_rel_err": {"mean": float(rel.mean()), "median": float(np.median(rel)), "p99": float(np.quantile(rel, 0.99))}}
    for b in (8, 32):
        a, g = bm(ref, b), bm(gpu, b)
        rb = np.abs(a - g) / np.maximum(np.abs(a), 1.0)
        entry[f"block{b}_rel_err"] = {"mean": float(rb.mean()), "p99": float(np.quantile(rb, 0.99)), "max": float(rb.max())}
    entry["note"] = ("both images are Monte-Carlo estimates with independent random streams (the reference's mt19937 is unseeded), so per-pixel "
                     "differences are sampling noise; block means and the global mean are the bias check")
    res[cfg] = entry
    print(cfg, json.dumps(entry))
    json.dump(res, open(out_path, "w"), indent=1)
