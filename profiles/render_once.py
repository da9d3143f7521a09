#!/usr/bin/env python
"""Renders the workload N times through the C ABI (no torch): the short command used for `ncu --set full` captures.
usage: python profiles/render_once.py [config] [frames] [--small]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rt = ge.load_package()
args = [a for a in sys.argv[1:] if not a.startswith("-")]
cfg = args[0] if args else "c4"
frames = int(args[1]) if len(args) > 1 else 2
full_kw, sample_kw = bench.WORKLOADS[cfg]
small = "--small" in sys.argv
xml = bench.make_scene(rt, cfg, sample_kw if small else full_kw, "small" if small else "full")
hs = rt.HostScene(xml)
sc = rt.Scene(hs.desc, 0)
cam = hs.camera(0)
integ = rt.RT_INTEGRATOR_PATH if cfg == "c5" else rt.RT_INTEGRATOR_REFERENCE
for _ in range(frames):
    img = sc.render(cam, rt.default_opts(integrator=integ))
st = sc.stats()
rays = st.rays_primary + st.rays_secondary + st.rays_shadow
print(f"{cfg}: {st.ms_total:.3f} ms/frame, {rays / st.ms_total / 1e3:.1f} Mrays/s, mean radiance {img.mean():.4f}")
