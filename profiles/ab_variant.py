#!/usr/bin/env python
"""One line per library variant (RT795_CORE_SO): frame time, per-stage times, SIMT diagnostics of the traversal kernels.
usage: RT795_CORE_SO=... python profiles/ab_variant.py [config] [frames] [spp]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rt = ge.load_package()
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kw = bench.WORKLOADS[cfg][0]
xml = bench.make_scene(rt, cfg, dict(kw, spp=spp) if spp else kw, f"full_spp{spp}" if spp else "full")
hs = rt.HostScene(xml)
sc = rt.Scene(hs.desc, 0)
cam = hs.camera(0)
integ = rt.RT_INTEGRATOR_PATH if cfg == "c5" else rt.RT_INTEGRATOR_REFERENCE
best = 1e9
for _ in range(frames):
    img = sc.render(cam, rt.default_opts(integrator=integ))
    best = min(best, sc.stats().ms_total)
sc.render(cam, rt.default_opts(integrator=integ, collect_counters=1))
st = sc.stats()
stage = f"extend {st.ms_extend:.2f} shade {st.ms_shade:.2f} shadow {st.ms_shadow:.2f}"
sc.render(cam, rt.default_opts(integrator=integ, collect_counters=3))
st = sc.stats()
d = st.diag
simt = (f"closest: node lanes {st.node_visits / max(1, d[0]):.1f} leaf lanes {d[3] / max(1, d[2]):.1f} leaf execs {d[2] / 1e6:.1f}M | "
        f"occl: node lanes {st.node_visits_shadow / max(1, d[4]):.1f} leaf lanes {d[7] / max(1, d[6]):.1f} leaf execs {d[6] / 1e6:.1f}M "
        f"nodes/ray {st.node_visits_shadow / max(1, st.rays_shadow):.2f} tris/ray {st.prim_tests_shadow / max(1, st.rays_shadow):.2f} | "
        f"closest nodes/ray {st.node_visits / max(1, st.rays_primary + st.rays_secondary):.2f} tris/ray {st.prim_tests / max(1, st.rays_primary + st.rays_secondary):.2f}")
print(f"{os.environ.get('RT795_CORE_SO', 'default')}: {cfg} frame {best:.2f} ms mean {img.mean():.5f} | {stage} | {simt}")
