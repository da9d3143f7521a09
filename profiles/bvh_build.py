#!/usr/bin/env python
"""Host binned-SAH vs device PLOC BLAS build on a workload: build time, tree size and the frame time the tree gives
(= its quality).  usage: python profiles/bvh_build.py [config] [frames]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rt = ge.load_package()
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
xml = bench.make_scene(rt, cfg, bench.WORKLOADS[cfg][0], "full")
t0 = time.perf_counter()
hs = rt.HostScene(xml)
t_load = (time.perf_counter() - t0) * 1e3
cam = hs.camera(0)
integ = rt.RT_INTEGRATOR_PATH if cfg == "c5" else rt.RT_INTEGRATOR_REFERENCE
rt.Scene(hs.desc, 0, bvh_builder=rt.RT_BVH_HOST_SAH).close()          # CUDA context + first-use costs out of the way
out = {"config": cfg, "ms_host_load_scene": t_load, "builders": {}}
for name, b in (("host_sah", rt.RT_BVH_HOST_SAH), ("device_ploc", rt.RT_BVH_DEVICE_PLOC), ("device_sah", rt.RT_BVH_DEVICE_SAH)):
    t0 = time.perf_counter()
    sc = rt.Scene(hs.desc, 0, bvh_builder=b)
    t_create = (time.perf_counter() - t0) * 1e3
    st = sc.stats()
    rec = {"ms_scene_create": t_create, "ms_bvh_build": st.ms_bvh_build, "ms_bvh_device": st.ms_bvh_device, "ms_upload": st.ms_upload,
           "device_meshes": st.bvh_device_meshes, "ploc_iterations": st.bvh_ploc_iterations, "nodes4": st.bvh_nodes, "nodes2": st.bvh2_nodes}
    best = 1e9
    for _ in range(frames):
        img = sc.render(cam, rt.default_opts(integrator=integ))
        best = min(best, sc.stats().ms_total)
    rec["ms_frame"] = best
    rec["mean_radiance"] = float(img.mean())
    sc.render(cam, rt.default_opts(integrator=integ, collect_counters=3))
    st = sc.stats()
    nq = max(1, st.rays_primary + st.rays_secondary)
    rec["closest_node_visits_per_ray"] = st.node_visits / nq
    rec["closest_tri_tests_per_ray"] = st.prim_tests / nq
    rec["shadow_node_visits_per_ray"] = st.node_visits_shadow / max(1, st.rays_shadow)
    rec["shadow_tri_tests_per_ray"] = st.prim_tests_shadow / max(1, st.rays_shadow)
    out["builders"][name] = rec
    sc.close()
print(json.dumps(out, indent=1))
