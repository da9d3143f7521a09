#!/usr/bin/env python
"""Measured (non-contract) traffic per ray, SURVEY §8d: mean BVH nodes visited and primitives tested per ray, by ray type.
usage: python profiles/counters.py [config] [--small]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rt = ge.load_package()
cfg = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "c4"
full_kw, sample_kw = bench.WORKLOADS[cfg]
kw = sample_kw if "--small" in sys.argv else full_kw
xml = bench.make_scene(rt, cfg, kw, "small" if "--small" in sys.argv else "full")
hs = rt.HostScene(xml)
sc = rt.Scene(hs.desc, 0)
cam = hs.camera(0)
integ = rt.RT_INTEGRATOR_PATH if cfg == "c5" else rt.RT_INTEGRATOR_REFERENCE
sc.render(cam, rt.default_opts(integrator=integ))
sc.render(cam, rt.default_opts(integrator=integ, collect_counters=3))
st = sc.stats()
ext = st.rays_primary + st.rays_secondary
out = {"config": cfg, "kw": str(kw), "rays": {"primary": st.rays_primary, "secondary": st.rays_secondary, "shadow": st.rays_shadow},
       "closest_hit": {"nodes_per_ray": st.node_visits / max(1, ext), "prims_per_ray": st.prim_tests / max(1, ext),
                       "measured_bytes_per_ray": 64 * st.node_visits / max(1, ext) + 48 * st.prim_tests / max(1, ext) + 48},
       "occlusion": {"nodes_per_ray": st.node_visits_shadow / max(1, st.rays_shadow), "prims_per_ray": st.prim_tests_shadow / max(1, st.rays_shadow),
                     "measured_bytes_per_ray": 64 * st.node_visits_shadow / max(1, st.rays_shadow) + 48 * st.prim_tests_shadow / max(1, st.rays_shadow) + 48},
       "simt": {k: {"lanes_at_node_step": nv / max(1, st.diag[o]), "lanes_alive_at_node_step": st.diag[o + 1] / max(1, st.diag[o]),
                    "lanes_in_leaf_phase": st.diag[o + 3] / max(1, st.diag[o + 2]), "node_step_execs": st.diag[o], "leaf_phase_execs": st.diag[o + 2]}
                for k, o, nv in (("closest_hit", 0, st.node_visits), ("occlusion", 4, st.node_visits_shadow))},
       "stage_ms_with_counters": {"extend": st.ms_extend, "shade": st.ms_shade, "shadow": st.ms_shadow},
       "bvh": {"nodes": st.bvh_nodes, "bytes": st.bvh_bytes, "tri_bytes": st.tri_bytes, "build_ms": st.ms_bvh_build}}
print(json.dumps(out))
