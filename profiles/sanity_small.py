#!/usr/bin/env python
"""Renders every small scene variant once (used under `compute-sanitizer --tool memcheck`)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

rt = ge.load_package()
with tempfile.TemporaryDirectory() as td:
    jobs = [("c1", rt.scenes.scene_c1(td + "/c1", res=64), 0), ("c2", rt.scenes.scene_c2(td + "/c2", res=48, spp=4), 0),
            ("c3", rt.scenes.scene_c3(td + "/c3", res=(64, 36), spp=4, n=24), 0), ("c4", rt.scenes.scene_c4(td + "/c4", res=(64, 36), spp=4, n=16), 0),
            ("c5", rt.scenes.scene_c5(td + "/c5", res=(64, 36), spp=4), 1)]
    for v in ["lights", "brdf", "textures", "stochastic", "env"]:
        jobs.append((v, rt.scenes.scene_features(td + "/f_" + v, v, res=(64, 48), spp=4), 0))
    for name, xml, integ in jobs:
        hs = rt.HostScene(xml)
        sc = rt.Scene(hs.desc, 0)
        cam = hs.camera(0)
        img = sc.render(cam, rt.default_opts(integrator=integ, queue_capacity=4096, wave_rays=2048))
        hits = sc.trace_primary(cam)
        ldr = sc.tonemap(img)
        print(name, float(img.mean()), int(hits["full"].sum()), int(ldr.max()))
        sc.close(); hs.close()
print("done")
