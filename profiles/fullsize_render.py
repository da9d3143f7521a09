#!/usr/bin/env python
"""Step 1 of the full-size parity check (runs on the GPU box): renders the full-size configurations through the C ABI
and writes the float framebuffers to gpurun_out/ as float16-free raw .f32 (int32 nx, ny + floats)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rt = ge.load_package()
lib = rt.load_core()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for cfg in sys.argv[1:] or ["c2", "c3", "c4"]:
    xml = bench.make_scene(rt, cfg, bench.WORKLOADS[cfg][0], "full")
    hs = rt.HostScene(xml)
    sc = rt.Scene(hs.desc, 0)
    cam = hs.camera(0)
    img = sc.render(cam)
    st = sc.stats()
    out = os.path.join(ROOT, "gpurun_out", f"full_{cfg}.f32")
    lib.rt_host_save_f32(out.encode(), img.ctypes.data, cam.nx, cam.ny)
    print(cfg, cam.nx, cam.ny, cam.num_samples, f"{st.ms_total:.2f} ms", float(img.mean()))
