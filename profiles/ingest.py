#!/usr/bin/env python
"""Scene ingest: XML + PLY + images parse (rt_host_load_scene) vs the prepared-scene blob (rt_host_load_scene_blob).  CPU only.
usage: python profiles/ingest.py [config]"""
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
xml = bench.make_scene(rt, cfg, bench.WORKLOADS[cfg][0], "full")
blob = xml + ".ingest_test.blob"


def best(f, n=3):
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); h = f(); t.append((time.perf_counter() - t0) * 1e3); h.close()
    return min(t)


hs = rt.HostScene(xml)
hs.save_blob(blob)
d = hs.desc.contents
out = {"config": cfg, "triangles": d.num_triangles, "vertices": d.num_vertices, "blob_bytes": os.path.getsize(blob),
       "ms_parse": best(lambda: rt.HostScene(xml)), "ms_blob": best(lambda: rt.HostScene(blob=blob))}
hs.close()
os.remove(blob)
print(json.dumps(out))
