#!/usr/bin/env python
"""Renders ONE rank's share of the tile-interleaved frame (tile_rank 0 of N) on one GPU, without NCCL: the per-rank device
time of an N-GPU run, used to split the N=8 frame time into render / frame reduce / fixed costs.
usage: python profiles/tile_share.py [config] [N] [frames]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rt = ge.load_package()
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frames = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 6
xml = bench.make_scene(rt, cfg, bench.WORKLOADS[cfg][0], "full")
hs = rt.HostScene(xml)
sc = rt.Scene(hs.desc, 0)
cam = hs.camera(0)
integ = rt.RT_INTEGRATOR_PATH if cfg == "c5" else rt.RT_INTEGRATOR_REFERENCE
if "--ranks" in sys.argv:                                   # per-rank device time: the load balance of the interleaved partition
    per = []
    for r in range(n):
        o = rt.default_opts(integrator=integ, tile_rank=r, tile_count=n)
        ms = min((sc.render(cam, o), sc.stats().ms_total)[1] for _ in range(frames))
        st = sc.stats()
        per.append(ms)
        print(f"rank {r}: {ms:.3f} ms, rays {st.rays_primary + st.rays_secondary + st.rays_shadow}")
    print(f"max/mean = {max(per) / (sum(per) / n):.4f}, mean {sum(per) / n:.3f} ms, max {max(per):.3f} ms")
    sys.exit(0)
for timing in (0, 1):
    o = rt.default_opts(integrator=integ, tile_rank=0, tile_count=n, collect_counters=timing)
    best = 1e9
    for _ in range(frames):
        t0 = time.perf_counter()
        sc.render(cam, o)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    st = sc.stats()
    print(f"{cfg} share 1/{n} timing={timing}: device {st.ms_total:.3f} ms, host call best {best:.3f} ms, waves {st.waves}, "
          f"launches {st.launches}, raygen {st.ms_raygen:.2f} extend {st.ms_extend:.2f} shade {st.ms_shade:.2f} "
          f"shadow {st.ms_shadow:.2f} resolve {st.ms_resolve:.2f}")
