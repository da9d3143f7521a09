#!/usr/bin/env python
"""bench.py - Mrays/s and ms/frame of the rendering hot path on N B200s, with roofline + CPU baseline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun, one rank per
GPU) prints ONE JSON line on rank 0.  A "step" is one frame of the workload: generate -> extend -> shade -> shadow
-> resolve for every pixel sample of the camera, image tiles interleaved over the N ranks, framebuffer reduced to
rank 0 over NCCL at frame end.  The workload at every N is BASELINE.json's headline configuration C4 (~870k-triangle
mesh x 16 instances, 1920x1080, 64 spp), so N>1 is STRONG scaling of one frame.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/ref_raytracer = the UNMODIFIED reference
compiled against stand-in headers; else the oracle port) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

# ALGORITHMIC bytes per ray, SURVEY.md §8d:  B_ray = 48 + 32*(ceil(log2 N_top) + ceil(log2(N_tri/4))) + 48*4 + 64*[instanced]
B_RAY = {"c1": 224, "c2": 400, "c3": 784, "c4": 1040, "c5": 416}
# ALGORITHMIC lane-ops per ray, SURVEY.md §8d ("ALGORITHMIC flops"): ~40 lane-ops per inner-node visit of the ideal binary descent
# (two child slabs) and ~50 per triangle test: OPS_ray = 40*(ceil(log2 N_top) + ceil(log2(N_tri/4))) + 50*4  (C1: one primitive).
# The traversal kernels are bound by the SM issue rate (instructions x active lanes), not by HBM: this is the roofline that binds.
OPS_RAY = {"c1": 40 * 4 + 50, "c2": 40 * 5 + 200, "c3": 40 * 17 + 200, "c4": 40 * 23 + 200, "c5": 40 * 5 + 200}
SM_COUNT, LANES_PER_SM = 148, 128            # B200: 148 SMs x 4 schedulers x 32 lanes, one warp instruction per scheduler and clock

WORKLOADS = {
    # name: (generator kwargs for the full configuration, kwargs for the bounded CPU sample).  The CPU sample keeps the full
    # RESOLUTION (same pixel footprint, same ray distribution over the scene) and cuts only the samples per pixel.
    "c4": (dict(res=(1920, 1080), spp=64, n=660), dict(res=(1920, 1080), spp=2, n=660)),
    "c3": (dict(res=(1920, 1080), spp=16, n=187), dict(res=(1920, 1080), spp=2, n=187)),
    "c2": (dict(res=1024, spp=100), dict(res=1024, spp=2)),
    "c1": (dict(res=800), dict(res=800)),
    "c5": (dict(res=(3840, 2160), spp=1024), dict(res=(480, 270), spp=4)),   # the reference has no path tracer: direct lighting only, not like-for-like
}
SMOKE_WORKLOADS = {"c4": dict(res=(480, 270), spp=4, n=660), "c3": dict(res=(480, 270), spp=4, n=187), "c2": dict(res=128, spp=16), "c1": dict(res=200),
                   "c5": dict(res=(240, 135), spp=4)}


def load_scenes():
    """The synthetic scene generators (raytracer-795_b200/scenes.py, numpy only) WITHOUT importing the product package: the
    reference arm must not load librt795.so or anything around it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("rt795_scenes", os.path.join(ROOT, "raytracer-795_b200", "scenes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def scene_cache_dir() -> str:
    d = os.environ.get("RT795_SCENE_CACHE", os.path.join(tempfile.gettempdir(), "rt795_scenes"))
    os.makedirs(d, exist_ok=True)
    return d


def make_scene(rt, config: str, kw: dict, tag: str) -> str:
    """`rt` is the product package or the bare scenes module (load_scenes())."""
    scenes = getattr(rt, "scenes", rt)
    d = os.path.join(scene_cache_dir(), f"{config}_{tag}")
    xml = os.path.join(d, f"{config}.xml")
    if not os.path.exists(xml):
        scenes.CONFIGS[config]["gen"](d, **kw)
    return xml


class ClockSampler:
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_kernel(config: str, kernel: str):
    """issue-slot utilisation / active lanes per instruction / pipe utilisation of `kernel` from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(config, {}).get(f"k_{kernel}_ncu")
    return None


def ncu_traffic(config: str, kernel: str):
    """dram bytes per launch of `kernel` ("extend" / "shadow") from the committed ncu --set full summary, if one exists for this workload."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(config, {}).get(f"k_{kernel}_dram_bytes_per_launch")
    return None


# ------------------------------------------------------------------------------------------------ reference arm
def ref_bench_binary() -> str | None:
    p = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
    return p if os.path.exists(p) else None


def run_reference_frames(xml: str, frames: int) -> tuple[list[float], list[int], str, int]:
    """`frames` renders of camera 0 by the reference's own CPU implementation on the host cores.  Returns (render-only seconds
    per frame, closest-hit queries per frame, kind, cores).

    kind "reference": oracle/_ref/ref_bench = the UNMODIFIED reference sources (compiled by oracle/Makefile against stand-in
    headers) driven by oracle/harness/bench_main.cpp: scene parsed and BVHs built once, then each frame is the reference's own
    8 hard-coded std::threads on Scene::ThreadedRendering (src/Scene.cpp:269-292, 340-356).  Rays are counted INSIDE that binary
    by a link-time wrapper around BVHMethods::FindIntersection (oracle/harness/count_wrap.cpp).  No product code is loaded.
    kind "port": oracle/_ref absent -> the oracle port (oracle/restate.cpp) with all host threads."""
    exe = ref_bench_binary()
    if exe:
        with tempfile.TemporaryDirectory() as td:
            r = subprocess.run([exe, xml, str(frames)], cwd=td, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True)
        recs = re.findall(r"RT795_FRAME \d+ seconds ([0-9.eE+-]+) queries (\d+)", r.stdout)
        assert len(recs) == frames, r.stdout[-2000:]
        return [float(a) for a, _ in recs], [int(b) for _, b in recs], "reference", 8
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as orc
    rt = ge.load_package()
    hs = rt.HostScene(xml)
    cores = os.cpu_count() or 1
    lib = orc.load_oracle()
    secs, qs = [], []
    for _ in range(frames):
        lib.oracle_ray_queries(1)
        t0 = time.perf_counter()
        orc.oracle_render(hs.desc, hs.camera(0), threads=cores)
        secs.append(time.perf_counter() - t0)
        qs.append(int(lib.oracle_ray_queries(1)))
    hs.close()
    return secs, qs, "port", cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scenes = load_scenes()
    full_kw, sample_kw = WORKLOADS[args.config]
    if args.small:
        sample_kw = SMOKE_WORKLOADS[args.config]
    tag = "refsample_small" if args.small else "refsample"
    if not args.small and args.steps + args.warmup > 12 and sample_kw.get("spp", 1) > 1:
        # a C4 frame at 1920x1080 costs the reference ~7 s per sample per pixel on the box's cores: with the driver's 20 + 5 frames the
        # per-step sample drops to 1 spp so that the whole run stays within a few minutes (same scene, same resolution, same ray mix)
        sample_kw = dict(sample_kw, spp=1)
        tag = "refsample_spp1"
    xml = make_scene(scenes, args.config, sample_kw, tag)
    secs, qs, kind, cores = run_reference_frames(xml, args.warmup + args.steps)
    secs, qs = secs[args.warmup:], qs[args.warmup:]
    ms = 1e3 * float(np.mean(secs))
    rays = float(np.mean(qs))
    val = rays / (ms * 1e-3) / 1e6
    label = scenes.CONFIGS[args.config]["label"]
    sample = (f"{args.config} scene, one frame per step at {sample_kw} (full workload: {full_kw}); rays = BVHMethods::FindIntersection calls counted "
              f"inside the {'unmodified reference binary (link-time wrapper)' if kind == 'reference' else 'oracle port'}; parse + BVH build outside the timed region")
    res = sample_kw["res"] if isinstance(sample_kw["res"], (tuple, list)) else (sample_kw["res"], sample_kw["res"])
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "resolution": [int(res[0]), int(res[1])], "spp": int(sample_kw.get("spp", 1)), "rays_per_frame": rays,
                       "integrator": "reference", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class _QuietStdout:
    """Keeps the process's stdout to the ONE JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints its
    version banner with printf, below Python), and the line is written to the saved descriptor at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self.fd, (text + "\n").encode())


def run_ours(args, rt):
    import torch
    import torch.distributed as dist

    out = _QuietStdout()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the rendering core has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                 # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    full_kw, sample_kw = WORKLOADS[args.config]
    if args.small:
        full_kw = sample_kw = SMOKE_WORKLOADS[args.config]
    tag = "small" if args.small else "full"
    if args.spp > 0 and "spp" in full_kw:
        full_kw = dict(full_kw, spp=args.spp)
        tag += f"_spp{args.spp}"
    if rank == 0:
        xml = make_scene(rt, args.config, full_kw, tag)
    barrier()
    xml = make_scene(rt, args.config, full_kw, tag)
    hs = rt.HostScene(xml)
    cam = hs.camera(0)
    # library default (AUTO): big meshes get the host builder's binned-SAH tree, built on the GPU (profiles/bench/r01_bvh_builders_c4.json)
    builder = {"host": rt.RT_BVH_HOST_SAH, "ploc": rt.RT_BVH_DEVICE_PLOC, "sah": rt.RT_BVH_DEVICE_SAH, "auto": rt.RT_BVH_AUTO}[args.bvh]
    # the first rt_scene_create of a process pays module load + first-use allocations of the device builder (cub temporaries): build
    # three times, report every build and keep the last scene
    build_ms = []
    sc = None
    for _ in range(3 if not args.small else 1):
        if sc is not None:
            sc.close()
        sc = rt.Scene(hs.desc, local, bvh_builder=builder)
        build_ms.append(float(sc.stats().ms_bvh_build))
    integrator = rt.RT_INTEGRATOR_PATH if args.config == "c5" else rt.RT_INTEGRATOR_REFERENCE
    npix = cam.nx * cam.ny
    # N>1: every rank's resolve kernel stores its tiles into rank 0's framebuffer over NVLink peer memory (multigpu.FrameGather;
    # RT795_GATHER=nccl selects the SUM-reduce fallback)
    gather = rt.multigpu.FrameGather(npix * 3, dev)
    opts = rt.default_opts(integrator=integrator, tile_rank=rank, tile_count=world, resolve_owned_only=gather.resolve_owned_only)
    frame = gather.frame
    host_frame = torch.empty(npix * 3, dtype=torch.float32).pin_memory()
    cam_host = torch.frombuffer(bytearray(bytes(cam)) + bytearray(bytes(opts)), dtype=torch.uint8).pin_memory()
    cam_dev = torch.empty_like(cam_host, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()

    mid_events = []

    def render_step(o, mid=None, release=True):
        sc.render_device(cam, o, gather.out_ptr, stream.cuda_stream)
        if mid is not None:
            mid.record(stream)
        gather.finish()                                       # p2p: device barrier after the peer stores; nccl: SUM reduce to rank 0
        if release:
            gather.release()                                  # p2p: nobody starts the next frame's stores before rank 0 is done with this one

    host_np = host_frame.numpy().reshape(cam.ny, cam.nx, 3)
    pageable_np = np.empty((cam.ny, cam.nx, 3), dtype=np.float32)   # what the reference hands over: new float[] (src/Image.cpp:8)
    pageable_np.fill(0.0)

    def e2e_pageable_step(o):
        sc.render(cam, o, out=pageable_np)

    def e2e_step(o):
        if world == 1:
            # the reference-facing call with HOST buffers: camera + options in host memory in, framebuffer in (pinned) host memory out;
            # rt_render enqueues on the library's stream and returns when the frame has landed in `host_np`
            sc.render(cam, o, out=host_np)
            return
        cam_dev.copy_(cam_host, non_blocking=True)            # this step's inputs (camera + options) from pinned host memory
        render_step(o, release=False)
        if rank == 0:
            host_frame.copy_(frame, non_blocking=True)        # the step's result: the gathered framebuffer
        gather.release()                                      # after rank 0's read of the frame is on the stream
        torch.cuda.synchronize()

    launches = 0
    traced = [0, 0, 0]

    def timed(step_fn, steps, wall=False):
        nonlocal launches, traced
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)                                    # L2 flush between timed iterations
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            if step_fn is render_step:
                em = torch.cuda.Event(enable_timing=True)
                step_fn(opts, em)
                mid_events.append((e0, em, e1))
            else:
                step_fn(opts)
            e1.record(stream)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            barrier()
            # device time from CUDA events on the launching stream; the host-buffer (e2e) call runs on the library's own stream and
            # is synchronous, so it is timed by the host clock around the call
            total += (t1 - t0) * 1e3 if wall else e0.elapsed_time(e1)
            st = sc.stats()
            launches += int(st.launches)
            traced = [int(st.rays_primary), int(st.rays_secondary), int(st.rays_shadow)]
        t = torch.tensor([total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                       # sampled over warm-up + timed region (all under load)
    for _ in range(max(args.warmup, 3)):
        render_step(opts)
    barrier()
    launches = 0
    ms_total = timed(render_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    timed_launches = launches
    # split of the device-timed step: this rank's render (rt_render_device, K8 stores included) vs what completes the gather
    split = torch.tensor([sum(a.elapsed_time(m) for a, m, _ in mid_events), sum(m.elapsed_time(b) for _, m, b in mid_events)],
                         dtype=torch.float64, device=dev) / max(1, len(mid_events))
    if world > 1:
        dist.all_reduce(split, op=dist.ReduceOp.MAX)
    ms_render_max, ms_gather_max = [float(x) for x in split.tolist()]
    gather_check = None
    if world > 1:
        # the gathered frame against rank 0 rendering the whole frame alone (float atomics reorder sums: compare to rounding)
        gathered = frame.clone() if rank == 0 else None
        barrier()
        if rank == 0:
            alone = torch.empty_like(frame)
            sc.render_device(cam, rt.default_opts(integrator=integrator), alone.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            denom = alone.abs().clamp_min(1e-3)
            gather_check = {"max_rel_diff_vs_single_gpu_frame": float(((gathered - alone).abs() / denom).max().item()),
                            "mean_frame": float(gathered.mean().item()), "mean_single": float(alone.mean().item())}
        barrier()
    e2e_step(opts)                                            # untimed: first-use allocations of the host-buffer path
    barrier()
    ms_e2e = timed(e2e_step, args.steps, wall=True)
    ms_e2e_pageable = None
    if world == 1:
        e2e_pageable_step(opts)
        ms_e2e_pageable = timed(e2e_pageable_step, args.steps, wall=True) / args.steps

    rays_local = torch.tensor(traced, dtype=torch.float64, device=dev)
    lt = torch.tensor([timed_launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rays_local)
        dist.all_reduce(lt)
    rays = [int(x) for x in rays_local.tolist()]
    rays_frame = sum(rays)
    ms_step = ms_total / args.steps
    value = rays_frame / (ms_step * 1e-3) / 1e6
    e2e_value = rays_frame / (ms_e2e / args.steps * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (k_extend): one extra frame with per-stage CUDA events (serialises stages)
    so = rt.default_opts(integrator=integrator, tile_rank=rank, tile_count=world, collect_counters=1,
                         resolve_owned_only=gather.resolve_owned_only)
    render_step(so)
    torch.cuda.synchronize()
    st = sc.stats()
    ext_rays = int(st.rays_primary + st.rays_secondary)
    peak, peak_src = measured_peaks()
    b_ray = B_RAY[args.config]
    stage = {"ms_extend": st.ms_extend, "ms_shade": st.ms_shade, "ms_shadow": st.ms_shadow, "ms_raygen": st.ms_raygen, "ms_resolve": st.ms_resolve,
             "launches_extend": int(st.launches_extend), "launches_shadow": int(st.launches_shadow)}

    # one more frame with the per-ray counters build: measured node visits / triangle tests and SIMT lane counts of the two traversal kernels
    co = rt.default_opts(integrator=integrator, tile_rank=rank, tile_count=world, collect_counters=2, resolve_owned_only=gather.resolve_owned_only)
    render_step(co)
    torch.cuda.synchronize()
    cs = sc.stats()
    ops_ray = OPS_RAY[args.config]
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = SM_COUNT * LANES_PER_SM * sm_mhz * 1e6 / 1e9             # G lane-ops/s at the clock observed under load
    issue_src = f"{SM_COUNT} SMs x {LANES_PER_SM} lanes x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)"

    def kernel_roofline(name, label, rays, ms, launches, visits, prims, diag):
        hbm = rays * b_ray / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        alg = rays * ops_ray / (ms * 1e-3) / 1e9 if ms > 0 else 0.0      # G lane-ops/s, algorithmic
        # measured work of the 4-wide tree: a 4-wide node visit tests four child slabs (2 x the binary visit's 40), a triangle test 50
        work = (visits * 80 + prims * 50) / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        extra = ncu_kernel(args.config, name)
        return {"kernel": label, "issue": {"achieved": alg, "peak": issue_peak, "unit": "Glaneops/s", "frac": alg / issue_peak, "measured_work_glaneops": work,
                                           "measured_work_frac": work / issue_peak, "node_visits_per_ray": visits / max(1, rays), "prim_tests_per_ray": prims / max(1, rays),
                                           "leaf_phase_lanes": (diag[3] / diag[2]) if diag[2] else None, "ncu": extra},
                "hbm": {"achieved": hbm, "peak": peak, "unit": "GB/s", "frac": hbm / peak, "traffic": ncu_traffic(args.config, name)},
                "rays_per_launch": rays / max(1, launches), "avg_launch_ms": ms / max(1, launches), "stage_ms": ms}

    k2 = kernel_roofline("extend", "k_raygen_trace / k_trace_persistent<closest-hit> (K1+K2 extend)", ext_rays, st.ms_extend, int(st.launches_extend),
                         int(cs.node_visits), int(cs.prim_tests), list(cs.diag)[0:4])
    k4 = kernel_roofline("shadow", "k_trace_persistent<any-hit> (K4 occlusion)", int(st.rays_shadow), st.ms_shadow, int(st.launches_shadow),
                         int(cs.node_visits_shadow), int(cs.prim_tests_shadow), list(cs.diag)[4:8])
    dom, other = (k4, k2) if st.ms_shadow >= st.ms_extend else (k2, k4)      # the dominant kernel = the traversal kernel with the larger stage time
    roofline = {"bound": "issue", "kernel": dom["kernel"], "achieved": dom["issue"]["achieved"], "peak": issue_peak, "unit": "Glaneops/s", "frac": dom["issue"]["frac"],
                "traffic": dom["hbm"]["traffic"], "peak_source": issue_src, "algorithmic_laneops_per_ray": ops_ray,
                "measured_work_frac": dom["issue"]["measured_work_frac"], "node_visits_per_ray": dom["issue"]["node_visits_per_ray"],
                "prim_tests_per_ray": dom["issue"]["prim_tests_per_ray"], "ncu": dom["issue"]["ncu"],
                "rays_per_launch": dom["rays_per_launch"], "avg_launch_ms": dom["avg_launch_ms"], "stage_ms": stage,
                "secondary": {"bound": "hbm", "achieved": dom["hbm"]["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["hbm"]["frac"], "peak_source": peak_src,
                              "algorithmic_bytes_per_ray": b_ray, "traffic": dom["hbm"]["traffic"],
                              "note": "the scene (BVH + triangles) is L2-resident: DRAM carries the ray queues only (ncu: DRAM throughput a few % of peak), so this line does not bind"},
                "other_traversal_kernel": other,
                "note": "software BVH traversal = divergent FP32 SIMT: the binding roofline is the SM issue rate (lane-ops/s); achieved = rays/s x algorithmic lane-ops per ray (SURVEY 8d)"}
    sstats = sc.stats()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sxml = make_scene(rt, args.config, sample_kw, "refsample_small" if args.small else "refsample")
        (secs,), (rrays,), kind, cores = run_reference_frames(sxml, 1)
        cpu_baseline = {"value": rrays / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind, "seconds": secs,
                        "sample": f"{args.config} scene at {sample_kw}: {rrays} closest-hit queries (BVHMethods::FindIntersection calls counted inside the "
                                  f"{'unmodified reference binary' if kind == 'reference' else 'oracle port'}) in {secs:.2f} s render-only"}

    if rank == 0:
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": rt.scenes.CONFIGS[args.config]["label"] + (" [--small sample]" if args.small else "") + (f" [--spp {args.spp}]" if args.spp else ""),
                           "resolution": [cam.nx, cam.ny], "spp": cam.num_samples, "rays_per_frame": {"primary": rays[0], "secondary": rays[1], "shadow": rays[2]},
                           "integrator": "path" if integrator else "reference", "parallelism": f"tiles16x16-interleaved-x{world}", "gather": {"mode": gather.mode, "fallback_reason": gather.why_not_p2p, "ms_render_max": ms_render_max, "ms_gather_max": ms_gather_max, "check": gather_check}, "l2": "flushed between steps (256 MB write)",
                           "scene_bytes": {"bvh": int(sstats.bvh_bytes), "triangles": int(sstats.tri_bytes)}, "bvh_build_ms": {"first_in_process": build_ms[0], "min": min(build_ms), "median": float(np.median(build_ms)), "all": build_ms},
                           "bvh_builder": args.bvh, "bvh_device_meshes": int(sstats.bvh_device_meshes)},
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(cam_host.numel()), "d2h_bytes_per_step": int(npix * 12),
                        "api": "rt_render (host camera/options in, host framebuffer out)" if world == 1 else ("rt_render_device (K8 stores owned tiles into rank 0's frame over NVLink peer memory) + device barrier" if gather.mode == "p2p" else "rt_render_device + NCCL reduce") + " + D2H on rank 0",
                        "timing": "host clock around the synchronous call, max over ranks",
                        "host_buffer": "page-locked (DMA straight into it)",
                        "pageable_host_buffer": None if ms_e2e_pageable is None else {"ms_per_step": ms_e2e_pageable, "value": rays_frame / (ms_e2e_pageable * 1e-3) / 1e6,
                            "note": "same call with an ordinary malloc'ed frame, as the reference's Image::_data is: staged through the library's pinned buffer + host memcpy"}},
                "gpu_launches": int(lt.item()), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline}
        out.emit(json.dumps(line))
    sc.close()
    hs.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--small", action="store_true", help="use the bounded sample configuration (smoke runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bvh", default="auto", choices=["host", "ploc", "sah", "auto"], help="per-mesh BVH builder (RtSceneDesc.bvh_builder)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel of the workload (exploration; the headline run uses the configuration's own)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)                                   # never imports the product package
    else:
        run_ours(args, ge.load_package())


if __name__ == "__main__":
    main()
