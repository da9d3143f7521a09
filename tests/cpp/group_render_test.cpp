// Multi-GPU through the C ABI, no Python / torch involved: renders camera 0 of a scene on ONE device (rt_render) and on a device
// group (rt_group_render: one host thread + context per device, tiles stored into the root's frame over peer memory) and compares
// the two frames.   usage: group_render_test scene.xml [device ...]      (no devices listed: all of rt_device_count())
// Prints one line:  GROUP devices=<n> peer=<k of n> ms_single=<..> ms_group=<..> ms_group_pinned=<..> max_rel_diff=<..>   and exits 0 iff
// max_rel_diff <= 1e-5.  ms_*: host clock around one call, frame delivered into an ordinary (pageable) std::vector; *_pinned: into
// rt_host_alloc_pinned memory.
#include "../../include/rt795.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s scene.xml [device ...]\n", argv[0]); return 2; }
    std::vector<int> devs;
    for (int i = 2; i < argc; i++) devs.push_back(atoi(argv[i]));
    if (devs.empty()) for (int i = 0; i < rt_device_count(); i++) devs.push_back(i);
    if (devs.empty()) { fprintf(stderr, "no CUDA device\n"); return 1; }
    RtHostScene* hs = nullptr;
    if (rt_host_load_scene(argv[1], &hs) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    const RtCameraDesc* cam = rt_host_camera(hs, 0);
    RtRenderOpts opts;
    if (rt_host_camera_render_opts(hs, 0, &opts) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    size_t n = (size_t)cam->nx * cam->ny * 3;
    std::vector<float> single(n), multi(n, -1.0f);
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    RtScene* sc = nullptr;
    if (rt_scene_create(rt_host_scene_desc(hs), devs[0], &sc) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    if (rt_render(sc, cam, &opts, single.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }   // warm-up (queues)
    auto t0 = now();
    if (rt_render(sc, cam, &opts, single.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    double ms_single = ms(t0, now());
    rt_scene_destroy(sc);
    RtSceneGroup* g = nullptr;
    if (rt_group_create(rt_host_scene_desc(hs), (int)devs.size(), devs.data(), &g) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    if (rt_group_render(g, cam, &opts, multi.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    auto t1 = now();
    if (rt_group_render(g, cam, &opts, multi.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    double ms_group = ms(t1, now());
    double ms_group_pinned = -1;
    if (float* pin = (float*)rt_host_alloc_pinned(n * sizeof(float))) {
        rt_group_render(g, cam, &opts, pin);
        auto t2 = now();
        if (rt_group_render(g, cam, &opts, pin) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
        ms_group_pinned = ms(t2, now());
        rt_host_free_pinned(pin);
    }
    int peer = 0;
    for (int i = 0; i < rt_group_num_devices(g); i++) peer += rt_group_uses_peer_memory(g, i);
    // a second camera-independent check: tile ownership must not leak into the API (per-device options are rejected)
    RtRenderOpts bad = opts; bad.tile_count = 2;
    int rejected = rt_group_render(g, cam, &bad, multi.data()) == RT_ERR_INVALID;
    if (rt_group_render(g, cam, &opts, multi.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    double worst = 0;
    for (size_t i = 0; i < n; i++) {
        double d = std::fabs((double)single[i] - (double)multi[i]) / std::fmax(1.0, std::fabs((double)single[i]));
        if (!(d <= worst)) worst = d;            // NaN-propagating max
    }
    printf("GROUP devices=%d peer=%d ms_single=%.3f ms_group=%.3f ms_group_pinned=%.3f rejected_tile_opts=%d max_rel_diff=%.3e\n", (int)devs.size(), peer, ms_single, ms_group,
           ms_group_pinned, rejected, worst);
    rt_group_destroy(g);
    rt_host_free_scene(hs);
    return (worst <= 1e-5 && rejected) ? 0 : 1;
}
