// CLI with the reference's contract (src/main.cpp:7-14): `raytracer scene.xml` renders every camera of the scene
// and writes each camera's <ImageName> (P3 PPM for "*.png*" names, half-float EXR otherwise, src/Image.cpp:26-34).
// Extra, optional knobs after the scene path:  --f32 (also dump <ImageName>.f32 raw floats)  --device N
// --gpus N (one frame split over devices N0..N0+N-1 of this process: rt_group_render, the in-process fan-out of src/Scene.cpp:340-356)
// A camera's <Renderer>/<RendererParams>/<Tonemap> elements (extensions) are honoured unless --integrator / --tonemap override them.
// --cache (keep / reuse <scene.xml>.rt795blob, the prepared scene, while its sources are unchanged)
// --integrator path|reference  --bvh auto|host|ploc|sah (per-mesh BVH builder)  --seed S  --tonemap KEY BURN SAT GAMMA (writes <ImageName>.tm.ppm)
#include "../../include/rt795.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char* argv[]) {
    if (argc < 2) { fprintf(stderr, "usage: %s scene.xml [--f32] [--cache] [--device N] [--gpus N] [--integrator path|reference] [--bvh auto|host|ploc|sah] [--seed S] [--tonemap key burn sat gamma]\n", argv[0]); return 2; }
    bool f32 = false, tonemap = false, cache = false, cli_integrator = false, cli_seed = false;
    int device = 0, gpus = 1, bvh = RT_BVH_AUTO;
    float tm[4] = {0.18f, 1.0f, 1.0f, 2.2f};
    RtRenderOpts opts;
    memset(&opts, 0, sizeof opts);
    opts.seed = 795; opts.pt_nee = 1; opts.pt_russian_roulette = 1; opts.pt_importance = 1;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "--f32")) f32 = true;
        else if (!strcmp(argv[i], "--cache")) cache = true;
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) { opts.seed = strtoull(argv[++i], nullptr, 10); cli_seed = true; }
        else if (!strcmp(argv[i], "--integrator") && i + 1 < argc) { opts.integrator = !strcmp(argv[++i], "path") ? RT_INTEGRATOR_PATH : RT_INTEGRATOR_REFERENCE; cli_integrator = true; }
        else if (!strcmp(argv[i], "--bvh") && i + 1 < argc) { const char* b = argv[++i]; bvh = !strcmp(b, "host") ? RT_BVH_HOST_SAH : !strcmp(b, "device") || !strcmp(b, "ploc") ? RT_BVH_DEVICE_PLOC : !strcmp(b, "sah") ? RT_BVH_DEVICE_SAH : RT_BVH_AUTO; }
        else if (!strcmp(argv[i], "--tonemap") && i + 4 < argc) { tonemap = true; for (int k = 0; k < 4; k++) tm[k] = (float)atof(argv[++i]); }
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    // One process renders on one GPU: hide the others before the CUDA runtime starts, so its initialisation touches a single
    // device (on an 8-GPU box the start-up of the runtime is the largest part of a one-frame run).
    if (gpus <= 1 && !getenv("CUDA_VISIBLE_DEVICES")) { setenv("CUDA_VISIBLE_DEVICES", std::to_string(device).c_str(), 1); device = 0; }
    auto t0 = now();
    RtHostScene* hs = nullptr;
    int hit = 0;
    if ((cache ? rt_host_load_scene_cached(argv[1], nullptr, &hs, &hit) : rt_host_load_scene(argv[1], &hs)) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    auto t1 = now();
    int ndev = rt_device_count();                                        // first CUDA call: runtime + driver initialisation
    auto t1b = now();
    if (ndev <= 0) { fprintf(stderr, "no CUDA device: %s\n", rt_last_error()); return 1; }
    RtScene* sc = nullptr;
    RtSceneGroup* group = nullptr;
    RtSceneDesc desc = *rt_host_scene_desc(hs);
    desc.bvh_builder = bvh;
    if (gpus > 1) {
        if (device + gpus > ndev) { fprintf(stderr, "--gpus %d from device %d: only %d CUDA device(s)\n", gpus, device, ndev); return 1; }
        std::vector<int> devs(gpus);
        for (int i = 0; i < gpus; i++) devs[i] = device + i;
        if (rt_group_create(&desc, gpus, devs.data(), &group) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
        sc = rt_group_scene(group, 0);                                    // the root replica: statistics, tone map
    } else if (rt_scene_create(&desc, device, &sc) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    auto t2 = now();
    printf("BVH construction complete.\n");
    {
        RtStats st;
        rt_get_stats(sc, &st);
        fprintf(stderr, "scene: load %.1f ms%s, CUDA start-up %.1f ms, create %.1f ms (BVH build %.1f ms, %d mesh(es) on the GPU; upload %.1f ms)\n", ms(t0, t1),
                cache ? (hit ? " (blob)" : " (parsed, blob written)") : "", ms(t1, t1b), ms(t1b, t2), st.ms_bvh_build, st.bvh_device_meshes, st.ms_upload);
    }
    for (int c = 0; c < rt_host_num_cameras(hs); c++) {
        const RtCameraDesc* cam = rt_host_camera(hs, c);
        std::vector<float> img((size_t)cam->nx * cam->ny * 3);
        // the camera's own <Renderer> / <RendererParams> (extension elements) unless the command line says otherwise
        RtRenderOpts co = opts;
        if (!cli_integrator) {
            RtRenderOpts xo;
            if (rt_host_camera_render_opts(hs, c, &xo) == RT_OK) { uint64_t seed = cli_seed ? opts.seed : xo.seed; co = xo; co.seed = seed; }
        }
        auto r0 = now();
        if ((group ? rt_group_render(group, cam, &co, img.data()) : rt_render(sc, cam, &co, img.data())) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
        auto r1 = now();
        RtStats st;
        rt_get_stats(sc, &st);
        if (group) {                                                      // whole-frame totals over the replicas; time = the slowest device
            for (int i = 1; i < rt_group_num_devices(group); i++) {
                RtStats si;
                rt_get_stats(rt_group_scene(group, i), &si);
                st.rays_primary += si.rays_primary; st.rays_secondary += si.rays_secondary; st.rays_shadow += si.rays_shadow;
                if (si.ms_total > st.ms_total) st.ms_total = si.ms_total;
            }
        }
        double rays = (double)(st.rays_primary + st.rays_secondary + st.rays_shadow);
        fprintf(stderr, "camera %d: %.3f ms, %.1f Mrays/s (%llu primary, %llu secondary, %llu shadow)\n", c, st.ms_total, rays / st.ms_total * 1e-3,
                (unsigned long long)st.rays_primary, (unsigned long long)st.rays_secondary, (unsigned long long)st.rays_shadow);
        const char* name = rt_host_camera_image_name(hs, c);
        if (f32 && rt_host_save_f32((std::string(name) + ".f32").c_str(), img.data(), cam->nx, cam->ny) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
        if (tonemap) {
            std::vector<uint8_t> ldr((size_t)cam->nx * cam->ny * 3);
            if (rt_tonemap(sc, img.data(), cam->nx, cam->ny, tm[0], tm[1], tm[2], tm[3], ldr.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
            std::string p = std::string(name) + ".tm.ppm";
            FILE* f = fopen(p.c_str(), "wb");
            if (f) { fprintf(f, "P6\n%d %d\n255\n", cam->nx, cam->ny); fwrite(ldr.data(), 1, ldr.size(), f); fclose(f); }
        }
        RtTonemap xt;
        if (!tonemap && rt_host_camera_tonemap(hs, c, &xt) == RT_OK && xt.enabled) {
            // <Tonemap> in the camera: the tone-mapped 8-bit frame goes next to the HDR one, "x.exr" -> "x.png" (written with the
            // reference's ".png" semantics, i.e. a P3 PPM, src/Image.cpp:62-103)
            std::vector<uint8_t> ldr((size_t)cam->nx * cam->ny * 3);
            if (rt_tonemap(sc, img.data(), cam->nx, cam->ny, xt.key, xt.burn_percent, xt.saturation, xt.gamma, ldr.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
            std::string p = name;
            size_t dot = p.rfind(".exr");
            p = dot != std::string::npos ? p.substr(0, dot) + ".png" : p + ".tm.png";
            std::vector<float> lf(ldr.begin(), ldr.end());
            if (rt_host_save_image(p.c_str(), lf.data(), cam->nx, cam->ny) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
        }
        auto w0 = now();
        if (rt_host_save_image(name, img.data(), cam->nx, cam->ny) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
        fprintf(stderr, "camera %d: rt_render call %.1f ms (first call allocates the ray queues), image write %.1f ms\n", c, ms(r0, r1), ms(w0, now()));
    }
    if (group) rt_group_destroy(group); else rt_scene_destroy(sc);
    rt_host_free_scene(hs);
    return 0;
}
