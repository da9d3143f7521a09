// CLI with the reference's contract (src/main.cpp:7-14): `raytracer scene.xml` renders every camera of the scene
// and writes each camera's <ImageName> (P3 PPM for "*.png*" names, half-float EXR otherwise, src/Image.cpp:26-34).
// Extra, optional knobs after the scene path:  --f32 (also dump <ImageName>.f32 raw floats)  --device N
// --cache (keep / reuse <scene.xml>.rt795blob, the prepared scene, while its sources are unchanged)
// --integrator path|reference  --bvh auto|host|device (per-mesh BVH builder)  --seed S  --tonemap KEY BURN SAT GAMMA (writes <ImageName>.tm.ppm)
#include "../../include/rt795.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char* argv[]) {
    if (argc < 2) { fprintf(stderr, "usage: %s scene.xml [--f32] [--cache] [--device N] [--integrator path|reference] [--bvh auto|host|device] [--seed S] [--tonemap key burn sat gamma]\n", argv[0]); return 2; }
    bool f32 = false, tonemap = false, cache = false;
    int device = 0, bvh = RT_BVH_AUTO;
    float tm[4] = {0.18f, 1.0f, 1.0f, 2.2f};
    RtRenderOpts opts;
    memset(&opts, 0, sizeof opts);
    opts.seed = 795; opts.pt_nee = 1; opts.pt_russian_roulette = 1; opts.pt_importance = 1;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "--f32")) f32 = true;
        else if (!strcmp(argv[i], "--cache")) cache = true;
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) opts.seed = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--integrator") && i + 1 < argc) opts.integrator = !strcmp(argv[++i], "path") ? RT_INTEGRATOR_PATH : RT_INTEGRATOR_REFERENCE;
        else if (!strcmp(argv[i], "--bvh") && i + 1 < argc) { const char* b = argv[++i]; bvh = !strcmp(b, "host") ? RT_BVH_HOST_SAH : !strcmp(b, "device") ? RT_BVH_DEVICE_PLOC : RT_BVH_AUTO; }
        else if (!strcmp(argv[i], "--tonemap") && i + 4 < argc) { tonemap = true; for (int k = 0; k < 4; k++) tm[k] = (float)atof(argv[++i]); }
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }
    RtHostScene* hs = nullptr;
    int hit = 0;
    if ((cache ? rt_host_load_scene_cached(argv[1], nullptr, &hs, &hit) : rt_host_load_scene(argv[1], &hs)) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    RtScene* sc = nullptr;
    RtSceneDesc desc = *rt_host_scene_desc(hs);
    desc.bvh_builder = bvh;
    if (rt_scene_create(&desc, device, &sc) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    printf("BVH construction complete.\n");
    for (int c = 0; c < rt_host_num_cameras(hs); c++) {
        const RtCameraDesc* cam = rt_host_camera(hs, c);
        std::vector<float> img((size_t)cam->nx * cam->ny * 3);
        if (rt_render(sc, cam, &opts, img.data()) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
        RtStats st;
        rt_get_stats(sc, &st);
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
        if (rt_host_save_image(name, img.data(), cam->nx, cam->ny) != RT_OK) { fprintf(stderr, "%s\n", rt_last_error()); return 1; }
    }
    rt_scene_destroy(sc);
    rt_host_free_scene(hs);
    return 0;
}
