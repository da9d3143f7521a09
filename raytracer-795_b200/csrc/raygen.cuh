// K1 device code: primary / sample ray of one work item.
//   Camera::getPrimaryRay (src/Camera.cpp:63-72), PixelLBCorner (:84-92), getSampleRay (:94-113), AddDepthOfField (:119-139),
//   enumerated like the loop of Scene::MultiSample (src/Scene.cpp:386-411).
// Shared by the stand-alone raygen kernel (kernels.cuh k_raygen) and by the fused raygen + closest-hit kernel
// (traverse_pt.cuh k_raygen_trace), where a lane generates its next primary ray the moment its previous one has finished.
#pragma once
#include "dscene.h"
#include "shade.cuh"

namespace rt795 {

struct RaygenArgs {
    DCamera cam;
    DTiling tl;
    uint64_t seed;
    int32_t max_depth;
    uint32_t base_rest, base_sample;   // first work item of the wave = base_rest * spp + base_sample (32-bit divisions only in the kernel)
    int32_t center_only;               // rt_trace_primary: pixel centres, one ray per pixel
};

// Work item w = work_base + i enumerates (pixel slot, sample) with the sample fastest; pixel slots run over this rank's tiles,
// inside a tile over 8x4 sub-blocks (one warp = one compact pixel block at 1 spp, 32 samples of one pixel at >= 32 spp).
// Returns false for slots outside the image (tiles cut by the right / bottom edge).
RT_D bool gen_camera_ray(const RaygenArgs& a, int i, float3& o, float3& d, float& time, uint32_t& pixel, uint32_t& sample) {
    const DCamera& cam = a.cam;
    const DTiling& tl = a.tl;
    int spp = a.center_only ? 1 : cam.spp;
    unsigned sw = a.base_sample + (unsigned)i, sq = sw / (unsigned)spp;
    sample = sw - sq * (unsigned)spp;
    unsigned rest = a.base_rest + sq;
    unsigned tile_px = (unsigned)(tl.tw * tl.th);
    int lt = (int)(rest / tile_px);
    int pit = (int)(rest - (unsigned)lt * tile_px);
    int k = lt * tl.count + tl.rank;
    int tx = k % tl.tiles_x, ty = k / tl.tiles_x;
    int sb = pit >> 5, ln = pit & 31, sbx = tl.tw >> 3;
    int x = tx * tl.tw + (sb % sbx) * 8 + (ln & 7), y = ty * tl.th + (sb / sbx) * 4 + (ln >> 3);
    if (x >= cam.nx || y >= cam.ny) return false;
    pixel = (uint32_t)(y * cam.nx + x);
    float3 pos = v3(cam.pos), gaze = v3(cam.gaze), up = v3(cam.up), right = v3(cam.right);
    if (spp <= 1) {                                          // getPrimaryRay
        float u = cam.l + (cam.r - cam.l) * (x + 0.5f) * cam.invnx;
        float v = cam.t - (cam.t - cam.b) * (y + 0.5f) * cam.invny;
        float3 m = pos + gaze * cam.dist;
        m = m + u * right;
        m = m + v * up;
        o = pos; d = (m - pos) / length(m - pos); time = 0.f;
    } else {                                                 // PixelLBCorner + getSampleRay (+ AddDepthOfField)
        float u = cam.l + x * cam.pw;
        float v = cam.t - (y + 1) * cam.ph;
        float3 m = pos + gaze * cam.dist;
        m = m + u * right;
        m = m + v * up;
        int si = (int)sample % cam.n, sj = (int)sample / cam.n;
        float4 r = philox_uniform4(a.seed, pixel, sample, 0u, DIM_CAMERA);
        m = m + (si + r.x) * cam.sw * right;
        m = m + (sj + r.y) * cam.sh * up;
        o = pos; d = (m - pos) / length(m - pos);
        if (cam.dof) {
            float3 q = pos;
            q = q + cam.aperture * (r.z - 0.5f) * right;
            q = q + cam.aperture * (r.w - 0.5f) * up;
            float3 dir = normalize0(m - pos);
            float t_fd = cam.focus / dot(dir, gaze);
            float3 p = o + d * t_fd;
            o = q; d = normalize0(p - q); time = 0.f;
        } else time = r.z;
    }
    return true;
}

RT_D int meta_pack(int depth, uint32_t flags, int beer_mat) { return (depth & 0xff) | (int)((flags & 0xffu) << 8) | (beer_mat << 16); }

}  // namespace rt795
