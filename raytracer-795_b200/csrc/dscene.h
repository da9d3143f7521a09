// Device-side scene layout (HBM) shared by core.cu (host upload code) and the kernels.
// Everything the traversal touches is 16-byte aligned and read as float4 (ld.global.nc.v4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt795 {

constexpr int kStackSize = 64;          // traversal stack entries per thread (node, t_entry); core.cu checks the BVH against it
constexpr int kMaxLights = 1 << 16;
constexpr int kMaxLevels = 256;         // ray-tree levels per wave (recursion depth <= 250, core.cu)
constexpr int kSmemTlasBytes = 8192;    // TLAS nodes (128 B) + top-level transform records (64 B) the traversal kernels stage in shared memory;
                                        // a larger top level is walked in global memory

// Traversal record of a top-level entry: 64 bytes = the "64-B instance record" of SURVEY §8d.
struct DObjectXf {
    float4 inv0, inv1, inv2;            // rows of M^-1 (3x4): o' = inv*(o - blur*time, 1), d' = inv*(d, 0)  (src/Helper.cpp:110-133)
    float4 blur_root;                   // blur.xyz ; w = bits(int): BLAS root in nodes4, or ~sphere index for spheres
};
// Shading record of a top-level entry.
struct DObjectSh {
    float nrm[9];                       // upper 3x3 of (M^-1)^T  (src/Helper.cpp:100-108)
    int32_t mat_index;                  // 1-based
    int32_t tex[2];                     // 1-based, 0 = none
    int32_t tex_offset;
    int32_t smooth;                     // mesh smooth flag
    int32_t first_tri;                  // mesh first triangle (global numbering), for primitive ids
    int32_t kind;                       // RT_OBJ_*
    int32_t geom;
    float radiance[3];                  // object light: emitted radiance (front side), else 0
    int32_t pad;
};
// Sampling record of an object light (RT_LIGHT_MESH / RT_LIGHT_SPHERE), indexed like DScene.lights; world space.
struct DLightGeom {
    float4 sphere;                      // sphere light: centre, radius
    int32_t first_tri, tri_count;       // mesh light: its triangles in DScene.light_tris
    float area;                         // mesh light: total world-space area
    int32_t pad;
};
struct DImage { const void* pixels; int32_t width, height, format, pad; };

struct DScene {
    // geometry for traversal
    const uint4* nodes4q;               // BvhNode64Q as 4 x uint4: the same 4-wide tree (same node indices) with 8-bit quantised boxes; NULL unless RT795_QNODES=1
    const float4* nodes4x;              // experiment RT795_NODES256: 256-byte nodes holding every axis pair in both orders (near,far picked by a 32-byte load address); else NULL
    const float4* nodes4;               // BvhNode128 as 8 x float4 (4-wide collapse of the SAH BVH2); every BLAS first, the TLAS nodes last
    const int32_t* obj_root4;           // per top-level entry: BLAS root in nodes4 (spheres: unused); the kernels read it from obj_xf, this copy serves rt_debug_read_bvh
    int32_t tlas_root4;                 // first TLAS node = its root; nodes [tlas_root4, tlas_root4 + tlas_nodes) are the TLAS
    int32_t tlas_nodes;
    const float4* tris_w;               // triangles in leaf order as ray -> (u, v, w) affine maps, 3 rows x float4 (traverse.cuh tri_test_affine)
    const float4* tris;                 // 3 x float4 per triangle in leaf order: (v0, bits prim), (e1, 0), (e2, 0); read by the shade stage
    const DObjectXf* obj_xf;
    const float4* spheres;              // center.xyz, R
    // shading data
    const DObjectSh* obj_sh;
    const int32_t* tri_idx;             // 3 ints per triangle (original order)
    const float* vertices;              // 3 per vertex
    const float* normals;               // 3 per vertex
    const float* texcoords;             // 2 per entry
    const struct RtMaterial* materials;
    const struct RtTexture* textures;
    const DImage* images;
    const struct RtLight* lights;
    const DLightGeom* light_geom;       // per light (zeros for the analytic light types)
    const float4* light_tris;           // mesh lights: 3 x float4 per triangle, world space: (a, cdf), (b, 0), (c, 0); cdf = cumulative area fraction
    int32_t num_lights, num_objects;
    int32_t max_depth;
    float shadow_eps, int_eps;
    float background[3], ambient[3];
    int32_t env_light, background_texture;
};

struct DCamera {                        // Camera::Camera, src/Camera.cpp:7-61
    float pos[3], gaze[3], up[3], right[3];
    float l, r, b, t, dist;
    int32_t nx, ny, spp, n;             // n = stratification grid edge
    float invnx, invny, pw, ph, sw, sh;
    int32_t dof; float focus, aperture;
};

struct DTiling {                        // interleaved tile partition across ranks
    int32_t tw, th, tiles_x, tiles_y, rank, count, local_tiles;
};

// One level of the ray queue, SoA of 16-byte records (coalesced float4 loads/stores).
struct DRayQueue {
    float4* o_time;                     // origin.xyz, time
    float4* d_tmax;                     // direction.xyz, tmax
    float4* w_pixel;                    // throughput weight rgb, bits(pixel index)
    int4* meta;                         // x sample, y depth|flags<<8|beer_mat<<16, z path node, w bits(beer_s)
    float4* hit;                        // t, beta, gamma, bits(prim)   (prim < 0: miss; sphere: prim = 0)
    int32_t* hit_obj;                   // top-level entry index
};
struct DShadowQueue {
    float4* o_time;
    float4* d_tmax;
    float4* c_pixel;                    // contribution rgb (added if unoccluded), bits(pixel)
};

enum : uint32_t { RAY_PRIMARY = 1u, RAY_BEER_SELF = 2u, RAY_BEER_SIBLING = 4u, RAY_SPECULAR = 8u,
                  RAY_DEAD = 16u };   // RAY_DEAD: level-0 slot outside the image (fused raygen files one record per work item)

// Per-pipeline device counters.  A wave's queue sizes live ONLY here: every kernel of the wave reads its input count from
// device memory, so the host enqueues a whole frame without a single round trip (core.cu).
struct DCounters {
    int32_t level_count[kMaxLevels + 1];    // rays in queue level L of the current wave (level 0: written by the host / raygen)
    int32_t shadow_count[kMaxLevels];       // shadow rays emitted by the shade pass of level L
    int32_t cursor_extend[kMaxLevels];      // persistent-kernel ray cursors, one per launch of the wave
    int32_t cursor_shadow[kMaxLevels];
    int32_t overflow;                       // a queue was full: the frame is redone with smaller waves (never silent)
    int32_t pad;
    unsigned long long total_secondary, total_shadow;           // folded in by k_wave_end
    unsigned long long node_visits, prim_tests;                 // closest-hit queries (COUNT builds)
    unsigned long long node_visits_shadow, prim_tests_shadow;   // occlusion queries
    unsigned long long diag[8];   // COUNT builds: [0/4] node-step warp executions, [1/5] sum of lanes alive (ray unfinished) at those, [2/6] leaf-phase executions, [3/7] lanes in them (closest / occlusion)
};

}  // namespace rt795
