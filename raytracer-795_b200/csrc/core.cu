// Core of librt795: scene flatten/upload and the wavefront render driver behind the C ABI (include/rt795.h).
//
// rt_scene_create  replaces BVH construction (src/Scene.cpp:321-323, src/BVH.cpp:53-110) with the host SAH builder
//                  (bvh_build.cpp) and uploads the flattened scene to HBM.
// rt_render*       replaces the 8-thread pixel loop (src/Scene.cpp:269-292, 336-356) with
//                  generate -> [extend -> shade -> shadow]* -> resolve over ray queues in HBM.
// There is no CPU rendering path in this library.
#include "../../include/rt795.h"
#include "bvh_build.h"
#include "bvh_device.h"
#include "kernels.cuh"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

namespace rt795 {
void SetLastError(const std::string& msg);
}
using namespace rt795;

namespace {

struct CudaFail { cudaError_t e; const char* what; int line; };
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw CudaFail{e_, #call, __LINE__}; } while (0)
struct Invalid { std::string msg; };

inline float BitsToFloat(int v) { float f; memcpy(&f, &v, 4); return f; }

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { Free(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    void Alloc(size_t count) {
        if (count <= n && p) return;
        Free();
        if (count == 0) count = 1;
        CK(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void Upload(const T* h, size_t count, cudaStream_t st) {
        Alloc(count);
        if (count) CK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void Free() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { Free(); }
};

struct QueueBufs {
    QueueBufs() = default;
    QueueBufs(QueueBufs&&) noexcept = default;
    QueueBufs& operator=(QueueBufs&&) noexcept = default;
    DevBuf<float4> o_time, d_tmax, w_pixel, hit;
    DevBuf<int4> meta;
    DevBuf<int32_t> hit_obj;
    void Alloc(size_t cap) { o_time.Alloc(cap); d_tmax.Alloc(cap); w_pixel.Alloc(cap); hit.Alloc(cap); meta.Alloc(cap); hit_obj.Alloc(cap); }
    DRayQueue View() { return DRayQueue{o_time.p, d_tmax.p, w_pixel.p, meta.p, hit.p, hit_obj.p}; }
};

}  // namespace

// Per-pipeline render state.  Alternate waves of a frame go to two pipelines (own stream, ray queues, shadow queue, counters),
// all enqueued by ONE host thread without a single round trip inside the frame: every kernel reads its input size from device
// memory (DCounters).  Two streams, so that the tail of one wave's persistent kernels (last rays finishing on a mostly idle
// GPU) overlaps with the other wave's work.
struct Pipe {
    cudaStream_t st = nullptr;
    std::vector<QueueBufs> levels;
    DevBuf<float4> sq_o, sq_d, sq_c;
    DevBuf<float4> sq_hit; DevBuf<int32_t> sq_hit_obj;   // RtRenderOpts.parity_exact only: closest hits of the shadow rays
    DevBuf<DCounters> counters;
    DCounters* h_counters = nullptr;   // pinned
    DevBuf<int32_t> hist;              // per wave of this pipeline: level / shadow-queue sizes (k_wave_end), read back once per frame
    int32_t* h_hist = nullptr; size_t h_hist_n = 0;   // pinned
    cudaEvent_t done = nullptr;
};
constexpr int kPipes = 4;      // upper bound; RtScene::npipes_wanted (default 2, RT795_PIPES) are used

typedef void (*TraceKernel)(DScene, TraceIO, DCounters*);
typedef void (*RaygenTraceKernel)(DScene, TraceIO, DCounters*, RaygenArgs);

struct RtScene {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    DScene d{};
    // geometry + shading buffers
    DevBuf<float4> nodes4, tris, spheres;
    DevBuf<float4> nodes4x;            // RT795_NODES256 builds only
    DevBuf<float4> nodes2_debug;       // binary BVH the 4-wide one is collapsed from: resident only with RT795_DEBUG_BVH=1 (rt_debug_read_bvh)
    DevBuf<uint4> nodes4q;             // 8-bit quantised twin of nodes4: resident only with RT795_QNODES=1 or RT795_DEBUG_BVH=1
    DevBuf<float4> tris_w;
    // RT795_QNODES=1: the occlusion kernel walks the 8-bit quantised copy of the 4-wide BVH (4 instead of 7 loads per node).
    // Measured on C4: 39.1 ms vs 34.9 ms with fp32 nodes - the byte decode (24 PRMT + 24 FADD per node) costs more on the
    // ALU-limited kernel than the saved L1 wavefronts return, so fp32 nodes stay the default.
    bool quantised_occlusion = false;
    bool keep_debug_bvh = false;
    DevBuf<int32_t> obj_root4;
    DevBuf<DObjectXf> obj_xf;
    DevBuf<DObjectSh> obj_sh;
    DevBuf<int32_t> tri_idx;
    DevBuf<float> vertices, normals, texcoords;
    DevBuf<RtMaterial> materials;
    DevBuf<RtTexture> textures;
    DevBuf<DImage> images;
    std::vector<DevBuf<unsigned char>> image_px;
    DevBuf<RtLight> lights;
    DevBuf<DLightGeom> light_geom;
    DevBuf<float4> light_tris;
    bool has_dielectric = false;
    bool has_specular = false;             // some material spawns continuation rays (mirror / conductor / dielectric)
    int bvh_depth = 0;
    int tlas_base2 = 0;
    // kernel variants (A/B switches, all default to the measured best): RT795_SMEM_TLAS, RT795_FUSED_RAYGEN
    bool packet_primary = false;           // RT795_PACKET=1: camera rays as warp packets (traverse_packet.cuh)
    bool smem_tlas = false, fused_raygen = true;   // measured on C4 (profiles/bench/r02_ab_driver.txt): staging 54.4 ms vs 53.7 ms without
    int trace_smem = 0;                    // dynamic shared memory of the traversal launches (staged top level), 0 = not staged
    TraceKernel k_closest = nullptr, k_closest_count = nullptr, k_any = nullptr, k_any_count = nullptr;
    RaygenTraceKernel k_rg_closest = nullptr, k_rg_closest_count = nullptr;
    int grid_extend = 0, grid_shadow = 0;   // resident CTAs: SM count x occupancy
    // TraceIO.flags for closest-hit / occlusion launches (RT795_FLAGS_CLOSEST / RT795_FLAGS_OCCLUSION / RT795_TRAVERSAL_FLAGS override):
    // postponed leaves (2) for both, dynamic fetch (1) for occlusion, CTA chunks of 256 / 512 queue entries (bits 16..23).  Chunks, C4 frame
    // (profiles/bench/r02_ab_chunks_c4.txt): none 52.77 ms, 256/512 51.26, 512/1024 50.92 - but from 1024 up the single-kernel stage times
    // grow (tail: occlusion 28.8 -> 29.3 ms at 1024, 36.6 at 4096), so the sizes that help the frame without lengthening a stage are used.
    int trav_flags = (1 << 16) | 2, trav_flags_shadow = (2 << 16) | 3;
    bool full_grid = false;                 // RT795_FULL_GRID=1: one thread per ray instead of resident warps
    // render state
    Pipe pipes[kPipes];
    size_t queue_cap = 0, shadow_cap = 0;
    // Frames are first enqueued optimistically (whole levels, two ping-pong queue buffers); a queue overflow flagged by the device
    // halves `wave_limit` and redoes the frame, and below kMinOptimisticWave the chunked, host-synchronised driver takes over.
    // Sizes of the deeper levels are known to the device only; what the previous frame saw for the same wave (its level and
    // shadow-queue sizes, [wave][2 * (nlev + 1)]) sizes this frame's grids.  Only a hint: every launch is correct for any size
    // (persistent loops, a looping remainder).  RT795_HINTS=0 switches them off (A/B).
    std::vector<int> hints;
    int hint_stride = 0, hint_wave = 0;
    bool use_hints = true;
    long long hint_key[6] = {0, 0, 0, 0, 0, 0};
    size_t wave_limit = 0;                  // 0 = not limited yet
    size_t learnt_cap = 0;                  // queue capacity wave_limit / force_chunked were learnt with
    bool force_chunked = false, env_chunked = false;   // RT795_DRIVER=chunked, or learnt from overflows
    DevBuf<float> accum, out_dev;
    float* h_stage = nullptr; size_t h_stage_n = 0;   // pinned staging for rt_render's host framebuffer
    DevBuf<RtHit> hits_dev;
    // tone map scratch
    DevBuf<double> tm_sum; DevBuf<float> tm_max; DevBuf<unsigned> tm_hist; DevBuf<float> tm_in; DevBuf<uint8_t> tm_out;
    RtStats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_start = nullptr;
    int npipes_wanted = 2;      // concurrent wave pipelines (RT795_PIPES=1..4 for A/B)
};

namespace {

void CameraSetup(const RtCameraDesc* c, DCamera* o) {     // Camera::Camera, src/Camera.cpp:7-61
    memset(o, 0, sizeof *o);
    int n = 1;
    for (int i = 1; i < 1000; i++) if (i * i >= c->num_samples) { n = i; break; }
    float3 gaze = v3(c->gaze), up = v3(c->up);
    float3 g = normalize0(gaze);
    float3 w = c->left_handed ? normalize0(gaze) : normalize0(-gaze);
    float3 right = normalize0(cross(up, w));
    float3 upv = cross(w, right);
    memcpy(o->pos, c->position, 12);
    o->gaze[0] = g.x; o->gaze[1] = g.y; o->gaze[2] = g.z;
    o->up[0] = upv.x; o->up[1] = upv.y; o->up[2] = upv.z;
    o->right[0] = right.x; o->right[1] = right.y; o->right[2] = right.z;
    o->l = c->left; o->r = c->right; o->b = c->bottom; o->t = c->top; o->dist = c->distance;
    o->nx = c->nx; o->ny = c->ny; o->spp = c->num_samples < 1 ? 1 : c->num_samples; o->n = n;
    o->invnx = 1.0f / c->nx; o->invny = 1.0f / c->ny;
    o->pw = (o->r - o->l) * o->invnx; o->ph = (o->t - o->b) * o->invny;
    o->sw = o->pw / n; o->sh = o->ph / n;
    o->dof = c->is_dof; o->focus = c->focus_distance; o->aperture = c->aperture_size;
}

void Validate(const RtSceneDesc* s) {
    auto bad = [](const std::string& m) { throw Invalid{m}; };
    if (!s) bad("null scene description");
    if (s->abi_version != RT795_ABI_VERSION) bad("abi_version mismatch");
    if (s->num_vertices < 0 || s->num_triangles < 0 || s->num_objects < 0 || s->num_meshes < 0 || s->num_spheres < 0 || s->num_textures < 0 ||
        s->num_images < 0 || s->num_lights < 0 || s->num_texcoords < 0) bad("negative count");
    if ((s->num_meshes && !s->meshes) || (s->num_spheres && !s->spheres) || (s->num_objects && !s->objects) || (s->num_textures && !s->textures) ||
        (s->num_images && !s->images) || (s->num_lights && !s->lights) || (s->num_texcoords && !s->texcoords)) bad("array pointer missing for a non-zero count");
    if (s->num_triangles > (1 << 28) || s->num_vertices > (1 << 30)) bad("scene too large (more than 2^28 triangles)");   // leaf refs carry 28-bit triangle indices
    if (s->num_vertices && !s->vertices) bad("vertices missing");
    if (s->num_triangles && !s->tri_indices) bad("tri_indices missing");
    if (s->num_materials < 1 || !s->materials) bad("at least one material is required");
    for (int i = 0; i < s->num_triangles * 3; i++) if (s->tri_indices[i] < 0 || s->tri_indices[i] >= s->num_vertices) bad("triangle vertex index out of range");
    for (int m = 0; m < s->num_meshes; m++) {
        const RtMesh& me = s->meshes[m];
        if (me.first_tri < 0 || me.tri_count < 1 || (long long)me.first_tri + me.tri_count > s->num_triangles) bad("mesh triangle range out of bounds");
        if (me.smooth && !s->vertex_normals) bad("smooth mesh without vertex_normals");
    }
    for (int t = 0; t < s->num_textures; t++) {
        const RtTexture& tx = s->textures[t];
        if (tx.type == RT_TEX_IMAGE && (tx.image < 0 || tx.image >= s->num_images)) bad("texture image index out of range");
    }
    for (int i = 0; i < s->num_images; i++) {
        if (s->images[i].width < 2 || s->images[i].height < 2 || !s->images[i].pixels) bad("image must be at least 2x2");
        if (s->images[i].format != RT_IMG_RGBA8 && s->images[i].format != RT_IMG_RGBA32F) bad("unknown image format");
        if ((long long)s->images[i].width * s->images[i].height > (1ll << 30)) bad("image too large");
    }
    for (int i = 0; i < s->num_objects; i++) {
        const RtObject& o = s->objects[i];
        if (o.kind == RT_OBJ_SPHERE) { if (o.geom < 0 || o.geom >= s->num_spheres) bad("object sphere index out of range"); }
        else if (o.kind == RT_OBJ_MESH) { if (o.geom < 0 || o.geom >= s->num_meshes) bad("object mesh index out of range"); }
        else bad("unknown object kind");
        if (o.mat_index < 1 || o.mat_index > s->num_materials) bad("object material index out of range");
        for (int k = 0; k < 2; k++) if (o.tex[k] < 0 || o.tex[k] > s->num_textures) bad("object texture id out of range");
        if (o.kind == RT_OBJ_MESH && (o.tex[0] || o.tex[1])) {
            const RtMesh& me = s->meshes[o.geom];
            for (int t = me.first_tri * 3; t < (me.first_tri + me.tri_count) * 3; t++) {
                int uv = s->tri_indices[t] + o.tex_offset;
                if (uv < 0 || uv >= s->num_texcoords) bad("textured mesh needs texture coordinates for every vertex");
            }
        }
    }
    for (int i = 0; i < s->num_lights; i++) {
        const RtLight& l = s->lights[i];
        if (l.type < RT_LIGHT_POINT || l.type > RT_LIGHT_SPHERE) bad("unknown light type");
        if (l.type == RT_LIGHT_ENVIRONMENT && (l.image < 0 || l.image >= s->num_images)) bad("environment light image index out of range");
        if (l.type == RT_LIGHT_MESH || l.type == RT_LIGHT_SPHERE) {
            if (l.object < 0 || l.object >= s->num_objects) bad("object light: object index out of range");
            if (s->objects[l.object].kind != (l.type == RT_LIGHT_MESH ? RT_OBJ_MESH : RT_OBJ_SPHERE)) bad("object light: the object is not of the light's kind");
        }
    }
    if (s->bvh_builder < RT_BVH_AUTO || s->bvh_builder > RT_BVH_DEVICE_SAH) bad("bvh_builder must be one of RT_BVH_*");
    if (s->env_light_index < -1 || s->env_light_index >= s->num_lights) bad("env_light_index out of range");
    if (s->background_texture < -1 || s->background_texture >= s->num_textures) bad("background_texture out of range");
}

constexpr int kDeviceBuildMinTris = 32768;   // RT_BVH_AUTO: smaller meshes build faster on the host than the ~40 launch round trips of the device build

// Returns false when device-built BLASes turned out too deep for the traversal stack: the caller retries with allow_device = false.
bool BuildAndUpload(RtScene* sc, const RtSceneDesc* s, bool allow_device) {
    auto t0 = std::chrono::steady_clock::now();
    cudaStream_t st = sc->stream;
    const float eps = s->int_test_eps;

    // ---- scene-wide vertex / index arrays first: the device BLAS builder reads them
    sc->tri_idx.Upload(s->tri_indices, (size_t)s->num_triangles * 3, st);
    sc->vertices.Upload(s->vertices, (size_t)s->num_vertices * 3, st);

    // ---- BLAS per mesh: large meshes on the device (bvh_device.cu), small ones on the host (bvh_build.cpp).  Every mesh yields
    // one segment of each scene-wide array (binary nodes, 4-wide nodes, quantised nodes, triangles), child references already
    // rebased to the segment's position.
    struct MeshSeg {
        std::vector<BvhNode64> h2; std::vector<BvhNode128> h4; std::vector<BvhNode64Q> h4q; std::vector<float4> htris;
        DeviceBlas dev; bool on_device = false;
        int n2 = 0, n4 = 0, ntris = 0;
        MeshSeg() = default;
        MeshSeg(const MeshSeg&) = delete;
        MeshSeg(MeshSeg&& o) noexcept : h2(std::move(o.h2)), h4(std::move(o.h4)), h4q(std::move(o.h4q)), htris(std::move(o.htris)), dev(o.dev), on_device(o.on_device), n2(o.n2), n4(o.n4), ntris(o.ntris) { o.dev = DeviceBlas(); }
        ~MeshSeg() { FreeDeviceBlas(&dev); }
    };
    std::vector<MeshSeg> segs;
    segs.reserve(s->num_meshes + 1);
    std::vector<int> mesh_root(s->num_meshes, 0), mesh_depth(s->num_meshes, 0), mesh_root4(s->num_meshes, 0), mesh_stack4(s->num_meshes, 0);
    std::vector<Aabb> mesh_bounds(s->num_meshes);
    int n2_total = 0, n4_total = 0, tri_total = 0;
    int builder = s->bvh_builder;
    if (const char* e = getenv("RT795_BVH_BUILDER")) builder = !strcmp(e, "host") ? RT_BVH_HOST_SAH : !strcmp(e, "device") || !strcmp(e, "ploc") ? RT_BVH_DEVICE_PLOC : !strcmp(e, "sah") ? RT_BVH_DEVICE_SAH : RT_BVH_AUTO;
    if (!allow_device) builder = RT_BVH_HOST_SAH;
    sc->stats.bvh_device_meshes = 0; sc->stats.ms_bvh_device = 0; sc->stats.bvh_ploc_iterations = 0;
    for (int m = 0; m < s->num_meshes; m++) {
        const RtMesh& me = s->meshes[m];
        segs.emplace_back();
        MeshSeg& sg = segs.back();
        const int node_base = n2_total, base4 = n4_total, tri_base = tri_total;
        mesh_root[m] = node_base; mesh_root4[m] = base4;
        bool want_device = (builder == RT_BVH_DEVICE_PLOC || builder == RT_BVH_DEVICE_SAH) ? me.tri_count >= 8 : builder == RT_BVH_AUTO && me.tri_count >= kDeviceBuildMinTris;
        int algo = builder == RT_BVH_DEVICE_PLOC ? kDeviceAlgoPloc : kDeviceAlgoSah;   // AUTO: the host's tree, built on the device
        if (want_device) {
            // depth limits leave room for a TLAS of the same order as the host's (checked exactly below; a miss rebuilds on the host)
            try {
                sg.on_device = BuildBlasOnDevice(sc->vertices.p, sc->tri_idx.p, me.first_tri, me.tri_count, eps, algo, node_base, base4, tri_base,
                                                 kStackSize - 12, kStackSize - 12, st, &sg.dev);
            } catch (const std::exception&) {
                // e.g. not enough free HBM for the builder's temporaries: the host builder needs none.  A sticky CUDA error
                // resurfaces at the next CK() below with its own message.
                cudaGetLastError();
                sg.on_device = false;
            }
            if (sg.on_device) {
                sg.n2 = sg.dev.n2; sg.n4 = sg.dev.n4; sg.ntris = sg.dev.tri_count;
                mesh_depth[m] = sg.dev.depth2; mesh_stack4[m] = sg.dev.max_stack4; mesh_bounds[m] = sg.dev.bounds;
                sc->stats.bvh_device_meshes++; sc->stats.ms_bvh_device += sg.dev.ms; sc->stats.bvh_ploc_iterations = sg.dev.ploc_iterations;
            }
        }
        if (!sg.on_device) {
            std::vector<Aabb> pb(me.tri_count);
            for (int t = 0; t < me.tri_count; t++) {
                const int* ix = s->tri_indices + 3 * (size_t)(me.first_tri + t);
                Aabb b; b.Reset();
                float ext = 0.f;
                for (int k = 0; k < 3; k++) b.Grow(s->vertices + 3 * (size_t)ix[k]);
                for (int k = 0; k < 3; k++) ext = std::max(ext, b.hi[k] - b.lo[k]);
                // the reference accepts beta, gamma >= -eps: hits up to eps*edge outside the triangle; pad so the box stays conservative
                for (int k = 0; k < 3; k++) {
                    float mag = std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k]));
                    float pad = 2.0f * eps * ext + 4.0e-7f * mag + 1e-30f;
                    b.lo[k] -= pad; b.hi[k] += pad;
                }
                pb[t] = b;
            }
            BuiltBvh bv;
            BuildSahBvh(pb.data(), me.tri_count, 4, &bv);
            BuiltBvh4 b4;
            CollapseToBvh4(bv, &b4);
            mesh_stack4[m] = b4.max_stack;
            auto rebase = [&](int32_t ch, int inner_base) {
                if (ch == kNoChild) return ch;
                if (ch >= 0) return ch + inner_base;
                int payload = ~ch; int first = (payload >> 3) + tri_base;
                return (int32_t)~((first << 3) | (payload & 7));
            };
            if (sc->quantised_occlusion || sc->keep_debug_bvh) QuantiseBvh4(b4, &sg.h4q);
            for (BvhNode64Q& nq : sg.h4q)
                for (int c = 0; c < 4; c++) { int32_t ch; memcpy(&ch, &nq.w[4 + c], 4); ch = rebase(ch, base4); memcpy(&nq.w[4 + c], &ch, 4); }
            sg.h4 = std::move(b4.nodes);
            for (BvhNode128& nd : sg.h4)
                for (int c = 0; c < 4; c++) { int32_t ch; memcpy(&ch, &nd.q[24 + c], 4); ch = rebase(ch, base4); memcpy(&nd.q[24 + c], &ch, 4); }
            mesh_depth[m] = bv.max_depth;
            mesh_bounds[m] = bv.bounds;
            sg.h2 = std::move(bv.nodes);
            for (BvhNode64& nd : sg.h2)
                for (int c = 0; c < 2; c++) { int32_t ch; memcpy(&ch, &nd.q[12 + c], 4); ch = rebase(ch, node_base); memcpy(&nd.q[12 + c], &ch, 4); }
            sg.htris.reserve((size_t)me.tri_count * 3);
            for (int k = 0; k < me.tri_count; k++) {
                int gtri = me.first_tri + bv.order[k];
                const int* ix = s->tri_indices + 3 * (size_t)gtri;
                const float *a = s->vertices + 3 * (size_t)ix[0], *b = s->vertices + 3 * (size_t)ix[1], *c = s->vertices + 3 * (size_t)ix[2];
                sg.htris.push_back(make_float4(a[0], a[1], a[2], BitsToFloat(gtri)));
                sg.htris.push_back(make_float4(b[0] - a[0], b[1] - a[1], b[2] - a[2], 0.f));
                sg.htris.push_back(make_float4(c[0] - a[0], c[1] - a[1], c[2] - a[2], 0.f));
            }
            sg.n2 = (int)sg.h2.size(); sg.n4 = (int)sg.h4.size(); sg.ntris = me.tri_count;
        }
        n2_total += sg.n2; n4_total += sg.n4; tri_total += sg.ntris;
    }

    // ---- top-level entries + TLAS
    std::vector<DObjectXf> oxf(s->num_objects);
    std::vector<DObjectSh> osh(s->num_objects);
    std::vector<Aabb> obounds(s->num_objects);
    std::vector<float4> spheres(s->num_spheres);
    for (int i = 0; i < s->num_spheres; i++) spheres[i] = make_float4(s->spheres[i].center[0], s->spheres[i].center[1], s->spheres[i].center[2], s->spheres[i].radius);
    int max_blas_depth = 0, max_blas_stack4 = 0;
    std::vector<int32_t> root4(std::max(1, s->num_objects), 0);
    struct Fwd { double M[3][3], T[3]; };
    std::vector<Fwd> fwd(s->num_objects);                                // forward (object -> world) transforms, for the object lights
    for (int i = 0; i < s->num_objects; i++) {
        const RtObject& o = s->objects[i];
        DObjectXf& x = oxf[i];
        x.inv0 = make_float4(o.inv[0], o.inv[1], o.inv[2], o.inv[3]);
        x.inv1 = make_float4(o.inv[4], o.inv[5], o.inv[6], o.inv[7]);
        x.inv2 = make_float4(o.inv[8], o.inv[9], o.inv[10], o.inv[11]);
        Aabb ob;   // object-space bounds
        int root;
        if (o.kind == RT_OBJ_SPHERE) {
            const RtSphere& sp = s->spheres[o.geom];
            for (int k = 0; k < 3; k++) { ob.lo[k] = sp.center[k] - sp.radius; ob.hi[k] = sp.center[k] + sp.radius; }
            root = ~o.geom;
        } else {
            ob = mesh_bounds[o.geom]; root = mesh_root4[o.geom]; max_blas_depth = std::max(max_blas_depth, mesh_depth[o.geom]);
            root4[i] = mesh_root4[o.geom]; max_blas_stack4 = std::max(max_blas_stack4, mesh_stack4[o.geom]);
        }
        x.blur_root = make_float4(o.blur[0], o.blur[1], o.blur[2], BitsToFloat(root));
        // forward matrix M = (M^-1)^-1 (affine): invert the 3x3 part in double
        double A[3][3], t[3];
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) A[r][c] = o.inv[r * 4 + c]; t[r] = o.inv[r * 4 + 3]; }
        double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        if (det == 0.0 || det != det) throw Invalid{"object has a singular inverse transformation"};
        double M[3][3];
        M[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det; M[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det; M[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
        M[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det; M[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det; M[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
        M[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det; M[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det; M[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
        double T[3];
        for (int r = 0; r < 3; r++) T[r] = -(M[r][0] * t[0] + M[r][1] * t[1] + M[r][2] * t[2]);
        memcpy(fwd[i].M, M, sizeof M); memcpy(fwd[i].T, T, sizeof T);
        Aabb wb; wb.Reset();
        for (int corner = 0; corner < 8; corner++) {
            double p[3] = {corner & 1 ? ob.hi[0] : ob.lo[0], corner & 2 ? ob.hi[1] : ob.lo[1], corner & 4 ? ob.hi[2] : ob.lo[2]};
            float q[3];
            for (int r = 0; r < 3; r++) q[r] = (float)(M[r][0] * p[0] + M[r][1] * p[1] + M[r][2] * p[2] + T[r]);
            wb.Grow(q);
            float qb[3] = {q[0] + o.blur[0], q[1] + o.blur[1], q[2] + o.blur[2]};   // swept by MotionBlur * time, time in [0,1]
            wb.Grow(qb);
        }
        for (int k = 0; k < 3; k++) {
            float mag = std::max(std::fabs(wb.lo[k]), std::fabs(wb.hi[k]));
            float pad = 1.0e-5f * mag + 1.0e-5f * (wb.hi[k] - wb.lo[k]) + 1e-30f;
            wb.lo[k] -= pad; wb.hi[k] += pad;
        }
        obounds[i] = wb;
        DObjectSh& h = osh[i];
        memcpy(h.nrm, o.inv_t, sizeof h.nrm);
        h.mat_index = o.mat_index; h.tex[0] = o.tex[0]; h.tex[1] = o.tex[1]; h.tex_offset = o.tex_offset;
        h.kind = o.kind; h.geom = o.geom;
        h.smooth = o.kind == RT_OBJ_MESH ? s->meshes[o.geom].smooth : 0;
        h.first_tri = o.kind == RT_OBJ_MESH ? s->meshes[o.geom].first_tri : 0;
        memcpy(h.radiance, o.radiance, 12); h.pad = 0;
    }
    // ---- object lights: world-space sampling records (a sphere light keeps its shape under a similarity transform only: its
    // radius is scaled by the cube root of |det M|; motion blur does not move emitters)
    std::vector<DLightGeom> lgeom(std::max(1, s->num_lights));
    std::vector<float4> ltris;
    memset(lgeom.data(), 0, lgeom.size() * sizeof(DLightGeom));
    for (int li = 0; li < s->num_lights; li++) {
        const RtLight& L = s->lights[li];
        if (L.type != RT_LIGHT_MESH && L.type != RT_LIGHT_SPHERE) continue;
        const RtObject& o = s->objects[L.object];
        const Fwd& f = fwd[L.object];
        auto xf = [&](const float* p, double* q) { for (int r = 0; r < 3; r++) q[r] = f.M[r][0] * p[0] + f.M[r][1] * p[1] + f.M[r][2] * p[2] + f.T[r]; };
        DLightGeom& g = lgeom[li];
        if (L.type == RT_LIGHT_SPHERE) {
            const RtSphere& sp = s->spheres[o.geom];
            double c[3]; xf(sp.center, c);
            double det = f.M[0][0] * (f.M[1][1] * f.M[2][2] - f.M[1][2] * f.M[2][1]) - f.M[0][1] * (f.M[1][0] * f.M[2][2] - f.M[1][2] * f.M[2][0]) + f.M[0][2] * (f.M[1][0] * f.M[2][1] - f.M[1][1] * f.M[2][0]);
            g.sphere = make_float4((float)c[0], (float)c[1], (float)c[2], (float)(sp.radius * std::cbrt(std::fabs(det))));
        } else {
            const RtMesh& me = s->meshes[o.geom];
            g.first_tri = (int)ltris.size() / 3; g.tri_count = me.tri_count;
            double total = 0;
            std::vector<double> areas(me.tri_count);
            for (int t = 0; t < me.tri_count; t++) {
                const int* ix = s->tri_indices + 3 * (size_t)(me.first_tri + t);
                double a[3], b[3], c[3];
                xf(s->vertices + 3 * (size_t)ix[0], a); xf(s->vertices + 3 * (size_t)ix[1], b); xf(s->vertices + 3 * (size_t)ix[2], c);
                double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
                double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                areas[t] = 0.5 * std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                total += areas[t];
                ltris.push_back(make_float4((float)a[0], (float)a[1], (float)a[2], 0.f));
                ltris.push_back(make_float4((float)b[0], (float)b[1], (float)b[2], 0.f));
                ltris.push_back(make_float4((float)c[0], (float)c[1], (float)c[2], 0.f));
            }
            double run = 0;
            for (int t = 0; t < me.tri_count; t++) { run += areas[t]; ltris[(size_t)(g.first_tri + t) * 3].w = t == me.tri_count - 1 ? 1.0f : (float)(total > 0 ? run / total : 1.0); }
            g.area = (float)total;
        }
    }
    BuiltBvh tl;
    BuildSahBvh(obounds.data(), s->num_objects, 1, &tl);
    const int tlas_base = n2_total, tlas_base4 = n4_total;
    BuiltBvh4 t4;
    CollapseToBvh4(tl, &t4);
    {
        int need4 = t4.max_stack + max_blas_stack4 + 4;
        int need2 = tl.max_depth + max_blas_depth + 3;
        if (need4 > kStackSize || need2 > kStackSize) {
            if (sc->stats.bvh_device_meshes > 0) return false;           // device-built trees ran deeper than the estimate: rebuild on the host
            if (need4 > kStackSize) throw Invalid{"4-wide BVH needs a deeper traversal stack (" + std::to_string(need4) + " > " + std::to_string(kStackSize) + ")"};
            throw Invalid{"BVH too deep for the traversal stack (" + std::to_string(need2) + " > " + std::to_string(kStackSize) + ")"};
        }
        sc->bvh_depth = need2;
    }
    std::vector<BvhNode64Q> tq;
    if (sc->quantised_occlusion || sc->keep_debug_bvh) QuantiseBvh4(t4, &tq);
    auto tlas_ref = [&](int32_t ch, int inner_base) {
        if (ch == kNoChild) return ch;
        if (ch >= 0) return ch + inner_base;
        int payload = ~ch; int obj = tl.order[payload >> 3];
        return (int32_t)~(obj << 3);                                     // TLAS leaves hold exactly one entry
    };
    for (BvhNode64& nd : tl.nodes)
        for (int c = 0; c < 2; c++) { int32_t ch; memcpy(&ch, &nd.q[12 + c], 4); ch = tlas_ref(ch, tlas_base); memcpy(&nd.q[12 + c], &ch, 4); }
    for (BvhNode64Q& nq : tq)
        for (int c = 0; c < 4; c++) { int32_t ch; memcpy(&ch, &nq.w[4 + c], 4); ch = tlas_ref(ch, tlas_base4); memcpy(&nq.w[4 + c], &ch, 4); }
    for (BvhNode128& nd : t4.nodes)
        for (int c = 0; c < 4; c++) { int32_t ch; memcpy(&ch, &nd.q[24 + c], 4); ch = tlas_ref(ch, tlas_base4); memcpy(&nd.q[24 + c], &ch, 4); }
    n2_total += (int)tl.nodes.size(); n4_total += (int)t4.nodes.size();
    auto t1 = std::chrono::steady_clock::now();

    // ---- upload / gather the segments into the scene-wide arrays
    static_assert(sizeof(BvhNode64) == 64, "node size");
    static_assert(sizeof(BvhNode128) == 128, "node4 size");
    static_assert(sizeof(BvhNode64Q) == 64, "quantised node size");
    // Resident for rendering: the 4-wide nodes, the triangle records (shading) and their affine maps (traversal).  The binary BVH
    // the 4-wide one was collapsed from and the quantised twin are uploaded only for rt_debug_read_bvh (RT795_DEBUG_BVH=1) / the
    // quantised occlusion variant (RT795_QNODES=1): 45 MB less next to C4's 71 MB traversal working set.
    const bool want2 = sc->keep_debug_bvh, want4q = sc->keep_debug_bvh || sc->quantised_occlusion;
    if (want2) sc->nodes2_debug.Alloc((size_t)n2_total * 4);
    sc->nodes4.Alloc((size_t)n4_total * 8);
    if (want4q) sc->nodes4q.Alloc((size_t)n4_total * 4);
    sc->tris.Alloc((size_t)tri_total * 3);
    {
        size_t o2 = 0, o4 = 0, ot = 0;
        auto put = [&](void* dst, const void* src, size_t bytes, bool dev) {
            if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        };
        for (MeshSeg& sg : segs) {
            bool d = sg.on_device;
            if (want2) put(sc->nodes2_debug.p + o2 * 4, d ? (const void*)sg.dev.nodes2 : (const void*)sg.h2.data(), (size_t)sg.n2 * 64, d);
            put(sc->nodes4.p + o4 * 8, d ? (const void*)sg.dev.nodes4 : (const void*)sg.h4.data(), (size_t)sg.n4 * 128, d);
            if (want4q) put(sc->nodes4q.p + o4 * 4, d ? (const void*)sg.dev.nodes4q : (const void*)sg.h4q.data(), (size_t)sg.n4 * 64, d);
            put(sc->tris.p + ot * 3, d ? (const void*)sg.dev.tris : (const void*)sg.htris.data(), (size_t)sg.ntris * 48, d);
            o2 += sg.n2; o4 += sg.n4; ot += sg.ntris;
        }
        if (want2) put(sc->nodes2_debug.p + o2 * 4, tl.nodes.data(), tl.nodes.size() * 64, false);
        put(sc->nodes4.p + o4 * 8, t4.nodes.data(), t4.nodes.size() * 128, false);
        if (want4q) put(sc->nodes4q.p + o4 * 4, tq.data(), tq.size() * 64, false);
    }
    sc->obj_root4.Upload(root4.data(), root4.size(), st);
    sc->spheres.Upload(spheres.data(), spheres.size(), st);
    sc->obj_xf.Upload(oxf.data(), oxf.size(), st);
    sc->obj_sh.Upload(osh.data(), osh.size(), st);
    if (s->vertex_normals) sc->normals.Upload(s->vertex_normals, (size_t)s->num_vertices * 3, st); else sc->normals.Alloc(1);
    if (s->texcoords) sc->texcoords.Upload(s->texcoords, (size_t)s->num_texcoords * 2, st); else sc->texcoords.Alloc(1);
    sc->materials.Upload(s->materials, s->num_materials, st);
    sc->textures.Upload(s->textures, s->num_textures, st);
    std::vector<RtLight> lights(s->lights, s->lights + s->num_lights);
    sc->lights.Upload(lights.data(), lights.size(), st);
    sc->light_geom.Upload(lgeom.data(), lgeom.size(), st);
    if (ltris.empty()) ltris.push_back(make_float4(0.f, 0.f, 0.f, 1.f));
    sc->light_tris.Upload(ltris.data(), ltris.size(), st);
    std::vector<DImage> dimg(s->num_images);
    sc->image_px.resize(s->num_images);
    for (int i = 0; i < s->num_images; i++) {
        const RtImage& im = s->images[i];
        size_t bytes = (size_t)im.width * im.height * (im.format == RT_IMG_RGBA8 ? 4 : 16);
        sc->image_px[i].Upload((const unsigned char*)im.pixels, bytes, st);
        dimg[i] = DImage{sc->image_px[i].p, im.width, im.height, im.format, 0};
    }
    sc->images.Upload(dimg.data(), dimg.size(), st);
    CK(cudaStreamSynchronize(st));
    auto t2 = std::chrono::steady_clock::now();

    DScene& d = sc->d;
    d.nodes4q = want4q ? sc->nodes4q.p : nullptr;
    sc->tris_w.Alloc((size_t)std::max(1, tri_total) * 3);
    if (tri_total > 0) {
        k_make_affine_tris<<<(tri_total + 255) / 256, 256, 0, st>>>(sc->tris.p, tri_total, sc->tris_w.p);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
    }
    d.tris_w = sc->tris_w.p;
    d.nodes4x = nullptr;
#ifdef RT795_NODES256
    {
        int n4_all = n4_total;
        sc->nodes4x.Alloc((size_t)n4_all * 16);
        k_expand_nodes256<<<(n4_all + 127) / 128, 128, 0, st>>>(sc->nodes4.p, n4_all, sc->nodes4x.p);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        d.nodes4x = sc->nodes4x.p;
    }
#endif
    d.nodes4 = sc->nodes4.p; d.obj_root4 = sc->obj_root4.p; d.tlas_root4 = tlas_base4; d.tlas_nodes = (int)t4.nodes.size();
    d.tris = sc->tris.p; d.obj_xf = sc->obj_xf.p; d.spheres = sc->spheres.p; sc->tlas_base2 = tlas_base;
    d.obj_sh = sc->obj_sh.p; d.tri_idx = sc->tri_idx.p; d.vertices = sc->vertices.p; d.normals = sc->normals.p; d.texcoords = sc->texcoords.p;
    d.materials = sc->materials.p; d.textures = sc->textures.p; d.images = sc->images.p; d.lights = sc->lights.p;
    d.light_geom = sc->light_geom.p; d.light_tris = sc->light_tris.p;
    d.num_lights = s->num_lights; d.num_objects = s->num_objects; d.max_depth = s->max_recursion_depth;
    d.shadow_eps = s->shadow_ray_eps; d.int_eps = s->int_test_eps;
    memcpy(d.background, s->background_color, 12); memcpy(d.ambient, s->ambient_light, 12);
    d.env_light = s->env_light_index; d.background_texture = s->background_texture;
    for (int i = 0; i < s->num_materials; i++) {
        if (s->materials[i].type == RT_MAT_DIELECTRIC) sc->has_dielectric = true;
        if (s->materials[i].type != RT_MAT_NORMAL) sc->has_specular = true;
    }

    sc->stats.bvh2_nodes = (uint64_t)n2_total;
    sc->stats.bvh_nodes = (uint64_t)n4_total; sc->stats.bvh_bytes = (uint64_t)n4_total * 128; sc->stats.tri_bytes = (uint64_t)tri_total * 48;   // the 4-wide BVH is the one the render kernels walk
    sc->stats.ms_bvh_build = std::chrono::duration<double, std::milli>(t1 - t0).count();
    sc->stats.ms_upload = std::chrono::duration<double, std::milli>(t2 - t1).count();
    return true;
}

inline int Blocks(long long n) { return (int)((n + kBlock - 1) / kBlock); }

constexpr size_t kMinOptimisticWave = 4096;   // below this wave size the chunked driver takes over

struct StageEvents { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[4]; };   // 0 raygen(+first extend when fused), 1 extend, 2 shade, 3 shadow

struct Driver {
    RtScene* sc;
    Pipe* pp;
    cudaStream_t st;
    DCamera cam;
    DTiling tl;
    RenderConsts rc;
    RaygenArgs rg;
    bool count;        // node / primitive-test counters (collect_counters & 2)
    bool stage_time;   // per-stage CUDA-event timing (collect_counters & 1)
    bool path;
    bool exact;        // RtRenderOpts.parity_exact
    int nlev;          // ray-tree levels a wave can have (level 0 = camera rays)
    int nbuf;          // queue buffers per pipeline (2 = ping-pong)
    bool branching;
    float* accum;
    DShadowQueue sq;
    uint64_t launches = 0, n_extend = 0, n_shade = 0, n_shadow = 0, secondary = 0, shadow = 0;
    StageEvents* sev = nullptr;

    template <typename F>
    void Timed(int stage, F&& f) {
        if (!stage_time) { f(); return; }
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        CK(cudaEventRecord(a, st));
        f();
        CK(cudaEventRecord(b, st));
        sev->ev[stage].push_back({a, b});
    }


    // expected size of a device-sized queue: the hint with headroom, capped by what the queue can hold
    static long long Expect(int hint, long long upper) { return hint >= 0 ? std::min<long long>(upper, (long long)hint + hint / 4 + 2048) : upper; }   // hint < 0: unknown

    void LaunchExtend(const DRayQueue& q, const int32_t* count_ptr, int first, int limit, int32_t* cursor, int hint = -1) {
        TraceIO io{};
        io.q = q; io.count_ptr = count_ptr; io.first = first; io.limit = limit; io.cursor = cursor; io.flags = sc->trav_flags;
        int grid = sc->full_grid ? Blocks(limit) : std::min(Blocks(count_ptr ? Expect(hint, limit) : limit), sc->grid_extend);
        if (exact) k_trace_persistent<false, false, TF_EXACT><<<std::max(1, grid), kBlock, 0, st>>>(sc->d, io, pp->counters.p);
        else (count ? sc->k_closest_count : sc->k_closest)<<<std::max(1, grid), kBlock, count ? 0 : sc->trace_smem, st>>>(sc->d, io, pp->counters.p);
        launches++; n_extend++;
    }
    void LaunchShadow(const int32_t* count_ptr, int limit, int32_t* cursor, int hint = -1) {
        TraceIO io{};
        io.sq = sq; io.count_ptr = count_ptr; io.first = 0; io.limit = limit; io.cursor = cursor; io.accum = accum; io.flags = sc->trav_flags_shadow;
        int grid = sc->full_grid ? Blocks(limit) : std::min(Blocks(count_ptr ? Expect(hint, limit) : limit), sc->grid_shadow);
        if (exact) {
            // the reference's IsShadow is a full closest-hit query (src/Light.cpp:188-204): same exact kernel over the shadow queue, then
            // the contributions of the rays that hit nothing inside their interval are added
            pp->sq_hit.Alloc(sc->shadow_cap); pp->sq_hit_obj.Alloc(sc->shadow_cap);
            io.q = DRayQueue{sq.o_time, sq.d_tmax, nullptr, nullptr, pp->sq_hit.p, pp->sq_hit_obj.p};
            io.flags = sc->trav_flags;
            k_trace_persistent<false, false, TF_EXACT><<<std::max(1, grid), kBlock, 0, st>>>(sc->d, io, pp->counters.p);
            k_shadow_resolve<<<std::max(1, Blocks(limit)), kBlock, 0, st>>>(sq, pp->sq_hit.p, count_ptr, limit, accum);
            launches += 2; n_shadow++;
            return;
        }
        (count ? sc->k_any_count : sc->k_any)<<<std::max(1, grid), kBlock, (count || sc->quantised_occlusion) ? 0 : sc->trace_smem, st>>>(sc->d, io, pp->counters.p);
        launches++; n_shadow++;
    }
    void LaunchShade(const DRayQueue& in, const DRayQueue& out, const int32_t* in_count, int first, int limit, int32_t* out_count, int32_t* shadow_count, int hint = -1) {
        ShadeIO io{};
        io.in = in; io.out = out; io.sq = sq; io.in_count = in_count; io.first = first; io.limit = limit;
        io.out_count = out_count; io.shadow_count = shadow_count; io.overflow = &pp->counters.p->overflow; io.accum = accum;
        auto launch = [&](bool loop, int grid) {
            if (loop) {
                if (path) k_shade<RT_INTEGRATOR_PATH, true><<<grid, kBlock, 0, st>>>(sc->d, cam, rc, io);
                else k_shade<RT_INTEGRATOR_REFERENCE, true><<<grid, kBlock, 0, st>>>(sc->d, cam, rc, io);
            } else {
                if (path) k_shade<RT_INTEGRATOR_PATH, false><<<grid, kBlock, 0, st>>>(sc->d, cam, rc, io);
                else k_shade<RT_INTEGRATOR_REFERENCE, false><<<grid, kBlock, 0, st>>>(sc->d, cam, rc, io);
            }
            launches++; n_shade++;
        };
        if (!in_count) { launch(false, std::max(1, Blocks(limit))); return; }     // size known to the host: one 128-ray group per CTA
        // size known to the device only.  With a hint: the one-group-per-CTA variant over the expected size, then the grid-stride
        // variant on a small grid for whatever lies beyond (its CTAs leave at once when nothing does).  Without: grid-stride only.
        if (hint >= 0) {
            int covered = Blocks(Expect(hint, limit)) * kBlock;
            io.limit = std::min(limit, covered);
            launch(false, Blocks(io.limit));
            if (covered >= limit) return;
            io.first = first + covered; io.limit = limit - covered;
            launch(true, std::max(1, std::min(Blocks(io.limit), sc->sm_count)));
        } else launch(true, std::max(1, std::min(Blocks(limit), sc->sm_count * 8)));
    }
    // Camera rays of one wave into level 0 + their closest hits.  Fused: one launch, slot = work item, `n` slots.  Unfused (A/B and
    // the chunked driver): k_raygen compacts into level 0 and counts into level_count[0].
    void LaunchPrimary(unsigned long long base, int n, bool fused) {
        DCounters* c = pp->counters.p;
        RaygenArgs a = rg;
        a.base_rest = (unsigned)(base / (unsigned long long)cam.spp); a.base_sample = (unsigned)(base % (unsigned long long)cam.spp);
        DRayQueue q0 = pp->levels[0].View();
        if (fused) {
            Timed(0, [&] {
                TraceIO io{};
                io.q = q0; io.count_ptr = nullptr; io.first = 0; io.limit = n; io.cursor = &c->cursor_extend[0]; io.flags = sc->trav_flags;
                int grid = sc->full_grid ? Blocks(n) : std::min(Blocks(n), sc->grid_extend);
                if (sc->packet_primary && !count) k_raygen_trace_packet<<<std::max(1, grid), kBlock, 0, st>>>(sc->d, io, c, a);
                else (count ? sc->k_rg_closest_count : sc->k_rg_closest)<<<std::max(1, grid), kBlock, count ? 0 : sc->trace_smem, st>>>(sc->d, io, c, a);
                launches++; n_extend++;
            });
        } else {
            Timed(0, [&] { k_raygen<<<Blocks(n), kBlock, 0, st>>>(a, n, q0, &c->level_count[0]); launches++; });
            Timed(1, [&] { LaunchExtend(q0, &c->level_count[0], 0, n, &c->cursor_extend[0]); });
        }
    }

    // ---- optimistic driver: whole levels, no host round trip.  Level L lives in queue buffer L % 2; its size is only known to
    // the device (level_count[L]); grids are sized by what the level can hold at most.
    void EnqueueWave(unsigned long long base, int n, const int* hint /* 2 * (nlev + 1) ints or NULL */, int32_t* hist) {
        DCounters* c = pp->counters.p;
        const bool fused = sc->fused_raygen && !exact;
        LaunchPrimary(base, n, fused);
        long long upper = n;
        const int nl = std::max(1, sc->d.num_lights);
        for (int lv = 0; lv < nlev; lv++) {
            DRayQueue q = pp->levels[lv % 2].View(), nq = pp->levels[(lv + 1) % 2].View();
            const int32_t* cnt = (lv == 0 && fused) ? nullptr : &c->level_count[lv];
            const int hl = hint ? hint[lv] : -1, hs = hint ? hint[nlev + 1 + lv] : -1;
            if (lv > 0) Timed(1, [&] { LaunchExtend(q, cnt, 0, (int)upper, &c->cursor_extend[lv], hl); });
            Timed(2, [&] { LaunchShade(q, nq, cnt, 0, (int)upper, &c->level_count[lv + 1], &c->shadow_count[lv], hl); });
            if (sc->d.num_lights > 0 || path) {
                long long sup = std::min<long long>((long long)rc.shadow_cap, upper * nl);
                Timed(3, [&] { LaunchShadow(&c->shadow_count[lv], (int)sup, &c->cursor_shadow[lv], hs); });
            }
            upper = std::min<long long>((long long)rc.queue_cap, branching ? upper * 2 : upper);
        }
        k_wave_end<<<1, 256, 0, st>>>(c, nlev, hist);
        launches++;
    }

    // ---- chunked driver (fallback): a level is processed in chunks sized so that neither the next level (<= 2 rays per ray when
    // a dielectric can branch) nor the shadow queue (<= L per ray) can overflow; deeper levels get their own buffers, and the host
    // reads the two queue sizes after every shade pass.  Recursion depth is bounded by the queue level count.
    void ProcessLevel(int lv, int n) {
        DCounters* c = pp->counters.p;
        QueueBufs& qb = pp->levels[lv % nbuf];
        QueueBufs& nb = pp->levels[(lv + 1) % nbuf];
        DRayQueue q = qb.View(), nq = nb.View();
        int nl = std::max(1, sc->d.num_lights);
        long long by_next = branching ? (long long)rc.queue_cap / 2 : rc.queue_cap;
        long long by_shadow = (long long)rc.shadow_cap / nl;
        int chunk = (int)std::max(1LL, std::min(by_next, by_shadow));
        for (int first = 0; first < n; first += chunk) {
            int m = std::min(chunk, n - first);
            CK(cudaMemsetAsync(&c->level_count[lv + 1], 0, sizeof(int32_t), st));
            CK(cudaMemsetAsync(&c->shadow_count[lv], 0, sizeof(int32_t), st));
            CK(cudaMemsetAsync(&c->cursor_extend[lv], 0, sizeof(int32_t), st));
            CK(cudaMemsetAsync(&c->cursor_shadow[lv], 0, sizeof(int32_t), st));
            Timed(1, [&] { LaunchExtend(q, nullptr, first, m, &c->cursor_extend[lv]); });
            Timed(2, [&] { LaunchShade(q, nq, nullptr, first, m, &c->level_count[lv + 1], &c->shadow_count[lv]); });
            CK(cudaMemcpyAsync(pp->h_counters, c, sizeof(DCounters), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            int next_n = pp->h_counters->level_count[lv + 1], shadow_n = pp->h_counters->shadow_count[lv];
            if (pp->h_counters->overflow) throw Invalid{"ray queue overflow (raise queue_capacity)"};
            if (shadow_n > 0) {
                Timed(3, [&] { LaunchShadow(nullptr, shadow_n, &c->cursor_shadow[lv]); });
                shadow += (uint64_t)shadow_n;
            }
            if (next_n > 0 && lv + 1 < nlev) {
                secondary += (uint64_t)next_n;
                ProcessLevel(lv + 1, next_n);
            }
        }
    }
    void ChunkedWave(unsigned long long base, int n) {
        DCounters* c = pp->counters.p;
        CK(cudaMemsetAsync(c, 0, offsetof(DCounters, overflow), st));
        LaunchPrimaryUnfusedOnly(base, n);
        CK(cudaMemcpyAsync(pp->h_counters, c, sizeof(DCounters), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int n0 = pp->h_counters->level_count[0];
        if (n0 > 0) ProcessLevel0(n0);
    }
    void LaunchPrimaryUnfusedOnly(unsigned long long base, int n) {
        RaygenArgs a = rg;
        a.base_rest = (unsigned)(base / (unsigned long long)cam.spp); a.base_sample = (unsigned)(base % (unsigned long long)cam.spp);
        Timed(0, [&] { k_raygen<<<Blocks(n), kBlock, 0, st>>>(a, n, pp->levels[0].View(), &pp->counters.p->level_count[0]); launches++; });
    }
    void ProcessLevel0(int n0) { ProcessLevel(0, n0); }
};

size_t QueueCapacity(const RtRenderOpts* opts) {
    // default 8 Mi rays per wave.  Every launch ends in a tail where the last rays finish on a mostly idle GPU, and the deeper levels
    // of a wave are sparse; bigger waves amortise both.  Round 2, two pipelines, ms per frame at 4 / 8 / 16 Mi (profiles/bench/
    // r02_queue_capacity.txt): C2 34.3 / 31.6 / 31.1, C4 53.8 / 53.4 / 53.5, C5 (32 spp) 140.8 / 137.5 / 135.9.
    size_t cap = opts->queue_capacity > 0 ? (size_t)opts->queue_capacity : ((size_t)1 << 23);
    if (opts->queue_capacity <= 0) if (const char* e = getenv("RT795_QUEUE_CAP")) { long long v = atoll(e); if (v > 0) cap = (size_t)v; }   // tuning override of the default
    return std::min<size_t>(std::max<size_t>(cap, 64), (size_t)1 << 28);
}

// Queue memory: 84 B per ray and level, 48 B per shadow-queue entry, per pipeline.  `nbuf` level buffers of `cap` rays each.  When
// deep recursion asks for many level buffers (chunked driver: one per depth), the capacity shrinks so that the queues stay inside
// a fixed share of the device memory that is free right now - MaxRecursionDepth 50 no longer asks for 37 GB.
void EnsureRenderBuffers(RtScene* sc, const RtRenderOpts* opts, int nbuf, size_t npix3, int npipe) {
    size_t cap = QueueCapacity(opts);
    const size_t lights = (size_t)std::min(8, std::max(2, sc->d.num_lights));
    if (nbuf > 2) {
        size_t free_b = 0, total_b = 0, held = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        for (Pipe& pp : sc->pipes) { for (auto& l : pp.levels) held += l.o_time.n * 84; held += pp.sq_o.n * 48; }
        size_t budget = (free_b + held) / 3;
        size_t per_ray = (size_t)npipe * ((size_t)nbuf * 84 + lights * 48);
        if (cap * per_ray > budget) cap = std::max<size_t>(1024, budget / per_ray);
    }
    size_t scap = lights * cap;   // a shade pass over `cap` rays emits <= num_lights shadow rays per ray
    sc->queue_cap = cap; sc->shadow_cap = scap;
    sc->accum.Alloc(npix3);
    for (int p = 0; p < npipe; p++) {
        Pipe& pp = sc->pipes[p];
        if (!pp.st) CK(cudaStreamCreateWithFlags(&pp.st, cudaStreamNonBlocking));
        if (!pp.done) CK(cudaEventCreateWithFlags(&pp.done, cudaEventDisableTiming));
        if ((int)pp.levels.size() < nbuf) pp.levels.resize(nbuf);
        for (int i = 0; i < nbuf; i++) pp.levels[i].Alloc(cap);
        pp.sq_o.Alloc(scap); pp.sq_d.Alloc(scap); pp.sq_c.Alloc(scap);
        pp.counters.Alloc(1);
        if (!pp.h_counters) CK(cudaMallocHost((void**)&pp.h_counters, sizeof(DCounters)));
    }
}

DTiling MakeTiling(const RtCameraDesc* cam, const RtRenderOpts* opts) {
    DTiling t;
    t.tw = opts->tile_w > 0 ? opts->tile_w : 16;
    t.th = opts->tile_h > 0 ? opts->tile_h : 16;
    if (t.tw % 8 || t.th % 4) throw Invalid{"tile_w must be a multiple of 8 and tile_h of 4"};
    t.tiles_x = (cam->nx + t.tw - 1) / t.tw; t.tiles_y = (cam->ny + t.th - 1) / t.th;
    t.count = opts->tile_count > 1 ? opts->tile_count : 1;
    t.rank = opts->tile_count > 1 ? opts->tile_rank : 0;
    if (t.rank < 0 || t.rank >= t.count) throw Invalid{"tile_rank out of range"};
    int ntiles = t.tiles_x * t.tiles_y;
    t.local_tiles = ntiles > t.rank ? (ntiles - t.rank + t.count - 1) / t.count : 0;
    return t;
}

// Pixels of this rank's tiles that lie inside the image (tiles are cut by the right / bottom edge).
unsigned long long OwnedPixels(const DTiling& t, int nx, int ny) {
    unsigned long long px = 0;
    for (int lt = 0; lt < t.local_tiles; lt++) {
        int k = lt * t.count + t.rank;
        int x0 = (k % t.tiles_x) * t.tw, y0 = (k / t.tiles_x) * t.th;
        px += (unsigned long long)std::max(0, std::min(t.tw, nx - x0)) * (unsigned long long)std::max(0, std::min(t.th, ny - y0));
    }
    return px;
}

void RenderDevice(RtScene* sc, const RtCameraDesc* cam, const RtRenderOpts* opts_in, float* d_out, cudaStream_t user_stream) {
    if (!cam || cam->nx < 1 || cam->ny < 1) throw Invalid{"bad camera resolution"};
    RtRenderOpts opts;
    if (opts_in) opts = *opts_in; else { memset(&opts, 0, sizeof opts); }
    CK(cudaSetDevice(sc->device));
    cudaStream_t st = user_stream ? user_stream : sc->stream;
    Driver proto;
    proto.sc = sc; proto.st = st; proto.pp = nullptr;
    CameraSetup(cam, &proto.cam);
    proto.tl = MakeTiling(cam, &opts);
    proto.count = (opts.collect_counters & 2) != 0;
    proto.stage_time = (opts.collect_counters & 1) != 0;
    bool path = opts.integrator == RT_INTEGRATOR_PATH;
    proto.path = path;
    proto.exact = opts.parity_exact != 0;
    int max_depth = path ? (opts.pt_max_depth > 0 ? opts.pt_max_depth : sc->d.max_depth) : sc->d.max_depth;
    if (max_depth < 0) max_depth = 0;
    if (max_depth > 250) throw Invalid{"recursion depth above 250 is not supported"};
    proto.branching = !path && sc->has_dielectric && max_depth > 0;
    // levels a wave can have: camera rays + one per bounce; a Whitted scene without mirror / conductor / dielectric has one
    proto.nlev = path ? max_depth + 1 : (sc->has_specular ? max_depth + 1 : 1);
    size_t npix3 = (size_t)cam->nx * cam->ny * 3;
    if ((unsigned long long)proto.tl.local_tiles * proto.tl.tw * proto.tl.th >= (1ull << 32)) throw Invalid{"frame too large: more than 2^32 pixel slots per rank"};
    const unsigned long long total = (unsigned long long)proto.tl.local_tiles * proto.tl.tw * proto.tl.th * (unsigned long long)proto.cam.spp;
    const unsigned long long owned = OwnedPixels(proto.tl, cam->nx, cam->ny) * (unsigned long long)proto.cam.spp;

    proto.rc.seed = opts.seed; proto.rc.integrator = opts.integrator; proto.rc.pt_max_depth = max_depth; proto.rc.pt_nee = opts.pt_nee; proto.rc.pt_rr = opts.pt_russian_roulette;
    proto.rc.pt_importance = opts.pt_importance;
    proto.rg.cam = proto.cam; proto.rg.tl = proto.tl; proto.rg.seed = opts.seed; proto.rg.max_depth = max_depth; proto.rg.base_rest = proto.rg.base_sample = 0; proto.rg.center_only = 0;

    const long long hint_key[6] = {cam->nx, cam->ny, proto.cam.spp, opts.integrator * 1000 + max_depth, (long long)proto.tl.rank * 65536 + proto.tl.count,
                                   (long long)opts.queue_capacity * 4 + (opts.pt_nee ? 1 : 0) + (opts.pt_russian_roulette ? 2 : 0)};
    struct FrameResult { int stride = 0, wave = 0; unsigned long long nwaves = 0; uint64_t waves = 0, launches = 0, n_extend = 0, n_shade = 0, n_shadow = 0, secondary = 0, shadow = 0; bool overflow = false; int npipe = 1; double ms[4] = {0, 0, 0, 0}; };

    // One attempt at the frame.  chunked = false: everything is enqueued without a host round trip and the device's overflow flag is
    // read once at the end.
    auto attempt = [&](bool chunked, FrameResult& fr) {
        const size_t cap_hint = QueueCapacity(&opts);
        unsigned long long wave_hint = opts.wave_rays > 0 ? (unsigned long long)opts.wave_rays : (unsigned long long)cap_hint;
        if (!chunked && sc->wave_limit) wave_hint = std::min<unsigned long long>(wave_hint, sc->wave_limit);
        // one pipeline when stages are timed / counted (clean per-stage numbers), in the chunked driver, or when the frame is a single wave
        int npipe = (proto.stage_time || proto.count || chunked || proto.exact) ? 1 : (int)std::min<unsigned long long>((unsigned long long)sc->npipes_wanted, (total + wave_hint - 1) / std::max<unsigned long long>(1, wave_hint));
        npipe = std::max(1, npipe);
        const int nbuf = chunked ? ((proto.branching || sc->d.num_lights > 8) ? proto.nlev + 1 : 2) : 2;
        EnsureRenderBuffers(sc, &opts, nbuf, npix3, npipe);
        proto.nbuf = nbuf;
        proto.rc.queue_cap = (int)sc->queue_cap; proto.rc.shadow_cap = (int)sc->shadow_cap;
        proto.accum = sc->accum.p;
        int wave = (int)std::min<unsigned long long>(wave_hint, (unsigned long long)sc->queue_cap);
        wave = std::max(32, wave / 32 * 32);
        fr = FrameResult();
        fr.npipe = npipe;

        CK(cudaMemsetAsync(sc->accum.p, 0, npix3 * sizeof(float), st));
        CK(cudaEventRecord(sc->ev_start, st));
        Driver drs[kPipes];
        StageEvents sev[kPipes];
        for (int p = 0; p < npipe; p++) {
            Driver& dr = drs[p];
            dr = proto;
            dr.pp = &sc->pipes[p];
            dr.st = dr.pp->st;
            dr.sev = &sev[p];
            dr.sq = DShadowQueue{dr.pp->sq_o.p, dr.pp->sq_d.p, dr.pp->sq_c.p};
            CK(cudaStreamWaitEvent(dr.st, sc->ev_start, 0));          // after the accumulator was cleared on the caller's stream
            CK(cudaMemsetAsync(dr.pp->counters.p, 0, sizeof(DCounters), dr.st));
        }
        const unsigned long long nwaves = (total + (unsigned long long)wave - 1) / (unsigned long long)wave;
        const int stride = 2 * (proto.nlev + 1);
        const size_t per_pipe = (size_t)((nwaves + npipe - 1) / npipe) * stride;
        // hints of the previous frame are valid if it had the same shape (camera, options, wave size)
        const bool hints_ok = !chunked && sc->use_hints && sc->hint_stride == stride && sc->hints.size() == (size_t)nwaves * stride && wave == sc->hint_wave &&
                              !memcmp(hint_key, sc->hint_key, sizeof hint_key);
        if (!chunked)
            for (int p = 0; p < npipe; p++) {
                Pipe& pp = sc->pipes[p];
                pp.hist.Alloc(per_pipe);
                if (pp.h_hist_n < per_pipe) { if (pp.h_hist) cudaFreeHost(pp.h_hist); pp.h_hist = nullptr; CK(cudaMallocHost((void**)&pp.h_hist, per_pipe * sizeof(int32_t))); pp.h_hist_n = per_pipe; }
            }
        unsigned long long w = 0;
        for (unsigned long long base = 0; base < total; base += (unsigned long long)wave, w++) {
            const int p = (int)(w % (unsigned long long)npipe);
            Driver& dr = drs[p];
            int n = (int)std::min<unsigned long long>((unsigned long long)wave, total - base);
            if (chunked) dr.ChunkedWave(base, n);
            else dr.EnqueueWave(base, n, hints_ok ? &sc->hints[(size_t)w * stride] : nullptr, sc->pipes[p].hist.p + (size_t)(w / (unsigned long long)npipe) * stride);
            fr.waves++;
        }
        for (int p = 0; p < npipe; p++) {
            Pipe& pp = sc->pipes[p];
            CK(cudaMemcpyAsync(pp.h_counters, pp.counters.p, sizeof(DCounters), cudaMemcpyDeviceToHost, pp.st));
            if (!chunked && per_pipe) CK(cudaMemcpyAsync(pp.h_hist, pp.hist.p, per_pipe * sizeof(int32_t), cudaMemcpyDeviceToHost, pp.st));
            CK(cudaEventRecord(pp.done, pp.st));
        }
        for (int p = 0; p < npipe; p++) CK(cudaStreamWaitEvent(st, sc->pipes[p].done, 0));
        for (int p = 0; p < npipe; p++) CK(cudaStreamSynchronize(sc->pipes[p].st));   // the one host wait of the frame (counters + overflow flag)
        CK(cudaGetLastError());
        for (int p = 0; p < npipe; p++) {
            const Driver& dr = drs[p];
            const DCounters* hc = sc->pipes[p].h_counters;
            fr.launches += dr.launches; fr.n_extend += dr.n_extend; fr.n_shade += dr.n_shade; fr.n_shadow += dr.n_shadow;
            if (chunked) { fr.secondary += dr.secondary; fr.shadow += dr.shadow; }
            else { fr.secondary += hc->total_secondary; fr.shadow += hc->total_shadow; }
            if (hc->overflow) fr.overflow = true;
            fr.stride = stride; fr.nwaves = nwaves; fr.wave = wave;
            for (int k = 0; k < 4; k++)
                for (auto& e : sev[p].ev[k]) {
                    float ms = 0; CK(cudaEventElapsedTime(&ms, e.first, e.second));
                    fr.ms[k] += ms;
                    cudaEventDestroy(e.first); cudaEventDestroy(e.second);
                }
        }
    };

    {   // what earlier frames learnt about overflows holds for the queue capacity it was learnt with
        size_t cap_now = QueueCapacity(&opts);
        if (cap_now != sc->learnt_cap) { sc->learnt_cap = cap_now; sc->wave_limit = 0; sc->force_chunked = sc->env_chunked; }
    }
    CK(cudaEventRecord(sc->ev0, st));
    FrameResult fr;
    bool chunked = sc->force_chunked;
    for (;;) {
        attempt(chunked, fr);
        if (!fr.overflow) break;
        if (chunked) throw Invalid{"ray queue overflow (raise queue_capacity)"};
        // a level outgrew its queue: redo the frame with half the wave (remembered for the scene's next frames)
        size_t cur = sc->wave_limit ? sc->wave_limit : (size_t)std::min<unsigned long long>(opts.wave_rays > 0 ? (unsigned long long)opts.wave_rays : ~0ull, (unsigned long long)sc->queue_cap);
        sc->wave_limit = cur / 2;
        if (sc->wave_limit < kMinOptimisticWave) { sc->force_chunked = true; chunked = true; }
    }
    if (!chunked) {
        // this frame's per-wave queue sizes become the next frame's hints (valid while camera / options / wave size repeat)
        sc->hints.assign((size_t)fr.nwaves * fr.stride, -1);
        for (unsigned long long w = 0; w < fr.nwaves; w++) {
            const int32_t* src = sc->pipes[w % (unsigned long long)fr.npipe].h_hist + (size_t)(w / (unsigned long long)fr.npipe) * fr.stride;
            for (int k = 0; k < fr.stride; k++) sc->hints[(size_t)w * fr.stride + k] = src[k];
        }
        sc->hint_stride = fr.stride; sc->hint_wave = fr.wave;
        memcpy(sc->hint_key, hint_key, sizeof hint_key);
    } else sc->hints.clear();
    double ms_resolve = 0;
    {
        cudaEvent_t ra = nullptr, rb = nullptr;
        if (proto.stage_time) { CK(cudaEventCreate(&ra)); CK(cudaEventCreate(&rb)); CK(cudaEventRecord(ra, st)); }
        float inv_spp = 1.0f / (float)proto.cam.spp;
        if (opts.resolve_owned_only && proto.tl.count > 1) {
            long long n = (long long)proto.tl.local_tiles * proto.tl.th * (proto.tl.tw >> 2);
            if (n > 0) k_resolve_owned<<<Blocks(n), kBlock, 0, st>>>(sc->accum.p, d_out, proto.tl, cam->nx, cam->ny, inv_spp);
        } else {
            k_resolve<<<Blocks((long long)npix3), kBlock, 0, st>>>(sc->accum.p, d_out, (int)npix3, inv_spp);
        }
        if (proto.stage_time) { CK(cudaEventRecord(rb, st)); CK(cudaEventSynchronize(rb)); float ms = 0; CK(cudaEventElapsedTime(&ms, ra, rb)); ms_resolve = ms; cudaEventDestroy(ra); cudaEventDestroy(rb); }
    }
    CK(cudaEventRecord(sc->ev1, st));
    CK(cudaEventSynchronize(sc->ev1));
    CK(cudaGetLastError());
    float ms = 0; CK(cudaEventElapsedTime(&ms, sc->ev0, sc->ev1));
    RtStats& s = sc->stats;
    s.rays_primary = owned; s.rays_secondary = fr.secondary; s.rays_shadow = fr.shadow; s.launches = fr.launches + 1 /* k_resolve */; s.waves = fr.waves;
    const bool fused_used = sc->fused_raygen && !chunked && !proto.exact;         // K1 ran inside the first closest-hit launch of each wave
    s.ms_raygen = fused_used ? 0.0 : fr.ms[0]; s.ms_extend = fr.ms[1] + (fused_used ? fr.ms[0] : 0.0); s.ms_shade = fr.ms[2]; s.ms_shadow = fr.ms[3];
    s.launches_extend = fr.n_extend; s.launches_shade = fr.n_shade; s.launches_shadow = fr.n_shadow;
    s.node_visits = s.prim_tests = s.node_visits_shadow = s.prim_tests_shadow = 0;
    for (int k = 0; k < 8; k++) s.diag[k] = 0;
    if (proto.count) {
        for (int p = 0; p < fr.npipe; p++) {
            const DCounters* hc = sc->pipes[p].h_counters;
            s.node_visits += hc->node_visits; s.prim_tests += hc->prim_tests;
            s.node_visits_shadow += hc->node_visits_shadow; s.prim_tests_shadow += hc->prim_tests_shadow;
            for (int k = 0; k < 8; k++) s.diag[k] += hc->diag[k];
        }
    }
    s.ms_total = ms; s.ms_resolve = ms_resolve;
}

// Device frame -> the caller's host frame.  The caller's frame is ordinary pageable memory in the reference (new float[],
// src/Image.cpp:8): a direct D2H copy into it goes through the driver's small bounce buffers.  It is staged through the library's
// own pinned buffer in 4 MiB pieces instead, the host memcpy of one piece overlapping the DMA of the next.  A caller that already
// hands over page-locked memory (cudaHostAlloc / cudaHostRegister) gets the plain DMA.
void FrameToHost(RtScene* sc, const float* d_src, float* out_rgb, size_t n) {
    {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, out_rgb) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
            CK(cudaMemcpyAsync(out_rgb, d_src, n * sizeof(float), cudaMemcpyDeviceToHost, sc->stream));
            CK(cudaStreamSynchronize(sc->stream));
            return;
        }
        cudaGetLastError();
    }
    if (sc->h_stage_n < n) {
        if (sc->h_stage) { cudaFreeHost(sc->h_stage); sc->h_stage = nullptr; sc->h_stage_n = 0; }
        CK(cudaMallocHost((void**)&sc->h_stage, n * sizeof(float)));
        sc->h_stage_n = n;
    }
    const size_t piece = (size_t)1 << 20;   // floats
    size_t npieces = (n + piece - 1) / piece;
    std::vector<cudaEvent_t> evs(npieces, nullptr);
    for (size_t k = 0; k < npieces; k++) {
        size_t off = k * piece, cnt = std::min(piece, n - off);
        CK(cudaMemcpyAsync(sc->h_stage + off, d_src + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, sc->stream));
        CK(cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming));
        CK(cudaEventRecord(evs[k], sc->stream));
    }
    cudaError_t first_err = cudaSuccess;
    for (size_t k = 0; k < npieces; k++) {
        size_t off = k * piece, cnt = std::min(piece, n - off);
        cudaError_t e = cudaEventSynchronize(evs[k]);
        cudaEventDestroy(evs[k]);
        if (e != cudaSuccess && first_err == cudaSuccess) first_err = e;
        if (first_err == cudaSuccess) memcpy(out_rgb + off, sc->h_stage + off, cnt * sizeof(float));
    }
    CK(first_err);
}

template <typename F>
int Guard(const char* fn, F&& f) {
    try { f(); return RT_OK; }
    catch (const CudaFail& c) {
        SetLastError(std::string(fn) + ": CUDA error '" + cudaGetErrorString(c.e) + "' at core.cu:" + std::to_string(c.line) + " in " + c.what);
        return c.e == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA;
    } catch (const Invalid& i) {
        SetLastError(std::string(fn) + ": " + i.msg);
        return i.msg.find("overflow") != std::string::npos ? RT_ERR_OVERFLOW : RT_ERR_INVALID;
    } catch (const std::bad_alloc&) {
        SetLastError(std::string(fn) + ": host allocation failed");
        return RT_ERR_OOM;
    } catch (const std::exception& e) {
        SetLastError(std::string(fn) + ": " + e.what());
        return RT_ERR_INVALID;
    }
}

}  // namespace

extern "C" {

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rt_scene_create(const RtSceneDesc* desc, int device, RtScene** out) {
    if (!out) { SetLastError("rt_scene_create: null out"); return RT_ERR_INVALID; }
    *out = nullptr;
    RtScene* sc = nullptr;
    int rcode = Guard("rt_scene_create", [&] {
        Validate(desc);
        int n = 0;
        CK(cudaGetDeviceCount(&n));
        if (device < 0 || device >= n) throw Invalid{"CUDA device " + std::to_string(device) + " not available (" + std::to_string(n) + " devices); there is no CPU fallback"};
        CK(cudaSetDevice(device));
        sc = new RtScene();
        sc->device = device;
        CK(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&sc->ev0)); CK(cudaEventCreate(&sc->ev1)); CK(cudaEventCreateWithFlags(&sc->ev_start, cudaEventDisableTiming));
        if (const char* e = getenv("RT795_PIPES")) sc->npipes_wanted = std::min(kPipes, std::max(1, atoi(e)));
        {
            if (const char* f = getenv("RT795_TRAVERSAL_FLAGS")) sc->trav_flags = sc->trav_flags_shadow = atoi(f);
            if (const char* f = getenv("RT795_FLAGS_CLOSEST")) sc->trav_flags = atoi(f);
            if (const char* f = getenv("RT795_FLAGS_OCCLUSION")) sc->trav_flags_shadow = atoi(f);
            if (const char* g = getenv("RT795_FULL_GRID")) sc->full_grid = atoi(g) != 0;
            if (const char* q = getenv("RT795_QNODES")) sc->quantised_occlusion = atoi(q) != 0;
            if (const char* q = getenv("RT795_DEBUG_BVH")) sc->keep_debug_bvh = atoi(q) != 0;
            if (const char* q = getenv("RT795_SMEM_TLAS")) sc->smem_tlas = atoi(q) != 0;
            if (const char* q = getenv("RT795_FUSED_RAYGEN")) sc->fused_raygen = atoi(q) != 0;
            if (const char* q = getenv("RT795_PACKET")) sc->packet_primary = atoi(q) != 0;
            if (const char* q = getenv("RT795_HINTS")) sc->use_hints = atoi(q) != 0;
            if (const char* q = getenv("RT795_DRIVER")) sc->env_chunked = sc->force_chunked = !strcmp(q, "chunked");
            sc->k_closest_count = k_trace_persistent<false, true, 0>; sc->k_any_count = k_trace_persistent<true, true, 0>;
            sc->k_rg_closest_count = k_raygen_trace<true, 0>;
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, device));
            sc->sm_count = prop.multiProcessorCount;
        }
        if (!BuildAndUpload(sc, desc, true)) BuildAndUpload(sc, desc, false);
        {
            // kernel variants: the top level (TLAS nodes + transform records) is staged in shared memory when it fits
            int stage_bytes = (sc->d.tlas_nodes * 8 + sc->d.num_objects * 4) * 16;
            bool smem = sc->smem_tlas && stage_bytes <= kSmemTlasBytes;
            sc->trace_smem = smem ? stage_bytes : 0;
            if (smem) {
                sc->k_closest = k_trace_persistent<false, false, TF_SMEM_TLAS>; sc->k_any = k_trace_persistent<true, false, TF_SMEM_TLAS>;
                sc->k_rg_closest = k_raygen_trace<false, TF_SMEM_TLAS>;
            } else {
                sc->k_closest = k_trace_persistent<false, false, 0>; sc->k_any = k_trace_persistent<true, false, 0>;
                sc->k_rg_closest = k_raygen_trace<false, 0>;
            }
            if (sc->quantised_occlusion) sc->k_any = k_trace_persistent<true, false, TF_QNODES>;
            int be = 1, bs = 1;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&be, sc->k_closest, kBlock, sc->trace_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bs, sc->k_any, kBlock, sc->quantised_occlusion ? 0 : sc->trace_smem));
            sc->grid_extend = sc->sm_count * std::max(1, be);
            sc->grid_shadow = sc->sm_count * std::max(1, bs);
        }
    });
    if (rcode != RT_OK) { if (sc) rt_scene_destroy(sc); return rcode; }
    *out = sc;
    return RT_OK;
}

void rt_scene_destroy(RtScene* sc) {
    if (!sc) return;
    cudaSetDevice(sc->device);
    if (sc->stream) cudaStreamSynchronize(sc->stream);
    for (Pipe& pp : sc->pipes) {
        if (pp.st) cudaStreamSynchronize(pp.st);
        if (pp.h_counters) cudaFreeHost(pp.h_counters);
        if (pp.h_hist) cudaFreeHost(pp.h_hist);
        if (pp.done) cudaEventDestroy(pp.done);
    }
    if (sc->h_stage) cudaFreeHost(sc->h_stage);
    if (sc->ev_start) cudaEventDestroy(sc->ev_start);
    if (sc->ev0) cudaEventDestroy(sc->ev0);
    if (sc->ev1) cudaEventDestroy(sc->ev1);
    cudaStream_t st = sc->stream;
    cudaStream_t pst[kPipes];
    for (int p = 0; p < kPipes; p++) pst[p] = sc->pipes[p].st;
    delete sc;
    if (st) cudaStreamDestroy(st);
    for (int p = 0; p < kPipes; p++) if (pst[p]) cudaStreamDestroy(pst[p]);
}

int rt_render_device(RtScene* sc, const RtCameraDesc* cam, const RtRenderOpts* opts, float* d_out_rgb, void* cuda_stream) {
    if (!sc || !d_out_rgb) { SetLastError("rt_render_device: null argument"); return RT_ERR_INVALID; }
    return Guard("rt_render_device", [&] { RenderDevice(sc, cam, opts, d_out_rgb, (cudaStream_t)cuda_stream); });
}

int rt_render(RtScene* sc, const RtCameraDesc* cam, const RtRenderOpts* opts, float* out_rgb) {
    if (!sc || !cam || !out_rgb) { SetLastError("rt_render: null argument"); return RT_ERR_INVALID; }
    return Guard("rt_render", [&] {
        CK(cudaSetDevice(sc->device));
        size_t n = (size_t)cam->nx * cam->ny * 3;
        if (opts && opts->resolve_owned_only) throw Invalid{"resolve_owned_only is for rt_render_device (the host frame of rt_render is written whole)"};
        sc->out_dev.Alloc(n);
        RenderDevice(sc, cam, opts, sc->out_dev.p, nullptr);
        FrameToHost(sc, sc->out_dev.p, out_rgb, n);
    });
}

// ---- multi-GPU inside one process: one RtScene, one context and one host thread per device, the frame cut into
// interleaved 16x16 tiles, every device's resolve kernel storing its own tiles straight into the ROOT device's frame over
// NVLink peer memory (k_resolve_owned) - the C ABI counterpart of the reference's 8-thread fan-out inside renderScene
// (src/Scene.cpp:340-356).  A device without peer access to the root resolves into its own memory; its frame is then copied
// to the root and merged by the same kernel.
struct RtSceneGroup {
    std::vector<RtScene*> scenes;
    std::vector<int> devices;
    std::vector<char> peer;                       // device i can store into the root's memory
    DevBuf<float> frame;                          // root device
    DevBuf<float> stage;                          // root device: landing buffer for frames of non-peer devices
    std::vector<float*> local;                    // non-peer devices: their own full-size frame (cudaMalloc on that device)
    size_t frame_n = 0;
    // workers
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    unsigned long long generation = 0;
    int pending = 0;
    bool quit = false;
    const RtCameraDesc* job_cam = nullptr;
    RtRenderOpts job_opts{};
    std::vector<int> rc;
    std::vector<std::string> err;
};

namespace {
void GroupWorker(RtSceneGroup* g, int i) {
    unsigned long long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_go.wait(lk, [&] { return g->quit || g->generation != seen; });
            if (g->quit) return;
            seen = g->generation;
        }
        RtRenderOpts o = g->job_opts;
        o.tile_rank = i; o.tile_count = (int)g->scenes.size(); o.resolve_owned_only = 1;
        float* dst = g->peer[i] ? g->frame.p : g->local[i];
        int rc = Guard("rt_group_render", [&] { RenderDevice(g->scenes[i], g->job_cam, &o, dst, nullptr); });
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->rc[i] = rc;
            if (rc != RT_OK) g->err[i] = rt_last_error();
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}
}  // namespace

int rt_group_create(const RtSceneDesc* desc, int num_devices, const int* devices, RtSceneGroup** out) {
    if (!out) { SetLastError("rt_group_create: null out"); return RT_ERR_INVALID; }
    *out = nullptr;
    if (num_devices < 1 || num_devices > 64) { SetLastError("rt_group_create: num_devices must be in 1..64"); return RT_ERR_INVALID; }
    RtSceneGroup* g = new RtSceneGroup();
    g->devices.resize(num_devices);
    for (int i = 0; i < num_devices; i++) g->devices[i] = devices ? devices[i] : i;
    g->scenes.assign(num_devices, nullptr);
    g->rc.assign(num_devices, RT_OK); g->err.assign(num_devices, std::string());
    // the scene is replicated: build every replica on its own host thread (each device builds its BLASes itself)
    {
        std::vector<std::thread> th;
        for (int i = 0; i < num_devices; i++)
            th.emplace_back([&, i] { g->rc[i] = rt_scene_create(desc, g->devices[i], &g->scenes[i]); if (g->rc[i] != RT_OK) g->err[i] = rt_last_error(); });
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < num_devices; i++)
        if (g->rc[i] != RT_OK) { int rc = g->rc[i]; SetLastError("rt_group_create: device " + std::to_string(g->devices[i]) + ": " + g->err[i]); rt_group_destroy(g); return rc; }
    int rcode = Guard("rt_group_create", [&] {
        const int root = g->devices[0];
        g->peer.assign(num_devices, 0);
        g->local.assign(num_devices, nullptr);
        for (int i = 0; i < num_devices; i++) {
            int dev = g->devices[i];
            if (i == 0) { g->peer[i] = 1; continue; }
            if (getenv("RT795_GROUP_NO_PEER")) continue;                    // test switch: every other replica takes the copy-and-merge path
            if (dev == root) { g->peer[i] = 1; continue; }
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, dev, root));
            if (can) {
                CK(cudaSetDevice(dev));
                cudaError_t e = cudaDeviceEnablePeerAccess(root, 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) { g->peer[i] = 1; cudaGetLastError(); }
                else cudaGetLastError();
            }
        }
        for (int i = 0; i < num_devices; i++) g->workers.emplace_back(GroupWorker, g, i);
    });
    if (rcode != RT_OK) { rt_group_destroy(g); return rcode; }
    *out = g;
    return RT_OK;
}

void rt_group_destroy(RtSceneGroup* g) {
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
    }
    g->cv_go.notify_all();
    for (auto& t : g->workers) t.join();
    for (size_t i = 0; i < g->local.size(); i++) if (g->local[i]) { cudaSetDevice(g->devices[i]); cudaFree(g->local[i]); }
    if (!g->devices.empty()) cudaSetDevice(g->devices[0]);
    g->frame.Free(); g->stage.Free();
    for (RtScene* s : g->scenes) rt_scene_destroy(s);
    delete g;
}

int rt_group_num_devices(const RtSceneGroup* g) { return g ? (int)g->scenes.size() : 0; }
RtScene* rt_group_scene(RtSceneGroup* g, int index) { return (g && index >= 0 && index < (int)g->scenes.size()) ? g->scenes[index] : nullptr; }
int rt_group_uses_peer_memory(const RtSceneGroup* g, int index) { return (g && index >= 0 && index < (int)g->peer.size()) ? g->peer[index] : 0; }

int rt_group_render(RtSceneGroup* g, const RtCameraDesc* cam, const RtRenderOpts* opts, float* out_rgb) {
    if (!g || !cam || !out_rgb) { SetLastError("rt_group_render: null argument"); return RT_ERR_INVALID; }
    if (opts && (opts->tile_count > 1 || opts->resolve_owned_only)) { SetLastError("rt_group_render: the group partitions the frame itself (leave tile_rank / tile_count / resolve_owned_only 0)"); return RT_ERR_INVALID; }
    const int n = (int)g->scenes.size();
    int rcode = Guard("rt_group_render", [&] {
        if (cam->nx < 1 || cam->ny < 1) throw Invalid{"bad camera resolution"};
        const size_t npix3 = (size_t)cam->nx * cam->ny * 3;
        CK(cudaSetDevice(g->devices[0]));
        g->frame.Alloc(npix3);
        for (int i = 0; i < n; i++)
            if (!g->peer[i] && g->frame_n < npix3) {
                CK(cudaSetDevice(g->devices[i]));
                if (g->local[i]) { cudaFree(g->local[i]); g->local[i] = nullptr; }
                CK(cudaMalloc((void**)&g->local[i], npix3 * sizeof(float)));
            }
        g->frame_n = std::max(g->frame_n, npix3);
        CK(cudaSetDevice(g->devices[0]));
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->job_cam = cam;
            if (opts) g->job_opts = *opts; else memset(&g->job_opts, 0, sizeof g->job_opts);
            g->pending = n;
            g->generation++;
        }
        g->cv_go.notify_all();
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_done.wait(lk, [&] { return g->pending == 0; });
        }
        for (int i = 0; i < n; i++) if (g->rc[i] != RT_OK) throw Invalid{"device " + std::to_string(g->devices[i]) + ": " + g->err[i]};
        // every worker returned after its device finished (RenderDevice waits for its last kernel): peer stores have landed.
        RtScene* root = g->scenes[0];
        CK(cudaSetDevice(root->device));
        for (int i = 0; i < n; i++) {
            if (g->peer[i]) continue;
            g->stage.Alloc(npix3);
            // stream-ordered on the root's stream: the merge kernel below must not start before the copy has landed
            CK(cudaMemcpyPeerAsync(g->stage.p, root->device, g->local[i], g->devices[i], npix3 * sizeof(float), root->stream));
            RtRenderOpts o; memset(&o, 0, sizeof o);
            if (opts) o = *opts;
            o.tile_rank = i; o.tile_count = n;
            DTiling tl = MakeTiling(cam, &o);
            long long m = (long long)tl.local_tiles * tl.th * (tl.tw >> 2);
            if (m > 0) k_resolve_owned<<<Blocks(m), kBlock, 0, root->stream>>>(g->stage.p, g->frame.p, tl, cam->nx, cam->ny, 1.0f);
            CK(cudaStreamSynchronize(root->stream));
        }
        FrameToHost(root, g->frame.p, out_rgb, npix3);
    });
    return rcode;
}

int rt_trace_primary(RtScene* sc, const RtCameraDesc* cam, RtHit* out_hits) {
    if (!sc || !cam || !out_hits) { SetLastError("rt_trace_primary: null argument"); return RT_ERR_INVALID; }
    return Guard("rt_trace_primary", [&] {
        // Same kernels and acceleration structures as a frame: fused raygen (pixel centres) + persistent closest-hit traversal of the
        // 4-wide BVH over the affine triangle maps, then one record per pixel.
        CK(cudaSetDevice(sc->device));
        cudaStream_t st = sc->stream;
        RtRenderOpts opts; memset(&opts, 0, sizeof opts);
        size_t npix = (size_t)cam->nx * cam->ny;
        opts.queue_capacity = (int)std::max<size_t>(sc->queue_cap, 1 << 20);
        EnsureRenderBuffers(sc, &opts, 2, npix * 3, 1);
        Pipe* pp = &sc->pipes[0];
        sc->hits_dev.Alloc(npix);
        CK(cudaMemsetAsync(sc->hits_dev.p, 0, npix * sizeof(RtHit), st));
        RaygenArgs rg; memset(&rg, 0, sizeof rg);
        CameraSetup(cam, &rg.cam);
        rg.tl = MakeTiling(cam, &opts);
        rg.center_only = 1;
        int wave = (int)sc->queue_cap / 32 * 32;
        unsigned long long total = (unsigned long long)rg.tl.local_tiles * rg.tl.tw * rg.tl.th;
        for (unsigned long long base = 0; base < total; base += (unsigned long long)wave) {
            int n = (int)std::min<unsigned long long>((unsigned long long)wave, total - base);
            CK(cudaMemsetAsync(pp->counters.p, 0, sizeof(DCounters), st));
            rg.base_rest = (unsigned)base; rg.base_sample = 0;
            TraceIO io{};
            io.q = pp->levels[0].View(); io.count_ptr = nullptr; io.first = 0; io.limit = n; io.cursor = &pp->counters.p->cursor_extend[0]; io.flags = sc->trav_flags;
            sc->k_rg_closest<<<std::max(1, std::min(Blocks(n), sc->grid_extend)), kBlock, sc->trace_smem, st>>>(sc->d, io, pp->counters.p, rg);
            k_primary_hits<<<Blocks(n), kBlock, 0, st>>>(sc->d, pp->levels[0].View(), n, sc->hits_dev.p);
        }
        CK(cudaMemcpyAsync(out_hits, sc->hits_dev.p, npix * sizeof(RtHit), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
    });
}

int rt_tonemap(RtScene* sc, const float* in_rgb, int nx, int ny, float key, float burn_percent, float saturation, float gamma, uint8_t* out_rgb8) {
    if (!sc || !in_rgb || !out_rgb8 || nx < 1 || ny < 1) { SetLastError("rt_tonemap: bad argument"); return RT_ERR_INVALID; }
    return Guard("rt_tonemap", [&] {
        CK(cudaSetDevice(sc->device));
        cudaStream_t st = sc->stream;
        int npix = nx * ny;
        sc->tm_in.Alloc((size_t)npix * 3); sc->tm_out.Alloc((size_t)npix * 3);
        sc->tm_sum.Alloc(1); sc->tm_max.Alloc(1); sc->tm_hist.Alloc(kTmBins);
        CK(cudaMemcpyAsync(sc->tm_in.p, in_rgb, (size_t)npix * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(sc->tm_sum.p, 0, sizeof(double), st));
        CK(cudaMemsetAsync(sc->tm_max.p, 0, sizeof(float), st));
        CK(cudaMemsetAsync(sc->tm_hist.p, 0, kTmBins * sizeof(unsigned), st));
        int grid = std::min(Blocks(npix), 148 * 8);
        k_tonemap_logsum<<<grid, kBlock, 0, st>>>(sc->tm_in.p, npix, sc->tm_sum.p, sc->tm_max.p);
        double sum = 0; float mx = 0;
        CK(cudaMemcpyAsync(&sum, sc->tm_sum.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&mx, sc->tm_max.p, sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        float lavg = (float)std::exp(sum / npix);
        float scale = key / std::max(lavg, 1e-20f);
        float max_scaled = std::max(mx * scale, 1e-20f);
        float lwhite = max_scaled;
        if (burn_percent > 0.f) {
            k_tonemap_hist<<<grid, kBlock, 0, st>>>(sc->tm_in.p, npix, scale, max_scaled, sc->tm_hist.p);
            std::vector<unsigned> hist(kTmBins);
            CK(cudaMemcpyAsync(hist.data(), sc->tm_hist.p, kTmBins * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            // luminance at the (100 - burn)% percentile becomes white
            unsigned long long target = (unsigned long long)((double)npix * (100.0 - burn_percent) / 100.0), acc = 0;
            int b = 0;
            for (; b < kTmBins; b++) { acc += hist[b]; if (acc >= target) break; }
            lwhite = std::max(1e-20f, (b + 0.5f) / (kTmBins - 1) * max_scaled);
        }
        k_tonemap_map<<<Blocks(npix), kBlock, 0, st>>>(sc->tm_in.p, npix, scale, lwhite, saturation, 1.0f / std::max(gamma, 1e-6f), sc->tm_out.p);
        CK(cudaMemcpyAsync(out_rgb8, sc->tm_out.p, (size_t)npix * 3, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
    });
}

void* rt_host_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); SetLastError("rt_host_alloc_pinned: cudaMallocHost failed"); return nullptr; }
    return p;
}
void rt_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int rt_get_stats(RtScene* sc, RtStats* out) {
    if (!sc || !out) { SetLastError("rt_get_stats: null argument"); return RT_ERR_INVALID; }
    *out = sc->stats;
    return RT_OK;
}

int rt_debug_read_bvh(RtScene* sc, int what, void* out, uint64_t capacity_bytes, uint64_t* size_bytes, int32_t* tlas_root) {
    if (!sc || !size_bytes) { SetLastError("rt_debug_read_bvh: null argument"); return RT_ERR_INVALID; }
    return Guard("rt_debug_read_bvh", [&] {
        CK(cudaSetDevice(sc->device));
        const void* src = nullptr; uint64_t bytes = 0; int32_t root = 0;
        switch (what) {
            case RT_DEBUG_BVH2:
                if (!sc->keep_debug_bvh) throw Invalid{"the binary BVH is not resident: set RT795_DEBUG_BVH=1 before rt_scene_create"};
                src = sc->nodes2_debug.p; bytes = (uint64_t)(sc->stats.bvh2_nodes) * 64; root = sc->tlas_base2; break;
            case RT_DEBUG_BVH4: src = sc->nodes4.p; bytes = sc->stats.bvh_nodes * 128; root = sc->d.tlas_root4; break;
            case RT_DEBUG_BVH4Q:
                if (!sc->nodes4q.p) throw Invalid{"the quantised nodes are not resident: set RT795_DEBUG_BVH=1 (or RT795_QNODES=1) before rt_scene_create"};
                src = sc->nodes4q.p; bytes = sc->stats.bvh_nodes * 64; root = sc->d.tlas_root4; break;
            case RT_DEBUG_TRIS: src = sc->tris.p; bytes = sc->stats.tri_bytes; break;
            case RT_DEBUG_OBJ_ROOT4: src = sc->obj_root4.p; bytes = (uint64_t)std::max(1, sc->d.num_objects) * 4; break;
            default: throw Invalid{"rt_debug_read_bvh: unknown array"};
        }
        *size_bytes = bytes;
        if (tlas_root) *tlas_root = root;
        if (!out) return;
        if (capacity_bytes < bytes) throw Invalid{"rt_debug_read_bvh: buffer too small"};
        CK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, sc->stream));
        CK(cudaStreamSynchronize(sc->stream));
    });
}

}  // extern "C"
