/*
 * rt795.h - C ABI of the B200-native rendering core for the CENG795 ray tracer hot path.
 *
 * The reference (badiba/raytracer-795) has NO plugin/FFI boundary: its hot path is the block
 * src/Scene.cpp:336-356 inside Scene::renderScene() - "given the prepared scene, fill Image::_data for one
 * camera" - reached only through private C++ methods and the global `Scene* pScene` (src/defs.h:34).
 * This header is the seam a maintainer would cut there: plain C structs that carry the
 * flattened output of the reference's scene preparation (src/Scene.cpp:296-323) and one render call that
 * fills a float RGB buffer in Image::_data layout (src/Image.cpp:17-24).  INTEGRATION.md shows the
 * reference-side binding.
 *
 * Conventions
 *   - C99-clean, no C++ or torch types; all indices 0-based unless a field says "1-based (reference)".
 *   - Every function returns RT_OK (0) or a negative RtStatus; rt_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing aborts or throws across this boundary.
 *   - The caller owns every input array (only read during the call that receives it) and every output
 *     buffer.  The library owns device memory behind the opaque RtScene handle.
 *   - One host thread per RtScene handle at a time.
 *   - There is NO CPU fallback: if no CUDA device is usable, rt_scene_create / rt_render fail with
 *     RT_ERR_CUDA.
 */
#ifndef RT795_H
#define RT795_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT795_ABI_VERSION 2

typedef enum RtStatus {
    RT_OK = 0,
    RT_ERR_INVALID = -1,   /* bad argument / inconsistent scene description            */
    RT_ERR_CUDA = -2,      /* CUDA runtime failure (message holds cudaGetErrorString)    */
    RT_ERR_OOM = -3,       /* device or host allocation failed                           */
    RT_ERR_IO = -4,        /* host shell: file could not be read / written               */
    RT_ERR_PARSE = -5,     /* host shell: malformed XML / PLY / EXR / PNG / JPEG / blob   */
    RT_ERR_OVERFLOW = -6   /* a wavefront queue would overflow (never silently dropped)  */
} RtStatus;

/* ---- enums mirror the reference's, value for value ---------------------------------------------------- */
/* MaterialType, src/Material.h:7 */
enum { RT_MAT_NORMAL = 0, RT_MAT_MIRROR = 1, RT_MAT_CONDUCTOR = 2, RT_MAT_DIELECTRIC = 3 };
/* BRDFType, src/Material.h:8 */
enum { RT_BRDF_NONE = 0, RT_BRDF_OBP = 1, RT_BRDF_MBP = 2, RT_BRDF_MBPN = 3, RT_BRDF_OP = 4, RT_BRDF_MP = 5,
       RT_BRDF_MPN = 6, RT_BRDF_TS = 7, RT_BRDF_TSF = 8 };
/* DecalMode, src/defs.h:8 */
enum { RT_DM_REPLACE_KD = 0, RT_DM_BLEND_KD = 1, RT_DM_BUMP_NORMAL = 2, RT_DM_REPLACE_NORMAL = 3,
       RT_DM_REPLACE_ALL = 4, RT_DM_REPLACE_BACKGROUND = 5, RT_DM_NONE = 6 };
/* Interpolation, src/defs.h:9 */
enum { RT_INTERP_NN = 0, RT_INTERP_BILINEAR = 1 };
/* TextureType, src/defs.h:10 */
enum { RT_TEX_IMAGE = 0, RT_TEX_PERLIN = 1 };
/* NoiseConversion, src/defs.h:11 */
enum { RT_NC_ABSVAL = 0, RT_NC_LINEAR = 1, RT_NC_NONE = 2 };
/* LightType, src/Light.h:12 */
enum { RT_LIGHT_POINT = 0, RT_LIGHT_AREA = 1, RT_LIGHT_DIRECTIONAL = 2, RT_LIGHT_SPOT = 3, RT_LIGHT_ENVIRONMENT = 4,
       /* object lights (course HW7, pages/Page7.md:9-13; no code in the reference snapshot): a top-level object that emits.  Rays that
        * hit it see its radiance; next-event estimation samples its surface: a mesh by triangle area then uniformly on the triangle,
        * a sphere uniformly over the cone of directions it subtends */
       RT_LIGHT_MESH = 5, RT_LIGHT_SPHERE = 6 };
/* top-level entry kinds */
enum { RT_OBJ_SPHERE = 0, RT_OBJ_MESH = 1 };
/* image pixel formats (what the reference's texture loaders produce, src/Texture.cpp:41-74) */
enum { RT_IMG_RGBA8 = 0, RT_IMG_RGBA32F = 1 };
/* integrators */
enum { RT_INTEGRATOR_REFERENCE = 0,   /* the reference's Whitted / distributed ray tracer (src/Scene.cpp:148-267) */
       RT_INTEGRATOR_PATH = 1 };      /* path tracer with NEE + Russian roulette (no reference code; pages/Page7.md) */

/* Material, src/Material.h:10-33.  Addressed by POSITION (materials[matIndex-1]); the XML id is ignored. */
typedef struct RtMaterial {
    int32_t type;            /* RT_MAT_*                                            */
    int32_t brdf;            /* RT_BRDF_*                                           */
    int32_t phong_exp;       /* int in the reference (Material.h:14)                */
    int32_t is_rough;        /* <Roughness> present (Parser.h:369-376)              */
    float ambient[3];
    float diffuse[3];
    float specular[3];
    float mirror[3];
    float refraction_index;
    float absorption_index;
    float roughness;
    float absorption_coeff[3];
    float emission[3];       /* extension (object lights for RT_INTEGRATOR_PATH); 0 for reference scenes */
} RtMaterial;

/* Texture map, src/Texture.h:15-31 (+ image binding). */
typedef struct RtTexture {
    int32_t type;            /* RT_TEX_*                                            */
    int32_t decal_mode;      /* RT_DM_*                                             */
    int32_t interpolation;   /* RT_INTERP_*                                         */
    int32_t noise_conversion;/* RT_NC_*                                             */
    int32_t normalizer;      /* int in the reference (Texture.h:30)                 */
    float noise_scale;
    float bump_factor;
    int32_t image;           /* 0-based index into RtSceneDesc.images, -1 for Perlin */
} RtTexture;

typedef struct RtImage {
    int32_t width, height;
    int32_t format;          /* RT_IMG_RGBA8: 4 bytes/pixel; RT_IMG_RGBA32F: 4 floats/pixel; row-major, top row first */
    const void* pixels;
} RtImage;

/* A triangle set with one BVH (the reference builds one BVH per object, src/Scene.cpp:321-323; mesh instances
 * share their base mesh's, src/Helper.cpp:58).  A standalone <Triangle> is a 1-triangle mesh with smooth=1
 * (it is ALWAYS smooth-shaded, src/Shape.cpp:262-276 + src/Scene.cpp:311-313). */
typedef struct RtMesh {
    int32_t first_tri;       /* into RtSceneDesc.tri_indices (triangle units)       */
    int32_t tri_count;
    int32_t smooth;          /* Triangle::isSmooth of its faces                     */
} RtMesh;

typedef struct RtSphere {
    float center[3];         /* pScene->vertices[cIndex-1], object space            */
    float radius;
} RtSphere;

/* One top-level entry, in the reference's loop order: objects (spheres, triangles, meshes; parse order), then
 * mesh instances (src/Helper.cpp:32-73).  Matrices are what ComputeObjectTransformations stores
 * (src/Helper.cpp:178-182, 220-224), row-major here. */
typedef struct RtObject {
    int32_t kind;            /* RT_OBJ_*                                            */
    int32_t geom;            /* index into spheres[] or meshes[]                    */
    int32_t mat_index;       /* 1-based (reference); the instance's own for instances (Helper.cpp:68) */
    int32_t tex[2];          /* 1-based texture ids (reference), 0 = none; instances inherit the base mesh's */
    int32_t tex_offset;      /* textureOffset - vertexOffset (Parser.h:1102,1147): uv index = vertex index + tex_offset */
    float inv[12];           /* rows 0..2 of M^-1 (3x4)                             */
    float inv_t[9];          /* upper 3x3 of (M^-1)^T, row-major (normal matrix)    */
    float blur[3];           /* MotionBlur translation (world units per unit time)  */
    float radiance[3];       /* object lights (<LightMesh> / <LightSphere> <Radiance>): emitted radiance, front side; 0 0 0 = not a light */
} RtObject;

/* Light, src/Light.h + constructors in src/Light.cpp.  Fields unused by a type are ignored. */
typedef struct RtLight {
    int32_t type;            /* RT_LIGHT_*                                          */
    float position[3];       /* point, spot, area                                   */
    float direction[3];      /* directional, spot: direction; area: normal (normalised by the library as the ctors do) */
    float intensity[3];      /* point/spot intensity; directional/area radiance     */
    float size;              /* area: edge length                                   */
    float coverage_deg;      /* spot: CoverageAngle (full angle, degrees)           */
    float falloff_deg;       /* spot: FalloffAngle (full angle, degrees)            */
    int32_t image;           /* environment: 0-based index into images[]            */
    int32_t object;          /* mesh / sphere light: index of the emitting top-level entry in objects[] (its radiance = `intensity`) */
} RtLight;

/* Per-mesh acceleration-structure builder (csrc/bvh_device.cu, csrc/bvh_build.cpp).  All produce the same node formats; any
 * conservative BVH returns the same closest hit up to ties.
 *   RT_BVH_HOST_SAH     binned SAH (16 bins x 3 axes) on the host cores.
 *   RT_BVH_DEVICE_SAH   the same algorithm run level-synchronously on the GPU: the same tree, ~10x sooner.
 *   RT_BVH_DEVICE_PLOC  Morton sort + parallel locally-ordered clustering + SAH top: ~4x sooner again, ~2.5 % slower to trace.
 *   RT_BVH_AUTO         DEVICE_SAH for meshes of >= 32768 triangles, HOST_SAH for smaller ones. */
enum { RT_BVH_AUTO = 0, RT_BVH_HOST_SAH = 1, RT_BVH_DEVICE_PLOC = 2, RT_BVH_DEVICE_SAH = 3 };

typedef struct RtSceneDesc {
    int32_t abi_version;     /* RT795_ABI_VERSION                                   */
    /* scene scalars, src/Scene.h:60-67, defaults src/Parser.h:23-25 */
    int32_t max_recursion_depth;
    float shadow_ray_eps;
    float int_test_eps;
    float background_color[3];
    float ambient_light[3];
    int32_t env_light_index;     /* index into lights[] or -1 (Parser.h:1297-1311)  */
    int32_t background_texture;  /* 0-based index into textures[] or -1 (Scene.cpp:494-500) */

    /* global vertex data (object space), src/Scene.h:77-82 */
    int32_t num_vertices;
    const float* vertices;        /* 3 per vertex                                   */
    const float* vertex_normals;  /* 3 per vertex, normalised smooth normals (Scene.cpp:303-318); may be NULL if no mesh is smooth */
    int32_t num_texcoords;
    const float* texcoords;       /* 2 per entry; may be NULL                       */

    int32_t num_triangles;
    const int32_t* tri_indices;   /* 3 per triangle, 0-based vertex indices         */

    int32_t num_meshes;   const RtMesh* meshes;
    int32_t num_spheres;  const RtSphere* spheres;
    int32_t num_objects;  const RtObject* objects;
    int32_t num_materials; const RtMaterial* materials;
    int32_t num_textures; const RtTexture* textures;
    int32_t num_images;   const RtImage* images;
    int32_t num_lights;   const RtLight* lights;

    int32_t bvh_builder;     /* RT_BVH_*: who builds the per-mesh BVH (replaces BVH::ConstructionHelper, src/BVH.cpp:64-135) */
} RtSceneDesc;

/* Camera, the arguments of Camera::Camera (src/Camera.cpp:7-61) after Parser::ParseCameras resolved
 * GazePoint / FovY (src/Parser.h:52-164). */
typedef struct RtCameraDesc {
    float position[3];
    float gaze[3];
    float up[3];
    float left, right, bottom, top;   /* image plane, src/Camera.h:8-17 */
    float distance;
    int32_t nx, ny;
    int32_t num_samples;      /* <NumSamples>; 1 -> SingleSample path (Scene.cpp:365-384) */
    float focus_distance;
    float aperture_size;
    int32_t is_dof;
    int32_t left_handed;
} RtCameraDesc;

typedef struct RtRenderOpts {
    int32_t integrator;       /* RT_INTEGRATOR_*                                    */
    uint64_t seed;            /* counter-based RNG seed (Philox4x32-10 keyed by seed; counter = pixel, sample, path node, dimension) */
    /* image partition across GPUs: the frame is cut into tile_w x tile_h tiles, tile k (row-major) belongs to
     * rank k % tile_count.  tile_count <= 1 renders the whole frame. */
    int32_t tile_rank, tile_count;
    int32_t tile_w, tile_h;   /* 0 -> default 16x16; tile_w must be a multiple of 8, tile_h of 4 */
    /* path tracer only */
    int32_t pt_max_depth;     /* 0 -> scene max_recursion_depth                     */
    int32_t pt_nee;           /* next-event estimation on/off                       */
    int32_t pt_russian_roulette; /* 0 off; survival probability keyed on 1 the path throughput, 2 the cosine between the sampled
                                    direction and the normal (the reference author's choice, pages/Page7.md:31); both unbiased */
    int32_t pt_importance;    /* 0 uniform hemisphere, 1 cosine-weighted            */
    /* engine knobs (0 -> defaults) */
    int32_t wave_rays;        /* primary rays generated per wave                    */
    int32_t queue_capacity;   /* entries per ray-queue level                        */
    int32_t collect_counters; /* bit 0: per-stage CUDA-event timing into RtStats.ms_* (serialises stages);
                                 bit 1: per-ray node-visit / primitive-test counters into RtStats (slower kernels) */
    int32_t resolve_owned_only; /* tile_count > 1 only.  0: the resolve writes the whole frame (pixels of other ranks' tiles = 0, so
                                 a SUM reduce gathers exactly).  1: the resolve stores ONLY this rank's tiles and leaves every other
                                 pixel of `out` untouched - `out` may then be a peer (NVLink) mapping of the root GPU's
                                 framebuffer and the resolve kernel itself is the gather (rt_render_device only) */
    int32_t parity_exact;     /* 1: every closest-hit query (camera, reflection, refraction AND shadow rays) follows the reference's
                                 per-object rules the fast path drops: inside an object the nearest accepted
                                 primitive by |t| wins, triangles are accepted from t >= -eps, the object counts only if Ray::gett
                                 (x-quotient, src/Ray.cpp:21-36) of that point lies in (0, best) - a hit just behind the origin masks the
                                 whole object - and shadow rays are full closest-hit queries (src/Light.cpp:188-204).  Slower (no any-hit
                                 early out, one pipeline); a debug / parity mode.  Ties still resolve in traversal order. */
} RtRenderOpts;

/* Primary-hit record for parity/debug (rt_trace_primary); mirrors what the oracle hit harness dumps. */
typedef struct RtHit {
    int32_t full;             /* 1 = hit                                            */
    int32_t object;           /* top-level entry index (reference loop order)       */
    int32_t primitive;        /* triangle index within its mesh (face order), 0 for spheres */
    float t;
    float point[3];           /* world                                              */
    float normal[3];          /* world, after texture/bump + (M^-1)^T, normalised   */
} RtHit;

typedef struct RtStats {
    uint64_t rays_primary, rays_secondary, rays_shadow;   /* queries issued in the last render */
    uint64_t launches;                                    /* kernels launched in the last render */
    uint64_t waves;
    uint64_t node_visits, prim_tests;                     /* closest-hit queries; only with collect_counters bit 1 */
    uint64_t node_visits_shadow, prim_tests_shadow;       /* occlusion queries; only with collect_counters bit 1 */
    uint64_t diag[8];         /* collect_counters bit 1, SIMT diagnostics of the traversal kernels (closest-hit [0..3], occlusion [4..7]):
                                 node-step warp executions, sum of lanes holding an unfinished ray at those, leaf-phase executions, lanes in them */
    double ms_total;          /* first launch -> framebuffer ready (CUDA events) */
    double ms_extend, ms_shade, ms_shadow, ms_raygen, ms_resolve;  /* per-stage sums, CUDA events (only with collect_counters bit 0) */
    uint64_t launches_extend, launches_shade, launches_shadow;     /* launches per stage in the last render */
    uint64_t bvh_nodes, bvh_bytes, tri_bytes;             /* scene footprint in HBM */
    double ms_bvh_build, ms_upload;                       /* rt_scene_create: BVH build (host + device, wall clock) / upload */
    double ms_bvh_device;                                 /* device time of the GPU BLAS builds inside ms_bvh_build (CUDA events) */
    int32_t bvh_device_meshes;                            /* meshes whose BLAS was built on the GPU */
    int32_t bvh_ploc_iterations;                          /* clustering iterations of the last GPU build */
    uint64_t bvh2_nodes;                                  /* binary twin of the BVH (rt_trace_primary, RT795_TRAVERSAL=classic) */
} RtStats;

typedef struct RtScene RtScene;   /* opaque */

/* ---- core: replaces src/Scene.cpp:336-356 (+ BVH build src/Scene.cpp:321-323) --------------------------- */

/* Copies + flattens the description, builds the two-level SAH BVH on the host (replaces BVH::BVH /
 * ConstructionHelper, src/BVH.cpp:53-110) and uploads everything to `device` (cuda ordinal). */
int rt_scene_create(const RtSceneDesc* desc, int device, RtScene** out);
void rt_scene_destroy(RtScene* scene);

/* Renders one camera into HOST memory: out_rgb[(row*nx + col)*3 + c], float32, pre-tone-map radiance
 * (Image::_data, src/Image.cpp:17-24).  Pixels of tiles owned by other ranks are written as 0.
 * Replaces the thread spawn/join block src/Scene.cpp:336-356 (ThreadedRendering :269-292). */
int rt_render(RtScene* scene, const RtCameraDesc* cam, const RtRenderOpts* opts, float* out_rgb);

/* Same, but the destination is DEVICE memory on the scene's device (nx*ny*3 floats) and the frame is ordered on
 * `cuda_stream` (a cudaStream_t passed as void*): it starts after the work already enqueued there and its last kernel
 * (the resolve) runs on it.  NULL = the library's own stream; to name the legacy default stream pass cudaStreamLegacy.
 * The call returns when the frame is complete. */
int rt_render_device(RtScene* scene, const RtCameraDesc* cam, const RtRenderOpts* opts, float* d_out_rgb, void* cuda_stream);

/* Primary visibility only (ray generation at pixel centres + closest hit), for parity tests:
 * Camera::getPrimaryRay (src/Camera.cpp:63-72) + BVHMethods::FindIntersection (src/Helper.cpp:18-80). */
int rt_trace_primary(RtScene* scene, const RtCameraDesc* cam, RtHit* out_hits);

/* Photographic (Reinhard global) tone map: HOST float RGB in -> 8-bit RGB out.  No reference code exists
 * (pages/Page5.md:47-53); operator documented in DESIGN.md. */
int rt_tonemap(RtScene* scene, const float* in_rgb, int nx, int ny, float key, float burn_percent, float saturation,
               float gamma, uint8_t* out_rgb8);

int rt_get_stats(RtScene* scene, RtStats* out);

/* Page-locked host memory for frames (cudaMallocHost / cudaFreeHost) for callers that do not link the CUDA runtime themselves:
 * rt_render / rt_group_render copy a finished frame straight into such a buffer by DMA; an ordinary malloc'ed frame (what the
 * reference's Image::_data is, src/Image.cpp:8) is staged through the library's own pinned buffer + a host memcpy (~2.5 ms for a
 * 1920x1080 frame).  NULL on failure. */
void* rt_host_alloc_pinned(uint64_t bytes);
void rt_host_free_pinned(void* p);

/* ---- multi-GPU inside ONE process: the counterpart of the reference's fan-out inside renderScene (8 std::threads over one
 * frame, src/Scene.cpp:340-356).  The scene is replicated on `num_devices` CUDA devices (devices[i], NULL = 0..n-1; devices[0]
 * is the root), one context + one host thread per device.  rt_group_render cuts the frame into interleaved tiles
 * (tile k -> device k % n), every device's resolve kernel stores its own tiles straight into the ROOT device's frame over
 * NVLink peer memory (cudaDeviceEnablePeerAccess; a device without peer access to the root is merged by a copy), and the
 * finished frame is copied to `out_rgb` (host, Image::_data layout) - same result as rt_render on one device, up to the order
 * of float additions.  opts: as rt_render; tile_rank / tile_count / resolve_owned_only must be 0. */
typedef struct RtSceneGroup RtSceneGroup;   /* opaque */
int rt_group_create(const RtSceneDesc* desc, int num_devices, const int* devices, RtSceneGroup** out);
void rt_group_destroy(RtSceneGroup* group);
int rt_group_render(RtSceneGroup* group, const RtCameraDesc* cam, const RtRenderOpts* opts, float* out_rgb);
int rt_group_num_devices(const RtSceneGroup* group);
/* The replica on devices[index] (owned by the group): for rt_get_stats / rt_tonemap. */
RtScene* rt_group_scene(RtSceneGroup* group, int index);
/* 1 if devices[index] stores its tiles into the root's frame directly (peer memory), 0 if it is merged by a copy. */
int rt_group_uses_peer_memory(const RtSceneGroup* group, int index);

/* Test / diagnostic aid: copies one of the scene's acceleration-structure arrays back to the host so a checker can walk it
 * (tests/test_gpu_bvh_device.py validates the device-built trees this way).  `out` may be NULL to query the size.  Formats:
 * csrc/bvh_build.h (BvhNode64, BvhNode128, BvhNode64Q; triangles as 3 x float4 {v0|global tri id, e1, e2}). */
enum { RT_DEBUG_BVH2 = 0, RT_DEBUG_BVH4 = 1, RT_DEBUG_BVH4Q = 2, RT_DEBUG_TRIS = 3, RT_DEBUG_OBJ_ROOT4 = 4 };
int rt_debug_read_bvh(RtScene* scene, int what, void* out, uint64_t capacity_bytes, uint64_t* size_bytes, int32_t* tlas_root);
const char* rt_last_error(void);
int rt_abi_version(void);
/* Number of usable CUDA devices (0 if none / driver missing); never fails. */
int rt_device_count(void);

/* ---- host shell: the reference's scene I/O surface, re-implemented (src/Parser.h, src/Scene.cpp:455-504,
 *      src/Helper.cpp:135-226, src/Scene.cpp:296-318, src/Image.cpp:26-107) ------------------------------- */
typedef struct RtHostScene RtHostScene;   /* opaque; owns the arrays behind the RtSceneDesc it exposes */

/* Parses a scene XML (+ PLY meshes, EXR/PNG images next to it), composes object transforms, computes smooth
 * normals, and flattens to an RtSceneDesc.  Pure host code (no CUDA). */
int rt_host_load_scene(const char* xml_path, RtHostScene** out);
void rt_host_free_scene(RtHostScene* hs);
/* Scene blob: the prepared scene (everything rt_host_load_scene produces, cameras and image names included) as one binary
 * file, so later runs skip the text parsing the reference does on every start (src/Parser.h:684-727, 1109-1143).  The blob
 * remembers size + mtime of the XML and of every PLY / image it was read from.
 *   rt_host_load_scene_blob: check_sources != 0 rejects (RT_ERR_PARSE, "stale") a blob whose sources changed or vanished.
 *   rt_host_load_scene_cached: loads `blob_path` (NULL -> "<xml_path>.rt795blob") when it is a fresh blob of xml_path, else
 *   parses xml_path and (best effort) writes the blob; *cache_hit (may be NULL) says which happened. */
int rt_host_save_scene_blob(const RtHostScene* hs, const char* blob_path);
int rt_host_load_scene_blob(const char* blob_path, int check_sources, RtHostScene** out);
int rt_host_load_scene_cached(const char* xml_path, const char* blob_path, RtHostScene** out, int* cache_hit);
const RtSceneDesc* rt_host_scene_desc(const RtHostScene* hs);
int rt_host_num_cameras(const RtHostScene* hs);
const RtCameraDesc* rt_host_camera(const RtHostScene* hs, int i);
const char* rt_host_camera_image_name(const RtHostScene* hs, int i);
/* Per-camera extensions beyond src/Parser.h, in the course's XML vocabulary (SURVEY App. B lists them as not parsed by the
 * reference snapshot): <Renderer>PathTracing</Renderer> + <RendererParams>NextEventEstimation RussianRoulette
 * ImportanceSampling</RendererParams> select RT_INTEGRATOR_PATH and its options (without <Renderer> the options describe the
 * reference integrator); <Tonemap><TMO>Photographic</TMO><TMOOptions>key burn%</TMOOptions><Saturation/><Gamma/></Tonemap>
 * asks for a tone-mapped 8-bit image next to the HDR one (rt_tonemap's parameters). */
typedef struct RtTonemap { int32_t enabled; float key, burn_percent, saturation, gamma; } RtTonemap;
int rt_host_camera_render_opts(const RtHostScene* hs, int i, RtRenderOpts* out);
int rt_host_camera_tonemap(const RtHostScene* hs, int i, RtTonemap* out);
/* Writers with the reference's file semantics: names containing ".png" get a clamped P3 ASCII PPM
 * (src/Image.cpp:62-103), anything else a half-float B,G,R EXR (src/Helper.cpp:361-412). */
int rt_host_save_image(const char* image_name, const float* rgb, int nx, int ny);
/* Raw float dump used by the tests ("<name>.f32": int32 nx, ny, then floats) - same format as the oracle harness. */
int rt_host_save_f32(const char* path, const float* rgb, int nx, int ny);

#ifdef __cplusplus
}
#endif
#endif /* RT795_H */
