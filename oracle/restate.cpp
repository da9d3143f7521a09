// CPU restatement of the reference's hot path - TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library.  It is the checker, never the product: the product (raytracer-795_b200/csrc) contains no CPU
// rendering path at all.
//
// What it is: a plain, recursive, scalar C++ re-statement of badiba/raytracer-795's per-pixel algorithm,
// function by function, consuming the same flattened RtSceneDesc / RtCameraDesc the CUDA core consumes
// (include/rt795.h).  Each function cites the reference code it follows.  It differs from the reference in
// exactly one deliberate way: every std::mt19937 draw (src/Camera.cpp:98-110, src/Scene.cpp:44-45,
// src/Light.cpp:523-531, :635-637 - unseeded, racy, irreproducible) is replaced by a
// counter-based Philox4x32-10 stream keyed (seed; pixel, sample, path-node, dimension), the same stream
// definition the CUDA kernels use, so that GPU and oracle can be compared sample for sample.
//
// Pinning: tests/test_oracle_pin.py checks this file against the UNMODIFIED reference compiled by
// oracle/Makefile (oracle/_ref/ref_raytracer, ref_hits) - bit-level on deterministic scenes (primary hits,
// Whitted radiance), statistically on stochastic ones - and against committed fixtures in tests/golden/.
//
// Path tracing (integrator 1) has NO reference code: that part restates DESIGN.md §"Path
// tracer", parity unpinned.
#include "../include/rt795.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------ small vector math
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return (&x)[i]; }
    float& operator[](int i) { return (&x)[i]; }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return a * s; }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline V3 cmul(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { float n = norm(a); return n > 0 ? a / n : a; }   // Eigen: unchanged when the norm is 0
inline V3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }
inline bool isnan3(V3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }

// ------------------------------------------------------------------------------------------ Philox4x32-10
// Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  Stream definition shared with the CUDA
// kernels by specification, implemented independently here.
struct Rng4 { float u[4]; };
inline Rng4 philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Rng4 o;
    uint32_t c[4] = {c0, c1, c2, c3};
    for (int i = 0; i < 4; i++) o.u[i] = (float)(c[i] >> 8) * (1.0f / 16777216.0f);   // [0,1)
    return o;
}
enum { DIM_CAMERA = 0, DIM_GLOSSY = 1, DIM_PT_BSDF = 2, DIM_PT_RR = 3, DIM_PT_LIGHT = 4, DIM_LIGHT0 = 0x100, DIM_LIGHT_STRIDE = 0x100 };

// ------------------------------------------------------------------------------------------ scene view
struct Ray { V3 o, d; float time; };

struct Hit {       // ReturnVal, src/defs.h:13-22 (+ ids for the parity harness)
    V3 point, normal;
    bool full = false;
    int matIndex = 0;
    int dm = RT_DM_NONE;
    V3 textureColor;
    float textureNormalizer = 1;
    int object = -1, prim = -1;
    float t = 0;
};

struct Bvh {       // per-mesh median-split BVH, src/BVH.cpp:53-135
    struct Node { V3 lo, hi; int left, right, start, end; };
    std::vector<Node> nodes;
    std::vector<int> prims;   // triangle indices (global triangle numbering), permuted in place like `primitives`
};

struct Ctx {
    const RtSceneDesc* s;
    std::vector<Bvh> bvh;     // one per mesh
    uint64_t seed;
    int integrator;
    RtRenderOpts opts;
    // camera (src/Camera.cpp:7-61)
    V3 pos, gaze, up, right;
    float l, r, b, t, dist;
    int nx, ny, spp, n;
    float invnx, invny, pw, ph, sw, sh;
    bool dof; float focus, aperture;
    // object lights (RT_LIGHT_MESH / RT_LIGHT_SPHERE; no reference code - DESIGN.md §7): world-space sampling records per light
    struct LightGeom { V3 c; float R = 0; std::vector<V3> tri; std::vector<float> cdf; float area = 0; };
    std::vector<LightGeom> lgeom;
};

V3 TriCenter(const RtSceneDesc* s, int tri) {   // Triangle::GetCenter, src/Shape.cpp:179-190
    const int* ix = s->tri_indices + 3 * tri;
    V3 a = ld3(s->vertices + 3 * ix[0]), b = ld3(s->vertices + 3 * ix[1]), c = ld3(s->vertices + 3 * ix[2]);
    return V3((a.x + b.x + c.x) / 3.0f, (a.y + b.y + c.y) / 3.0f, (a.z + b.z + c.z) / 3.0f);
}

int BuildNode(const RtSceneDesc* s, Bvh& bv, int start, int end, int split, int depth) {
    // BVH::ConstructionHelper, src/BVH.cpp:64-110
    if (start == end) return -1;
    if (start == end - 1 || depth >= 30) {
        Bvh::Node n; n.left = n.right = -1; n.start = start; n.end = end;
        n.lo = V3(0, 0, 0); n.hi = V3(0, 0, 0);      // leaves carry no box in the reference
        bv.nodes.push_back(n);
        return (int)bv.nodes.size() - 1;
    }
    if (split > 2) split = 0;
    Bvh::Node n;
    float mx = std::numeric_limits<float>::max();
    n.lo = V3(mx, mx, mx); n.hi = V3(-mx, -mx, -mx);
    std::vector<float> centers;
    centers.reserve(end - start);
    for (int i = start; i < end; i++) {
        const int* ix = s->tri_indices + 3 * bv.prims[i];
        for (int k = 0; k < 3; k++) {
            V3 p = ld3(s->vertices + 3 * ix[k]);
            for (int a = 0; a < 3; a++) { n.lo[a] = std::min(n.lo[a], p[a]); n.hi[a] = std::max(n.hi[a], p[a]); }
        }
        centers.push_back(TriCenter(s, bv.prims[i])[split]);
    }
    std::vector<float> sorted = centers;
    std::sort(sorted.begin(), sorted.end());
    int len = end - start, mi = len / 2;
    float median = (len % 2 == 0) ? (sorted[mi - 1] + sorted[mi]) * 0.5f : sorted[mi];   // FindMedian, src/BVH.cpp:117-135
    int swap = start;
    for (int i = start; i < end; i++) {
        float c = TriCenter(s, bv.prims[i])[split];
        if (c < median) { std::swap(bv.prims[swap], bv.prims[i]); swap++; }
    }
    n.start = start; n.end = end; n.left = n.right = -1;
    int me = (int)bv.nodes.size();
    bv.nodes.push_back(n);
    int L = BuildNode(s, bv, start, swap, split + 1, depth + 1);
    int R = BuildNode(s, bv, swap, end, split + 1, depth + 1);
    bv.nodes[me].left = L; bv.nodes[me].right = R;
    return me;
}

// ------------------------------------------------------------------------------------------ Perlin, src/Perlin.cpp
const float kPerlinTable[16][3] = {{1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1},
                                   {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}, {1, 1, 0}, {-1, 1, 0}, {0, -1, 1}, {0, -1, -1}};
const int kPerlinShuffle[16] = {12, 7, 15, 6, 11, 0, 4, 9, 13, 3, 14, 8, 2, 5, 1, 10};
inline int Shuffled(int i) { int k = i % 16; if (k < 0) k += 16; return kPerlinShuffle[k]; }     // src/Perlin.cpp:91-97
inline float PWeight(float x) {                                                                 // src/Perlin.cpp:27-30 (pow in double)
    x = std::fabs(x);
    return (float)(((-6) * pow((double)x, 5)) + (15 * pow((double)x, 4)) - (10 * pow((double)x, 3)) + 1);
}
float ComputePerlin(V3 p, float scale, int nc) {                                                // src/Perlin.cpp:52-84
    V3 point = p * scale;
    int ii = (int)floor(point.x), jj = (int)floor(point.y), kk = (int)floor(point.z);
    float value = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                int lx = ii + i, ly = jj + j, lz = kk + k;
                int idx = Shuffled(lx + Shuffled(ly + Shuffled(lz)));
                V3 g(kPerlinTable[idx][0], kPerlinTable[idx][1], kPerlinTable[idx][2]);
                V3 d = point - V3((float)lx, (float)ly, (float)lz);
                value += dot(g, d) * (PWeight(d.x) * PWeight(d.y) * PWeight(d.z));
            }
    if (nc == RT_NC_LINEAR) value = (value + 1) * 0.5f;
    else if (nc == RT_NC_ABSVAL) value = std::fabs(value);
    return value;
}
V3 PerlinGradient(V3 p, float scale, int nc) {                                                  // src/Perlin.cpp:36-50
    const float eps = 0.001f;
    float v0 = ComputePerlin(p, scale, nc);
    float x = (ComputePerlin(V3(p.x + eps, p.y, p.z), scale, nc) - v0) / eps;
    float y = (ComputePerlin(V3(p.x, p.y + eps, p.z), scale, nc) - v0) / eps;
    float z = (ComputePerlin(V3(p.x, p.y, p.z + eps), scale, nc) - v0) / eps;
    return V3(x, y, z);
}

// ------------------------------------------------------------------------------------------ image textures, src/Texture.cpp:41-131
V3 ColorAtPixel(const RtImage& im, int i, int j) {
    if (i < 0) i = 0; else if (i >= im.width) i = im.width - 1;
    if (j < 0) j = 0; else if (j >= im.height) j = im.height - 1;
    size_t o = ((size_t)j * im.width + i) * 4;
    if (im.format == RT_IMG_RGBA8) { const unsigned char* p = (const unsigned char*)im.pixels + o; return V3((float)p[0], (float)p[1], (float)p[2]); }
    const float* p = (const float*)im.pixels + o;
    return V3(p[0], p[1], p[2]);
}
V3 ColorAtCoordinates(const RtImage& im, int interp, float u, float v) {
    u = u - floor(u); v = v - floor(v);
    float i = u * im.width, j = v * im.height;
    if (interp == RT_INTERP_NN) return ColorAtPixel(im, (int)i, (int)j);
    int li = (int)floor(i), lj = (int)floor(j);
    float a = i - li, b = j - lj;
    return (1 - a) * (1 - b) * ColorAtPixel(im, li, lj) + (1 - a) * b * ColorAtPixel(im, li, lj + 1) + a * (1 - b) * ColorAtPixel(im, li + 1, lj) +
           a * b * ColorAtPixel(im, li + 1, lj + 1);
}
void ChangeAtCoordinates(const RtImage& im, float u, float v, float* du, float* dv) {           // src/Texture.cpp:76-109
    u = u - floor(u); v = v - floor(v);
    int i = (int)(u * im.width), j = (int)(v * im.height);
    if (i < 0) i = 0; else if (i >= im.width - 1) i = im.width - 2;
    if (j < 0) j = 0; else if (j >= im.height - 1) j = im.height - 2;
    V3 c = ColorAtPixel(im, i, j), cu = ColorAtPixel(im, i + 1, j), cv = ColorAtPixel(im, i, j + 1);
    float dede = (cu.x + cu.y + cu.z) / 3.0f, nene = (c.x + c.y + c.z) / 3.0f;
    V3 d = cv - c;
    *du = dede - nene;
    *dv = (d.x + d.y + d.z) / 3.0f;
}

// ------------------------------------------------------------------------------------------ intersection
int AbsSmallestIndex(V3 v) {                                                                    // src/Helper.cpp:322-336
    float a = std::fabs(v.x), b = std::fabs(v.y), c = std::fabs(v.z);
    if (a <= b && a <= c) return 0;
    if (b <= a && b <= c) return 1;
    return 2;
}
V3 OrthonormalU(V3 v) {                                                                         // src/Helper.cpp:338-343
    V3 nl = v;
    nl[AbsSmallestIndex(nl)] = 1;
    return normalized(cross(v, nl));
}

inline float Det3(V3 c0, V3 c1, V3 c2) {   // determinant of the matrix with COLUMNS c0,c1,c2 (src/Shape.cpp:305-308)
    return c0.x * (c1.y * c2.z - c2.y * c1.z) - c1.x * (c0.y * c2.z - c2.y * c0.z) + c2.x * (c0.y * c1.z - c1.y * c0.z);
}

// Triangle::TextureComputation, src/Shape.cpp:505-616
void TriTextures(const Ctx& cx, const RtObject& ob, const int* ix, V3 e1, V3 e2, float beta, float gamma, Hit& ret) {
    const RtSceneDesc* s = cx.s;
    ret.dm = RT_DM_NONE;
    int ntex = (ob.tex[0] ? 1 : 0) + (ob.tex[1] ? 1 : 0);
    if (ntex == 0) return;
    float alpha = 1 - beta - gamma;
    const float* t0 = s->texcoords + 2 * (ix[0] + ob.tex_offset);
    const float* t1 = s->texcoords + 2 * (ix[1] + ob.tex_offset);
    const float* t2 = s->texcoords + 2 * (ix[2] + ob.tex_offset);
    float u = t0[0] * alpha + t1[0] * beta + t2[0] * gamma, v = t0[1] * alpha + t1[1] * beta + t2[1] * gamma;
    for (int k = 0; k < 2; k++) {
        if (!ob.tex[k]) continue;
        const RtTexture& tx = s->textures[ob.tex[k] - 1];
        if (tx.type == RT_TEX_IMAGE) {
            const RtImage& im = s->images[tx.image];
            if (tx.decal_mode == RT_DM_REPLACE_KD || tx.decal_mode == RT_DM_BLEND_KD || tx.decal_mode == RT_DM_REPLACE_ALL) {
                ret.dm = tx.decal_mode;
                ret.textureColor = ColorAtCoordinates(im, tx.interpolation, u, v);
                ret.textureNormalizer = (float)tx.normalizer;
            } else if (tx.decal_mode == RT_DM_REPLACE_NORMAL || tx.decal_mode == RT_DM_BUMP_NORMAL) {
                float a00 = t1[0] - t0[0], a01 = t1[1] - t0[1], a10 = t2[0] - t0[0], a11 = t2[1] - t0[1];
                float det = a00 * a11 - a01 * a10, inv = 1.0f / det;
                float i00 = a11 * inv, i01 = -a01 * inv, i10 = -a10 * inv, i11 = a00 * inv;
                V3 T = e1 * i00 + e2 * i01, B = e1 * i10 + e2 * i11;
                if (tx.decal_mode == RT_DM_REPLACE_NORMAL) {
                    V3 rn = ColorAtCoordinates(im, tx.interpolation, u, v) / 255.0f;
                    rn = normalized(rn - V3(0.5f, 0.5f, 0.5f));
                    ret.normal = T * rn.x + B * rn.y + ret.normal * rn.z;     // TBN * rNormal (not re-normalised here)
                } else {
                    float du, dv;
                    ChangeAtCoordinates(im, u, v, &du, &dv);
                    du *= tx.bump_factor; dv *= tx.bump_factor;
                    V3 pu = T + du * ret.normal, pv = B + dv * ret.normal;
                    V3 nn = normalized(cross(pv, pu));
                    ret.normal = dot(ret.normal, nn) > 0 ? nn : -nn;
                }
            }
        } else {
            if (tx.decal_mode == RT_DM_REPLACE_KD) {
                ret.dm = tx.decal_mode;
                float p = ComputePerlin(ret.point, tx.noise_scale, tx.noise_conversion);
                ret.textureColor = V3(p, p, p);
                ret.textureNormalizer = 1;
            } else if (tx.decal_mode == RT_DM_BUMP_NORMAL) {
                V3 g = PerlinGradient(ret.point, tx.noise_scale, tx.noise_conversion);
                V3 gp = dot(g, ret.normal) * ret.normal;
                V3 nn = ret.normal - (g - gp) * tx.bump_factor;
                ret.normal = dot(ret.normal, nn) > 0 ? nn : -nn;
                ret.normal = normalized(ret.normal);
            }
        }
    }
}

// Triangle::bvhIntersect, src/Shape.cpp:297-345
bool TriIntersect(const Ctx& cx, const RtObject& ob, const RtMesh& mesh, int tri, const Ray& ray, Hit& ret) {
    const RtSceneDesc* s = cx.s;
    const int* ix = s->tri_indices + 3 * tri;
    V3 a = ld3(s->vertices + 3 * ix[0]), b = ld3(s->vertices + 3 * ix[1]), c = ld3(s->vertices + 3 * ix[2]);
    float det = Det3(a - b, a - c, ray.d);
    float beta = Det3(a - ray.o, a - c, ray.d) / det;
    float gamma = Det3(a - b, a - ray.o, ray.d) / det;
    float t = Det3(a - b, a - c, a - ray.o) / det;
    float eps = s->int_test_eps;
    if (!(t >= -eps && (beta + gamma <= 1) && beta >= -eps && gamma >= -eps)) return false;
    V3 normal;
    if (mesh.smooth) {
        float alpha = 1 - beta - gamma;
        normal = ld3(s->vertex_normals + 3 * ix[0]) * alpha + ld3(s->vertex_normals + 3 * ix[1]) * beta + ld3(s->vertex_normals + 3 * ix[2]) * gamma;
    } else normal = cross(c - b, a - b);
    ret = Hit();
    ret.normal = normal / norm(normal);
    ret.point = ray.o + ray.d * t;
    ret.t = t;
    TriTextures(cx, ob, ix, b - a, c - a, beta, gamma, ret);
    ret.full = true;
    return true;
}

// Sphere::bvhIntersect + Sphere::TextureComputation, src/Shape.cpp:347-503
bool SphereIntersect(const Ctx& cx, const RtObject& ob, const RtSphere& sp, const Ray& ray, Hit& ret) {
    const RtSceneDesc* s = cx.s;
    V3 d = ray.d, o = ray.o, c = ld3(sp.center);
    float R = sp.radius;
    float disc = (dot(d, o - c) * dot(d, o - c) - dot(d, d) * (dot(o - c, o - c) - R * R));
    if (disc < s->int_test_eps) return false;
    float t1 = (-dot(d, o - c) + std::sqrt(disc)) / dot(d, d);
    float t2 = (-dot(d, o - c) - std::sqrt(disc)) / dot(d, d);
    float t;
    if (t1 >= 0 && t2 < 0) t = t1;
    else if (t2 >= 0 && t1 < 0) t = t2;
    else if (t1 < 0 && t2 < 0) return false;
    else t = t1 < t2 ? t1 : t2;
    ret = Hit();
    ret.point = o + d * t;
    ret.t = t;
    ret.normal = (ret.point - c) / norm(ret.point - c);
    ret.dm = RT_DM_NONE;
    for (int k = 0; k < 2; k++) {
        if (!ob.tex[k]) continue;
        const RtTexture& tx = s->textures[ob.tex[k] - 1];
        if (tx.type == RT_TEX_IMAGE) {
            const RtImage& im = s->images[tx.image];
            V3 lc = ret.point - c;
            float theta = acos(lc.y / R), phi = atan2(lc.z, lc.x);
            float tu = (float)((-phi + M_PI) / (2 * M_PI)), tv = (float)(theta / M_PI);
            if (tx.decal_mode == RT_DM_REPLACE_KD || tx.decal_mode == RT_DM_BLEND_KD || tx.decal_mode == RT_DM_REPLACE_ALL) {
                ret.dm = tx.decal_mode;
                ret.textureColor = ColorAtCoordinates(im, tx.interpolation, tu, tv);
                ret.textureNormalizer = (float)tx.normalizer;
            } else if (tx.decal_mode == RT_DM_REPLACE_NORMAL || tx.decal_mode == RT_DM_BUMP_NORMAL) {
                float pi = (float)M_PI;
                V3 dpdu(lc.z * 2 * pi, 0, lc.x * (-2) * pi);
                V3 dpdv(lc.y * cos(phi) * pi, (-1) * R * sin(theta) * pi, lc.y * sin(phi) * pi);
                if (tx.decal_mode == RT_DM_REPLACE_NORMAL) {
                    V3 rn = ColorAtCoordinates(im, tx.interpolation, tu, tv) / 255.0f;
                    rn = normalized(rn - V3(0.5f, 0.5f, 0.5f));
                    V3 T = normalized(dpdu), B = normalized(dpdv);
                    ret.normal = T * rn.x + B * rn.y + ret.normal * rn.z;
                } else {
                    float du, dv;
                    ChangeAtCoordinates(im, tu, tv, &du, &dv);
                    du *= tx.bump_factor; dv *= tx.bump_factor;
                    V3 pu = dpdu + du * ret.normal, pv = dpdv + dv * ret.normal;
                    V3 nn = normalized(cross(pv, pu));
                    ret.normal = dot(ret.normal, nn) > 0 ? nn : -nn;
                }
            }
        } else {
            if (tx.decal_mode == RT_DM_REPLACE_KD) {
                ret.dm = tx.decal_mode;
                float p = ComputePerlin(ret.point, tx.noise_scale, tx.noise_conversion);
                ret.textureColor = V3(p, p, p);
                ret.textureNormalizer = 1;
            } else if (tx.decal_mode == RT_DM_BUMP_NORMAL) {
                V3 g = PerlinGradient(ret.point, tx.noise_scale, tx.noise_conversion);
                V3 gp = dot(g, ret.normal) * ret.normal;
                V3 nn = ret.normal - (g - gp) * tx.bump_factor;
                ret.normal = dot(ret.normal, nn) > 0 ? nn : -nn;
                ret.normal = normalized(ret.normal);
            }
        }
    }
    ret.full = true;
    return true;
}

// BVH::RayBBoxIntersection, src/BVH.cpp:212-266
bool RayBox(const Ray& ray, V3 lo, V3 hi) {
    float te[3], tl[3];
    for (int a = 0; a < 3; a++) {
        if (ray.d[a] > 0) { te[a] = (lo[a] - ray.o[a]) / ray.d[a]; tl[a] = (hi[a] - ray.o[a]) / ray.d[a]; }
        else { te[a] = (hi[a] - ray.o[a]) / ray.d[a]; tl[a] = (lo[a] - ray.o[a]) / ray.d[a]; }
    }
    auto min3 = [](float a, float b, float c) { if (a <= b && a <= c) return a; if (b <= a && b <= c) return b; return c; };
    auto max3 = [](float a, float b, float c) { if (a >= b && a >= c) return a; if (b >= a && b >= c) return b; return c; };
    return !(min3(tl[0], tl[1], tl[2]) < max3(te[0], te[1], te[2]));
}

// BVH::FindIntersectionWithBVH, src/BVH.cpp:137-210 (both children always visited; nearer by Euclidean distance, ties -> right)
Hit BvhIntersect(const Ctx& cx, const RtObject& ob, const RtMesh& mesh, const Bvh& bv, int node, const Ray& ray) {
    Hit none;
    if (node < 0) return none;
    const Bvh::Node& n = bv.nodes[node];
    if (n.left < 0 && n.right < 0) {
        Hit nearest;
        float best = std::numeric_limits<float>::max();
        for (int i = n.start; i < n.end; i++) {
            Hit h;
            if (TriIntersect(cx, ob, mesh, bv.prims[i], ray, h)) {
                float dist = norm(h.point - ray.o);
                if (dist < best) { best = dist; nearest = h; nearest.matIndex = ob.mat_index; nearest.prim = bv.prims[i] - mesh.first_tri; }
            }
        }
        return nearest;
    }
    if (RayBox(ray, n.lo, n.hi)) {
        Hit L = BvhIntersect(cx, ob, mesh, bv, n.left, ray);
        Hit R = BvhIntersect(cx, ob, mesh, bv, n.right, ray);
        if (L.full && !R.full) return L;
        if (!L.full && R.full) return R;
        if (L.full && R.full) return norm(L.point - ray.o) < norm(R.point - ray.o) ? L : R;
    }
    return none;
}

// Ray::gett, src/Ray.cpp:21-36
float Gett(const Ray& r, V3 p) {
    float t = (p.x - r.o.x) / r.d.x; if (t == t) return t;
    t = (p.y - r.o.y) / r.d.y; if (t == t) return t;
    t = (p.z - r.o.z) / r.d.z; if (t == t) return t;
    return t;   // the reference falls off the end (UB); NaN fails both comparisons below either way
}

// BVHMethods::FindIntersection, src/Helper.cpp:18-80 (+ TransformRay :110-133, TransformNormal :100-108)
std::atomic<unsigned long long> g_ray_queries{0};   // every closest-hit query the algorithm issues (bench.py: rays of the reference arm)
Hit FindIntersection(const Ctx& cx, const Ray& ray) {
    const RtSceneDesc* s = cx.s;
    thread_local unsigned local = 0;
    if (++local == 1024) { g_ray_queries.fetch_add(1024, std::memory_order_relaxed); local = 0; }
    Hit nearest;
    float nearestDistance = std::numeric_limits<float>::max();
    if (isnan3(ray.o) || isnan3(ray.d)) return nearest;
    const RtObject* winner = nullptr;
    for (int i = 0; i < s->num_objects; i++) {
        const RtObject& ob = s->objects[i];
        Ray tr;
        tr.time = ray.time;
        V3 o(ray.o.x - ob.blur[0] * ray.time, ray.o.y - ob.blur[1] * ray.time, ray.o.z - ob.blur[2] * ray.time);
        const float* m = ob.inv;
        tr.o = V3(m[0] * o.x + m[1] * o.y + m[2] * o.z + m[3], m[4] * o.x + m[5] * o.y + m[6] * o.z + m[7], m[8] * o.x + m[9] * o.y + m[10] * o.z + m[11]);
        tr.d = V3(m[0] * ray.d.x + m[1] * ray.d.y + m[2] * ray.d.z, m[4] * ray.d.x + m[5] * ray.d.y + m[6] * ray.d.z, m[8] * ray.d.x + m[9] * ray.d.y + m[10] * ray.d.z);
        Hit ret;
        if (ob.kind == RT_OBJ_SPHERE) {
            Hit h;
            if (SphereIntersect(cx, ob, s->spheres[ob.geom], tr, h)) { ret = h; ret.matIndex = ob.mat_index; ret.prim = 0; }
        } else {
            const RtMesh& mesh = s->meshes[ob.geom];
            ret = BvhIntersect(cx, ob, mesh, cx.bvh[ob.geom], 0, tr);
        }
        if (ret.full) {
            float distance = Gett(tr, ret.point);
            if (distance < nearestDistance && distance > 0) {
                nearestDistance = distance;
                winner = &ob;
                ret.point = ray.o + ray.d * distance;
                ret.t = distance;
                ret.object = i;
                ret.matIndex = ob.mat_index;     // instances override (Helper.cpp:68); for objects it already is the object's
                nearest = ret;
            }
        }
    }
    if (nearest.full) {
        const float* m = winner->inv_t;
        V3 n = nearest.normal;
        nearest.normal = normalized(V3(m[0] * n.x + m[1] * n.y + m[2] * n.z, m[3] * n.x + m[4] * n.y + m[5] * n.z, m[6] * n.x + m[7] * n.y + m[8] * n.z));
    }
    return nearest;
}

// ------------------------------------------------------------------------------------------ lights, src/Light.cpp
float ConductorFresnel(float n_t, float k_t, V3 d, V3 normal) {                                 // src/Scene.cpp:135-146 == src/Light.cpp:18-28
    float cos_t = -dot(d, normal);
    float twoNtCost = 2 * n_t * cos_t;
    float cosSquared = (float)pow(cos_t, 2);
    float nt_ktSquare = (float)(pow(n_t, 2) + pow(k_t, 2));
    float rs = (nt_ktSquare - twoNtCost + cosSquared) / (nt_ktSquare + twoNtCost + cosSquared);
    float rp = (nt_ktSquare * cosSquared - twoNtCost + 1) / (nt_ktSquare * cosSquared + twoNtCost + 1);
    return 0.5f * (rs + rp);
}

V3 TermBRDF(const RtMaterial& mat, V3 wi, V3 wo, V3 n) {                                        // src/Light.cpp:62-155
    const float eps = 0.001f;
    V3 kd = ld3(mat.diffuse), ks = ld3(mat.specular);
    int p = mat.phong_exp;
    switch (mat.brdf) {
        case RT_BRDF_MP: case RT_BRDF_OP: case RT_BRDF_MPN: {
            float n_wi = dot(n, wi);
            V3 wr = -wi + n * 2 * n_wi;
            wr = wr / norm(wr);
            float ca = std::max(0.0f, dot(wr, wo));
            if (mat.brdf == RT_BRDF_MP) return kd + ks * (float)pow(ca, p);
            if (mat.brdf == RT_BRDF_OP) {
                float ci = std::max(0.0f, dot(wi, n));
                if (ci < eps) return V3(0, 0, 0);
                return kd + (ks * (float)pow(ca, p)) / ci;
            }
            return kd / (float)M_PI + ks * (float)((p + 2) / (2 * M_PI)) * (float)pow(ca, p);
        }
        case RT_BRDF_MBP: case RT_BRDF_OBP: case RT_BRDF_MBPN: {
            V3 h = normalized(wo + wi);
            float ca = std::max(0.0f, dot(n, h));
            if (mat.brdf == RT_BRDF_MBP) return kd + ks * (float)pow(ca, p);
            if (mat.brdf == RT_BRDF_OBP) {
                float ci = std::max(0.0f, dot(wi, n));
                if (ci < eps) return V3(0, 0, 0);
                return kd + (ks * (float)pow(ca, p)) / ci;
            }
            return kd / (float)M_PI + ks * (float)((p + 8) / (8 * M_PI)) * (float)pow(ca, p);
        }
        case RT_BRDF_TS: case RT_BRDF_TSF: {
            V3 wh = normalized(wo + wi);
            V3 diffusePart = kd / (float)M_PI;
            float f = 1;
            if (mat.brdf == RT_BRDF_TSF) {
                f = ConductorFresnel(mat.refraction_index, mat.absorption_index, -wo, n);
                diffusePart = diffusePart * (1 - f);
            }
            float cosAlpha = dot(wh, n), cosTheta = dot(wi, n), cosPhi = dot(wo, n);
            float left = 2.0f * dot(n, wh) * dot(n, wo); left = left / dot(wo, wh);
            float right = 2.0f * dot(n, wh) * dot(n, wi); right = right / dot(wi, wh);
            float g = std::min(1.0f, std::min(left, right));
            float x = (float)((p + 2.0f) / (2.0f * M_PI));
            float dd = x * (float)pow(cosAlpha, p);
            V3 sp = ks * g * dd;
            sp = sp / (4 * cosPhi * cosTheta);
            if (mat.brdf == RT_BRDF_TSF) sp = sp * f;
            return diffusePart + sp;
        }
    }
    return V3(0, 0, 0);
}
V3 BRDF(const RtMaterial& mat, V3 wi, V3 wo, V3 n, V3 radiance) {                               // src/Light.cpp:157-162
    V3 term = TermBRDF(mat, wi, wo, n);
    float ca = std::max(0.0f, dot(wi, n));
    return cmul(radiance, term) * ca;
}
// *Light::Diffuse + Specular (identical in every light class, e.g. src/Light.cpp:206-236) for irradiance E, direction wi
V3 DiffuseSpecular(const RtMaterial& mat, const Ray& ray, const Hit& ret, V3 E, V3 wi) {
    float alpha = std::max(0.0f, dot(ret.normal, wi));
    V3 diffuse;
    if (ret.dm == RT_DM_REPLACE_KD) diffuse = cmul(E, (ret.textureColor / ret.textureNormalizer) * alpha);
    else if (ret.dm == RT_DM_BLEND_KD) diffuse = cmul(E, ((ld3(mat.diffuse) + (ret.textureColor / ret.textureNormalizer)) * 0.5f) * alpha);
    else diffuse = cmul(E, ld3(mat.diffuse) * alpha);
    V3 wo = -ray.d;
    V3 h = (wo + wi) / norm(wo + wi);
    float a2 = std::max(0.0f, dot(ret.normal, h));
    V3 specular = cmul(E, ld3(mat.specular) * (float)pow(a2, mat.phong_exp));
    return diffuse + specular;
}

V3 LightShading(const Ctx& cx, int li, const Ray& ray, const Hit& ret, const RtMaterial& mat, uint32_t pixel, uint32_t sample, uint32_t node) {
    const RtSceneDesc* s = cx.s;
    const RtLight& L = s->lights[li];
    V3 p = ret.point, n = ret.normal;
    V3 origin = p + n * s->shadow_ray_eps;
    auto shade = [&](V3 wi, V3 E) { return mat.brdf != RT_BRDF_NONE ? BRDF(mat, wi, -ray.d, n, E) : DiffuseSpecular(mat, ray, ret, E, wi); };
    auto blocked_before = [&](V3 target) {   // Point/Spot/Area IsShadow, e.g. src/Light.cpp:188-204
        V3 dir = target - p;
        Ray sr{origin, dir / norm(dir), ray.time};
        Hit h = FindIntersection(cx, sr);
        return h.full && norm(p - target) > norm(p - h.point);
    };
    switch (L.type) {
        case RT_LIGHT_POINT: {                                                                   // src/Light.cpp:168-250
            V3 pos = ld3(L.position);
            if (blocked_before(pos)) return V3(0, 0, 0);
            float d = norm(p - pos);
            return shade(normalized(pos - p), ld3(L.intensity) / (d * d));
        }
        case RT_LIGHT_DIRECTIONAL: {                                                             // src/Light.cpp:256-321
            V3 dir = normalized(ld3(L.direction));
            Ray sr{origin, -dir, ray.time};
            if (FindIntersection(cx, sr).full) return V3(0, 0, 0);
            return shade(-dir, ld3(L.intensity));
        }
        case RT_LIGHT_SPOT: {                                                                    // src/Light.cpp:327-436
            V3 pos = ld3(L.position), dir = normalized(ld3(L.direction));
            float coverage = (float)((L.coverage_deg * 0.5f) * (M_PI / 180.0f)), fall = (float)((L.falloff_deg * 0.5f) * (M_PI / 180.0f));
            if (blocked_before(pos)) return V3(0, 0, 0);
            float angle = acos(dot(normalized(p - pos), dir));
            float d = norm(p - pos);
            V3 E = ld3(L.intensity) / (d * d);
            if (angle < fall) return shade(normalized(pos - p), E);
            if (angle < coverage) {
                float cf = cos(fall), cc = cos(coverage);
                float fo = (float)pow((cos(angle) - cc) / (cf - cc), 4);
                return shade(normalized(pos - p), E) * fo;
            }
            return V3(0, 0, 0);
        }
        case RT_LIGHT_AREA: {                                                                    // src/Light.cpp:442-545
            V3 pos = ld3(L.position), ln = normalized(ld3(L.direction));
            V3 u = OrthonormalU(ln), v = cross(ln, u);
            Rng4 r = philox(cx.seed, pixel, sample, node, DIM_LIGHT0 + li * DIM_LIGHT_STRIDE);
            float uChi = r.u[0] - 0.5f, vChi = r.u[1] - 0.5f;
            V3 sp = pos + (u * L.size * uChi) + (v * L.size * vChi);
            if (blocked_before(sp)) return V3(0, 0, 0);
            V3 dtp = normalized(p - sp);
            float cosTheta = std::fabs(dot(dtp, ln));
            float dsq = norm(p - sp); dsq = dsq * dsq;
            V3 E = ld3(L.intensity) * ((L.size * L.size) * (cosTheta / dsq));
            return shade(normalized(sp - p), E);
        }
        case RT_LIGHT_SPHERE: case RT_LIGHT_MESH: {                                              // object lights: twin of csrc/shade.cuh light_eval
            const Ctx::LightGeom& g = cx.lgeom[li];
            const float pi = 3.14159265358979323846f;
            Rng4 r = philox(cx.seed, pixel, sample, node, DIM_LIGHT0 + li * DIM_LIGHT_STRIDE);
            V3 target, E;
            float object_tmax = 3.0e38f;                          // where the shadow ray (from p + n eps) meets the emitter's own surface
            if (L.type == RT_LIGHT_SPHERE) {
                V3 pc = g.c - p;
                float R = g.R, d2 = dot(pc, pc);
                if (!(d2 > R * R * 1.0002f)) return V3(0, 0, 0);
                float d = std::sqrt(d2), cosmax = std::sqrt(std::max(0.f, 1.f - R * R / d2));
                float ct = 1.f - r.u[0] * (1.f - cosmax), st = std::sqrt(std::max(0.f, 1.f - ct * ct));
                float ph = 2.0f * pi * r.u[1];
                V3 w = pc / d, u = OrthonormalU(w), v = cross(w, u);
                V3 dir = normalized(u * (st * std::cos(ph)) + v * (st * std::sin(ph)) + w * ct);
                float th = d * ct - std::sqrt(std::max(0.f, R * R - d2 * st * st));
                target = p + dir * th;
                E = ld3(L.intensity) * (2.0f * pi * (1.f - cosmax));
                V3 oc = origin - g.c;
                float bq = dot(oc, dir), cq = dot(oc, oc) - R * R, dq = bq * bq - cq;
                object_tmax = dq > 0.f ? std::max(0.f, -bq - std::sqrt(dq)) : 3.0e38f;
            } else {
                int ntri = (int)g.cdf.size();
                if (ntri <= 0 || !(g.area > 0.f)) return V3(0, 0, 0);
                int lo = 0, hi = ntri - 1;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (r.u[0] < g.cdf[mid]) hi = mid; else lo = mid + 1; }
                V3 a = g.tri[3 * lo], b = g.tri[3 * lo + 1], c = g.tri[3 * lo + 2];
                float su = std::sqrt(r.u[1]), b1 = 1.f - su, b2 = r.u[2] * su;
                target = a * b1 + b * b2 + c * (1.f - b1 - b2);
                V3 ln = normalized(cross(c - b, a - b));
                V3 dv = p - target;
                float dsq = dot(dv, dv);
                float cosl = dot(ln, dv) / std::sqrt(std::max(dsq, 1e-30f));
                if (!(cosl > 0.f)) return V3(0, 0, 0);
                E = ld3(L.intensity) * (g.area * cosl / dsq);
                V3 wdir = -dv / std::sqrt(std::max(dsq, 1e-30f));
                float den = dot(ln, wdir);
                object_tmax = den < 0.f ? dot(ln, a - origin) / den : 3.0e38f;
            }
            V3 dv = target - p;
            float D = norm(dv);
            V3 wi = dv / D;
            float eps = s->shadow_ray_eps, sdot = dot(n, wi);
            float tmax = -eps * sdot + std::sqrt(std::max(0.f, eps * eps * sdot * sdot - eps * eps + D * D));
            tmax = std::min(tmax * 1.0002f, object_tmax) * 0.9998f;
            V3 rad = shade(wi, E);
            if (rad.x == 0 && rad.y == 0 && rad.z == 0) return rad;
            Ray sr{origin, wi, ray.time};
            Hit h = FindIntersection(cx, sr);
            if (h.full && dot(h.point - origin, wi) < tmax) return V3(0, 0, 0);                 // something before the emitter's own surface
            return rad;
        }
        case RT_LIGHT_ENVIRONMENT: {                                                             // src/Light.cpp:551-660
            V3 u = OrthonormalU(n), w = cross(n, u);
            V3 dir = n;
            {
                // The reference rejection-samples the unit ball and keeps the upper half (:634-648); the normalised direction is
                // uniform over the hemisphere about n.  Restated as the direct draw of that same distribution (cos(theta) uniform
                // in [0,1], phi uniform) because the product kernel does so (csrc/shade.cuh); the reference's RNG is
                // unseeded, so only the distribution can be pinned - tests/test_oracle_pin.py checks it against the reference binary.
                Rng4 r = philox(cx.seed, pixel, sample, node, DIM_LIGHT0 + li * DIM_LIGHT_STRIDE);
                float c = r.u[0], sn = std::sqrt(std::max(0.f, 1.f - c * c));
                float ph = 2.0f * 3.14159265358979323846f * r.u[1];
                float sp_ = std::sin(ph), cp_ = std::cos(ph);
                dir = normalized(u * (sn * cp_) + n * c + w * (sn * sp_));
            }
            Ray sr{origin, dir, ray.time};
            if (FindIntersection(cx, sr).full) return V3(0, 0, 0);
            float theta = acos(dir.y), phi = atan2(dir.z, dir.x);
            float tu = (float)((-phi + M_PI) / (2 * M_PI)), tv = (float)(theta / M_PI);
            V3 rad = ColorAtCoordinates(s->images[L.image], RT_INTERP_BILINEAR, tu, tv);
            V3 E = rad * 2.0f * (float)M_PI;
            return shade(dir, E);
        }
    }
    return V3(0, 0, 0);
}

// Scene::BasicShading, src/Scene.cpp:243-267 (+ ambient :22-30)
V3 BasicShading(const Ctx& cx, const Ray& ray, const Hit& ret, const RtMaterial& mat, uint32_t pixel, uint32_t sample, uint32_t node) {
    V3 raw = cmul(ld3(cx.s->ambient_light), ld3(mat.ambient));
    for (int i = 0; i < cx.s->num_lights; i++) raw = raw + LightShading(cx, i, ray, ret, mat, pixel, sample, node);
    return raw;
}

// Ray-tree node ids of the RNG counter (root 1, reflected 2n, refracted 2n + 1; hashed on beyond 30 bounces) - twin of csrc/kernels.cuh child_node
inline uint32_t ChildNode(uint32_t node, uint32_t branch) { return node < 0x40000000u ? node * 2u + branch : node * 2654435761u + 0x9E3779B9u + branch; }

V3 NanCheck(V3 v) { return isnan3(v) ? V3(0, 0, 0) : v; }                                       // src/Scene.cpp:221-228

struct Spawn { Ray ray; Hit ret; };

// Scene::MirrorReflectance, src/Scene.cpp:32-55
Spawn MirrorReflectance(const Ctx& cx, const Ray& ray, const Hit& ret, const RtMaterial& mat, uint32_t pixel, uint32_t sample, uint32_t node) {
    V3 wo = -ray.d;
    float n_wo = dot(ret.normal, wo);
    V3 wr = -wo + ret.normal * 2 * n_wo;
    wr = wr / norm(wr);
    if (mat.is_rough) {
        V3 u = OrthonormalU(wr), v = cross(wr, u);
        Rng4 r = philox(cx.seed, pixel, sample, node, DIM_GLOSSY);
        float uChi = r.u[0] - 0.5f, vChi = r.u[1] - 0.5f;
        wr = normalized(wr + ((u * uChi + v * vChi) * mat.roughness));
    }
    Spawn sp;
    sp.ray = Ray{ret.point + ret.normal * cx.s->shadow_ray_eps, wr, ray.time};
    sp.ret = FindIntersection(cx, sp.ray);
    return sp;
}

V3 RecursiveShading(const Ctx& cx, const Ray& ray, const Hit& ret, int depth, uint32_t pixel, uint32_t sample, uint32_t node);

// Radiance an object light (<LightMesh> / <LightSphere>) sends back along the ray that hit its front side (extension).
inline V3 Emitted(const Ctx& cx, const Ray& ray, const Hit& ret) {
    if (!ret.full || ret.object < 0) return V3(0, 0, 0);
    V3 le = ld3(cx.s->objects[ret.object].radiance);
    if ((le.x != 0 || le.y != 0 || le.z != 0) && dot(ret.normal, ray.d) < 0) return le;
    return V3(0, 0, 0);
}
V3 RecursiveShadingCore(const Ctx& cx, const Ray& ray, const Hit& ret, int depth, uint32_t pixel, uint32_t sample, uint32_t node);
V3 RecursiveShading(const Ctx& cx, const Ray& ray, const Hit& ret, int depth, uint32_t pixel, uint32_t sample, uint32_t node) {
    return RecursiveShadingCore(cx, ray, ret, depth, pixel, sample, node) + Emitted(cx, ray, ret);
}

inline V3 Recurse(const Ctx& cx, const Spawn& sp, int depth, uint32_t pixel, uint32_t sample, uint32_t node) {
    return RecursiveShading(cx, sp.ray, sp.ret, depth, pixel, sample, node);
}

// Scene::RecursiveShading, src/Scene.cpp:148-219 (+ DielectricRefraction :57-118, FresnelReflectance :120-128, BeerLaw :130-133)
V3 RecursiveShadingCore(const Ctx& cx, const Ray& ray, const Hit& ret, int depth, uint32_t pixel, uint32_t sample, uint32_t node) {
    if (!ret.full) return V3(0, 0, 0);
    const RtSceneDesc* s = cx.s;
    const RtMaterial& mat = s->materials[ret.matIndex - 1];
    if (mat.type == RT_MAT_NORMAL || depth <= 0) return BasicShading(cx, ray, ret, mat, pixel, sample, node);
    if (mat.type == RT_MAT_MIRROR) {
        Spawn sc = MirrorReflectance(cx, ray, ret, mat, pixel, sample, node);
        V3 refl = Recurse(cx, sc, depth - 1, pixel, sample, ChildNode(node, 0));
        refl = cmul(ld3(mat.mirror), refl);
        return BasicShading(cx, ray, ret, mat, pixel, sample, node) + refl;
    }
    if (mat.type == RT_MAT_DIELECTRIC) {
        float dp = dot(ray.d, ret.normal);
        float nt = mat.refraction_index, snell, n_t, n_i;
        V3 normal;
        bool entering;
        if (dp < 0) { snell = 1.0f / nt; normal = ret.normal; n_t = nt; n_i = 1; entering = true; }
        else { snell = nt; normal = -ret.normal; n_t = 1; n_i = nt; entering = false; }
        float cosTheta = -dot(ray.d, normal);
        V3 left = (ray.d + normal * cosTheta) * snell;
        float sq = (float)(1 - pow(snell, 2) * (1 - pow(cosTheta, 2)));
        bool tir = sq < 0;
        sq = std::sqrt(sq);
        V3 tdir = normalized(left - (normal * sq));
        Spawn tr;
        tr.ray = Ray{ret.point - s->shadow_ray_eps * normal, tdir, ray.time};
        tr.ret = FindIntersection(cx, tr.ray);
        float cos_t = -dot(tr.ray.d, normal), cos_i = -dot(ray.d, normal);
        float rPar = (n_t * cos_i - n_i * cos_t) / (n_t * cos_i + n_i * cos_t);
        float rPer = (n_i * cos_i - n_t * cos_t) / (n_i * cos_i + n_t * cos_t);
        float fresnel = 0.5f * (float)(pow(rPar, 2) + pow(rPer, 2));
        // Beer distance: |hit(refracted).point - p|.  On a miss / TIR the reference reads an indeterminate point
        // (SURVEY A.7); BOTH this oracle and the CUDA core define that distance as 0 (beer = 1).
        float beerDistance = tr.ret.full ? norm(tr.ret.point - ret.point) : 0.0f;
        V3 beer(exp(-mat.absorption_coeff[0] * beerDistance), exp(-mat.absorption_coeff[1] * beerDistance), exp(-mat.absorption_coeff[2] * beerDistance));
        if (entering) {
            V3 inside = Recurse(cx, tr, depth - 1, pixel, sample, ChildNode(node, 1));
            inside = (1 - fresnel) * inside;
            inside = cmul(beer, inside);
            Spawn sc = MirrorReflectance(cx, ray, ret, mat, pixel, sample, node);
            V3 refl = Recurse(cx, sc, depth - 1, pixel, sample, ChildNode(node, 0));
            refl = fresnel * refl;
            inside = NanCheck(inside); refl = NanCheck(refl);
            return BasicShading(cx, ray, ret, mat, pixel, sample, node) + inside + refl;
        }
        if (tir) {
            Spawn sc = MirrorReflectance(cx, ray, ret, mat, pixel, sample, node);
            V3 ir = Recurse(cx, sc, depth - 1, pixel, sample, ChildNode(node, 0));
            ir = cmul(beer, ir);
            return NanCheck(ir);
        }
        V3 outside = Recurse(cx, tr, depth - 1, pixel, sample, ChildNode(node, 1));
        outside = (1 - fresnel) * outside;
        Spawn sc = MirrorReflectance(cx, ray, ret, mat, pixel, sample, node);
        V3 refl = Recurse(cx, sc, depth - 1, pixel, sample, ChildNode(node, 0));
        refl = fresnel * refl;
        refl = cmul(beer, refl);
        outside = NanCheck(outside); refl = NanCheck(refl);
        return outside + refl;
    }
    // conductor
    float fresnel = ConductorFresnel(mat.refraction_index, mat.absorption_index, ray.d, ret.normal);
    Spawn sc = MirrorReflectance(cx, ray, ret, mat, pixel, sample, node);
    V3 refl = Recurse(cx, sc, depth - 1, pixel, sample, ChildNode(node, 0));
    refl = fresnel * refl;
    refl = cmul(ld3(mat.mirror), refl);
    return BasicShading(cx, ray, ret, mat, pixel, sample, node) + refl;
}

// Scene::Shading, src/Scene.cpp:230-241
V3 Shading(const Ctx& cx, const Ray& ray, const Hit& ret, uint32_t pixel, uint32_t sample) {
    if (ret.dm == RT_DM_REPLACE_ALL) return ret.textureColor + Emitted(cx, ray, ret);
    return RecursiveShading(cx, ray, ret, cx.s->max_recursion_depth, pixel, sample, 1);
}

// Scene::GetBackgroundColor, src/Scene.cpp:413-435.  (row, col) are passed exactly as the two callers pass them:
// SingleSample hands over (x, y) - a transposed lookup at 1 spp - MultiSample (y, x).
V3 Background(const Ctx& cx, int row, int col, const Ray& ray) {
    const RtSceneDesc* s = cx.s;
    if (s->env_light_index != -1) {
        const RtLight& L = s->lights[s->env_light_index];
        if (L.type != RT_LIGHT_ENVIRONMENT) return ld3(s->background_color);
        float theta = acos(ray.d.y), phi = atan2(ray.d.z, ray.d.x);
        float tu = (float)((-phi + M_PI) / (2 * M_PI)), tv = (float)(theta / M_PI);
        return ColorAtCoordinates(s->images[L.image], RT_INTERP_BILINEAR, tu, tv);
    }
    if (s->background_texture == -1) return ld3(s->background_color);
    float u = ((float)col) / cx.nx, v = ((float)row) / cx.ny;
    const RtTexture& tx = s->textures[s->background_texture];
    if (tx.type != RT_TEX_IMAGE) return ld3(s->background_color);
    return ColorAtCoordinates(s->images[tx.image], tx.interpolation, u, v);
}

// ------------------------------------------------------------------------------------------ camera, src/Camera.cpp
void SetupCamera(Ctx& cx, const RtCameraDesc* c) {
    int i = 1;
    cx.n = 1;
    while (i < 1000) { if (i * i >= c->num_samples) { cx.n = i; break; } i++; }                 // :21-28
    V3 gaze = ld3(c->gaze), up = ld3(c->up);
    cx.gaze = normalized(gaze);
    V3 w = c->left_handed ? normalized(gaze) : normalized(-gaze);
    cx.right = normalized(cross(up, w));
    cx.up = cross(w, cx.right);
    cx.pos = ld3(c->position);
    cx.l = c->left; cx.r = c->right; cx.b = c->bottom; cx.t = c->top; cx.dist = c->distance;
    cx.nx = c->nx; cx.ny = c->ny; cx.spp = c->num_samples;
    cx.invnx = 1.0f / c->nx; cx.invny = 1.0f / c->ny;
    cx.pw = (cx.r - cx.l) * cx.invnx; cx.ph = (cx.t - cx.b) * cx.invny;
    cx.sw = cx.pw / cx.n; cx.sh = cx.ph / cx.n;
    cx.dof = c->is_dof != 0; cx.focus = c->focus_distance; cx.aperture = c->aperture_size;
}
Ray PrimaryRay(const Ctx& cx, int col, int row) {                                               // :63-72
    float u = cx.l + (cx.r - cx.l) * (col + 0.5f) * cx.invnx;
    float v = cx.t - (cx.t - cx.b) * (row + 0.5f) * cx.invny;
    V3 m = cx.pos + cx.gaze * cx.dist;
    m = m + u * cx.right;
    m = m + v * cx.up;
    return Ray{cx.pos, (m - cx.pos) / norm(m - cx.pos), 0};
}
Ray SampleRay(const Ctx& cx, int col, int row, int si) {                                        // PixelLBCorner :84-92, getSampleRay :94-113, AddDepthOfField :119-139
    float u = cx.l + col * cx.pw;
    float v = cx.t - (row + 1) * cx.ph;
    V3 lb = cx.pos + cx.gaze * cx.dist;
    lb = lb + u * cx.right;
    lb = lb + v * cx.up;
    int i = si % cx.n, j = si / cx.n;
    Rng4 r = philox(cx.seed, (uint32_t)(row * cx.nx + col), (uint32_t)si, 0, DIM_CAMERA);
    V3 m = lb;
    m = m + (i + r.u[0]) * cx.sw * cx.right;
    m = m + (j + r.u[1]) * cx.sh * cx.up;
    Ray ray{cx.pos, (m - cx.pos) / norm(m - cx.pos), 0};
    if (cx.dof) {
        float xChi = r.u[2] - 0.5f, yChi = r.u[3] - 0.5f;
        V3 q = cx.pos;
        q = q + cx.aperture * xChi * cx.right;
        q = q + cx.aperture * yChi * cx.up;
        V3 dir = normalized(m - cx.pos);
        float t_fd = cx.focus / dot(dir, cx.gaze);
        V3 p = ray.o + ray.d * t_fd;
        return Ray{q, normalized(p - q), 0};
    }
    ray.time = r.u[2];
    return ray;
}

// ------------------------------------------------------------------------------------------ path tracer (no reference code; DESIGN.md §Path tracer)
V3 PathTrace(const Ctx& cx, Ray ray, uint32_t pixel, uint32_t sample, int row, int col);

// Scene::SingleSample / MultiSample, src/Scene.cpp:365-411
V3 RenderPixel(const Ctx& cx, int col, int row) {
    uint32_t pixel = (uint32_t)(row * cx.nx + col);
    if (cx.integrator == RT_INTEGRATOR_PATH) {
        V3 color(0, 0, 0);
        for (int i = 0; i < cx.spp; i++) {
            Ray r = cx.spp > 1 ? SampleRay(cx, col, row, i) : PrimaryRay(cx, col, row);
            color = color + PathTrace(cx, r, pixel, (uint32_t)i, row, col);
        }
        return cx.spp > 1 ? color / (float)cx.spp : color;
    }
    if (cx.spp <= 1) {
        Ray ray = PrimaryRay(cx, col, row);
        Hit h = FindIntersection(cx, ray);
        if (h.full) return Shading(cx, ray, h, pixel, 0);
        return Background(cx, /*row=*/col, /*col=*/row, ray);     // SingleSample(x, y) forwards its mis-named (row=x, col=y)
    }
    V3 color(0, 0, 0);
    for (int i = 0; i < cx.spp; i++) {
        Ray ray = SampleRay(cx, col, row, i);
        Hit h = FindIntersection(cx, ray);
        if (h.full) color = color + Shading(cx, ray, h, pixel, (uint32_t)i);
        else color = color + Background(cx, row, col, ray);
    }
    return color / (float)cx.spp;
}

// ---- path tracer restatement of DESIGN.md §Path tracer (kept in lock-step with csrc/kernels: same decisions, same RNG dims)
V3 PathTrace(const Ctx& cx, Ray ray, uint32_t pixel, uint32_t sample, int row, int col) {
    const RtSceneDesc* s = cx.s;
    int max_depth = cx.opts.pt_max_depth > 0 ? cx.opts.pt_max_depth : s->max_recursion_depth;
    V3 L(0, 0, 0), beta(1, 1, 1);
    bool count_emission = true;      // emission is counted on camera rays and after specular bounces (NEE covers the rest)
    for (uint32_t node = 1; (int)node <= max_depth + 1; node++) {
        Hit h = FindIntersection(cx, ray);
        if (!h.full) {
            // background lookup convention of the renderer: transposed pixel-uv at 1 spp, as the kernels do
            int br = cx.spp <= 1 ? col : row, bc = cx.spp <= 1 ? row : col;
            if (node == 1 || (s->env_light_index != -1 && (count_emission || !cx.opts.pt_nee))) L = L + cmul(beta, Background(cx, br, bc, ray));
            break;
        }
        const RtMaterial& mat = s->materials[h.matIndex - 1];
        V3 Le = ld3(mat.emission) + (h.object >= 0 ? ld3(s->objects[h.object].radiance) : V3(0, 0, 0));
        if ((Le.x > 0 || Le.y > 0 || Le.z > 0) && (count_emission || !cx.opts.pt_nee) && dot(h.normal, ray.d) < 0) L = L + cmul(beta, Le);
        if ((int)node == max_depth + 1) break;
        V3 n = h.normal;
        if (mat.type == RT_MAT_MIRROR || mat.type == RT_MAT_CONDUCTOR) {
            V3 wo = -ray.d;
            V3 wr = normalized(-wo + n * 2 * dot(n, wo));
            float f = mat.type == RT_MAT_CONDUCTOR ? ConductorFresnel(mat.refraction_index, mat.absorption_index, ray.d, n) : 1.0f;
            beta = cmul(beta, ld3(mat.mirror)) * f;
            ray = Ray{h.point + n * s->shadow_ray_eps, wr, ray.time};
            count_emission = true;
            continue;
        }
        if (mat.type == RT_MAT_DIELECTRIC) {
            float dp = dot(ray.d, n), nt = mat.refraction_index;
            float snell = dp < 0 ? 1.0f / nt : nt, n_t = dp < 0 ? nt : 1.0f, n_i = dp < 0 ? 1.0f : nt;
            V3 nn = dp < 0 ? n : -n;
            float ci = -dot(ray.d, nn);
            float sq = 1 - snell * snell * (1 - ci * ci);
            Rng4 r = philox(cx.seed, pixel, sample, node, DIM_PT_BSDF);
            float F = 1.0f;
            V3 tdir;
            if (sq >= 0) {
                tdir = normalized((ray.d + nn * ci) * snell - nn * std::sqrt(sq));
                float ct = -dot(tdir, nn);
                float rPar = (n_t * ci - n_i * ct) / (n_t * ci + n_i * ct), rPer = (n_i * ci - n_t * ct) / (n_i * ci + n_t * ct);
                F = 0.5f * (rPar * rPar + rPer * rPer);
            }
            if (r.u[0] < F) {
                V3 wr = normalized(ray.d + nn * 2 * ci);
                ray = Ray{h.point + nn * s->shadow_ray_eps, wr, ray.time};
            } else {
                if (dp >= 0) {   // leaving the medium: Beer over the inside segment just travelled
                    beta = cmul(beta, V3(exp(-mat.absorption_coeff[0] * h.t), exp(-mat.absorption_coeff[1] * h.t), exp(-mat.absorption_coeff[2] * h.t)));
                }
                ray = Ray{h.point - nn * s->shadow_ray_eps, tdir, ray.time};
            }
            count_emission = true;
            continue;
        }
        // diffuse/glossy surface: NEE + BSDF sampling
        if (dot(n, ray.d) > 0) n = -n;   // shade the side we see
        Hit hs = h; hs.normal = n;
        if (cx.opts.pt_nee) {
            V3 direct(0, 0, 0);
            for (int i = 0; i < s->num_lights; i++) direct = direct + LightShading(cx, i, ray, hs, mat, pixel, sample, node);
            L = L + cmul(beta, direct);
        }
        Rng4 r = philox(cx.seed, pixel, sample, node, DIM_PT_BSDF);
        V3 u = OrthonormalU(n), w = cross(n, u);
        V3 wi; float pdf;
        if (cx.opts.pt_importance) {
            float rr = std::sqrt(r.u[0]), ph = 2.0f * (float)M_PI * r.u[1];
            float x = rr * cos(ph), z = rr * sin(ph), y = std::sqrt(std::max(0.0f, 1.0f - r.u[0]));
            wi = normalized(u * x + n * y + w * z);
            pdf = std::max(1e-8f, dot(wi, n)) / (float)M_PI;
        } else {
            float y = r.u[0], rr = std::sqrt(std::max(0.0f, 1.0f - y * y)), ph = 2.0f * (float)M_PI * r.u[1];
            wi = normalized(u * (rr * cos(ph)) + n * y + w * (rr * sin(ph)));
            pdf = 1.0f / (2.0f * (float)M_PI);
        }
        float cosi = std::max(0.0f, dot(wi, n));
        V3 f = mat.brdf != RT_BRDF_NONE ? TermBRDF(mat, wi, -ray.d, n) : ld3(mat.diffuse) / (float)M_PI;
        beta = cmul(beta, f) * (cosi / pdf);
        if (isnan3(beta)) break;
        count_emission = false;
        if (cx.opts.pt_russian_roulette && node >= 2) {
            float q = cx.opts.pt_russian_roulette == 2 ? cosi : std::max(beta.x, std::max(beta.y, beta.z));   // mode 2: keyed on the cosine (pages/Page7.md:31)
            q = std::min(0.95f, std::max(0.05f, q));
            Rng4 rr = philox(cx.seed, pixel, sample, node, DIM_PT_RR);
            if (rr.u[0] >= q) break;
            beta = beta / q;
        }
        ray = Ray{h.point + n * s->shadow_ray_eps, wi, ray.time};
    }
    return L;
}

void BuildCtx(Ctx& cx, const RtSceneDesc* s, const RtCameraDesc* cam, const RtRenderOpts* opts) {
    cx.s = s;
    cx.seed = opts ? opts->seed : 0;
    cx.integrator = opts ? opts->integrator : 0;
    if (opts) cx.opts = *opts; else memset(&cx.opts, 0, sizeof cx.opts);
    cx.bvh.resize(s->num_meshes);
    for (int m = 0; m < s->num_meshes; m++) {
        Bvh& bv = cx.bvh[m];
        bv.prims.resize(s->meshes[m].tri_count);
        for (int i = 0; i < s->meshes[m].tri_count; i++) bv.prims[i] = s->meshes[m].first_tri + i;
        BuildNode(s, bv, 0, s->meshes[m].tri_count, 0, 0);
    }
    SetupCamera(cx, cam);
    // object lights: world-space sampling records (forward transform = inverse of the stored M^-1, in double; a sphere's radius scales
    // with the cube root of |det M|), as csrc/core.cu builds them
    cx.lgeom.assign(std::max(1, s->num_lights), Ctx::LightGeom());
    for (int li = 0; li < s->num_lights; li++) {
        const RtLight& L = s->lights[li];
        if (L.type != RT_LIGHT_MESH && L.type != RT_LIGHT_SPHERE) continue;
        const RtObject& o = s->objects[L.object];
        double A[3][3], t[3], M[3][3], T[3];
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) A[r][c] = o.inv[r * 4 + c]; t[r] = o.inv[r * 4 + 3]; }
        double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
        M[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / det; M[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / det; M[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / det;
        M[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / det; M[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / det; M[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / det;
        M[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / det; M[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / det; M[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / det;
        for (int r = 0; r < 3; r++) T[r] = -(M[r][0] * t[0] + M[r][1] * t[1] + M[r][2] * t[2]);
        auto xf = [&](const float* p) { double q[3]; for (int r = 0; r < 3; r++) q[r] = M[r][0] * p[0] + M[r][1] * p[1] + M[r][2] * p[2] + T[r]; return V3((float)q[0], (float)q[1], (float)q[2]); };
        Ctx::LightGeom& g = cx.lgeom[li];
        if (L.type == RT_LIGHT_SPHERE) {
            const RtSphere& sp = s->spheres[o.geom];
            double detM = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
            g.c = xf(sp.center); g.R = (float)(sp.radius * std::cbrt(std::fabs(detM)));
        } else {
            const RtMesh& me = s->meshes[o.geom];
            std::vector<double> areas(me.tri_count);
            double total = 0;
            for (int k = 0; k < me.tri_count; k++) {
                const int* ix = s->tri_indices + 3 * (size_t)(me.first_tri + k);
                double P[3][3];
                for (int v = 0; v < 3; v++) for (int r = 0; r < 3; r++) { const float* p = s->vertices + 3 * (size_t)ix[v]; P[v][r] = M[r][0] * p[0] + M[r][1] * p[1] + M[r][2] * p[2] + T[r]; }
                double e1[3] = {P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]}, e2[3] = {P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2]};
                double nn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                areas[k] = 0.5 * std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
                total += areas[k];
                for (int v = 0; v < 3; v++) g.tri.push_back(V3((float)P[v][0], (float)P[v][1], (float)P[v][2]));
            }
            double run = 0;
            g.cdf.resize(me.tri_count);
            for (int k = 0; k < me.tri_count; k++) { run += areas[k]; g.cdf[k] = k == me.tri_count - 1 ? 1.0f : (float)(total > 0 ? run / total : 1.0); }
            g.area = (float)total;
        }
    }
}

}  // namespace

extern "C" {

// Renders rows [row0,row1) x cols [col0,col1) (whole frame if row1/col1 <= 0) into out_rgb (full-frame layout
// (row*nx+col)*3; untouched elsewhere) with `threads` host threads, column-interleaved like the reference
// (src/Scene.cpp:269-292).  Returns 0.
int oracle_render(const RtSceneDesc* s, const RtCameraDesc* cam, const RtRenderOpts* opts, int row0, int row1, int col0, int col1, int threads, float* out_rgb) {
    Ctx cx;
    BuildCtx(cx, s, cam, opts);
    if (row1 <= 0) { row0 = 0; row1 = cam->ny; }
    if (col1 <= 0) { col0 = 0; col1 = cam->nx; }
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++)
        pool.emplace_back([&, k]() {
            for (int y = row0; y < row1; y++)
                for (int x = col0 + k; x < col1; x += threads) {
                    V3 c = RenderPixel(cx, x, y);
                    float* o = out_rgb + ((size_t)y * cam->nx + x) * 3;
                    o[0] = c.x; o[1] = c.y; o[2] = c.z;
                }
        });
    for (auto& t : pool) t.join();
    return 0;
}

// Primary visibility at pixel centres: getPrimaryRay + FindIntersection, like oracle/harness/hits_main.cpp does
// over the real reference.
int oracle_trace_primary(const RtSceneDesc* s, const RtCameraDesc* cam, int threads, RtHit* out) {
    Ctx cx;
    RtRenderOpts o; memset(&o, 0, sizeof o);
    BuildCtx(cx, s, cam, &o);
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++)
        pool.emplace_back([&, k]() {
            for (int y = 0; y < cam->ny; y++)
                for (int x = k; x < cam->nx; x += threads) {
                    Ray ray = PrimaryRay(cx, x, y);
                    Hit h = FindIntersection(cx, ray);
                    RtHit& r = out[(size_t)y * cam->nx + x];
                    memset(&r, 0, sizeof r);
                    r.full = h.full;
                    if (h.full) {
                        r.object = h.object; r.primitive = h.prim; r.t = dot(h.point - ray.o, ray.d);
                        r.point[0] = h.point.x; r.point[1] = h.point.y; r.point[2] = h.point.z;
                        r.normal[0] = h.normal.x; r.normal[1] = h.normal.y; r.normal[2] = h.normal.z;
                    }
                }
        });
    for (auto& t : pool) t.join();
    return 0;
}

// Photographic tone map: CPU statement of DESIGN.md §7 "Tone map" in double precision.  NO reference code exists
// (pages/Page5.md:47-53 names a "global tonemapping operator", :101 and pages/Page7.md:35 say the author's own was
// buggy and never shipped): this restates the Reinhard et al. 2002 global operator with the course's TMOOptions
// (key, burn %), Saturation and Gamma - parity unpinned by the reference, pinned against this independent statement.
//   Lbar = exp(mean log(1e-5 + L)), Ls = key L / Lbar, L_white = centre of the 4096-bin histogram cell (over
//   [0, max Ls]) in which the cumulative pixel count first reaches (100 - burn)% (burn <= 0: max Ls),
//   Ld = Ls (1 + Ls / L_white^2) / (1 + Ls), c' = (c / L)^saturation Ld, out = round(255 clamp(c')^(1/gamma)).
int oracle_tonemap(const float* rgb, int nx, int ny, float key, float burn_percent, float saturation, float gamma, unsigned char* out_rgb8) {
    const int bins = 4096;
    const size_t npix = (size_t)nx * ny;
    auto lum = [&](size_t i) { double L = 0.2126 * (double)rgb[3 * i] + 0.7152 * (double)rgb[3 * i + 1] + 0.0722 * (double)rgb[3 * i + 2]; return L >= 0.0 ? L : 0.0; };
    double sum = 0.0, mx = 0.0;
    for (size_t i = 0; i < npix; i++) { double L = lum(i); sum += std::log(1e-5 + L); mx = std::max(mx, L); }
    double lbar = std::exp(sum / (double)npix);
    double scale = (double)key / std::max(lbar, 1e-20);
    double max_scaled = std::max(mx * scale, 1e-20);
    double lwhite = max_scaled;
    if (burn_percent > 0.f) {
        std::vector<unsigned long long> hist(bins, 0);
        for (size_t i = 0; i < npix; i++) {
            int b = (int)(lum(i) * scale / max_scaled * (bins - 1));
            hist[std::min(bins - 1, std::max(0, b))]++;
        }
        unsigned long long target = (unsigned long long)((double)npix * (100.0 - (double)burn_percent) / 100.0), acc = 0;
        int b = 0;
        for (; b < bins; b++) { acc += hist[b]; if (acc >= target) break; }
        lwhite = std::max(1e-20, (b + 0.5) / (bins - 1) * max_scaled);
    }
    double inv_gamma = 1.0 / std::max((double)gamma, 1e-6);
    for (size_t i = 0; i < npix; i++) {
        double L = 0.2126 * (double)rgb[3 * i] + 0.7152 * (double)rgb[3 * i + 1] + 0.0722 * (double)rgb[3 * i + 2];
        if (!(L > 0.0)) { out_rgb8[3 * i] = out_rgb8[3 * i + 1] = out_rgb8[3 * i + 2] = 0; continue; }
        double Ls = L * scale;
        double Ld = Ls * (1.0 + Ls / (lwhite * lwhite)) / (1.0 + Ls);
        for (int k = 0; k < 3; k++) {
            double v = std::pow(std::max((double)rgb[3 * i + k], 0.0) / L, (double)saturation) * Ld;
            v = std::min(1.0, std::max(0.0, v));
            out_rgb8[3 * i + k] = (unsigned char)(255.0 * std::pow(v, inv_gamma) + 0.5);
        }
    }
    return 0;
}

// Number of closest-hit queries issued since the last reset, to within 1024 per thread (used to express the
// reference arm's throughput in rays/s: the reference itself cannot count).
unsigned long long oracle_ray_queries(int reset) { unsigned long long v = g_ray_queries.load(); if (reset) g_ray_queries.store(0); return v; }

// Known-answer hooks for unit tests (closed forms listed in SURVEY.md §4).
float oracle_perlin(float x, float y, float z, float scale, int nc) { return ComputePerlin(V3(x, y, z), scale, nc); }
void oracle_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, float* out4) { Rng4 r = philox(seed, c0, c1, c2, c3); memcpy(out4, r.u, 16); }
float oracle_conductor_fresnel(float n, float k, float cos_i) { return ConductorFresnel(n, k, V3(0, 0, -1), V3(0, 0, cos_i)); }

}  // extern "C"
