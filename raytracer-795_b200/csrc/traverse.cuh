// K2/K4 primitives: sphere / triangle intersection, ray transform into an instance, slab-test helpers (device).
// The traversal loop itself is traverse_pt.cuh.
//
// Replaces, per ray: BVHMethods::FindIntersection (src/Helper.cpp:18-80), Transforming::TransformRay (:110-133),
// BVH::FindIntersectionWithBVH + RayBBoxIntersection (src/BVH.cpp:137-266), the geometric part of
// Triangle::bvhIntersect (src/Shape.cpp:297-345) and Sphere::bvhIntersect (:347-398).
//
// Semantics kept: object-space direction is NOT renormalised, so t is the world parameter;
// triangle accepted iff beta >= -eps, gamma >= -eps, beta+gamma <= 1 (eps = IntersectionTestEpsilon); sphere misses
// when the discriminant < eps and takes the smaller non-negative root; a hit counts only for 0 < t < best.
// Deliberate deviations: the true t is used instead of Ray::gett's x-quotient; a triangle hit
// with t in [-eps, 0] does not mask the rest of its object; ties resolve in traversal order.
#pragma once
#include "dscene.h"
#include "vecmath.cuh"

namespace rt795 {

struct HitRec {
    float t, beta, gamma;
    int prim;       // index into the leaf-ordered triangle array (>=0), sphere: 0; miss: -1
    int obj;        // top-level entry
};

constexpr int kBlasExit = 0x7ffffffe;   // stack marker: leaving an instance, restore the world-space ray
constexpr int kSentinel = 0x7fffffff;
constexpr int kNoChildDev = 0x7ffffffd; // == rt795::kNoChild (bvh_build.h)

RT_D float4 ldg4(const float4* p) { return __ldg(p); }
// Ray-queue records are written once and read once (they stream through L2, ~700 MB per wave against the 60 MB scene): they are
// marked evict-first (ld.global.cs / st.global.cs) so that they displace less of the scene.  C4 50.6 -> 50.3 ms, C2 29.8 -> 29.4
// (profiles/bench/r02_ab_stream_hints.txt); -DRT795_NO_STREAM_HINTS builds the plain loads / stores.
#ifndef RT795_NO_STREAM_HINTS
#define RT_QLD(P) __ldcs(P)
#define RT_QST(P, V) __stcs((P), (V))
#else
#define RT_QLD(P) (*(P))
#define RT_QST(P, V) (*(P) = (V))
#endif
// 32 bytes in one instruction (LDG.E.256, sm_100): two adjacent float4s of a 32-byte-aligned pair
RT_D void ldg8(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// Triangle test on a precomputed affine map (Woop et al. 2005) - algebraically the reference's Cramer solve (src/Shape.cpp:297-345: beta = weight of
// b, gamma of c): rows (r_u, r_v, r_w) take a world point p to the triangle's own
// coordinates, p = A + u e1 + v e2 + w n.  The hit is where the ray crosses w = 0; (u, v) there are the reference's (beta, gamma).
// 22 FMA-class operations instead of ~40 (two cross products, four dot products), same predicate.
RT_D bool tri_test_affine(const float4* __restrict__ rows, float3 o, float3 d, float eps, float tmax, float& t, float& u, float& v) {
    float4 rw = ldg4(rows + 2), ru = ldg4(rows), rv = ldg4(rows + 1);
    float ow = rw.x * o.x + rw.y * o.y + rw.z * o.z + rw.w, dw = rw.x * d.x + rw.y * d.y + rw.z * d.z;
    t = __fdividef(-ow, dw);                      // MUFU.RCP + FMUL (<= 2 ulp) instead of the IEEE division with its slow path
    float ou = ru.x * o.x + ru.y * o.y + ru.z * o.z + ru.w, du = ru.x * d.x + ru.y * d.y + ru.z * d.z;
    float ov = rv.x * o.x + rv.y * o.y + rv.z * o.z + rv.w, dv = rv.x * d.x + rv.y * d.y + rv.z * d.z;
    u = ou + t * du; v = ov + t * dv;
    return (t > 0.f) && (t < tmax) && (u >= -eps) && (v >= -eps) && (u + v <= 1.0f);   // NaN (degenerate triangle or parallel ray) fails every test
}

// TF_EXACT (traverse_pt.cuh): the reference's acceptance from t >= -eps, candidates ordered by |t| (`best` = the object's current |t|).
RT_D bool tri_test_affine_exact(const float4* __restrict__ rows, float3 o, float3 d, float eps, float best, float& t, float& u, float& v) {
    float4 rw = ldg4(rows + 2), ru = ldg4(rows), rv = ldg4(rows + 1);
    float ow = rw.x * o.x + rw.y * o.y + rw.z * o.z + rw.w, dw = rw.x * d.x + rw.y * d.y + rw.z * d.z;
    t = -ow / dw;
    float ou = ru.x * o.x + ru.y * o.y + ru.z * o.z + ru.w, du = ru.x * d.x + ru.y * d.y + ru.z * d.z;
    float ov = rv.x * o.x + rv.y * o.y + rv.z * o.z + rv.w, dv = rv.x * d.x + rv.y * d.y + rv.z * d.z;
    u = ou + t * du; v = ov + t * dv;
    return (t >= -eps) && (fabsf(t) < best) && (u >= -eps) && (v >= -eps) && (u + v <= 1.0f);
}

// Sphere::bvhIntersect, src/Shape.cpp:347-398, in object space with the un-normalised direction.
RT_D bool sphere_test(float4 cr, float3 o, float3 d, float eps, float& t) {
    float3 oc = o - v3(cr);
    float doc = dot(d, oc), dd = dot(d, d);
    float disc = doc * doc - dd * (dot(oc, oc) - cr.w * cr.w);
    if (disc < eps || !(disc == disc)) return false;
    float s = sqrtf(disc);
    float t1 = (-doc + s) / dd, t2 = (-doc - s) / dd;
    if (t1 >= 0.f && t2 < 0.f) t = t1;
    else if (t2 >= 0.f && t1 < 0.f) t = t2;
    else if (t1 < 0.f && t2 < 0.f) return false;
    else t = t1 < t2 ? t1 : t2;
    return t == t;
}

RT_D void xf_ray_rows(float4 r0, float4 r1, float4 r2, float4 br, float3 o, float3 d, float time, float3& oo, float3& od, int& root) {
    float3 ob = v3(o.x - br.x * time, o.y - br.y * time, o.z - br.z * time);
    oo = v3(r0.x * ob.x + r0.y * ob.y + r0.z * ob.z + r0.w, r1.x * ob.x + r1.y * ob.y + r1.z * ob.z + r1.w, r2.x * ob.x + r2.y * ob.y + r2.z * ob.z + r2.w);
    od = v3(r0.x * d.x + r0.y * d.y + r0.z * d.z, r1.x * d.x + r1.y * d.y + r1.z * d.z, r2.x * d.x + r2.y * d.y + r2.z * d.z);
    root = __float_as_int(br.w);
}
RT_D void xf_ray(const DObjectXf* __restrict__ x, float3 o, float3 d, float time, float3& oo, float3& od, int& root) {
    xf_ray_rows(ldg4(&x->inv0), ldg4(&x->inv1), ldg4(&x->inv2), ldg4(&x->blur_root), o, d, time, oo, od, root);
}

RT_D float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // MUFU.RCP, <= 1 ulp; no IEEE slow path

RT_D float3 safe_rcp(float3 d) {
    // 1/0 = +-inf is what the slab test wants; avoid 0*inf = NaN at the box faces by nudging exact zeros.
    // Approximate reciprocals: the slab test then walks the boxes with a direction that is off by <= 1 ulp per component,
    // the same size as the FMA roundings of the test itself and far inside the padding of the leaf boxes (core.cu);
    // the IEEE division (MUFU.RCP + Newton step + FCHK + slow-path call, x3 at every instance entry and exit) showed up
    // as 13 CALL sites in the SASS of each traversal kernel.
    const float tiny = 1e-30f;
    return v3(fast_rcp(fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), fast_rcp(fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
              fast_rcp(fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
}

}  // namespace rt795
