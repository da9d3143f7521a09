// K2/K4 core: two-level BVH2 traversal + sphere / triangle intersection (device).
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

#define RT_COUNT_NODE() do { if (COUNT) ++cnt_nodes; } while (0)
#define RT_COUNT_PRIM() do { if (COUNT) ++cnt_prims; } while (0)

constexpr int kBlasExit = 0x7ffffffe;   // stack marker: leaving an instance, restore the world-space ray
constexpr int kSentinel = 0x7fffffff;
constexpr int kNoChildDev = 0x7ffffffd; // == rt795::kNoChild (bvh_build.h)

RT_D float4 ldg4(const float4* p) { return __ldg(p); }

// Möller–Trumbore on the {v0, e1, e2} record; algebraically the reference's Cramer solve (beta = weight of b, gamma of c).
RT_D bool tri_test(const float4* __restrict__ tri, float3 o, float3 d, float eps, float tmax, float& t, float& u, float& v, int& prim) {
    float4 a = ldg4(tri), b = ldg4(tri + 1), c = ldg4(tri + 2);
    float3 v0 = v3(a), e1 = v3(b), e2 = v3(c);
    float3 p = cross(d, e2);
    float det = dot(e1, p);
    float inv = 1.0f / det;
    float3 tv = o - v0;
    u = dot(tv, p) * inv;
    float3 q = cross(tv, e1);
    v = dot(d, q) * inv;
    t = dot(e2, q) * inv;
    prim = __float_as_int(a.w);
    return (t > 0.f) && (t < tmax) && (u >= -eps) && (v >= -eps) && (u + v <= 1.0f);   // NaN (det == 0) fails every test
}

// The same test on a precomputed affine map (Woop et al. 2005): rows (r_u, r_v, r_w) take a world point p to the triangle's own
// coordinates, p = A + u e1 + v e2 + w n.  The hit is where the ray crosses w = 0; (u, v) there are the reference's (beta, gamma).
// 22 FMA-class operations instead of ~40 (two cross products, four dot products), same predicate.
RT_D bool tri_test_affine(const float4* __restrict__ rows, float3 o, float3 d, float eps, float tmax, float& t, float& u, float& v) {
    float4 rw = ldg4(rows + 2), ru = ldg4(rows), rv = ldg4(rows + 1);
    float ow = rw.x * o.x + rw.y * o.y + rw.z * o.z + rw.w, dw = rw.x * d.x + rw.y * d.y + rw.z * d.z;
    t = -ow / dw;
    float ou = ru.x * o.x + ru.y * o.y + ru.z * o.z + ru.w, du = ru.x * d.x + ru.y * d.y + ru.z * d.z;
    float ov = rv.x * o.x + rv.y * o.y + rv.z * o.z + rv.w, dv = rv.x * d.x + rv.y * d.y + rv.z * d.z;
    u = ou + t * du; v = ov + t * dv;
    return (t > 0.f) && (t < tmax) && (u >= -eps) && (v >= -eps) && (u + v <= 1.0f);   // NaN (degenerate triangle or parallel ray) fails every test
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

RT_D void xf_ray(const DObjectXf* __restrict__ x, float3 o, float3 d, float time, float3& oo, float3& od, int& root) {
    float4 r0 = ldg4(&x->inv0), r1 = ldg4(&x->inv1), r2 = ldg4(&x->inv2), br = ldg4(&x->blur_root);
    float3 ob = v3(o.x - br.x * time, o.y - br.y * time, o.z - br.z * time);
    oo = v3(r0.x * ob.x + r0.y * ob.y + r0.z * ob.z + r0.w, r1.x * ob.x + r1.y * ob.y + r1.z * ob.z + r1.w, r2.x * ob.x + r2.y * ob.y + r2.z * ob.z + r2.w);
    od = v3(r0.x * d.x + r0.y * d.y + r0.z * d.z, r1.x * d.x + r1.y * d.y + r1.z * d.z, r2.x * d.x + r2.y * d.y + r2.z * d.z);
    root = __float_as_int(br.w);
}

RT_D float3 safe_rcp(float3 d) {
    // 1/0 = +-inf is what the slab test wants; avoid 0*inf = NaN at the box faces by nudging exact zeros
    const float tiny = 1e-30f;
    return v3(1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)), 1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
              1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
}

// ANY_HIT: stop at the first accepted hit with t < tmax (shadow rays).  Returns true if something was hit.
// COUNT: per-ray node-visit / primitive-test counters (RtRenderOpts.collect_counters; SURVEY §8d "measured traffic").
template <bool ANY_HIT, bool COUNT>
RT_D bool trace(const DScene& sc, float3 wo, float3 wd, float time, float tmax, HitRec& hit, unsigned& cnt_nodes, unsigned& cnt_prims) {
    int stack_n[kStackSize];
    float stack_t[kStackSize];
    int sp = 0;
    stack_n[sp] = kSentinel; stack_t[sp] = 0.f; sp++;

    float3 o = wo, d = wd;
    float3 id = safe_rcp(d);
    float3 ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);
    int node = sc.tlas_root;
    int cur_obj = -1;
    hit.prim = -1; hit.obj = -1; hit.t = tmax; hit.beta = 0.f; hit.gamma = 0.f;
    const float eps = sc.int_eps;
    bool found = false;

    while (node != kSentinel) {
        if (node == kBlasExit) {
            o = wo; d = wd; id = safe_rcp(d); ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);
            cur_obj = -1;
            --sp; node = stack_n[sp];
            while (node != kSentinel && node != kBlasExit && stack_t[sp] >= tmax) { --sp; node = stack_n[sp]; }
            continue;
        }
        if (node >= 0) {
            RT_COUNT_NODE();
            const float4* np = sc.nodes + (size_t)node * 4;
            float4 q0 = ldg4(np), q1 = ldg4(np + 1), q2 = ldg4(np + 2), q3 = ldg4(np + 3);
            // slabs: (plane - o) * id = plane*id - o*id
            float c0lox = q0.x * id.x - ood.x, c0hix = q0.y * id.x - ood.x, c0loy = q0.z * id.y - ood.y, c0hiy = q0.w * id.y - ood.y;
            float c1lox = q1.x * id.x - ood.x, c1hix = q1.y * id.x - ood.x, c1loy = q1.z * id.y - ood.y, c1hiy = q1.w * id.y - ood.y;
            float c0loz = q2.x * id.z - ood.z, c0hiz = q2.y * id.z - ood.z, c1loz = q2.z * id.z - ood.z, c1hiz = q2.w * id.z - ood.z;
            float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.f));
            float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tmax));
            float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.f));
            float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tmax));
            int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            bool h0 = (t0n <= t0f) && c0 != kNoChildDev, h1 = (t1n <= t1f) && c1 != kNoChildDev;
            if (h0 && h1) {
                bool swap = t1n < t0n;
                int nearc = swap ? c1 : c0, farc = swap ? c0 : c1;
                float fart = swap ? t0n : t1n;
                stack_n[sp] = farc; stack_t[sp] = fart; sp++;
                node = nearc;
            } else if (h0) node = c0;
            else if (h1) node = c1;
            else {
                --sp; node = stack_n[sp];
                while (node != kSentinel && node != kBlasExit && stack_t[sp] >= tmax) { --sp; node = stack_n[sp]; }
            }
            continue;
        }
        // leaf
        int payload = ~node;
        if (cur_obj < 0) {
            // TLAS leaf: one top-level entry
            int obj = payload >> 3;
            float3 oo, od; int root;
            xf_ray(sc.obj_xf + obj, wo, wd, time, oo, od, root);
            if (root < 0) {
                RT_COUNT_PRIM();
                float t;
                if (sphere_test(ldg4(sc.spheres + (~root)), oo, od, eps, t) && t > 0.f && t < tmax) {
                    tmax = t; hit.t = t; hit.prim = 0; hit.obj = obj; hit.beta = 0.f; hit.gamma = 0.f; found = true;
                    if (ANY_HIT) return true;
                }
                --sp; node = stack_n[sp];
                while (node != kSentinel && node != kBlasExit && stack_t[sp] >= tmax) { --sp; node = stack_n[sp]; }
            } else {
                o = oo; d = od; id = safe_rcp(d); ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);
                cur_obj = obj;
                stack_n[sp] = kBlasExit; stack_t[sp] = -1.f; sp++;
                node = root;
            }
            continue;
        }
        // BLAS leaf: up to 8 triangles
        int first = payload >> 3, count = (payload & 7) + 1;
        const float4* tp = sc.tris + (size_t)first * 3;
        for (int i = 0; i < count; i++, tp += 3) {
            RT_COUNT_PRIM();
            float t, u, v; int prim;
            if (tri_test(tp, o, d, eps, tmax, t, u, v, prim)) {
                tmax = t; hit.t = t; hit.beta = u; hit.gamma = v; hit.prim = first + i; hit.obj = cur_obj; found = true;
                if (ANY_HIT) return true;
            }
        }
        --sp; node = stack_n[sp];
        while (node != kSentinel && node != kBlasExit && stack_t[sp] >= tmax) { --sp; node = stack_n[sp]; }
    }
    return found;
}

}  // namespace rt795
