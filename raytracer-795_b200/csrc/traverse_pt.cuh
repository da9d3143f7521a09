// K2/K4, persistent-warp variant: speculative while-while traversal with postponed leaves and dynamic ray fetch
// (Aila & Laine, "Understanding the Efficiency of Ray Traversal on GPUs", HPG 2009; Aila, Laine & Karras 2012),
// adapted to the two-level (TLAS -> instance transform -> BLAS) layout of dscene.h.
//
// Why: ncu on the first version (profiles/r01a_*): k_extend ran 18.0 and k_shadow 13.7 active threads per warp
// instruction - lanes idle while siblings finish their triangles or their whole ray.  Here (1) a lane that reaches a
// triangle leaf postpones it and keeps descending until every lane of the warp holds a leaf, so box tests and triangle
// tests run warp-wide; (2) warps are persistent and lanes whose ray has terminated pull the next ray from a global
// cursor as soon as fewer than kFetchThreshold lanes are still busy.
// Semantics are exactly those of traverse.cuh (same box / triangle / sphere tests, same hit record).
#pragma once
#include "traverse.cuh"

namespace rt795 {

constexpr int kFetchThreshold = 20;

struct TraceIO {
    const float4* o_time;      // ray queue (read)
    const float4* d_tmax;
    int first, count;          // rays [first, first+count)
    int32_t* cursor;           // global next-ray cursor (zeroed before the launch)
    // closest-hit outputs
    float4* hit;
    int32_t* hit_obj;
    // shadow outputs
    const float4* c_pixel;
    float* accum;
};

RT_D void pt_accum(float* accum, float4 c) {
    float* p = accum + (size_t)(uint32_t)__float_as_int(c.w) * 3;
    if (c.x == c.x && c.y == c.y && c.z == c.z) {
        if (c.x != 0.f) atomicAdd(p, c.x);
        if (c.y != 0.f) atomicAdd(p + 1, c.y);
        if (c.z != 0.f) atomicAdd(p + 2, c.z);
    }
}

template <bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(128) k_trace_persistent(DScene sc, TraceIO io, DCounters* ctr) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const float eps = sc.int_eps;

    int stack_n[kStackSize];
    float stack_t[kStackSize];
    int sp = 0;
    int node = kSentinel, cur_obj = -1, rayidx = -1;
    float3 o = v3(0.f, 0.f, 0.f), d = o, id = o, ood = o, wo = o, wd = o;
    float time = 0.f, tmax = 0.f;
    float hb = 0.f, hg = 0.f;
    int hprim = -1, hobj = -1;
    bool exhausted = false;
    unsigned cnt_nodes = 0, cnt_prims = 0;

#define RT_POP()                                                                                       \
    do {                                                                                               \
        --sp; node = stack_n[sp];                                                                      \
    } while (node != kSentinel && node != kBlasExit && stack_t[sp] >= tmax)

    while (true) {
        // ---------------- retire finished rays, fetch new ones
        bool idle = node == kSentinel;
        if (idle && rayidx >= 0) {
            if (ANY_HIT) { if (hprim < 0) pt_accum(io.accum, __ldg(io.c_pixel + rayidx)); }
            else { io.hit[rayidx] = make_float4(tmax, hb, hg, __int_as_float(hprim)); io.hit_obj[rayidx] = hobj; }
            rayidx = -1;
        }
        unsigned m = __ballot_sync(full, idle);
        if (m && !exhausted) {
            int nreq = __popc(m), mine = __popc(m & lt), leader = __ffs(m) - 1, base = 0;
            if ((int)lane == leader) base = atomicAdd(io.cursor, nreq);
            base = __shfl_sync(full, base, leader);
            if (base + nreq >= io.count) exhausted = true;
            if (idle && base + mine < io.count) {
                rayidx = io.first + base + mine;
                float4 ot = io.o_time[rayidx], dt = io.d_tmax[rayidx];
                wo = v3(ot); wd = v3(dt); time = ot.w; tmax = dt.w;
                hprim = -1; hobj = -1; hb = hg = 0.f;
                if (any_nan(wo) || any_nan(wd)) { node = kSentinel; if (!ANY_HIT) tmax = 0.f; }   // src/Helper.cpp:28-30: NaN ray = no hit
                else {
                    o = wo; d = wd; id = safe_rcp(d); ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);
                    sp = 0; stack_n[sp] = kSentinel; stack_t[sp] = 0.f; sp++;
                    node = sc.tlas_root; cur_obj = -1;
                }
            }
        }
        if (__all_sync(full, node == kSentinel && rayidx < 0)) break;
        // ---------------- traverse
        while (node != kSentinel) {
            bool searching = true;
            int leaf = 0;   // pending triangle leaf (negative ref) or 0
            while (true) {
                if (node == kSentinel) break;
                if (node == kBlasExit) {
                    if (leaf < 0) break;   // the pending leaf lives in this instance's object space: test it before leaving
                    o = wo; d = wd; id = safe_rcp(d); ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);
                    cur_obj = -1;
                    RT_POP();
                } else if (node >= 0) {
                    if (COUNT) ++cnt_nodes;
                    const float4* np = sc.nodes + (size_t)node * 4;
                    float4 q0 = ldg4(np), q1 = ldg4(np + 1), q2 = ldg4(np + 2), q3 = ldg4(np + 3);
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
                        stack_n[sp] = swap ? c0 : c1; stack_t[sp] = swap ? t0n : t1n; sp++;
                        node = swap ? c1 : c0;
                    } else if (h0) node = c0;
                    else if (h1) node = c1;
                    else RT_POP();
                } else if (cur_obj < 0) {
                    // TLAS leaf: one top-level entry
                    int obj = (~node) >> 3;
                    float3 oo, od; int root;
                    xf_ray(sc.obj_xf + obj, wo, wd, time, oo, od, root);
                    if (root < 0) {
                        if (COUNT) ++cnt_prims;
                        float t;
                        if (sphere_test(ldg4(sc.spheres + (~root)), oo, od, eps, t) && t > 0.f && t < tmax) {
                            tmax = t; hprim = 0; hobj = obj; hb = hg = 0.f;
                            if (ANY_HIT) { node = kSentinel; leaf = 0; break; }
                        }
                        RT_POP();
                    } else {
                        o = oo; d = od; id = safe_rcp(d); ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);
                        cur_obj = obj;
                        stack_n[sp] = kBlasExit; stack_t[sp] = -1.f; sp++;
                        node = root;
                    }
                } else {
                    // triangle leaf: postpone the first one, stop at the second
                    if (leaf < 0) break;
                    leaf = node; searching = false;
                    RT_POP();
                }
                if (!__any_sync(__activemask(), searching)) break;
            }
            // ---------------- postponed leaves
            while (leaf < 0) {
                int payload = ~leaf;
                int first = payload >> 3, count = (payload & 7) + 1;
                const float4* tp = sc.tris + (size_t)first * 3;
                bool done = false;
                for (int i = 0; i < count; i++, tp += 3) {
                    if (COUNT) ++cnt_prims;
                    float t, u, v; int prim;
                    if (tri_test(tp, o, d, eps, tmax, t, u, v, prim)) {
                        tmax = t; hb = u; hg = v; hprim = first + i; hobj = cur_obj;
                        if (ANY_HIT) { done = true; break; }
                    }
                }
                leaf = 0;
                if (ANY_HIT && done) { node = kSentinel; break; }
                if (node < 0 && cur_obj >= 0) { leaf = node; RT_POP(); }   // a second leaf was reached while the first was pending
            }
            if (!exhausted && __popc(__activemask()) < kFetchThreshold) break;
        }
    }
#undef RT_POP
    if (COUNT) {
        atomicAdd(ANY_HIT ? &ctr->node_visits_shadow : &ctr->node_visits, (unsigned long long)cnt_nodes);
        atomicAdd(ANY_HIT ? &ctr->prim_tests_shadow : &ctr->prim_tests, (unsigned long long)cnt_prims);
    }
}

}  // namespace rt795
