// K1 + K2 for camera rays as WARP PACKETS: the 32 rays of a warp (32 samples of one pixel at >= 32 spp, an 8x4 pixel block at 1 spp)
// walk the BVH together on ONE traversal stack.
//
// Why: the per-ray kernel (traverse_pt.cuh) runs 20.7 of 32 lanes per instruction on C4's camera rays (ncu r02) although those rays are
// as coherent as rays get - the triangles are smaller than a pixel, so the samples of a pixel end in different leaves and the lanes drift
// apart in the loop's phases (node step / instance entry and exit / leaf / pop), each with its own stack in local memory.  Here the
// control flow is warp-uniform: every node the packet visits is visited by all live lanes (each tests its OWN ray against the four child
// boxes; a child is entered if ANY lane hits it, nearest-first by the warp-minimum entry distance, culled against the warp-maximum
// t_max), every leaf's triangles are tested by all lanes, node and triangle loads are broadcasts (one L1 wavefront per warp instead of
// one per lane), and there is one stack per warp (kept redundantly in every lane's local memory: coalesced 128-byte pushes / pops).
// A lane does more box and triangle tests than it would alone (the union over the packet), but every instruction runs 32 wide.
// Results: the same closest hit per ray (same box / triangle / sphere tests, same predicate; only the visiting order differs, i.e. ties).
// Only for rays that ARE coherent: level 0 of a wave.  Reflection / refraction / path-tracing bounces and shadow rays keep the per-ray kernel.
#pragma once
#include "traverse_pt.cuh"

namespace rt795 {

__global__ void __launch_bounds__(128, 7) k_raygen_trace_packet(DScene sc, TraceIO io, DCounters* ctr, RaygenArgs rg) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const float eps = sc.int_eps;
    const int total = io.limit;
    int stk_n[kStackSize];
    unsigned stk_t[kStackSize];

    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(io.cursor, 32);
        base = __shfl_sync(full, base, 0);
        if (base >= total) break;
        const int rayidx = io.first + base + (int)lane;
        bool live = base + (int)lane < total;
        float3 wo = v3(0.f, 0.f, 0.f), wd = v3(0.f, 0.f, 1.f);
        float time = 0.f;
        if (live) {
            uint32_t pixel = 0, sample = 0;
            bool in_image = gen_camera_ray(rg, rayidx, wo, wd, time, pixel, sample);
            RT_QST(io.q.o_time + rayidx, make_float4(wo.x, wo.y, wo.z, time));
            RT_QST(io.q.d_tmax + rayidx, make_float4(wd.x, wd.y, wd.z, 3.0e38f));
            RT_QST(io.q.w_pixel + rayidx, make_float4(1.f, 1.f, 1.f, __int_as_float((int)pixel)));
            RT_QST(io.q.meta + rayidx, make_int4((int)sample, meta_pack(rg.max_depth, in_image ? (RAY_PRIMARY | RAY_SPECULAR) : RAY_DEAD, 0), 1, 0));
            live = in_image && !any_nan(wo) && !any_nan(wd);               // src/Helper.cpp:28-30: NaN ray = no hit
        }
        const bool record = base + (int)lane < total;
        float tmax = live ? 3.0e38f : 0.f, hb = 0.f, hg = 0.f;
        int hprim = -1, hobj = -1;
        float3 o = wo, d = wd, id = safe_rcp(wd), ood = v3(wo.x * id.x, wo.y * id.y, wo.z * id.z);
        int sgx = id.x < 0.f, sgy = id.y < 0.f, sgz = id.z < 0.f;
        int sp = 0, node = __any_sync(full, live) ? sc.tlas_root4 : kSentinel, cur_obj = -1;

#define PK_SET_RAY(O, D)                                                                               \
    do {                                                                                               \
        o = (O); d = (D); id = safe_rcp(d); ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);              \
        sgx = id.x < 0.f; sgy = id.y < 0.f; sgz = id.z < 0.f;                                          \
    } while (0)
        // pop: skips entries whose warp-minimum entry distance is not below the warp-maximum t_max any more
#define PK_POP()                                                                                       \
    do {                                                                                               \
        unsigned wmax_ = __reduce_max_sync(full, __float_as_uint(tmax));                               \
        node = kSentinel;                                                                              \
        while (sp > 0) {                                                                               \
            --sp;                                                                                      \
            int n_ = stk_n[sp]; unsigned t_ = stk_t[sp];                                               \
            if (n_ == kBlasExit || t_ < wmax_) { node = n_; break; }                                   \
        }                                                                                              \
    } while (0)

        while (node != kSentinel) {
            if (node == kBlasExit) {
                PK_SET_RAY(wo, wd);
                cur_obj = -1;
                PK_POP();
            } else if (node >= 0) {
                const float4* np = sc.nodes4 + (size_t)node * 8;
                float4 nx4 = ldg4(np + sgx), fx4 = ldg4(np + (1 - sgx)), ny4 = ldg4(np + 2 + sgy), fy4 = ldg4(np + (3 - sgy));
                float4 nz4 = ldg4(np + 4 + sgz), fz4 = ldg4(np + (5 - sgz)), cr = ldg4(np + 6);
                float tn[4]; int rf[4];
#define PK_SLAB(K, C)                                                                                                               \
                {                                                                                                                   \
                    float n_ = fmaxf(fmaxf(nx4.C * id.x - ood.x, ny4.C * id.y - ood.y), fmaxf(nz4.C * id.z - ood.z, 0.f));         \
                    float f_ = fminf(fminf(fx4.C * id.x - ood.x, fy4.C * id.y - ood.y), fminf(fz4.C * id.z - ood.z, tmax));        \
                    rf[K] = __float_as_int(cr.C);                                                                                   \
                    tn[K] = (n_ <= f_ && rf[K] != kNoChildDev) ? n_ : 3.0e38f;      /* a dead lane has tmax = 0: never hits */      \
                }
                PK_SLAB(0, x) PK_SLAB(1, y) PK_SLAB(2, z) PK_SLAB(3, w)
#undef PK_SLAB
                // warp-minimum entry distance per child (non-negative floats order like their bit patterns)
                unsigned m[4];
#pragma unroll
                for (int k = 0; k < 4; k++) m[k] = __reduce_min_sync(full, __float_as_uint(tn[k]));
                const unsigned miss = __float_as_uint(3.0e38f);
                int nh = (m[0] < miss) + (m[1] < miss) + (m[2] < miss) + (m[3] < miss);
                if (nh == 0) PK_POP();
                else {
                    // sort ascending (5-comparator network on warp-uniform values); misses sink to the end
#define PK_CSWAP(A, B) { bool s_ = m[B] < m[A]; unsigned ta_ = s_ ? m[B] : m[A], tb_ = s_ ? m[A] : m[B]; int ra_ = s_ ? rf[B] : rf[A], rb_ = s_ ? rf[A] : rf[B]; m[A] = ta_; m[B] = tb_; rf[A] = ra_; rf[B] = rb_; }
                    PK_CSWAP(0, 1) PK_CSWAP(2, 3) PK_CSWAP(0, 2) PK_CSWAP(1, 3) PK_CSWAP(1, 2)
#undef PK_CSWAP
                    if (nh > 3) { stk_n[sp] = rf[3]; stk_t[sp] = m[3]; sp++; }
                    if (nh > 2) { stk_n[sp] = rf[2]; stk_t[sp] = m[2]; sp++; }
                    if (nh > 1) { stk_n[sp] = rf[1]; stk_t[sp] = m[1]; sp++; }
                    node = rf[0];
                }
            } else if (cur_obj < 0) {
                // TLAS leaf: one top-level entry, entered by the whole packet (each lane with its own time)
                int obj = (~node) >> 3;
                float3 oo, od; int root;
                xf_ray(sc.obj_xf + obj, wo, wd, time, oo, od, root);
                if (root < 0) {
                    float t;
                    if (live && sphere_test(ldg4(sc.spheres + (~root)), oo, od, eps, t) && t > 0.f && t < tmax) { tmax = t; hprim = 0; hobj = obj; hb = hg = 0.f; }
                    PK_POP();
                } else {
                    PK_SET_RAY(oo, od);
                    cur_obj = obj;
                    stk_n[sp] = kBlasExit; stk_t[sp] = 0u; sp++;
                    node = root;
                }
            } else {
                // triangle leaf: every live lane tests every triangle (broadcast loads)
                int payload = ~node;
                int first = payload >> 3, count = (payload & 7) + 1;
                const float4* tp = sc.tris_w + (size_t)first * 3;
                for (int i = 0; i < count; i++, tp += 3) {
                    float t, u, v;
                    if (tri_test_affine(tp, o, d, eps, tmax, t, u, v)) { tmax = t; hb = u; hg = v; hprim = first + i; hobj = cur_obj; }   // dead lanes: tmax = 0 rejects
                }
                PK_POP();
            }
        }
#undef PK_POP
#undef PK_SET_RAY
        if (record) { RT_QST(io.q.hit + rayidx, make_float4(live ? tmax : 0.f, hb, hg, __int_as_float(hprim))); RT_QST(io.q.hit_obj + rayidx, hobj); }
    }
}

}  // namespace rt795
