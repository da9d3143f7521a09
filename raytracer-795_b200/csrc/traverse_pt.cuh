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
//
// Measured dead end (kept here so it is not retried blindly): making the phases strictly warp-synchronous (every vote over
// the full warp, blocked lanes idling inside the phase) raises the leaf-phase batch from 3.0 to 10.8 lanes (occlusion) and
// 12.9 to 18.9 (closest hit) but drops the node-step batch from 15.9 to 11.8 and 24.2 to 20.0 lanes because leaf holders
// wait for the slowest searcher: C4 frame 68.7 -> 77.0 ms.  The loose votes over __activemask() below are faster.
// Fifth dead end: prefetch.global.L1 of a postponed leaf's triangles (when the leaf is recorded) and / or of the nodes pushed on
// the stack: frame +0.2 ms (leaves) to +1.5 ms (leaves + pushed nodes); the leaf phases run with few lanes but are not
// waiting on those lines.
// Fourth dead end: sign-specialised 256-byte nodes (every axis stored as [lo4|hi4] and again as [hi4|lo4], the ray's sign bit picks
// the 32-byte half, one LDG.E.256 per axis: 4 load instructions per node instead of 7, no selects).  Correct, but slower:
// occlusion 30.8 -> 33.1 ms (76 instead of 72 registers: 6 instead of 7 CTAs/SM), closest hit 17.6 -> 17.85 ms at unchanged
// occupancy - the node array doubles (59 MB) and with it the L1/L2 footprint per visit; the wavefronts saved do not pay for it.
// Third dead end: per-warp private batches of 32..256 queue entries (one visit to the global ray cursor per batch instead of per
// fetch) - unlike the queue-slot reservations (kernels.cuh block_reserve), the cursor's same-address atomics are not what
// limits these kernels: occlusion stage 30.8 -> 33.4..35.7 ms, closest hit 17.6 -> 17.7..24.3 ms (tail imbalance grows with
// the batch).
// Second dead end: 2 / 4 pending-leaf slots per lane instead of one (sibling leaves of a 4-wide node arrive in groups).  Leaf
// phases drop 64.6 M -> 23.5 M (occlusion, 4 slots) and their batch rises 2.5 -> 4.3 lanes, but the extra speculative node
// visits (8.28 -> 8.92 per ray) cost as much: occlusion stage +1.0 ms, closest hit -0.5 ms, frame +1 ms.
#pragma once
#include "raygen.cuh"
#include "shade.cuh"
#include "traverse.cuh"

namespace rt795 {

constexpr int kFetchThreshold = 20;   // default; TraceIO.flags bits 8..15 override (tuning)

// Compile-time variants of the traversal kernel (template parameter FLAGS)
enum : int {
    TF_QNODES = 1,      // occlusion only: walk the 8-bit quantised copy of the nodes (DScene.nodes4q)
    TF_SMEM_TLAS = 2,   // stage the TLAS nodes and the top-level transform records in shared memory (broadcast LDS instead of LDG)
    TF_EXACT = 4,       // closest hit only, RtRenderOpts.parity_exact: the reference's per-object rules, see below
};

// TF_EXACT - the reference's closest-hit rules the fast path deliberately drops (deviations 1 and 2):
//  * inside one object the winner is the accepted primitive nearest by |t|, and a triangle is accepted from t >= -eps
//    (src/Shape.cpp:330-331, src/BVH.cpp:165-171, 193-203): a hit just BEHIND the origin wins and then
//  * the object counts only if Ray::gett of the winner's point (the x-quotient (p.x - o.x) / d.x, y then z on NaN,
//    src/Ray.cpp:21-36) lies in (0, best) (src/Helper.cpp:41-43) - so that winner masks the whole object;
//  * the hit's t IS that quotient (the world point is o + t' d, src/Helper.cpp:47), rounding noise of the x-quotient included.
// The best hit over the objects visited so far lives in (g_t, g_prim, g_obj, ...); (tmax, hprim, ...) hold the running winner of the
// object the lane is inside (tmax = |t|, tsig = t) and are +inf / none in the TLAS: like the reference, one object's search is not
// bounded by another object's hit (its t is a different, noisier quantity).  Leaving an object merges by the reference's rule; a tie
// between objects goes to the earlier top-level entry.  Equal |t| INSIDE an object still resolves in traversal order, not in the
// order of the reference's own median-split BVH.

struct TraceIO {
    DRayQueue q;               // closest hit: o_time / d_tmax read, hit / hit_obj written (fused raygen: the whole record is written)
    DShadowQueue sq;           // occlusion: shadow queue (read)
    const int32_t* count_ptr;  // device address of the queue size (the producing kernel's counter); NULL: exactly `limit` rays
    int first, limit;          // rays [first, first + min(limit, *count_ptr - first))
    int32_t* cursor;           // global next-ray cursor (zero before the launch)
    float* accum;              // occlusion: pixel accumulator
    int flags;                 // bit 0: dynamic ray fetch, bit 1: postpone leaves, bit 2: L2 prefetch of upcoming records, bits 8..15: fetch
                               // threshold, bits 16..23: CTA chunk size / 256 (0: every warp pulls from the global cursor); tuning knobs -
                               // results do not depend on them
};

// Resident CTAs per SM the compiler must leave room for (A/B knobs, profiles/bench/r02_ab_occupancy_c4.txt).  Closest hit: 7 x 128
// threads = 72 registers (80 -> 72: stage 17.2 -> 16.65 ms on C4 in round 1; 8 CTAs / 64 registers no better).  Occlusion: 8 CTAs
// (64 registers, 16 bytes spilled) 28.68 -> 28.25 ms; 6 CTAs 28.83.
#ifndef RT795_MINB_ANY
#define RT795_MINB_ANY 8
#endif
#ifndef RT795_MINB_CLOSEST
#define RT795_MINB_CLOSEST 7
#endif

template <bool ANY_HIT, bool COUNT, int FLAGS, bool RAYGEN>
RT_D void trace_body(const DScene& sc, const TraceIO& io, DCounters* ctr, const RaygenArgs* rg) {
    constexpr bool QNODES = (FLAGS & TF_QNODES) != 0;
    constexpr bool SMEM = (FLAGS & TF_SMEM_TLAS) != 0;
    constexpr bool EXACT = (FLAGS & TF_EXACT) != 0 && !ANY_HIT;
    const float tnear_min = EXACT ? -sc.int_eps : 0.f;      // EXACT: primitives up to eps behind the origin are candidates
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const float eps = sc.int_eps;
    const float4* const q_o = ANY_HIT ? io.sq.o_time : io.q.o_time;
    const float4* const q_d = ANY_HIT ? io.sq.d_tmax : io.q.d_tmax;
    int total = io.limit;
    if (io.count_ptr) total = max(0, min(io.limit, *io.count_ptr - io.first));

    // north_star "BVH top levels staged in shared memory": every ray starts in the TLAS and re-enters it after each instance,
    // so the TLAS nodes (C4: 6 nodes) and the 64-byte transform records of the top-level entries (C4: 17) are the hottest lines
    // of the scene.  Staged once per CTA in DYNAMIC shared memory sized to exactly these records (C4: 1.9 KB per CTA; a fixed
    // 8 KB array took 56 KB per SM out of the L1 and cost 2 ms per C4 frame).  The host selects this variant only when the top
    // level fits kSmemTlasBytes; otherwise the top level is walked in global memory like the BLASes.
    extern __shared__ float4 s_stage[];
    float4* const s_tlas = s_stage;
    float4* const s_xf = s_stage + (SMEM ? sc.tlas_nodes * 8 : 0);
    const bool tlas_smem = SMEM;
    if (SMEM) {
        const float4* src = sc.nodes4 + (size_t)sc.tlas_root4 * 8;
        for (int i = threadIdx.x; i < sc.tlas_nodes * 8; i += blockDim.x) s_tlas[i] = __ldg(src + i);
        const float4* xs = reinterpret_cast<const float4*>(sc.obj_xf);
        for (int i = threadIdx.x; i < sc.num_objects * 4; i += blockDim.x) s_xf[i] = __ldg(xs + i);
        __syncthreads();
    }

    // Traversal stack: (node, t_entry) packed into 64 bits (one LDL.64 / STL.64 per entry) with the TOP entry cached in
    // registers, so a pop hands out the cached entry and the reload of the next one overlaps the work on the popped node
    // (ncu r01b: 26 % of the samples sat in the pop loop waiting on two dependent local loads with ~7 lanes active).
    // Occlusion queries never cull stack entries (every t_entry is 0), so their stack holds bare 32-bit node refs: half the
    // local-memory wavefronts per push / pop (the L1 wavefront rate is that kernel's busiest unit, ncu r01e).  C4: occlusion
    // stage 35.3 -> 33.5 ms.  (Keeping the first 8..40 entries of each lane's stack in shared memory, [entry][thread] so that
    // divergent depths stay conflict-free, measured SLOWER: occlusion 33.5 -> 37.5 ms, closest hit 17.5 -> 17.8 ms - the
    // shared/local select per access and the L1 capacity given up to the carve-out cost more than the wavefronts saved.)
    unsigned long long stack[ANY_HIT ? 1 : kStackSize];
    int stack32[ANY_HIT ? kStackSize : 1];
    int sp = 0;
    int top_n = kSentinel; float top_t = 0.f;
    int sgx = 0, sgy = 0, sgz = 0;   // 1 where the current (world or object space) direction component is negative
    int node = kSentinel, cur_obj = -1, rayidx = -1;
    float3 o = v3(0.f, 0.f, 0.f), d = o, id = o, ood = o, wo = o, wd = o;
    float time = 0.f, tmax = 0.f;
    float hb = 0.f, hg = 0.f;
    int hprim = -1, hobj = -1;
    float g_t = 0.f, g_b = 0.f, g_g = 0.f, tsig = 0.f;      // EXACT: parked best over earlier objects; signed t of the object's winner
    int g_prim = -1, g_obj = -1;
    bool exhausted = false;
    const int fetch_threshold = (io.flags >> 8) & 0xff ? (io.flags >> 8) & 0xff : kFetchThreshold;
    // CTA chunks: the CTA takes `chunk` consecutive queue entries from the global cursor at a time and its warps pull from that chunk
    // through a shared-memory cursor (one 64-bit word: end << 32 | next), so the four warps of a CTA work on neighbouring rays - rays
    // of neighbouring pixels walk the same subtrees (L1) - and the global cursor sees one atomic per chunk instead of one per refill.
    const int chunk = ((io.flags >> 16) & 0xff) * 256;
    __shared__ unsigned long long s_chunk;
    __shared__ int s_lock, s_done;
    if (chunk) { if (threadIdx.x == 0) { s_chunk = 0ull; s_lock = 0; s_done = 0; } __syncthreads(); }
    unsigned cnt_nodes = 0, cnt_prims = 0;
    // COUNT builds only: which lanes of each warp currently hold an unfinished ray (shared bitmask per warp), to split the
    // inactive lanes at a node step into "busy elsewhere" and "ray finished, waiting for a refill"
    __shared__ unsigned s_alive[4];
    unsigned d_exec = 0, d_alive = 0, d_leafexec = 0, d_leaflanes = 0;
    if (COUNT) { if (lane == 0) s_alive[threadIdx.x >> 5] = 0; __syncwarp(); }

#define RT_SET_RAY(O, D)                                                                               \
    do {                                                                                               \
        o = (O); d = (D); id = safe_rcp(d); ood = v3(o.x * id.x, o.y * id.y, o.z * id.z);              \
        sgx = id.x < 0.f; sgy = id.y < 0.f; sgz = id.z < 0.f;                                          \
    } while (0)
#define RT_PUSH(N, T)                                                                                  \
    do {                                                                                               \
        if (ANY_HIT) { stack32[sp++] = top_n; top_n = (N); }                                           \
        else {                                                                                         \
            stack[sp++] = ((unsigned long long)__float_as_uint(top_t) << 32) | (unsigned)top_n;        \
            top_n = (N); top_t = (T);                                                                  \
        }                                                                                              \
    } while (0)
#define RT_POP1()                                                                                      \
    do {                                                                                               \
        node = top_n;                                                                                  \
        if (node != kSentinel) top_n = stack32[--sp];                                                  \
    } while (0)
#define RT_POP()                                                                                       \
    do { if (ANY_HIT) { RT_POP1(); } else {                                                            \
      bool more_ = true;                                                                               \
      while (more_) {                                                                                  \
        node = top_n;                                                                                  \
        if (node == kSentinel) break;                                                                  \
        float tt_ = top_t;                                                                             \
        unsigned long long e_ = stack[--sp];                                                           \
        top_n = (int)(unsigned)e_; top_t = __uint_as_float((unsigned)(e_ >> 32));                      \
        more_ = !(node == kBlasExit || tt_ < tmax);                                                    \
      } } } while (0)

    // EXACT: Ray::gett on the object-space ray (src/Ray.cpp:21-36); products and sums rounded separately like the reference's build
    auto gett = [&](float3 oo_, float3 dd_, float t_) {
        float px = __fadd_rn(oo_.x, __fmul_rn(dd_.x, t_)), py = __fadd_rn(oo_.y, __fmul_rn(dd_.y, t_)), pz = __fadd_rn(oo_.z, __fmul_rn(dd_.z, t_));
        float q = __fdiv_rn(__fsub_rn(px, oo_.x), dd_.x);
        if (q != q) q = __fdiv_rn(__fsub_rn(py, oo_.y), dd_.y);
        if (q != q) q = __fdiv_rn(__fsub_rn(pz, oo_.z), dd_.z);
        return q;
    };
    // EXACT: leaving an object - its winner (if any) against the parked best, src/Helper.cpp:41-49
#define RT_MERGE_OBJECT()                                                                              \
    do {                                                                                               \
        if (hprim >= 0) {                                                                              \
            float tq_ = gett(o, d, tsig);                                                              \
            /* strict <: the EARLIER top-level entry keeps a tie (the reference loops over entries in order) */ \
            if (tq_ > 0.f && (tq_ < g_t || (tq_ == g_t && g_prim >= 0 && cur_obj < g_obj))) {          \
                g_t = tq_; g_prim = hprim; g_obj = cur_obj; g_b = hb; g_g = hg;                        \
            }                                                                                          \
        }                                                                                              \
        tmax = 3.0e38f; hprim = -1;                                                                    \
    } while (0)

    while (true) {
        // ---------------- retire finished rays, fetch new ones
        bool idle = node == kSentinel;
        if (COUNT && idle) atomicAnd(&s_alive[threadIdx.x >> 5], ~(1u << lane));
        if (ANY_HIT) {
            bool add = idle && rayidx >= 0 && hprim < 0;   // unoccluded: the light's contribution reaches the pixel
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (add) c = RT_QLD(io.sq.c_pixel + rayidx);
            warp_accum(io.accum, add, (uint32_t)__float_as_int(c.w), v3(c));
            if (idle) rayidx = -1;
        } else if (idle && rayidx >= 0) {
            if (EXACT) { RT_QST(io.q.hit + rayidx, make_float4(g_t, g_b, g_g, __int_as_float(g_prim))); RT_QST(io.q.hit_obj + rayidx, g_obj); }
            else { RT_QST(io.q.hit + rayidx, make_float4(tmax, hb, hg, __int_as_float(hprim))); RT_QST(io.q.hit_obj + rayidx, hobj); }
            rayidx = -1;
        }
        unsigned m = __ballot_sync(full, idle);
        if (m && !exhausted) {
            int nreq = __popc(m), mine = __popc(m & lt), leader = __ffs(m) - 1, base = 0, got = 0;
            if (chunk == 0) {
                if ((int)lane == leader) base = atomicAdd(io.cursor, nreq);
                base = __shfl_sync(full, base, leader);
                if (base + nreq >= total) exhausted = true;
                got = max(0, min(nreq, total - base));
            } else {
                if ((int)lane == leader) {
                    unsigned long long w = atomicAdd(&s_chunk, (unsigned long long)nreq);
                    int nx = (int)(unsigned)w, en = (int)(unsigned)(w >> 32);
                    if (nx < en) { base = nx; got = min(nreq, en - nx); }
                    else if (*(volatile int*)&s_done) got = -1;                       // the queue has been handed out completely
                    else if (atomicCAS(&s_lock, 0, 1) == 0) {
                        unsigned long long w2 = *(volatile unsigned long long*)&s_chunk;   // under the lock: somebody may have refilled since
                        if ((int)(unsigned)w2 >= (int)(unsigned)(w2 >> 32)) {
                            int g = atomicAdd(io.cursor, chunk);
                            if (g >= total) { *(volatile int*)&s_done = 1; got = -1; }
                            else {
                                int e = min(g + chunk, total);
                                base = g; got = min(nreq, e - g);
                                atomicExch(&s_chunk, ((unsigned long long)(unsigned)e << 32) | (unsigned)(g + got));
                            }
                        }
                        __threadfence_block();
                        atomicExch(&s_lock, 0);
                    }
                }
                base = __shfl_sync(full, base, leader);
                got = __shfl_sync(full, got, leader);
                if (got < 0) { exhausted = true; got = 0; }
            }
            if (idle && mine < got) {
                rayidx = io.first + base + mine;
                bool live = true;
                if (RAYGEN) {
                    // K1 fused into K2: the lane generates its next primary ray itself and files the queue record the shade stage
                    // reads; slot = work item (no compaction: slots outside the image are filed as RAY_DEAD)
                    uint32_t pixel = 0, sample = 0;
                    live = gen_camera_ray(*rg, rayidx, wo, wd, time, pixel, sample);
                    tmax = 3.0e38f;
                    RT_QST(io.q.o_time + rayidx, make_float4(wo.x, wo.y, wo.z, time));
                    RT_QST(io.q.d_tmax + rayidx, make_float4(wd.x, wd.y, wd.z, 3.0e38f));
                    RT_QST(io.q.w_pixel + rayidx, make_float4(1.f, 1.f, 1.f, __int_as_float((int)pixel)));
                    RT_QST(io.q.meta + rayidx, make_int4((int)sample, meta_pack(rg->max_depth, live ? (RAY_PRIMARY | RAY_SPECULAR) : RAY_DEAD, 0), 1, 0));
                } else {
                    float4 ot = RT_QLD(q_o + rayidx), dt = RT_QLD(q_d + rayidx);
                    if (io.flags & 4) {
                        // the queue is streamed once from HBM and every lane of the warp waits here for its record (ncu r2d: 6.7 % of
                        // the occlusion kernel's samples sit on the first use of `ot`): pull the records that will be fetched
                        // ~2 us from now into L2 while this one is in flight
                        int ahead = rayidx + 16384;
                        if (ahead < io.first + total) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(q_o + ahead));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(q_d + ahead));
                            if (ANY_HIT) asm volatile("prefetch.global.L2 [%0];" ::"l"(io.sq.c_pixel + ahead));
                        }
                    }
                    wo = v3(ot); wd = v3(dt); time = ot.w; tmax = dt.w;
                }
                hprim = -1; hobj = -1; hb = hg = 0.f;
                if (EXACT) { g_t = tmax; g_prim = -1; g_obj = -1; g_b = g_g = 0.f; tmax = 3.0e38f; }   // the ray's own bound applies at the merges only
                if (!live || any_nan(wo) || any_nan(wd)) { node = kSentinel; if (!ANY_HIT) { tmax = 0.f; g_t = 0.f; } }   // src/Helper.cpp:28-30: NaN ray = no hit
                else {
                    RT_SET_RAY(wo, wd);
                    if (COUNT) atomicOr(&s_alive[threadIdx.x >> 5], 1u << lane);
                    sp = 0; top_n = kSentinel; top_t = 0.f;
                    node = sc.tlas_root4; cur_obj = -1;
                }
            }
        }
        if (exhausted && __all_sync(full, node == kSentinel && rayidx < 0)) break;   // (a fetch can come back empty while another warp refills the CTA's chunk)
        // ---------------- traverse
        while (node != kSentinel) {
            bool searching = true;
            int leaf = 0;   // pending triangle leaf (negative ref) or 0
            while (true) {
                if (node == kSentinel) break;
                if (node == kBlasExit) {
                    if (leaf < 0) break;   // the pending leaf lives in this instance's object space: test it before leaving
                    if (EXACT) RT_MERGE_OBJECT();
                    RT_SET_RAY(wo, wd);
                    cur_obj = -1;
                    RT_POP();
                } else if (node >= 0) {
                    // 4-wide inner node: one 128-byte line, four independent slab tests
                    if (COUNT) {
                        ++cnt_nodes;
                        unsigned am = __activemask();
                        if ((int)lane == __ffs(am) - 1) { d_exec++; d_alive += __popc(*(volatile unsigned*)&s_alive[threadIdx.x >> 5]); }
                    }
                    float tn[4]; int rf[4];
                    if (QNODES) {
                        // quantised node: 4 loads; plane = origin + byte * scale, so t = byte * (scale * 1/d) + (origin - o) * 1/d.
                        // byte -> float exactly: the byte is spliced into the mantissa of 2^23 (PRMT), minus 2^23 (FADD on the FMA pipe).
                        const uint4* qp = sc.nodes4q + (size_t)node * 4;
                        uint4 a0 = __ldg(qp), a1 = __ldg(qp + 1), a2 = __ldg(qp + 2), a3 = __ldg(qp + 3);
                        float ax = __uint_as_float(a3.z) * id.x, ay = __uint_as_float(a3.w) * id.y, az = __uint_as_float(a0.w) * id.z;
                        float bx = __uint_as_float(a0.x) * id.x - ood.x, by = __uint_as_float(a0.y) * id.y - ood.y, bz = __uint_as_float(a0.z) * id.z - ood.z;
                        unsigned nxw = sgx ? a2.w : a2.x, fxw = sgx ? a2.x : a2.w, nyw = sgy ? a3.x : a2.y, fyw = sgy ? a2.y : a3.x;
                        unsigned nzw = sgz ? a3.y : a2.z, fzw = sgz ? a2.z : a3.y;
                        rf[0] = (int)a1.x; rf[1] = (int)a1.y; rf[2] = (int)a1.z; rf[3] = (int)a1.w;
#define RT_QB(W, K) (__uint_as_float(__byte_perm((W), 0x4B000000u, 0x7650u | (K))) - 8388608.0f)
#define RT_SLABQ(K)                                                                                                                     \
                        {                                                                                                               \
                            float n_ = fmaxf(fmaxf(RT_QB(nxw, K) * ax + bx, RT_QB(nyw, K) * ay + by), fmaxf(RT_QB(nzw, K) * az + bz, 0.f)); \
                            float f_ = fminf(fminf(RT_QB(fxw, K) * ax + bx, RT_QB(fyw, K) * ay + by), fminf(RT_QB(fzw, K) * az + bz, tmax)); \
                            tn[K] = (n_ <= f_) ? n_ : 3.0e38f;      /* an absent child has the inverted byte range 255..0 */           \
                        }
                        RT_SLABQ(0) RT_SLABQ(1) RT_SLABQ(2) RT_SLABQ(3)
#undef RT_SLABQ
#undef RT_QB
                    } else {
                        // near / far planes are picked by the ray's direction signs through the LOAD ADDRESS (lo and hi of an axis are
                        // adjacent float4s of the same 128-byte line), so no min/max per axis is needed: 6 FMA + 2 three-input min/max
                        // + 2 clamps per child instead of 6 FMA + 10 min/max (the ALU pipe is the busiest pipe of this kernel, ncu r01c).
                        float4 nx4, fx4, ny4, fy4, nz4, fz4, cr;
                        if (SMEM && tlas_smem && cur_obj < 0) {
                            const float4* np = s_tlas + (node - sc.tlas_root4) * 8;
                            nx4 = np[sgx]; fx4 = np[1 - sgx]; ny4 = np[2 + sgy]; fy4 = np[3 - sgy]; nz4 = np[4 + sgz]; fz4 = np[5 - sgz]; cr = np[6];
                        } else {
#if defined(RT795_NODES256)
                            // experiment: 256-byte node with (lo,hi) and (hi,lo) of every axis: 3 x LDG.256 picked by address + 1 x LDG.128
                            const float4* xp = sc.nodes4x + (size_t)node * 16;
                            ldg8(xp + 2 * sgx, nx4, fx4); ldg8(xp + 4 + 2 * sgy, ny4, fy4); ldg8(xp + 8 + 2 * sgz, nz4, fz4); cr = ldg4(xp + 12);
#else
                            const float4* np = sc.nodes4 + (size_t)node * 8;
                            nx4 = ldg4(np + sgx); fx4 = ldg4(np + (1 - sgx)); ny4 = ldg4(np + 2 + sgy); fy4 = ldg4(np + (3 - sgy));
                            nz4 = ldg4(np + 4 + sgz); fz4 = ldg4(np + (5 - sgz)); cr = ldg4(np + 6);
#endif
                        }
#define RT_SLAB(K, C)                                                                                                                   \
                        {                                                                                                                   \
                            float n_ = fmaxf(fmaxf(nx4.C * id.x - ood.x, ny4.C * id.y - ood.y), fmaxf(nz4.C * id.z - ood.z, tnear_min));       \
                            float f_ = fminf(fminf(fx4.C * id.x - ood.x, fy4.C * id.y - ood.y), fminf(fz4.C * id.z - ood.z, tmax));            \
                            rf[K] = __float_as_int(cr.C);                                                                                   \
                            tn[K] = (n_ <= f_ && rf[K] != kNoChildDev) ? n_ : 3.0e38f;                                                      \
                        }
                        RT_SLAB(0, x) RT_SLAB(1, y) RT_SLAB(2, z) RT_SLAB(3, w)
#undef RT_SLAB
                    }
                    if (ANY_HIT) {
                        // occlusion query: visiting order cannot change the answer and stack entries are never culled (t = 0), so
                        // the whole step is straight-line: predicated pushes of all hit children but one, else one pop.
                        bool h0 = tn[0] < 3.0e38f, h1 = tn[1] < 3.0e38f, h2 = tn[2] < 3.0e38f, h3 = tn[3] < 3.0e38f;
                        if (h3 && (h0 || h1 || h2)) RT_PUSH(rf[3], 0.f);
                        if (h2 && (h0 || h1)) RT_PUSH(rf[2], 0.f);
                        if (h1 && h0) RT_PUSH(rf[1], 0.f);
                        if (h0 || h1 || h2 || h3) node = h0 ? rf[0] : (h1 ? rf[1] : (h2 ? rf[2] : rf[3]));
                        else RT_POP1();
                    } else {
                        int nh = (tn[0] < 3.0e38f) + (tn[1] < 3.0e38f) + (tn[2] < 3.0e38f) + (tn[3] < 3.0e38f);
                        if (nh == 0) RT_POP();
                        else {
                            if (nh > 1 || tn[0] >= 3.0e38f) {
                                // sort ascending by entry distance (5-comparator network); misses (3e38) sink to the end
#define RT_CSWAP(A, B) { bool s_ = tn[B] < tn[A]; float ta_ = s_ ? tn[B] : tn[A], tb_ = s_ ? tn[A] : tn[B]; int ra_ = s_ ? rf[B] : rf[A], rb_ = s_ ? rf[A] : rf[B]; tn[A] = ta_; tn[B] = tb_; rf[A] = ra_; rf[B] = rb_; }
                                RT_CSWAP(0, 1) RT_CSWAP(2, 3) RT_CSWAP(0, 2) RT_CSWAP(1, 3) RT_CSWAP(1, 2)
#undef RT_CSWAP
                                if (nh > 3) RT_PUSH(rf[3], tn[3]);
                                if (nh > 2) RT_PUSH(rf[2], tn[2]);
                                if (nh > 1) RT_PUSH(rf[1], tn[1]);
                            }
                            node = rf[0];
                        }
                    }
                } else if (cur_obj < 0) {
                    // TLAS leaf: one top-level entry
                    int obj = (~node) >> 3;
                    float3 oo, od; int root;
                    if (SMEM && tlas_smem) xf_ray_rows(s_xf[obj * 4], s_xf[obj * 4 + 1], s_xf[obj * 4 + 2], s_xf[obj * 4 + 3], wo, wd, time, oo, od, root);
                    else xf_ray(sc.obj_xf + obj, wo, wd, time, oo, od, root);
                    if (root < 0) {
                        if (COUNT) ++cnt_prims;
                        float t;
                        if (EXACT) {
                            // Sphere::bvhIntersect returns its non-negative root; the object then counts iff gett of that point is in (0, best)
                            if (sphere_test(ldg4(sc.spheres + (~root)), oo, od, eps, t)) {
                                float tq = gett(oo, od, t);
                                if (tq > 0.f && (tq < g_t || (tq == g_t && g_prim >= 0 && obj < g_obj))) { g_t = tq; g_prim = 0; g_obj = obj; g_b = g_g = 0.f; }
                            }
                        } else if (sphere_test(ldg4(sc.spheres + (~root)), oo, od, eps, t) && t > 0.f && t < tmax) {
                            tmax = t; hprim = 0; hobj = obj; hb = hg = 0.f;
                            if (ANY_HIT) { node = kSentinel; leaf = 0; break; }
                        }
                        RT_POP();
                    } else {
                        RT_SET_RAY(oo, od);
                        cur_obj = obj;
                        if (EXACT) { tmax = 3.0e38f; hprim = -1; }     // this object's own search, unbounded by the others' hits
                        RT_PUSH(kBlasExit, -1.f);
                        node = root;                       // the entry's BLAS root in nodes4 (DObjectXf.blur_root.w)
                    }
                } else {
                    // triangle leaf: postpone the first one, stop at the second
                    if (leaf < 0) break;
                    leaf = node; searching = false;
                    RT_POP();
                    if (!(io.flags & 2)) break;
                }
                if ((io.flags & 2) && !__any_sync(__activemask(), searching)) break;
            }
            // ---------------- postponed leaves
            while (leaf < 0) {
                if (COUNT) { unsigned am = __activemask(); if ((int)lane == __ffs(am) - 1) { d_leafexec++; d_leaflanes += __popc(am); } }
                int payload = ~leaf;
                int first = payload >> 3, count = (payload & 7) + 1;
                const float4* tp = sc.tris_w + (size_t)first * 3;
                bool done = false;
                for (int i = 0; i < count; i++, tp += 3) {
                    if (COUNT) ++cnt_prims;
                    float t, u, v;
                    if (EXACT) {
                        if (tri_test_affine_exact(tp, o, d, eps, tmax, t, u, v)) { tmax = fabsf(t); tsig = t; hb = u; hg = v; hprim = first + i; hobj = cur_obj; }
                    } else if (tri_test_affine(tp, o, d, eps, tmax, t, u, v)) {
                        tmax = t; hb = u; hg = v; hprim = first + i; hobj = cur_obj;
                        if (ANY_HIT) { done = true; break; }
                    }
                }
                leaf = 0;
                if (ANY_HIT && done) { node = kSentinel; break; }
                if (node < 0 && cur_obj >= 0) { leaf = node; RT_POP(); }   // a second leaf was reached while the first was pending
            }
            if ((io.flags & 1) && !exhausted && __popc(__activemask()) < fetch_threshold) break;
        }
    }
#undef RT_MERGE_OBJECT
#undef RT_POP
#undef RT_POP1
#undef RT_PUSH
#undef RT_SET_RAY
    if (COUNT) {
        atomicAdd(ANY_HIT ? &ctr->node_visits_shadow : &ctr->node_visits, (unsigned long long)cnt_nodes);
        atomicAdd(ANY_HIT ? &ctr->prim_tests_shadow : &ctr->prim_tests, (unsigned long long)cnt_prims);
        int b = ANY_HIT ? 4 : 0;
        atomicAdd(&ctr->diag[b + 0], (unsigned long long)d_exec); atomicAdd(&ctr->diag[b + 1], (unsigned long long)d_alive);
        atomicAdd(&ctr->diag[b + 2], (unsigned long long)d_leafexec); atomicAdd(&ctr->diag[b + 3], (unsigned long long)d_leaflanes);
    }
}

template <bool ANY_HIT, bool COUNT, int FLAGS>
__global__ void __launch_bounds__(128, (COUNT || (FLAGS & TF_QNODES)) ? 0 : (ANY_HIT ? RT795_MINB_ANY : RT795_MINB_CLOSEST)) k_trace_persistent(DScene sc, TraceIO io, DCounters* ctr) {
    trace_body<ANY_HIT, COUNT, FLAGS, false>(sc, io, ctr, nullptr);
}

// K1 + K2 in one launch (level 0 of a wave): work items come from the cursor, rays from the camera.
template <bool COUNT, int FLAGS>
__global__ void __launch_bounds__(128, COUNT ? 0 : RT795_MINB_CLOSEST) k_raygen_trace(DScene sc, TraceIO io, DCounters* ctr, RaygenArgs rg) {
    trace_body<false, COUNT, FLAGS, true>(sc, io, ctr, &rg);
}

}  // namespace rt795
