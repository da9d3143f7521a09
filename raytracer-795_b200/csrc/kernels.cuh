// The wavefront kernels K1..K8 for sm_100a.
//   K1 k_raygen   Camera::getPrimaryRay / PixelLBCorner / getSampleRay / AddDepthOfField (src/Camera.cpp:63-139); raygen.cuh.
//                 In a frame K1 runs fused into the first closest-hit launch (traverse_pt.cuh k_raygen_trace); the stand-alone
//                 kernel serves rt_trace_primary and the A/B switch RT795_FUSED_RAYGEN=0
//   K2 k_trace_persistent<closest>   closest hit for a ray queue (traverse_pt.cuh)
//   K3 k_shade    surface evaluation + direct-light shadow-ray emission + continuation rays (shade.cuh)
//   K4 k_trace_persistent<any-hit>   occlusion query + accumulate
//   K5            compaction is fused into K3: warp-ballot slot reservation + one atomicAdd per warp per queue
//   K6 k_shade<RT_INTEGRATOR_PATH>   path-tracing bounce (NEE + Russian roulette)
//   K7 k_tonemap_*  photographic tone map (log-average reduction, histogram white point, per-pixel map)
//   K8 k_resolve  box-filter mean (src/Scene.cpp:409) into Image::_data layout (src/Image.cpp:17-24)
#pragma once
#include "shade.cuh"
#include "traverse_pt.cuh"
#include "traverse_packet.cuh"

namespace rt795 {

constexpr int kBlock = 128;

// ---- queue slot reservation (K5): ballot + prefix inside the warp, per-warp totals in shared memory, ONE atomicAdd per CTA per queue.
// (First version: one atomicAdd per warp.)  A queue counter is a single address, served by a single L2 slice at
// about one atomic per clock; with one atomic per warp the counters of a 4 Mi-ray launch took 130 k (raygen) to 390 k (shade,
// two lights) serialised atomics - raygen 3.6 -> 1.9 ms per C4 frame, and the other pipeline's kernels stop queueing behind them.
// Must be called by every thread of the CTA (kBlock threads); `scratch` is kBlock/32 + 1 ints of shared memory.
RT_D int block_reserve(int32_t* counter, int want /*0,1,2*/, int* scratch) {
    const unsigned full = 0xffffffffu;
    unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned b1 = __ballot_sync(full, want >= 1), b2 = __ballot_sync(full, want >= 2);
    unsigned lt = (1u << lane) - 1u;
    if (lane == 0) scratch[warp] = __popc(b1) + __popc(b2);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) { int c = scratch[w]; scratch[w] = tot; tot += c; }
        scratch[kBlock / 32] = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    return scratch[kBlock / 32] + scratch[warp] + __popc(b1 & lt) + __popc(b2 & lt);
}

// ---- K1 (stand-alone; compacting: slots outside the image take no queue entry)
__global__ void __launch_bounds__(kBlock) k_raygen(RaygenArgs a, int count, DRayQueue q, int32_t* q_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 o = v3(0.f, 0.f, 0.f), d = o;
    float time = 0.f;
    uint32_t pixel = 0, sample = 0;
    bool active = i < count && gen_camera_ray(a, i, o, d, time, pixel, sample);
    __shared__ int s_res[kBlock / 32 + 1];
    int slot = block_reserve(q_count, active ? 1 : 0, s_res);
    if (active) {
        RT_QST(q.o_time + slot, make_float4(o.x, o.y, o.z, time));
        RT_QST(q.d_tmax + slot, make_float4(d.x, d.y, d.z, 3.0e38f));
        RT_QST(q.w_pixel + slot, make_float4(1.f, 1.f, 1.f, __int_as_float((int)pixel)));
        RT_QST(q.meta + slot, make_int4((int)sample, meta_pack(a.max_depth, RAY_PRIMARY | RAY_SPECULAR, 0), 1, 0));
    }
}

// End of a wave: fold the wave's queue sizes into the frame totals and clear the per-wave counters for the next wave on this
// pipeline (one thread block; replaces a handful of memsets and the host's bookkeeping).
// `hist` (2 * (nlevels + 1) ints, may be NULL) receives the wave's level and shadow-queue sizes: the next frame's grid hints.
__global__ void __launch_bounds__(256) k_wave_end(DCounters* c, int nlevels, int32_t* hist) {
    unsigned long long sec = 0, sh = 0;
    for (int l = threadIdx.x; l < nlevels; l += blockDim.x) {
        if (l > 0) sec += (unsigned long long)c->level_count[l];
        sh += (unsigned long long)c->shadow_count[l];
    }
    for (int o = 16; o; o >>= 1) { sec += __shfl_xor_sync(0xffffffffu, sec, o); sh += __shfl_xor_sync(0xffffffffu, sh, o); }
    __shared__ unsigned long long s_sec[8], s_sh[8];
    if ((threadIdx.x & 31) == 0) { s_sec[threadIdx.x >> 5] = sec; s_sh[threadIdx.x >> 5] = sh; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += s_sec[w]; b += s_sh[w]; }
        c->total_secondary += a;
        c->total_shadow += b;
    }
    __syncthreads();
    for (int l = threadIdx.x; l <= nlevels; l += blockDim.x) {
        if (hist) { hist[l] = c->level_count[l]; hist[nlevels + 1 + l] = l < nlevels ? c->shadow_count[l] : 0; }
        c->level_count[l] = 0;
        if (l < nlevels) { c->shadow_count[l] = 0; c->cursor_extend[l] = 0; c->cursor_shadow[l] = 0; }
    }
}

// Ray-tree node ids for the RNG counter: root 1, reflected child 2n, refracted child 2n + 1 - unique while they fit 32 bits (30
// bounces); deeper nodes continue with a multiplicative hash instead of shifting the id out (which would hand every node below
// depth 32 the same stream).  oracle/restate.cpp ChildNode is the twin.
RT_D uint32_t child_node(uint32_t node, uint32_t branch) {
    return node < 0x40000000u ? node * 2u + branch : node * 2654435761u + 0x9E3779B9u + branch;
}

// ---- K3 / K6
struct RayOut { float3 o, d, w; float time; int sample, meta, node; float beer_s; };

RT_D void write_ray(const DRayQueue& q, int slot, uint32_t pixel, const RayOut& r) {
    RT_QST(q.o_time + slot, make_float4(r.o.x, r.o.y, r.o.z, r.time));
    RT_QST(q.d_tmax + slot, make_float4(r.d.x, r.d.y, r.d.z, 3.0e38f));
    RT_QST(q.w_pixel + slot, make_float4(r.w.x, r.w.y, r.w.z, __int_as_float((int)pixel)));
    RT_QST(q.meta + slot, make_int4(r.sample, r.meta, r.node, __float_as_int(r.beer_s)));
}

// 6 resident CTAs (80 registers, 68 bytes spilled): C4 shade stage 10.99 -> 10.34 ms; 5 CTAs (96 registers, no spills) 10.99; 4 CTAs 12.70
// (profiles/bench/r02_ab_occupancy_c4.txt)
#ifndef RT795_SHADE_MIN_BLOCKS
#define RT795_SHADE_MIN_BLOCKS 6
#endif
// Shade pass over rays [first, first + min(limit, *in_count - first)) of queue level `in`; continuation rays go to `out`
// (count in *out_count), shadow rays to `sq` (count in *shadow_count).  The ray count is device-resident; the grid covers `limit`.
struct ShadeIO {
    DRayQueue in, out;
    DShadowQueue sq;
    const int32_t* in_count;   // NULL: exactly `limit` rays
    int first, limit;
    int32_t* out_count;
    int32_t* shadow_count;
    int32_t* overflow;
    float* accum;
};

// LOOP = false: one 128-ray group per CTA, the grid covers `limit` (used when the host knows the ray count: level 0 of a wave, the
// chunked driver).  LOOP = true: grid-stride over the groups with a small fixed grid, for the deeper levels whose size only the
// device knows - sizing the grid by what such a level could hold at most (queue capacity = 32 768 CTAs) costs ~60 us of empty
// CTAs per launch, 10 ms per C2 frame.  ANY loop around this body costs 300 bytes of spills at the 96-register cap (ptxas -v), so
// the looping variant is compiled for 4 resident CTAs (128 registers, no spills).
template <int INTEGRATOR, bool LOOP>
__global__ void __launch_bounds__(kBlock, LOOP ? 4 : RT795_SHADE_MIN_BLOCKS) k_shade(DScene sc, DCamera cam, RenderConsts rc, ShadeIO io) {
    const DRayQueue& in = io.in;
    const DRayQueue& out = io.out;
    const DShadowQueue& sq = io.sq;
    float* const accum = io.accum;
    int count = io.limit;
    if (io.in_count) count = max(0, min(io.limit, *io.in_count - io.first));
    if ((int)(blockIdx.x * kBlock) >= count) return;
  for (int group = blockIdx.x * kBlock; group < count; group += LOOP ? (int)(gridDim.x * kBlock) : 0x40000000) {
    int i = group + threadIdx.x;
    bool active = i < count;
    i += io.first;

    // per-lane state that survives to the (warp-convergent) queue pushes
    int n_out = 0;
    RayOut r0, r1;
    bool lit = false;                 // evaluate lights at this hit
    float3 pend = v3(0.f, 0.f, 0.f);  // this lane's direct accumulate (background / replace_all / ambient / emission), added warp-aggregated below
    bool has_pend = false;
    Surface s;
    RtMaterial mat;
    float3 W = v3(0.f, 0.f, 0.f), rd = W;
    float rtime = 0.f;
    uint32_t pixel = 0, sample = 0, node = 0;

    if (active) {
        float4 ot = RT_QLD(in.o_time + i), dt = RT_QLD(in.d_tmax + i), wp = RT_QLD(in.w_pixel + i), hr = RT_QLD(in.hit + i);
        int4 me = RT_QLD(in.meta + i);
        float3 ro = v3(ot);
        rd = v3(dt); rtime = ot.w;
        W = v3(wp);
        pixel = (uint32_t)__float_as_int(wp.w);
        sample = (uint32_t)me.x; node = (uint32_t)me.z;
        int depth = me.y & 0xff;
        uint32_t flags = ((uint32_t)me.y >> 8) & 0xffu;
        int beer_mat = (int)((uint32_t)me.y >> 16);
        HitRec h;
        h.t = hr.x; h.beta = hr.y; h.gamma = hr.z; h.prim = __float_as_int(hr.w); h.obj = RT_QLD(in.hit_obj + i);

        if (flags & RAY_DEAD) {
            // level-0 slot outside the image (fused raygen): nothing to shade
        } else if (INTEGRATOR == RT_INTEGRATOR_REFERENCE) {
            // ---- Beer factors whose distance is only known after this ray's (or its sibling's) extend step (src/Scene.cpp:110-115)
            if (flags & (RAY_BEER_SELF | RAY_BEER_SIBLING)) {
                float t_b = h.t; bool hit_b = h.prim >= 0;
                float s_b = __int_as_float(me.w);
                if (flags & RAY_BEER_SIBLING) { float4 sh = in.hit[i - 1]; t_b = sh.x; hit_b = __float_as_int(sh.w) >= 0; }
                if (hit_b) {
                    float e = sc.shadow_eps;
                    float dist = sqrtf(fmaxf(0.f, e * e + 2.f * t_b * e * s_b + t_b * t_b));   // |hit - p| with origin p -/+ eps*n
                    const RtMaterial bm = sc.materials[beer_mat - 1];
                    W = W * v3(expf(-bm.absorption_coeff[0] * dist), expf(-bm.absorption_coeff[1] * dist), expf(-bm.absorption_coeff[2] * dist));
                }   // miss / TIR: indeterminate in the reference; defined as distance 0 here and in the oracle
            }
            if (h.prim < 0) {
                if (flags & RAY_PRIMARY) { pend = W * background(sc, cam, pixel, rd); has_pend = true; }   // secondary miss = black (src/Scene.cpp:150-153)
            } else {
                s = eval_surface(sc, ro, rd, rtime, h);
                mat = sc.materials[s.mat_index - 1];
                float3 Le = v3(0.f, 0.f, 0.f);
                if (!is_zero(s.Le) && dot(s.N, rd) < 0.f) Le = W * s.Le;                               // object light seen by this ray (extension; the reference snapshot has none)
                if ((flags & RAY_PRIMARY) && s.dm == RT_DM_REPLACE_ALL) {                              // Scene::Shading, src/Scene.cpp:232-234
                    pend = W * s.tex_color + Le; has_pend = true;
                } else {
                    const float eps = sc.shadow_eps;
                    auto reflect_ray = [&](float3 weight, uint32_t fl, int bmat, float bs) {            // MirrorReflectance, src/Scene.cpp:32-55
                        float3 wo = -rd;
                        float3 wr = -wo + s.N * 2.f * dot(s.N, wo);
                        wr = wr / length(wr);
                        if (mat.is_rough) {
                            float3 u = orthonormal_u(wr), v = cross(wr, u);
                            float4 r = philox_uniform4(rc.seed, pixel, sample, node, DIM_GLOSSY);
                            wr = normalize0(wr + ((u * (r.x - 0.5f) + v * (r.y - 0.5f)) * mat.roughness));
                        }
                        RayOut r;
                        r.o = s.P + s.N * eps; r.d = wr; r.w = weight; r.time = rtime; r.sample = (int)sample;
                        r.meta = meta_pack(depth - 1, fl, bmat); r.node = (int)child_node(node, 0u); r.beer_s = bs;
                        return r;
                    };
                    if (mat.type == RT_MAT_NORMAL || depth <= 0) lit = true;
                    else if (mat.type == RT_MAT_MIRROR) {
                        lit = true;
                        float3 w = W * v3(mat.mirror);
                        if (!is_zero(w) && !any_nan(w)) { r0 = reflect_ray(w, 0u, 0, 0.f); n_out = 1; }
                    } else if (mat.type == RT_MAT_CONDUCTOR) {
                        lit = true;
                        float f = conductor_fresnel(mat.refraction_index, mat.absorption_index, rd, s.N);
                        float3 w = (W * f) * v3(mat.mirror);
                        if (!is_zero(w) && !any_nan(w)) { r0 = reflect_ray(w, 0u, 0, 0.f); n_out = 1; }
                    } else {   // dielectric: DielectricRefraction, src/Scene.cpp:57-118 + RecursiveShading :166-209
                        float dp = dot(rd, s.N);
                        float nt = mat.refraction_index, snell, n_t, n_i;
                        float3 nn;
                        bool entering = dp < 0.f;
                        if (entering) { snell = 1.0f / nt; nn = s.N; n_t = nt; n_i = 1.f; }
                        else { snell = nt; nn = -s.N; n_t = 1.f; n_i = nt; }
                        float cosT = -dot(rd, nn);
                        float3 left = (rd + nn * cosT) * snell;
                        float sq = 1.f - (snell * snell) * (1.f - cosT * cosT);
                        bool tir = sq < 0.f;
                        float3 tdir = normalize0(left - nn * sqrtf(sq));
                        float cos_t = -dot(tdir, nn), cos_i = cosT;
                        float rpar = (n_t * cos_i - n_i * cos_t) / (n_t * cos_i + n_i * cos_t);
                        float rper = (n_i * cos_i - n_t * cos_t) / (n_i * cos_i + n_t * cos_t);
                        float F = 0.5f * (rpar * rpar + rper * rper);
                        RayOut tr;
                        tr.o = s.P - eps * nn; tr.d = tdir; tr.time = rtime; tr.sample = (int)sample; tr.node = (int)child_node(node, 1u);
                        // offset.dir for the Beer distance: origin - p = -eps*nn
                        tr.beer_s = -dot(nn, tdir);
                        if (entering) {
                            lit = true;
                            float3 wt = W * (1.f - F), wr = W * F;
                            tr.w = wt; tr.meta = meta_pack(depth - 1, RAY_BEER_SELF, s.mat_index);
                            bool okt = !is_zero(wt) && !any_nan(wt), okr = !is_zero(wr) && !any_nan(wr);
                            if (okt) { r0 = tr; n_out = 1; }
                            if (okr) { RayOut rr = reflect_ray(wr, 0u, 0, 0.f); if (n_out) r1 = rr; else r0 = rr; n_out++; }
                        } else if (tir) {
                            // beer * Rec(reflected) with the Beer distance of a NaN ray: defined as 0 -> weight unchanged
                            if (!is_zero(W) && !any_nan(W)) { r0 = reflect_ray(W, 0u, 0, 0.f); n_out = 1; }
                        } else {
                            // outside = (1-F) Rec(refracted); reflected = F * beer(|hit(refracted) - p|) * Rec(reflected): the
                            // reflected sibling reads the refracted ray's hit distance from slot-1 in the next shade pass
                            float3 wt = W * (1.f - F), wr = W * F;
                            tr.w = wt; tr.meta = meta_pack(depth - 1, 0u, 0);
                            bool okr = !is_zero(wr) && !any_nan(wr);
                            bool okt = (!is_zero(wt) && !any_nan(wt)) || okr;     // keep the refracted ray alive if its sibling needs its distance
                            if (any_nan(wt)) tr.w = v3(0.f, 0.f, 0.f);
                            if (okt) { r0 = tr; n_out = 1; }
                            if (okr) { r1 = reflect_ray(wr, RAY_BEER_SIBLING, s.mat_index, tr.beer_s); n_out = 2; }
                        }
                    }
                    if (lit) { pend = W * (v3(sc.ambient) * v3(mat.ambient)); has_pend = true; }       // Scene::ambient, src/Scene.cpp:22-30
                    if (!is_zero(Le)) { pend = pend + Le; has_pend = true; }
                }
            }
        } else {
            // ---- K6 path-tracing bounce (no reference code; DESIGN.md §Path tracer; oracle/restate.cpp PathTrace is the twin)
            bool count_emission = (flags & RAY_SPECULAR) != 0;
            int max_depth = rc.pt_max_depth;
            if (h.prim < 0) {
                if (node == 1u || (sc.env_light != -1 && (count_emission || !rc.pt_nee))) { pend = W * background(sc, cam, pixel, rd); has_pend = true; }
            } else {
                s = eval_surface(sc, ro, rd, rtime, h);
                mat = sc.materials[s.mat_index - 1];
                float3 Le = v3(mat.emission) + s.Le;
                if ((Le.x > 0.f || Le.y > 0.f || Le.z > 0.f) && (count_emission || !rc.pt_nee) && dot(s.N, rd) < 0.f) { pend = W * Le; has_pend = true; }
                if ((int)node < max_depth + 1) {
                    const float eps = sc.shadow_eps;
                    float3 n = s.N;
                    RayOut r;
                    r.time = rtime; r.sample = (int)sample; r.node = (int)(node + 1u); r.beer_s = 0.f;
                    if (mat.type == RT_MAT_MIRROR || mat.type == RT_MAT_CONDUCTOR) {
                        float3 wo = -rd;
                        float3 wr = normalize0(-wo + n * 2.f * dot(n, wo));
                        float f = mat.type == RT_MAT_CONDUCTOR ? conductor_fresnel(mat.refraction_index, mat.absorption_index, rd, n) : 1.0f;
                        r.w = (W * v3(mat.mirror)) * f;
                        r.o = s.P + n * eps; r.d = wr; r.meta = meta_pack(0, RAY_SPECULAR, 0);
                        if (!is_zero(r.w) && !any_nan(r.w)) { r0 = r; n_out = 1; }
                    } else if (mat.type == RT_MAT_DIELECTRIC) {
                        float dp = dot(rd, n), nt = mat.refraction_index;
                        float snell = dp < 0.f ? 1.0f / nt : nt, n_t = dp < 0.f ? nt : 1.0f, n_i = dp < 0.f ? 1.0f : nt;
                        float3 nn = dp < 0.f ? n : -n;
                        float ci = -dot(rd, nn);
                        float sq = 1.f - snell * snell * (1.f - ci * ci);
                        float4 u4 = philox_uniform4(rc.seed, pixel, sample, node, DIM_PT_BSDF);
                        float F = 1.0f;
                        float3 tdir = v3(0.f, 0.f, 0.f);
                        if (sq >= 0.f) {
                            tdir = normalize0((rd + nn * ci) * snell - nn * sqrtf(sq));
                            float ct = -dot(tdir, nn);
                            float rpar = (n_t * ci - n_i * ct) / (n_t * ci + n_i * ct), rper = (n_i * ci - n_t * ct) / (n_i * ci + n_t * ct);
                            F = 0.5f * (rpar * rpar + rper * rper);
                        }
                        r.w = W;
                        if (u4.x < F) { r.o = s.P + nn * eps; r.d = normalize0(rd + nn * 2.f * ci); }
                        else {
                            if (dp >= 0.f) r.w = W * v3(expf(-mat.absorption_coeff[0] * h.t), expf(-mat.absorption_coeff[1] * h.t), expf(-mat.absorption_coeff[2] * h.t));
                            r.o = s.P - nn * eps; r.d = tdir;
                        }
                        r.meta = meta_pack(0, RAY_SPECULAR, 0);
                        if (!is_zero(r.w) && !any_nan(r.w)) { r0 = r; n_out = 1; }
                    } else {
                        if (dot(n, rd) > 0.f) n = -n;
                        s.N = n;
                        lit = rc.pt_nee != 0;
                        float4 u4 = philox_uniform4(rc.seed, pixel, sample, node, DIM_PT_BSDF);
                        const float pi = 3.14159265358979323846f;
                        float3 u = orthonormal_u(n), w = cross(n, u);
                        float3 wi; float pdf;
                        if (rc.pt_importance) {
                            float rr = sqrtf(u4.x), ph = 2.0f * pi * u4.y;
                            wi = normalize0(u * (rr * cosf(ph)) + n * sqrtf(fmaxf(0.f, 1.f - u4.x)) + w * (rr * sinf(ph)));
                            pdf = fmaxf(1e-8f, dot(wi, n)) / pi;
                        } else {
                            float y = u4.x, rr = sqrtf(fmaxf(0.f, 1.f - y * y)), ph = 2.0f * pi * u4.y;
                            wi = normalize0(u * (rr * cosf(ph)) + n * y + w * (rr * sinf(ph)));
                            pdf = 1.0f / (2.0f * pi);
                        }
                        float cosi = fmaxf(0.f, dot(wi, n));
                        float3 f = mat.brdf != RT_BRDF_NONE ? term_brdf(mat, wi, -rd, n) : v3(mat.diffuse) / pi;
                        float3 beta = (W * f) * (cosi / pdf);
                        bool alive = !any_nan(beta);
                        if (alive && rc.pt_rr && node >= 2u) {
                            // survival probability: the path throughput (mode 1) or, as the reference's author chose (pages/Page7.md:31), the
                            // cosine between the sampled direction and the normal (mode 2); survivors are divided by it either way (unbiased)
                            float q = rc.pt_rr == 2 ? fminf(0.95f, fmaxf(0.05f, cosi)) : fminf(0.95f, fmaxf(0.05f, fmaxf(beta.x, fmaxf(beta.y, beta.z))));
                            float4 rr4 = philox_uniform4(rc.seed, pixel, sample, node, DIM_PT_RR);
                            if (rr4.x >= q) alive = false; else beta = beta / q;
                        }
                        if (alive && !is_zero(beta)) {
                            r.w = beta; r.o = s.P + n * eps; r.d = wi; r.meta = meta_pack(0, 0u, 0);
                            r0 = r; n_out = 1;
                        }
                    }
                }
            }
        }
    }

    warp_accum(accum, has_pend, pixel, pend);
    // ---- continuation rays -> next queue (pairs stay adjacent: slot, slot+1)
    __shared__ int s_res[kBlock / 32 + 1];
    {
        int slot = block_reserve(io.out_count, n_out, s_res);
        if (n_out) {
            if (slot + n_out > rc.queue_cap) atomicExch(io.overflow, 1);
            else { write_ray(out, slot, pixel, r0); if (n_out == 2) write_ray(out, slot + 1, pixel, r1); }
        }
    }
    // ---- direct lighting: one shadow ray per light with a non-zero unoccluded contribution (Scene::BasicShading loop, src/Scene.cpp:252-264)
    // (CTA-uniform loop: every thread takes part in each light's slot reservation; a light's rays of one CTA end up contiguous
    // in the shadow queue, so the occlusion kernel's warps see rays towards one light from neighbouring hits)
    if (__syncthreads_or(lit)) {
        for (int li = 0; li < sc.num_lights; li++) {
            LightSample ls;
            bool emit = false;
            float3 c = v3(0.f, 0.f, 0.f);
            if (lit && light_eval(sc, li, mat, s, rd, rc.seed, pixel, sample, node, ls)) {
                c = W * ls.radiance;
                emit = !is_zero(c) && !any_nan(c);
            }
            int slot = block_reserve(io.shadow_count, emit ? 1 : 0, s_res);
            if (emit) {
                if (slot >= rc.shadow_cap) atomicExch(io.overflow, 1);
                else {
                    float3 so = s.P + s.N * sc.shadow_eps;
                    RT_QST(sq.o_time + slot, make_float4(so.x, so.y, so.z, rtime));
                    RT_QST(sq.d_tmax + slot, make_float4(ls.wi.x, ls.wi.y, ls.wi.z, ls.tmax));
                    RT_QST(sq.c_pixel + slot, make_float4(c.x, c.y, c.z, __int_as_float((int)pixel)));
                }
            }
        }
    }
    if (!LOOP) break;
    __syncthreads();                                               // s_res is reused by the next group
  }
}

// ---- experiment RT795_NODES256: 128-byte node -> 256-byte node with every axis pair in both orders (see traverse_pt.cuh)
__global__ void k_expand_nodes256(const float4* __restrict__ in, int n, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* s = in + (size_t)i * 8;
    float4* o = out + (size_t)i * 16;
    for (int a = 0; a < 3; a++) {
        float4 lo = s[2 * a], hi = s[2 * a + 1];
        o[4 * a] = lo; o[4 * a + 1] = hi; o[4 * a + 2] = hi; o[4 * a + 3] = lo;
    }
    o[12] = s[6]; o[13] = s[7]; o[14] = o[15] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- {v0, e1, e2} records -> affine ray-to-triangle maps (dscene.h tris_w), one thread per triangle, in double for the inverse
__global__ void k_make_affine_tris(const float4* __restrict__ tris, int n, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = tris[3 * (size_t)i], b = tris[3 * (size_t)i + 1], c = tris[3 * (size_t)i + 2];
    double A[3] = {a.x, a.y, a.z}, e1[3] = {b.x, b.y, b.z}, e2[3] = {c.x, c.y, c.z};
    double nn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    double n2 = nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2];
    float4* o = out + 3 * (size_t)i;
    if (!(n2 > 0.0) || !(n2 < 1.0e300)) {                                // zero-area triangle: w = 1 everywhere, never hit
        o[0] = o[1] = make_float4(0.f, 0.f, 0.f, 0.f); o[2] = make_float4(0.f, 0.f, 0.f, 1.f);
        return;
    }
    double ru[3] = {(e2[1] * nn[2] - e2[2] * nn[1]) / n2, (e2[2] * nn[0] - e2[0] * nn[2]) / n2, (e2[0] * nn[1] - e2[1] * nn[0]) / n2};
    double rv[3] = {(nn[1] * e1[2] - nn[2] * e1[1]) / n2, (nn[2] * e1[0] - nn[0] * e1[2]) / n2, (nn[0] * e1[1] - nn[1] * e1[0]) / n2};
    double rw[3] = {nn[0] / n2, nn[1] / n2, nn[2] / n2};
    o[0] = make_float4((float)ru[0], (float)ru[1], (float)ru[2], (float)-(ru[0] * A[0] + ru[1] * A[1] + ru[2] * A[2]));
    o[1] = make_float4((float)rv[0], (float)rv[1], (float)rv[2], (float)-(rv[0] * A[0] + rv[1] * A[1] + rv[2] * A[2]));
    o[2] = make_float4((float)rw[0], (float)rw[1], (float)rw[2], (float)-(rw[0] * A[0] + rw[1] * A[1] + rw[2] * A[2]));
}

// ---- RtRenderOpts.parity_exact: shadow rays were traced as full closest-hit queries (hits in `hit`); a light's contribution reaches
// the pixel iff nothing was hit inside the ray's interval (src/Light.cpp:188-204: occluded iff |p - L| > |p - hit|)
__global__ void __launch_bounds__(kBlock) k_shadow_resolve(DShadowQueue sq, const float4* __restrict__ hit, const int32_t* count_ptr, int limit, float* accum) {
    int count = limit;
    if (count_ptr) count = max(0, min(limit, *count_ptr));
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool add = i < count && __float_as_int(hit[i].w) < 0;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (add) c = sq.c_pixel[i];
    warp_accum(accum, add, (uint32_t)__float_as_int(c.w), v3(c));
}

// ---- K8
__global__ void k_resolve(const float* __restrict__ accum, float* __restrict__ out, int n3, float inv_spp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) out[i] = accum[i] * inv_spp;
}

// K8 restricted to the tiles this rank owns (RtRenderOpts.resolve_owned_only): `out` may be a peer mapping of the root GPU's
// framebuffer, so the resolve IS the multi-GPU gather - each rank stores its own tiles over NVLink, the other pixels of
// `out` are never touched.  One thread per (owned tile, row, 4-pixel group): 48 contiguous bytes = three 16-byte stores.
__global__ void k_resolve_owned(const float* __restrict__ accum, float* __restrict__ out, DTiling tl, int nx, int ny, float inv_spp) {
    int gx = tl.tw >> 2;                                             // 4-pixel groups per tile row (tile_w is a multiple of 8)
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)tl.local_tiles * tl.th * gx;
    if (i >= n) return;
    int g = (int)(i % gx);
    int row = (int)((i / gx) % tl.th);
    int lt = (int)(i / ((long long)gx * tl.th));
    int k = lt * tl.count + tl.rank;
    int x = (k % tl.tiles_x) * tl.tw + g * 4, y = (k / tl.tiles_x) * tl.th + row;
    if (y >= ny || x >= nx) return;
    size_t base = ((size_t)y * nx + x) * 3;
    if (x + 4 <= nx && (base & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(accum)) & 15) == 0) {
        const float4* a = reinterpret_cast<const float4*>(accum + base);
        float4* o = reinterpret_cast<float4*>(out + base);
#pragma unroll
        for (int j = 0; j < 3; j++) { float4 v = a[j]; o[j] = make_float4(v.x * inv_spp, v.y * inv_spp, v.z * inv_spp, v.w * inv_spp); }
    } else {
        int m = min(4, nx - x) * 3;
        for (int j = 0; j < m; j++) out[base + j] = accum[base + j] * inv_spp;
    }
}

// ---- primary-hit dump for rt_trace_primary (parity harness)
__global__ void k_primary_hits(DScene sc, DRayQueue q, int count, RtHit* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if ((((uint32_t)q.meta[i].y >> 8) & 0xffu) & RAY_DEAD) return;   // slot outside the image
    float4 ot = q.o_time[i], dt = q.d_tmax[i], hr = q.hit[i];
    uint32_t pixel = (uint32_t)__float_as_int(q.w_pixel[i].w);
    RtHit r;
    r.full = 0; r.object = 0; r.primitive = 0; r.t = 0.f;
    r.point[0] = r.point[1] = r.point[2] = 0.f; r.normal[0] = r.normal[1] = r.normal[2] = 0.f;
    HitRec h;
    h.t = hr.x; h.beta = hr.y; h.gamma = hr.z; h.prim = __float_as_int(hr.w); h.obj = q.hit_obj[i];
    if (h.prim >= 0) {
        Surface s = eval_surface(sc, v3(ot), v3(dt), ot.w, h);
        r.full = 1; r.object = h.obj; r.primitive = s.prim_in_mesh; r.t = h.t;
        r.point[0] = s.P.x; r.point[1] = s.P.y; r.point[2] = s.P.z;
        r.normal[0] = s.N.x; r.normal[1] = s.N.y; r.normal[2] = s.N.z;
    }
    out[pixel] = r;
}

// ---- K7 photographic tone map (Reinhard et al. 2002 global operator; DESIGN.md §Tone map)
RT_D float luminance(float r, float g, float b) { return 0.2126f * r + 0.7152f * g + 0.0722f * b; }

__global__ void k_tonemap_logsum(const float* __restrict__ rgb, int npix, double* sum_log, float* max_lum) {
    float acc = 0.f, mx = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        float L = luminance(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        if (!(L >= 0.f)) L = 0.f;
        acc += logf(1e-5f + L);
        mx = fmaxf(mx, L);
    }
    for (int o = 16; o; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(sum_log, (double)acc); atomicMax((int*)max_lum, __float_as_int(mx)); }
}
constexpr int kTmBins = 4096;
__global__ void k_tonemap_hist(const float* __restrict__ rgb, int npix, float scale, float max_scaled, unsigned* hist) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        float L = luminance(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        if (!(L >= 0.f)) L = 0.f;
        float Ls = L * scale;
        int b = (int)(Ls / max_scaled * (kTmBins - 1));
        b = b < 0 ? 0 : (b >= kTmBins ? kTmBins - 1 : b);
        atomicAdd(&hist[b], 1u);
    }
}
__global__ void k_tonemap_map(const float* __restrict__ rgb, int npix, float scale, float lwhite, float saturation, float inv_gamma, uint8_t* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    float L = luminance(r, g, b);
    if (!(L > 0.f)) { out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = 0; return; }
    float Ls = L * scale;
    float Ld = Ls * (1.f + Ls / (lwhite * lwhite)) / (1.f + Ls);
    float c[3] = {r, g, b};
    for (int k = 0; k < 3; k++) {
        float v = powf(fmaxf(c[k], 0.f) / L, saturation) * Ld;
        v = fminf(1.f, fmaxf(0.f, v));
        out[3 * i + k] = (uint8_t)(255.f * powf(v, inv_gamma) + 0.5f);
    }
}

}  // namespace rt795
