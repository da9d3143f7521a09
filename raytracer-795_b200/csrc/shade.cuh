// K3 device code: direct lighting, BRDFs, Whitted continuation rays, path-tracing bounce.
//
// Follows Scene::Shading / RecursiveShading / BasicShading (src/Scene.cpp:148-267), MirrorReflectance (:32-55),
// DielectricRefraction (:57-118), FresnelReflectance / BeerLaw / ConductorFresnel (:120-146), the five light classes
// (src/Light.cpp:168-660) and Light::BRDF / TermBRDF (:62-162).  The reference's recursion becomes weighted rays:
// a ray carries the product of the factors the reference multiplies its subtree by (mirrorRef, Fresnel, 1-Fresnel,
// Beer); leaves add weight * radiance into the pixel accumulator.
#pragma once
#include "surface.cuh"

namespace rt795 {

struct RenderConsts {
    uint64_t seed;
    int integrator;
    int pt_max_depth, pt_nee, pt_rr, pt_importance;
    int queue_cap, shadow_cap;
};

RT_D float3 orthonormal_u(float3 v) {                         // GeometryHelpers::GetOrthonormalUVector, src/Helper.cpp:322-343
    float a = fabsf(v.x), b = fabsf(v.y), c = fabsf(v.z);
    float3 nl = v;
    if (a <= b && a <= c) nl.x = 1.f; else if (b <= a && b <= c) nl.y = 1.f; else nl.z = 1.f;
    return normalize0(cross(v, nl));
}

RT_D float conductor_fresnel(float n_t, float k_t, float3 d, float3 normal) {   // src/Scene.cpp:135-146, src/Light.cpp:18-28
    float cos_t = -dot(d, normal);
    float two = 2.f * n_t * cos_t, c2 = cos_t * cos_t, nk = n_t * n_t + k_t * k_t;
    float rs = (nk - two + c2) / (nk + two + c2);
    float rp = (nk * c2 - two + 1.f) / (nk * c2 + two + 1.f);
    return 0.5f * (rs + rp);
}

// pow(float, int) of the reference (promoted to double there).  Phong exponents are small non-negative integers:
// square-and-multiply (<= 2 log2 p multiplies, within a few ulp of pow) instead of the ~45-instruction powf.
RT_D float ipow(float x, int p) {
    if (p < 0 || p > 65536) return powf(x, (float)p);
    float r = 1.0f;
    unsigned e = (unsigned)p;
    while (e) {
        if (e & 1u) r *= x;
        x *= x;
        e >>= 1;
    }
    return r;
}

RT_D float3 term_brdf(const RtMaterial& m, float3 wi, float3 wo, float3 n) {     // Light::TermBRDF, src/Light.cpp:62-155
    const float pi = 3.14159265358979323846f, eps = 0.001f;
    float3 kd = v3(m.diffuse), ks = v3(m.specular);
    int p = m.phong_exp;
    switch (m.brdf) {
        case RT_BRDF_MP: case RT_BRDF_OP: case RT_BRDF_MPN: {
            float3 wr = -wi + n * 2.f * dot(n, wi);
            wr = wr / length(wr);
            float ca = fmaxf(0.f, dot(wr, wo));
            if (m.brdf == RT_BRDF_MP) return kd + ks * ipow(ca, p);
            if (m.brdf == RT_BRDF_OP) {
                float ci = fmaxf(0.f, dot(wi, n));
                if (ci < eps) return v3(0.f, 0.f, 0.f);
                return kd + (ks * ipow(ca, p)) / ci;
            }
            return kd / pi + ks * ((p + 2) / (2.f * pi)) * ipow(ca, p);
        }
        case RT_BRDF_MBP: case RT_BRDF_OBP: case RT_BRDF_MBPN: {
            float3 h = normalize0(wo + wi);
            float ca = fmaxf(0.f, dot(n, h));
            if (m.brdf == RT_BRDF_MBP) return kd + ks * ipow(ca, p);
            if (m.brdf == RT_BRDF_OBP) {
                float ci = fmaxf(0.f, dot(wi, n));
                if (ci < eps) return v3(0.f, 0.f, 0.f);
                return kd + (ks * ipow(ca, p)) / ci;
            }
            return kd / pi + ks * ((p + 8) / (8.f * pi)) * ipow(ca, p);
        }
        case RT_BRDF_TS: case RT_BRDF_TSF: {
            float3 wh = normalize0(wo + wi);
            float3 dpart = kd / pi;
            float f = 1.f;
            if (m.brdf == RT_BRDF_TSF) { f = conductor_fresnel(m.refraction_index, m.absorption_index, -wo, n); dpart = dpart * (1.f - f); }
            float cosA = dot(wh, n), cosT = dot(wi, n), cosP = dot(wo, n);
            float left = 2.f * dot(n, wh) * dot(n, wo) / dot(wo, wh);
            float right = 2.f * dot(n, wh) * dot(n, wi) / dot(wi, wh);
            float g = fminf(1.f, fminf(left, right));
            float dd = ((p + 2.f) / (2.f * pi)) * ipow(cosA, p);
            float3 sp = (ks * g * dd) / (4.f * cosP * cosT);
            if (m.brdf == RT_BRDF_TSF) sp = sp * f;
            return dpart + sp;
        }
        default: return v3(0.f, 0.f, 0.f);
    }
}

// Light::BRDF (src/Light.cpp:157-162) or *Light::Diffuse + Specular (e.g. :206-236): outgoing radiance for irradiance E from wi.
RT_D float3 surface_response(const RtMaterial& m, const Surface& s, float3 ray_d, float3 wi, float3 E) {
    if (m.brdf != RT_BRDF_NONE) return (E * term_brdf(m, wi, -ray_d, s.N)) * fmaxf(0.f, dot(wi, s.N));
    float alpha = fmaxf(0.f, dot(s.N, wi));
    float3 kd;
    if (s.dm == RT_DM_REPLACE_KD) kd = s.tex_color / s.tex_normalizer;
    else if (s.dm == RT_DM_BLEND_KD) kd = (v3(m.diffuse) + (s.tex_color / s.tex_normalizer)) * 0.5f;
    else kd = v3(m.diffuse);
    float3 diffuse = E * (kd * alpha);
    float3 wo = -ray_d;
    float3 h = (wo + wi) / length(wo + wi);
    float a2 = fmaxf(0.f, dot(s.N, h));
    float3 specular = E * (v3(m.specular) * ipow(a2, m.phong_exp));
    return diffuse + specular;
}

struct LightSample { float3 wi; float3 radiance; float tmax; };   // radiance = reflected radiance if unoccluded (before path weight)

// *Light::BasicShading (src/Light.cpp:238-250, 306-321, 405-436, 522-545, 628-660) minus the shadow query, which is
// emitted as a shadow ray (direction wi from P + N*eps, occluding interval (0, tmax)) - SURVEY A.5.
RT_D bool light_eval(const DScene& sc, int li, const RtMaterial& m, const Surface& s, float3 ray_d, uint64_t seed, uint32_t pixel, uint32_t sample,
                     uint32_t node, LightSample& out) {
    const float pi = 3.14159265358979323846f;
    const RtLight L = sc.lights[li];
    const float eps = sc.shadow_eps;
    float3 P = s.P, N = s.N;
    float3 target;
    bool finite = true, object_light = false;
    float object_tmax = 3.0e38f;
    float falloff = 1.f;
    float3 E;
    switch (L.type) {
        case RT_LIGHT_POINT: {
            target = v3(L.position);
            float dist = length(P - target);
            E = v3(L.intensity) / (dist * dist);
        } break;
        case RT_LIGHT_DIRECTIONAL: {
            out.wi = -normalize0(v3(L.direction));
            E = v3(L.intensity);
            finite = false;
        } break;
        case RT_LIGHT_SPOT: {
            target = v3(L.position);
            float3 dir = normalize0(v3(L.direction));
            float coverage = (L.coverage_deg * 0.5f) * (pi / 180.0f), fall = (L.falloff_deg * 0.5f) * (pi / 180.0f);
            float angle = acosf(dot(normalize0(P - target), dir));
            if (angle < fall) falloff = 1.f;
            else if (angle < coverage) { float cf = cosf(fall), cc = cosf(coverage); float q = (cosf(angle) - cc) / (cf - cc); falloff = (q * q) * (q * q); }
            else return false;
            float dist = length(P - target);
            E = v3(L.intensity) / (dist * dist);
        } break;
        case RT_LIGHT_AREA: {
            float3 ln = normalize0(v3(L.direction));
            float3 u = orthonormal_u(ln), v = cross(ln, u);
            float4 r = philox_uniform4(seed, pixel, sample, node, DIM_LIGHT0 + li * DIM_LIGHT_STRIDE);
            target = v3(L.position) + (u * L.size * (r.x - 0.5f)) + (v * L.size * (r.y - 0.5f));
            float3 dtp = normalize0(P - target);
            float cosT = fabsf(dot(dtp, ln));
            float dsq = length(P - target); dsq = dsq * dsq;
            E = v3(L.intensity) * ((L.size * L.size) * (cosT / dsq));
        } break;
        case RT_LIGHT_SPHERE: {
            // object light, sphere: a direction uniform over the cone the sphere subtends from P (pdf = 1 / (2 pi (1 - cos theta_max)));
            // the shadow interval ends just before the sphere's own surface.  (No reference code: pages/Page7.md:9-13, 137-153.)
            const DLightGeom g = sc.light_geom[li];
            float3 pc = v3(g.sphere) - P;
            float R = g.sphere.w, d2 = dot(pc, pc);
            if (!(d2 > R * R * 1.0002f)) return false;                                 // P on / inside the emitter: no cone
            float d = sqrtf(d2), cosmax = sqrtf(fmaxf(0.f, 1.f - R * R / d2));
            float4 r = philox_uniform4(seed, pixel, sample, node, DIM_LIGHT0 + li * DIM_LIGHT_STRIDE);
            float ct = 1.f - r.x * (1.f - cosmax), st = sqrtf(fmaxf(0.f, 1.f - ct * ct)), sp_, cp_;
            sincosf(2.0f * pi * r.y, &sp_, &cp_);
            float3 w = pc / d, u = orthonormal_u(w), v = cross(w, u);
            float3 dir = normalize0(u * (st * cp_) + v * (st * sp_) + w * ct);
            float th = d * ct - sqrtf(fmaxf(0.f, R * R - d2 * st * st));              // first intersection of P + t dir with the sphere
            target = P + dir * th;
            E = v3(L.intensity) * (2.0f * pi * (1.f - cosmax));
            // where the SHADOW ray (origin P + N eps) meets the sphere: the offset moves that point a lot for grazing directions
            float3 oc = P + N * eps - v3(g.sphere);
            float bq = dot(oc, dir), cq = dot(oc, oc) - R * R, dq = bq * bq - cq;
            object_tmax = dq > 0.f ? fmaxf(0.f, -bq - sqrtf(dq)) : 3.0e38f;           // the offset ray passes the sphere by: nothing of its own to avoid
            object_light = true;
        } break;
        case RT_LIGHT_MESH: {
            // object light, mesh: a triangle with probability proportional to its area, a uniform point on it; only the front side emits
            const DLightGeom g = sc.light_geom[li];
            if (g.tri_count <= 0 || !(g.area > 0.f)) return false;
            float4 r = philox_uniform4(seed, pixel, sample, node, DIM_LIGHT0 + li * DIM_LIGHT_STRIDE);
            const float4* lt = sc.light_tris + (size_t)g.first_tri * 3;
            int lo = 0, hi = g.tri_count - 1;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (r.x < __ldg(lt + 3 * (size_t)mid).w) hi = mid; else lo = mid + 1; }
            float4 a4 = __ldg(lt + 3 * (size_t)lo), b4 = __ldg(lt + 3 * (size_t)lo + 1), c4 = __ldg(lt + 3 * (size_t)lo + 2);
            float su = sqrtf(r.y), b1 = 1.f - su, b2 = r.z * su;
            target = v3(a4) * b1 + v3(b4) * b2 + v3(c4) * (1.f - b1 - b2);
            float3 ln = normalize0(cross(v3(c4) - v3(b4), v3(a4) - v3(b4)));          // the reference's face normal, (c-b) x (a-b)
            float3 dv = P - target;
            float dsq = dot(dv, dv);
            float cosl = dot(ln, dv) * rsqrtf(fmaxf(dsq, 1e-30f));
            if (!(cosl > 0.f)) return false;                                          // P behind the emitting side
            E = v3(L.intensity) * (g.area * cosl / dsq);
            {   // where the shadow ray (origin P + N eps) meets the emitting triangle's plane
                float3 wdir = -dv * rsqrtf(fmaxf(dsq, 1e-30f));
                float den = dot(ln, wdir);
                object_tmax = den < 0.f ? dot(ln, v3(a4) - (P + N * eps)) / den : 3.0e38f;
            }
            object_light = true;
        } break;
        default: {   // RT_LIGHT_ENVIRONMENT
            float3 u = orthonormal_u(N), w = cross(N, u);
            float3 dir = N;
            {
                // The reference draws a point in the unit ball by rejection and keeps the upper half (src/Light.cpp:634-648): the
                // normalised result is uniform over the hemisphere.  Drawn directly here (cos(theta) uniform in [0,1], phi uniform):
                // same distribution, one Philox call, no loop whose trip count diverges across the warp (max of 32 geometric
                // draws ~ 12 iterations); C5 shade stage 94 -> 65 ms.  oracle/restate.cpp does the same.  DESIGN.md §6.
                float4 r = philox_uniform4(seed, pixel, sample, node, DIM_LIGHT0 + li * DIM_LIGHT_STRIDE);
                float c = r.x, sn = sqrtf(fmaxf(0.f, 1.f - c * c)), sp_, cp_;
                sincosf(2.0f * pi * r.y, &sp_, &cp_);
                dir = normalize0(u * (sn * cp_) + N * c + w * (sn * sp_));
            }
            out.wi = dir;
            E = env_lookup(sc, L.image, dir) * 2.0f * pi;
            finite = false;
        } break;
    }
    if (finite) {
        float3 dv = target - P;
        float D = length(dv);
        out.wi = dv / D;
        float sdot = dot(N, out.wi);
        // occluded iff the closest hit h satisfies |P-h| < |P-target| with h = P + N*eps + t*wi  
        out.tmax = -eps * sdot + sqrtf(fmaxf(0.f, eps * eps * sdot * sdot - eps * eps + D * D));
        // the emitter is geometry: the occluding interval ends just before the shadow ray reaches the emitter's own surface
        if (object_light) out.tmax = fminf(out.tmax * 1.0002f, object_tmax) * 0.9998f;
    } else out.tmax = 3.0e38f;
    // the reference normalises (target - P) again for wi (e.g. src/Light.cpp:243); same vector
    out.radiance = surface_response(m, s, ray_d, out.wi, E) * falloff;
    return true;
}

// Scene::GetBackgroundColor, src/Scene.cpp:413-435.  At 1 spp the reference's pixel-uv lookup is transposed.
RT_D float3 background(const DScene& sc, const DCamera& cam, uint32_t pixel, float3 d) {
    if (sc.env_light != -1) {
        const RtLight L = sc.lights[sc.env_light];
        if (L.type != RT_LIGHT_ENVIRONMENT) return v3(sc.background);
        return env_lookup(sc, L.image, d);
    }
    if (sc.background_texture == -1) return v3(sc.background);
    const RtTexture tx = sc.textures[sc.background_texture];
    if (tx.type != RT_TEX_IMAGE) return v3(sc.background);
    int x = (int)(pixel % (uint32_t)cam.nx), y = (int)(pixel / (uint32_t)cam.nx);
    float u, v;
    if (cam.spp <= 1) { u = ((float)y) / cam.nx; v = ((float)x) / cam.ny; }
    else { u = ((float)x) / cam.nx; v = ((float)y) / cam.ny; }
    return image_lookup(sc.images[tx.image], tx.interpolation, u, v);
}

// Warp-aggregated accumulate.  MUST be called by all 32 lanes of the warp (converged).  With >= 32 samples per pixel the
// lanes of a warp mostly add to the SAME pixel; 32 same-address float atomics serialise in one L2 slice (measured: the
// accumulate was 13 % of the occlusion kernel), so lanes that share the leader's pixel are summed with a shuffle butterfly
// and added once; the others fall back to their own atomics.  Only the order of float additions changes.
RT_D void warp_accum(float* accum, bool valid, uint32_t pixel, float3 c) {
    const unsigned full = 0xffffffffu;
    valid = valid && !any_nan(c) && !is_zero(c);
    unsigned vm = __ballot_sync(full, valid);
    if (!vm) return;
    int leader = __ffs(vm) - 1;
    uint32_t lp = __shfl_sync(full, pixel, leader);
    bool same = valid && pixel == lp;
    unsigned sm = __ballot_sync(full, same);
    if (__popc(sm) >= 4) {
        float3 v = same ? c : v3(0.f, 0.f, 0.f);
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            v.x += __shfl_xor_sync(full, v.x, off); v.y += __shfl_xor_sync(full, v.y, off); v.z += __shfl_xor_sync(full, v.z, off);
        }
        if ((int)(threadIdx.x & 31) == leader) {
            float* p = accum + (size_t)lp * 3;
            if (v.x != 0.f) atomicAdd(p, v.x);
            if (v.y != 0.f) atomicAdd(p + 1, v.y);
            if (v.z != 0.f) atomicAdd(p + 2, v.z);
        }
        valid = valid && !same;
    }
    if (valid) {
        float* p = accum + (size_t)pixel * 3;
        if (c.x != 0.f) atomicAdd(p, c.x);
        if (c.y != 0.f) atomicAdd(p + 1, c.y);
        if (c.z != 0.f) atomicAdd(p + 2, c.z);
    }
}

RT_D void accum_add(float* accum, uint32_t pixel, float3 c) {
    if (any_nan(c)) return;                      // NanCheck (src/Scene.cpp:221-228) applied per contribution
    float* p = accum + (size_t)pixel * 3;
    if (c.x != 0.f) atomicAdd(p, c.x);
    if (c.y != 0.f) atomicAdd(p + 1, c.y);
    if (c.z != 0.f) atomicAdd(p + 2, c.z);
}

}  // namespace rt795
