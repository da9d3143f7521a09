// Surface attributes at the closest hit (device), evaluated ONCE per hit in the shade stage.
//
// The reference evaluates these for every candidate hit inside bvhIntersect (src/Shape.cpp:297-616); deferring them
// to the closest hit gives the identical result.  Follows:
//   geometric / smooth normal            Triangle::bvhIntersect src/Shape.cpp:318-334, Sphere :392
//   Triangle::TextureComputation         src/Shape.cpp:505-616
//   Sphere::TextureComputation           src/Shape.cpp:400-503
//   Texture::GetColorAtPixel / GetColorAtCoordinates / GetChangeAtCoordinates   src/Texture.cpp:41-131
//   Perlin::ComputePerlin / GetGradient  src/Perlin.cpp:27-97
//   Transforming::TransformNormal        src/Helper.cpp:100-108
#pragma once
#include "../../include/rt795.h"
#include "dscene.h"
#include "traverse.cuh"
#include "vecmath.cuh"

namespace rt795 {

struct Surface {
    float3 P;            // world hit point  (ray.getPoint(t), src/Helper.cpp:47)
    float3 N;            // world shading normal, normalised
    float3 tex_color;    // ReturnVal::textureColor
    float tex_normalizer;
    int dm;              // ReturnVal::dm
    int mat_index;       // 1-based
    int prim_in_mesh;    // triangle index within its mesh (face order); 0 for spheres
    float3 Le;           // emitted radiance of an object light (<LightMesh> / <LightSphere>), else 0
};

// Perlin tables of the reference (src/Perlin.cpp:4-25) packed into immediates: no constant-memory lookups with
// lane-divergent indices (those serialise), just shifts.
//   gradient component = ((bits >> (2*idx)) & 3) - 1, idx in [0,16):  table = {1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{1,0,1},{-1,0,1},{1,0,-1},
//   {-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1},{1,1,0},{-1,1,0},{0,-1,1},{0,-1,-1}
//   permutation = 4-bit nibbles of {12,7,15,6,11,0,4,9 | 13,3,14,8,2,5,1,10}
RT_D int perlin_shuffled(int i) {                      // i % 16 folded to [0,16) (src/Perlin.cpp:91-97)
    unsigned k = (unsigned)i & 15u;
    unsigned w = (k & 8u) ? 0xa1528e3du : 0x940b6f7cu;
    return (int)((w >> ((k & 7u) * 4u)) & 15u);
}
RT_D float3 perlin_grad(int idx) {
    // 2-bit codes: 0 -> -1, 1 -> 0, 2 -> +1
    unsigned sh = (unsigned)idx * 2u;
    const unsigned X = 0x52552222u, Y = 0xa22550au, Z = 0x250a0a55u;
    return v3((float)((int)((X >> sh) & 3u) - 1), (float)((int)((Y >> sh) & 3u) - 1), (float)((int)((Z >> sh) & 3u) - 1));
}
RT_D float perlin_weight(float x) {
    x = fabsf(x);
    float x3 = x * x * x;
    return (-6.0f * x3 * x * x) + (15.0f * x3 * x) - (10.0f * x3) + 1.0f;
}
// Perlin::ComputePerlin, src/Perlin.cpp:52-84.  Same corner order and the same per-corner arithmetic
// (dot(g, d) * ((wx*wy)*wz), accumulated i-major, k-minor); only the shared sub-expressions (per-axis weights, the
// nested permutation lookups) are evaluated once instead of once per corner.  Split in three so that the bump gradient
// below can share what its four evaluations have in common:
//   perlin_codes: the eight corner gradient indices of a lattice cell, 4 bits each (corner c = 4i + 2j + k in bits 4c..4c+3)
//   perlin_sum:   the weighted corner sum            perlin_convert: the <NoiseConversion>
RT_D unsigned perlin_codes(int ii, int jj, int kk) {
    int pz[2] = {perlin_shuffled(kk), perlin_shuffled(kk + 1)};
    unsigned codes = 0u;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++)
                codes |= (unsigned)perlin_shuffled(ii + i + perlin_shuffled(jj + j + pz[k])) << (4 * (4 * i + 2 * j + k));
    return codes;
}
RT_D float perlin_sum(unsigned codes, const float* dx, const float* dy, const float* dz, const float* wx, const float* wy, const float* wz) {
    float value = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float3 g = perlin_grad((int)((codes >> (4 * (4 * i + 2 * j + k))) & 15u));
                value += dot(g, v3(dx[i], dy[j], dz[k])) * (wx[i] * wy[j] * wz[k]);
            }
    return value;
}
RT_D float perlin_convert(float value, int nc) {
    if (nc == RT_NC_LINEAR) value = (value + 1.0f) * 0.5f;
    else if (nc == RT_NC_ABSVAL) value = fabsf(value);
    return value;
}
// one axis of a sample point: lattice index, the two offsets and their weights
RT_D void perlin_axis(float x, int& cell, float* d, float* w) {
    cell = (int)floorf(x);
    d[0] = x - (float)cell; d[1] = x - (float)(cell + 1);
    w[0] = perlin_weight(d[0]); w[1] = perlin_weight(d[1]);
}
RT_D float perlin(float3 p, float scale, int nc) {
    float3 pt = p * scale;
    int ii, jj, kk;
    float dx[2], dy[2], dz[2], wx[2], wy[2], wz[2];
    perlin_axis(pt.x, ii, dx, wx); perlin_axis(pt.y, jj, dy, wy); perlin_axis(pt.z, kk, dz, wz);
    return perlin_convert(perlin_sum(perlin_codes(ii, jj, kk), dx, dy, dz, wx, wy, wz), nc);
}
// Bump gradient by forward differences (src/Shape.cpp:486-494: four ComputePerlin calls, eps = 0.001).  The three displaced points
// differ from p in one coordinate, by much less than a lattice cell: each re-evaluates only that axis (offsets + weights) and re-uses
// the other two axes and - unless it crossed into the next cell - the cell's eight gradient indices.  Same values as four full
// evaluations (same expressions per corner), about 40 % of their instructions.
RT_D float3 perlin_gradient(float3 p, float scale, int nc) {
    const float eps = 0.001f;
    float3 pt = p * scale;
    int ii, jj, kk, c2;
    float dx[2], dy[2], dz[2], wx[2], wy[2], wz[2], d2[2], w2[2];
    perlin_axis(pt.x, ii, dx, wx); perlin_axis(pt.y, jj, dy, wy); perlin_axis(pt.z, kk, dz, wz);
    const unsigned codes = perlin_codes(ii, jj, kk);
    float v0 = perlin_convert(perlin_sum(codes, dx, dy, dz, wx, wy, wz), nc);
    perlin_axis((p.x + eps) * scale, c2, d2, w2);
    float x = (perlin_convert(perlin_sum(c2 == ii ? codes : perlin_codes(c2, jj, kk), d2, dy, dz, w2, wy, wz), nc) - v0) / eps;
    perlin_axis((p.y + eps) * scale, c2, d2, w2);
    float y = (perlin_convert(perlin_sum(c2 == jj ? codes : perlin_codes(ii, c2, kk), dx, d2, dz, wx, w2, wz), nc) - v0) / eps;
    perlin_axis((p.z + eps) * scale, c2, d2, w2);
    float z = (perlin_convert(perlin_sum(c2 == kk ? codes : perlin_codes(ii, jj, c2), dx, dy, d2, wx, wy, w2), nc) - v0) / eps;
    return v3(x, y, z);
}

RT_D float3 image_pixel(const DImage& im, int i, int j) {
    i = i < 0 ? 0 : (i >= im.width ? im.width - 1 : i);
    j = j < 0 ? 0 : (j >= im.height ? im.height - 1 : j);
    size_t o = (size_t)j * im.width + i;
    if (im.format == RT_IMG_RGBA8) {
        uchar4 p = __ldg((const uchar4*)im.pixels + o);
        return v3((float)p.x, (float)p.y, (float)p.z);
    }
    float4 p = __ldg((const float4*)im.pixels + o);
    return v3(p);
}
RT_D float3 image_lookup(const DImage& im, int interp, float u, float v) {
    u = u - floorf(u); v = v - floorf(v);
    float i = u * im.width, j = v * im.height;
    if (interp == RT_INTERP_NN) return image_pixel(im, (int)i, (int)j);
    int li = (int)floorf(i), lj = (int)floorf(j);
    float a = i - li, b = j - lj;
    return (1 - a) * (1 - b) * image_pixel(im, li, lj) + (1 - a) * b * image_pixel(im, li, lj + 1) + a * (1 - b) * image_pixel(im, li + 1, lj) +
           a * b * image_pixel(im, li + 1, lj + 1);
}
RT_D void image_change(const DImage& im, float u, float v, float& du, float& dv) {
    u = u - floorf(u); v = v - floorf(v);
    int i = (int)(u * im.width), j = (int)(v * im.height);
    i = i < 0 ? 0 : (i >= im.width - 1 ? im.width - 2 : i);
    j = j < 0 ? 0 : (j >= im.height - 1 ? im.height - 2 : j);
    float3 c = image_pixel(im, i, j), cu = image_pixel(im, i + 1, j), cv = image_pixel(im, i, j + 1);
    du = (cu.x + cu.y + cu.z) / 3.0f - (c.x + c.y + c.z) / 3.0f;
    float3 d = cv - c;
    dv = (d.x + d.y + d.z) / 3.0f;
}
// lat-long direction lookup shared by EnvironmentLight::ComputeLightContribution (src/Light.cpp:565-575) and
// Scene::GetBackgroundColor (src/Scene.cpp:418-425)
RT_D float3 env_lookup(const DScene& sc, int image, float3 d) {
    const float pi = 3.14159265358979323846f;
    float theta = acosf(d.y), phi = atan2f(d.z, d.x);
    return image_lookup(sc.images[image], RT_INTERP_BILINEAR, (-phi + pi) / (2.0f * pi), theta / pi);
}

// Perlin entries of the texture loop, identical for spheres and triangles (src/Shape.cpp:476-499, 589-612)
RT_D void perlin_texture(const RtTexture& tx, float3 p_obj, float3& n, Surface& s) {
    if (tx.decal_mode == RT_DM_REPLACE_KD) {
        s.dm = RT_DM_REPLACE_KD;
        float v = perlin(p_obj, tx.noise_scale, tx.noise_conversion);
        s.tex_color = v3(v, v, v);
        s.tex_normalizer = 1.0f;
    } else if (tx.decal_mode == RT_DM_BUMP_NORMAL) {
        float3 g = perlin_gradient(p_obj, tx.noise_scale, tx.noise_conversion);
        float3 gp = dot(g, n) * n;
        float3 nn = n - (g - gp) * tx.bump_factor;
        n = normalize0(dot(n, nn) > 0.f ? nn : -nn);
    }
}

// (ray, hit) -> Surface.  `o`, `d`, `time` are the WORLD ray.
RT_D Surface eval_surface(const DScene& sc, float3 o, float3 d, float time, const HitRec& h) {
    Surface s;
    s.P = o + d * h.t;
    s.dm = RT_DM_NONE;
    s.tex_color = v3(0.f, 0.f, 0.f);
    s.tex_normalizer = 1.0f;
    const DObjectSh& ob = sc.obj_sh[h.obj];
    s.mat_index = ob.mat_index;
    s.Le = v3(ob.radiance);
    float3 oo, od; int root;
    xf_ray(sc.obj_xf + h.obj, o, d, time, oo, od, root);
    float3 p_obj = oo + od * h.t;      // object-space hit point (the reference textures in object space)
    float3 n;
    const int t0 = ob.tex[0], t1 = ob.tex[1];
    if (ob.kind == RT_OBJ_SPHERE) {
        float4 cr = __ldg(sc.spheres + ob.geom);
        float3 lc = p_obj - v3(cr);
        n = lc / length(lc);
        s.prim_in_mesh = 0;
        for (int k = 0; k < 2; k++) {
            int tid = k == 0 ? t0 : t1;
            if (!tid) continue;
            const RtTexture tx = sc.textures[tid - 1];
            if (tx.type == RT_TEX_IMAGE) {
                const float pi = 3.14159265358979323846f;
                const DImage im = sc.images[tx.image];
                float R = cr.w;
                float theta = acosf(lc.y / R), phi = atan2f(lc.z, lc.x);
                float tu = (-phi + pi) / (2.0f * pi), tv = theta / pi;
                if (tx.decal_mode == RT_DM_REPLACE_KD || tx.decal_mode == RT_DM_BLEND_KD || tx.decal_mode == RT_DM_REPLACE_ALL) {
                    s.dm = tx.decal_mode;
                    s.tex_color = image_lookup(im, tx.interpolation, tu, tv);
                    s.tex_normalizer = (float)tx.normalizer;
                } else if (tx.decal_mode == RT_DM_REPLACE_NORMAL || tx.decal_mode == RT_DM_BUMP_NORMAL) {
                    float3 dpdu = v3(lc.z * 2.0f * pi, 0.f, lc.x * (-2.0f) * pi);
                    float3 dpdv = v3(lc.y * cosf(phi) * pi, -R * sinf(theta) * pi, lc.y * sinf(phi) * pi);
                    if (tx.decal_mode == RT_DM_REPLACE_NORMAL) {
                        float3 rn = image_lookup(im, tx.interpolation, tu, tv) / 255.0f;
                        rn = normalize0(rn - v3(0.5f, 0.5f, 0.5f));
                        n = normalize0(dpdu) * rn.x + normalize0(dpdv) * rn.y + n * rn.z;
                    } else {
                        float du, dv;
                        image_change(im, tu, tv, du, dv);
                        du *= tx.bump_factor; dv *= tx.bump_factor;
                        float3 nn = normalize0(cross(dpdv + dv * n, dpdu + du * n));
                        n = dot(n, nn) > 0.f ? nn : -nn;
                    }
                }
            } else perlin_texture(tx, p_obj, n, s);
        }
    } else {
        const float4* tp = sc.tris + (size_t)h.prim * 3;
        float4 ra = __ldg(tp), rb = __ldg(tp + 1), rc = __ldg(tp + 2);
        int gtri = __float_as_int(ra.w);                 // triangle index in the original (face-order) numbering
        s.prim_in_mesh = gtri - ob.first_tri;
        float3 e1 = v3(rb), e2 = v3(rc);
        const int* ix = sc.tri_idx + (size_t)gtri * 3;
        int i0 = 0, i1 = 0, i2 = 0;
        if (ob.smooth || t0 || t1) { i0 = __ldg(ix); i1 = __ldg(ix + 1); i2 = __ldg(ix + 2); }
        float alpha = 1.0f - h.beta - h.gamma;
        if (ob.smooth) {
            float3 n0 = v3(sc.normals + 3 * (size_t)i0), n1 = v3(sc.normals + 3 * (size_t)i1), n2 = v3(sc.normals + 3 * (size_t)i2);
            n = n0 * alpha + n1 * h.beta + n2 * h.gamma;
        } else n = cross(e1, e2);                        // == (c-b) x (a-b)
        n = n / length(n);
        if (t0 || t1) {
            const float* uv0 = sc.texcoords + 2 * (size_t)(i0 + ob.tex_offset);
            const float* uv1 = sc.texcoords + 2 * (size_t)(i1 + ob.tex_offset);
            const float* uv2 = sc.texcoords + 2 * (size_t)(i2 + ob.tex_offset);
            float u0 = __ldg(uv0), v0 = __ldg(uv0 + 1), u1 = __ldg(uv1), v1 = __ldg(uv1 + 1), u2 = __ldg(uv2), v2 = __ldg(uv2 + 1);
            float u = u0 * alpha + u1 * h.beta + u2 * h.gamma, v = v0 * alpha + v1 * h.beta + v2 * h.gamma;
            for (int k = 0; k < 2; k++) {
                int tid = k == 0 ? t0 : t1;
                if (!tid) continue;
                const RtTexture tx = sc.textures[tid - 1];
                if (tx.type == RT_TEX_IMAGE) {
                    const DImage im = sc.images[tx.image];
                    if (tx.decal_mode == RT_DM_REPLACE_KD || tx.decal_mode == RT_DM_BLEND_KD || tx.decal_mode == RT_DM_REPLACE_ALL) {
                        s.dm = tx.decal_mode;
                        s.tex_color = image_lookup(im, tx.interpolation, u, v);
                        s.tex_normalizer = (float)tx.normalizer;
                    } else if (tx.decal_mode == RT_DM_REPLACE_NORMAL || tx.decal_mode == RT_DM_BUMP_NORMAL) {
                        float a00 = u1 - u0, a01 = v1 - v0, a10 = u2 - u0, a11 = v2 - v0;
                        float inv = 1.0f / (a00 * a11 - a01 * a10);
                        float i00 = a11 * inv, i01 = -a01 * inv, i10 = -a10 * inv, i11 = a00 * inv;
                        float3 T = e1 * i00 + e2 * i01, B = e1 * i10 + e2 * i11;
                        if (tx.decal_mode == RT_DM_REPLACE_NORMAL) {
                            float3 rn = image_lookup(im, tx.interpolation, u, v) / 255.0f;
                            rn = normalize0(rn - v3(0.5f, 0.5f, 0.5f));
                            n = T * rn.x + B * rn.y + n * rn.z;
                        } else {
                            float du, dv;
                            image_change(im, u, v, du, dv);
                            du *= tx.bump_factor; dv *= tx.bump_factor;
                            float3 nn = normalize0(cross(B + dv * n, T + du * n));
                            n = dot(n, nn) > 0.f ? nn : -nn;
                        }
                    }
                } else perlin_texture(tx, p_obj, n, s);
            }
        }
    }
    const float* m = ob.nrm;
    s.N = normalize0(v3(m[0] * n.x + m[1] * n.y + m[2] * n.z, m[3] * n.x + m[4] * n.y + m[5] * n.z, m[6] * n.x + m[7] * n.y + m[8] * n.z));
    return s;
}

}  // namespace rt795
