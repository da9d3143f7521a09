// float3 helpers + Philox4x32-10 for the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt795 {

#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__

RT_HD float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
RT_HD float3 v3(const float* p) { return make_float3(p[0], p[1], p[2]); }
RT_HD float3 v3(float4 a) { return make_float3(a.x, a.y, a.z); }
RT_HD float3 operator+(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD float3 operator-(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD float3 operator-(float3 a) { return v3(-a.x, -a.y, -a.z); }
RT_HD float3 operator*(float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
RT_HD float3 operator*(float s, float3 a) { return v3(a.x * s, a.y * s, a.z * s); }
RT_HD float3 operator*(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }   // component-wise (cwiseProduct)
RT_HD float3 operator/(float3 a, float s) { float r = 1.0f / s; return v3(a.x * r, a.y * r, a.z * r); }   // one IEEE reciprocal instead of three divisions (<= 1 ulp from x/s)
RT_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_HD float3 cross(float3 a, float3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RT_HD float length(float3 a) { return sqrtf(dot(a, a)); }
RT_HD float3 normalize0(float3 a) {                                                           // Eigen normalized(): unchanged at 0
#ifdef __CUDA_ARCH__
    float n2 = dot(a, a);
    return n2 > 0.f ? a * rsqrtf(n2) : a;      // MUFU.RSQ (<= 2 ulp) instead of sqrt + reciprocal
#else
    float n = length(a); return n > 0.f ? a / n : a;
#endif
}
RT_HD bool any_nan(float3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }
RT_HD bool is_zero(float3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
RT_HD float comp(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// Philox4x32-10 (Salmon et al., SC'11).  Stream: key = seed, counter = (pixel, sample, path node, dimension).
// Same definition as oracle/restate.cpp (specified in DESIGN.md §RNG); implemented independently.
RT_HD float4 philox_uniform4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const float s = 1.0f / 16777216.0f;
    return make_float4((float)(c0 >> 8) * s, (float)(c1 >> 8) * s, (float)(c2 >> 8) * s, (float)(c3 >> 8) * s);
}
enum { DIM_CAMERA = 0, DIM_GLOSSY = 1, DIM_PT_BSDF = 2, DIM_PT_RR = 3, DIM_PT_LIGHT = 4, DIM_LIGHT0 = 0x100, DIM_LIGHT_STRIDE = 0x100 };

}  // namespace rt795
