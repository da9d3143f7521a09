// One node of the 8-bit quantised 4-wide BVH (layout in bvh_build.h, BvhNode64Q), shared by the host builder
// (QuantiseBvh4) and the device builder (bvh_device.cu) so both produce identical bytes from the same fp32 node.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define RT795_HD __host__ __device__
#else
#define RT795_HD
#endif

namespace rt795 {

// q: the 32 floats of a BvhNode128; w: the 16 words of the BvhNode64Q.  kNoChildRef = bvh_build.h kNoChild.
RT795_HD inline void QuantiseNode4(const float* q, uint32_t* w, int32_t no_child_ref) {
    int32_t ref[4]; bool has[4];
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int k = 0; k < 4; k++) {
        memcpy(&ref[k], &q[24 + k], 4);
        has[k] = ref[k] != no_child_ref;
        if (!has[k]) continue;
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], q[a * 8 + k]); hi[a] = fmaxf(hi[a], q[a * 8 + 4 + k]); }
    }
    float scale[3];
    uint32_t plo[3], phi[3];
    for (int a = 0; a < 3; a++) {
        if (!(hi[a] >= lo[a])) { lo[a] = 0.f; hi[a] = 0.f; }
        float ext = hi[a] - lo[a];
        int e;
        frexpf(ext > 0.f ? ext / 255.0f : 1e-30f, &e);              // ext/255 = m * 2^e, m in [0.5, 1)  ->  2^e >= ext/255
        for (;; e++) {                                               // grow the grid until every plane fits in a byte
            float sc = ldexpf(1.0f, e);
            bool ok = true;
            uint32_t pl = 0, ph = 0;
            for (int k = 0; k < 4 && ok; k++) {
                if (!has[k]) { pl |= 255u << (8 * k); continue; }    // absent child: inverted range lo = 255, hi = 0
                float clo = q[a * 8 + k], chi = q[a * 8 + 4 + k];
                int l = (int)floorf((clo - lo[a]) / sc), h = (int)ceilf((chi - lo[a]) / sc);
                l = l < 0 ? 0 : (l > 255 ? 255 : l);
                while (l > 0 && lo[a] + (float)l * sc > clo) l--;    // decoded in float exactly as the kernel does: stay conservative
                while (h < 256 && lo[a] + (float)h * sc < chi) h++;
                if (h > 255 || lo[a] + (float)l * sc > clo) { ok = false; break; }
                pl |= (uint32_t)l << (8 * k); ph |= (uint32_t)h << (8 * k);
            }
            if (ok) { scale[a] = sc; plo[a] = pl; phi[a] = ph; break; }
        }
    }
    memcpy(&w[0], &lo[0], 4); memcpy(&w[1], &lo[1], 4); memcpy(&w[2], &lo[2], 4); memcpy(&w[3], &scale[2], 4);
    for (int k = 0; k < 4; k++) memcpy(&w[4 + k], &ref[k], 4);
    w[8] = plo[0]; w[9] = plo[1]; w[10] = plo[2]; w[11] = phi[0];
    w[12] = phi[1]; w[13] = phi[2]; memcpy(&w[14], &scale[0], 4); memcpy(&w[15], &scale[1], 4);
}

}  // namespace rt795
