// D1: device BLAS builder - see bvh_device.h for the pipeline.  sm_100a only.
#include "bvh_device.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "bvh_quant.h"

namespace rt795 {
namespace {

#define BCK(x)                                                                                                          \
    do {                                                                                                                \
        cudaError_t e_ = (x);                                                                                           \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("bvh_device: ") + #x + ": " + cudaGetErrorString(e_)); \
    } while (0)

constexpr int kPlocRadius = 16;          // nearest-neighbour search window (+-) in the Morton-ordered cluster array (RT795_PLOC_RADIUS overrides, <= kMaxPlocRadius)
constexpr int kMaxPlocRadius = 64;
constexpr int kTopClusters = 4096;       // clustering stops here; the levels above are built with binned SAH (RT795_PLOC_TOP overrides; 1 = pure PLOC)
constexpr int kTB = 256;
constexpr int kMaxPlocIterations = 4096; // a healthy build of 1 M triangles takes ~40; beyond the cap the host builder takes over
constexpr int kCntMask = 0x3fffffff, kLeafBit = 0x40000000;
enum { CTR_QOUT = 0, CTR_ALLOC = 1, CTR_STACK = 2, CTR_DEPTH = 3, CTR_N = 4 };

inline int Grid(long long n) { return (int)((n + kTB - 1) / kTB); }

// One cudaMalloc for all temporaries: Reserve() records, Commit() allocates and hands the pointers out.
class Arena {
public:
    template <class T> void Reserve(T** slot, size_t n) {
        off_ = (off_ + 255) & ~(size_t)255;
        recs_.push_back({reinterpret_cast<void**>(slot), off_});
        off_ += (n ? n : 1) * sizeof(T);
    }
    void Commit() {
        BCK(cudaMalloc(&base_, off_ ? off_ : 256));
        for (auto& r : recs_) *r.slot = static_cast<char*>(base_) + r.off;
    }
    ~Arena() { if (base_) cudaFree(base_); }
private:
    struct Rec { void** slot; size_t off; };
    std::vector<Rec> recs_;
    size_t off_ = 0;
    void* base_ = nullptr;
};

__device__ __forceinline__ float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 min3(float3 a, float3 b) { return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ float3 max3(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
__device__ __forceinline__ float half_area(float3 lo, float3 hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
// float min/max atomics on the ordered-integer view of IEEE floats (target initialised to +-3e38)
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v)); else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v)); else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}

// Monotone float <-> int maps, so __reduce_min_sync / __reduce_max_sync (integer only) can reduce floats inside a lane group.
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---- B1: padded triangle boxes + bounds of their centres
__global__ void __launch_bounds__(kTB) k_prim_bounds(const float* __restrict__ v, const int32_t* __restrict__ ix, int n, float eps,
                                                     float4* __restrict__ plo, float4* __restrict__ phi, float* cb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 cmin = make_float3(3.0e38f, 3.0e38f, 3.0e38f), cmax = make_float3(-3.0e38f, -3.0e38f, -3.0e38f);
    if (i < n) {
        float3 lo = cmin, hi = cmax;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float* p = v + 3 * (size_t)ix[3 * (size_t)i + k];
            float3 q = make_float3(p[0], p[1], p[2]);
            lo = min3(lo, q); hi = max3(hi, q);
        }
        float ext = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
        // same rule as the host builder (core.cu BuildAndUpload): the reference accepts beta, gamma >= -eps, i.e. hits up to
        // eps*edge outside the triangle, so the box is padded to stay conservative
        // (explicit roundings, no FMA contraction: bit-identical to the host's boxes, so the device SAH sees the host's input)
        float pe = __fmul_rn(__fmul_rn(2.0f, eps), ext);
        float px = __fadd_rn(__fadd_rn(pe, __fmul_rn(4.0e-7f, fmaxf(fabsf(lo.x), fabsf(hi.x)))), 1e-30f);
        float py = __fadd_rn(__fadd_rn(pe, __fmul_rn(4.0e-7f, fmaxf(fabsf(lo.y), fabsf(hi.y)))), 1e-30f);
        float pz = __fadd_rn(__fadd_rn(pe, __fmul_rn(4.0e-7f, fmaxf(fabsf(lo.z), fabsf(hi.z)))), 1e-30f);
        lo = make_float3(lo.x - px, lo.y - py, lo.z - pz);
        hi = make_float3(hi.x + px, hi.y + py, hi.z + pz);
        plo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
        phi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
        cmin = cmax = make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        cmin.x = fminf(cmin.x, __shfl_xor_sync(full, cmin.x, s)); cmin.y = fminf(cmin.y, __shfl_xor_sync(full, cmin.y, s)); cmin.z = fminf(cmin.z, __shfl_xor_sync(full, cmin.z, s));
        cmax.x = fmaxf(cmax.x, __shfl_xor_sync(full, cmax.x, s)); cmax.y = fmaxf(cmax.y, __shfl_xor_sync(full, cmax.y, s)); cmax.z = fmaxf(cmax.z, __shfl_xor_sync(full, cmax.z, s));
    }
    if ((threadIdx.x & 31) == 0 && cmin.x <= cmax.x) {
        atomic_min_f(cb + 0, cmin.x); atomic_min_f(cb + 1, cmin.y); atomic_min_f(cb + 2, cmin.z);
        atomic_max_f(cb + 3, cmax.x); atomic_max_f(cb + 4, cmax.y); atomic_max_f(cb + 5, cmax.z);
    }
}

// ---- B2: 63-bit Morton key of the box centre
__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__global__ void __launch_bounds__(kTB) k_morton(const float4* __restrict__ plo, const float4* __restrict__ phi, int n, const float* __restrict__ cb,
                                                unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 lo = plo[i], hi = phi[i];
    float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    unsigned long long q[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float ext = cb[3 + a] - cb[a];
        float t = ext > 0.f ? (c[a] - cb[a]) / ext : 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        q[a] = (unsigned long long)fminf(t * 2097152.0f, 2097151.0f);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = (uint32_t)i;
}

// ---- B3: PLOC.  Cluster i = (box, node id in lo.w).  Leaves are nodes 0..n-1 in Morton order, inner nodes n..2n-2.
__global__ void __launch_bounds__(kTB) k_init_clusters(const uint32_t* __restrict__ vals, const float4* __restrict__ plo, const float4* __restrict__ phi, int n,
                                                       float4* __restrict__ clo, float4* __restrict__ chi, float4* __restrict__ nlo, float4* __restrict__ nhi,
                                                       int32_t* __restrict__ parent, int32_t* __restrict__ meta, float* __restrict__ cst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = vals[i];
    float4 lo = plo[p], hi = phi[p];
    lo.w = __int_as_float(i);
    clo[i] = lo; chi[i] = hi; nlo[i] = lo; nhi[i] = hi;
    parent[i] = -1; meta[i] = 1 | kLeafBit; cst[i] = 1.0f;
}

// Nearest neighbour of every cluster inside the window, under the symmetric total order on pairs
// (merged half-area, not-buddy, min index, max index) - "buddy" = index ^ 1.  Symmetry guarantees that the globally
// smallest pair is mutual, so every iteration merges at least one pair; the buddy preference makes runs of identical
// boxes (all distances tie) pair up instead of chaining.
__global__ void __launch_bounds__(kTB) k_nearest(const float4* __restrict__ clo, const float4* __restrict__ chi, int m, int radius, int32_t* __restrict__ nn) {
    __shared__ float4 slo[kTB + 2 * kMaxPlocRadius], shi[kTB + 2 * kMaxPlocRadius];
    int base = blockIdx.x * kTB - radius;
    for (int t = threadIdx.x; t < kTB + 2 * radius; t += kTB) {
        int g = base + t;
        if (g >= 0 && g < m) { slo[t] = clo[g]; shi[t] = chi[g]; }
    }
    __syncthreads();
    int i = blockIdx.x * kTB + threadIdx.x;
    if (i >= m) return;
    float3 lo = f3(slo[threadIdx.x + radius]), hi = f3(shi[threadIdx.x + radius]);
    float best = 3.4e38f;
    int bestj = -1;
#pragma unroll 4
    for (int d = -radius; d <= radius; d++) {
        int j = i + d;
        if (d == 0 || j < 0 || j >= m) continue;
        int t = threadIdx.x + radius + d;
        float a = half_area(min3(lo, f3(slo[t])), max3(hi, f3(shi[t])));
        if (a < best || (a == best && j == (i ^ 1))) { best = a; bestj = j; }
    }
    nn[i] = bestj;
}

__global__ void __launch_bounds__(kTB) k_merge_flags(const int32_t* __restrict__ nn, int m, unsigned long long* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int j = nn[i];
    bool mutual = j >= 0 && nn[j] == i;
    unsigned long long keep = !(mutual && i > j), merge = mutual && i < j;
    flags[i] = (keep << 32) | merge;                                  // one scan yields the compacted position and the new node id
}

__global__ void __launch_bounds__(kTB) k_merge_emit(const float4* __restrict__ clo, const float4* __restrict__ chi, const int32_t* __restrict__ nn,
                                                    const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ scan, int m, int next_id,
                                                    float4* __restrict__ clo_out, float4* __restrict__ chi_out, float4* __restrict__ nlo, float4* __restrict__ nhi,
                                                    int2* __restrict__ child, int32_t* __restrict__ parent, int32_t* __restrict__ meta, float* __restrict__ cst,
                                                    uint32_t* __restrict__ totals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    unsigned long long f = flags[i], s = scan[i];
    uint32_t keep = (uint32_t)(f >> 32), merge = (uint32_t)(f & 1u);
    if (i == m - 1) { totals[0] = (uint32_t)(s >> 32) + keep; totals[1] = (uint32_t)(s & 0xffffffffu) + merge; }
    if (!keep) return;
    float4 lo = clo[i], hi = chi[i];
    if (merge) {
        int j = nn[i];
        float4 lo2 = clo[j], hi2 = chi[j];
        int l = __float_as_int(lo.w), r = __float_as_int(lo2.w);
        int id = next_id + (int)(s & 0xffffffffu);
        float3 ulo = min3(f3(lo), f3(lo2)), uhi = max3(f3(hi), f3(hi2));
        float area = half_area(ulo, uhi), al = half_area(f3(lo), f3(hi)), ar = half_area(f3(lo2), f3(hi2));
        int c = (meta[l] & kCntMask) + (meta[r] & kCntMask);
        // SAH, bottom-up: cost(node) = min(#triangles if <= 4, C_t + sum_child area_child / area * cost(child)) - the host builder's
        // constants (traversal kSahTraversalCost, intersection 1, bvh_build.h) and leaf rule
        float split = area > 0.f ? kSahTraversalCost + (al * cst[l] + ar * cst[r]) / area : 3.0e38f;
        split = fminf(split, 3.0e38f);
        bool leaf = c <= 4 && (float)c <= split;
        cst[id] = leaf ? (float)c : split;
        meta[id] = c | (leaf ? kLeafBit : 0);
        child[id] = make_int2(l, r);
        parent[l] = id; parent[r] = id; parent[id] = -1;
        lo = make_float4(ulo.x, ulo.y, ulo.z, __int_as_float(id));
        hi = make_float4(uhi.x, uhi.y, uhi.z, 0.f);
        nlo[id] = lo; nhi[id] = hi;
    }
    uint32_t pos = (uint32_t)(s >> 32);
    clo_out[pos] = lo; chi_out[pos] = hi;
}

// ---- B3b: top of the tree.  Once the clusters are few (kTopClusters) their Morton neighbourhoods say little about which
// should merge; the last levels - the ones every ray walks - are built by the host's binned-SAH builder over the cluster
// boxes (a few thousand "primitives", < 1 ms) and spliced in as ordinary inner nodes.
__global__ void __launch_bounds__(kTB) k_gather_clusters(const float4* __restrict__ clo, int m, const int32_t* __restrict__ meta, const float* __restrict__ cst,
                                                         int32_t* __restrict__ out_meta, float* __restrict__ out_cst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int id = __float_as_int(clo[i].w);
    out_meta[i] = meta[id]; out_cst[i] = cst[id];
}
__global__ void __launch_bounds__(kTB) k_set_parents(int first_id, int count, const int2* __restrict__ child, int32_t* __restrict__ parent, int root) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    int id = first_id + t;
    int2 ch = child[id];
    parent[ch.x] = id; parent[ch.y] = id;
    if (id == root) parent[id] = -1;
}

// ---- B4: one walk to the root per node: DFS offset of its first triangle, state (0 live inner, 1 leaf of the final tree, 2 buried
// inside a collapsed leaf), depth
__global__ void __launch_bounds__(kTB) k_layout(int total, const int32_t* __restrict__ parent, const int2* __restrict__ child, const int32_t* __restrict__ meta,
                                                int32_t* __restrict__ first, uint8_t* __restrict__ state, int32_t* ctr) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    int depth = 0;
    if (id < total) {
        int off = 0, cur = id;
        bool buried = false;
        for (int p = parent[cur]; p >= 0; p = parent[cur]) {
            int2 ch = child[p];
            int mp = meta[p];
            if (ch.y == cur) off += meta[ch.x] & kCntMask;
            buried |= (mp & kLeafBit) != 0;
            cur = p; depth++;
        }
        first[id] = off;
        uint8_t st = buried ? 2 : ((meta[id] & kLeafBit) ? 1 : 0);
        state[id] = st;
        if (st == 2) depth = 0;
    }
    int wmax = __reduce_max_sync(0xffffffffu, depth);
    if ((threadIdx.x & 31) == 0 && wmax > 0) atomicMax(&ctr[CTR_DEPTH], wmax);
}

// ---- B5
__global__ void __launch_bounds__(kTB) k_emit_tris(int n, const int32_t* __restrict__ first, const uint32_t* __restrict__ vals, const float* __restrict__ v,
                                                   const int32_t* __restrict__ ix, int first_tri, float4* __restrict__ tris) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = vals[i];
    const int32_t* t = ix + 3 * (size_t)p;
    const float *a = v + 3 * (size_t)t[0], *b = v + 3 * (size_t)t[1], *c = v + 3 * (size_t)t[2];
    float4* o = tris + 3 * (size_t)first[i];
    o[0] = make_float4(a[0], a[1], a[2], __int_as_float(first_tri + (int)p));
    o[1] = make_float4(b[0] - a[0], b[1] - a[1], b[2] - a[2], 0.f);
    o[2] = make_float4(c[0] - a[0], c[1] - a[1], c[2] - a[2], 0.f);
}

// The finished binary tree lives in an "id space": ids [id_base, id_base + id_count) may be inner nodes (state 0), leaves of the
// final tree (state 1: triangles [first, first + count) of the leaf-ordered array) or unused (state 2).  PLOC: id_base = n
// (triangle leaves are ids 0..n-1), root = last id, numbered in reverse creation order.  Device SAH: id_base = 0, root = 0.
__global__ void __launch_bounds__(kTB) k_live_flags(int id_base, int id_count, const uint8_t* __restrict__ state, int32_t* __restrict__ live) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < id_count) live[t] = state[id_base + t] == 0 ? 1 : 0;
}

__device__ __forceinline__ int32_t encode_leaf(int32_t first, int32_t count) { return ~((first << 3) | (count - 1)); }

// Binary nodes, numbered by the scan of the live flags (PLOC: reversed, so the root - created last - becomes node 0 and nodes
// merged in the same iteration, neighbours on the Morton curve, stay neighbours in memory).
__global__ void __launch_bounds__(kTB) k_emit_bvh2(int id_base, int id_count, int reverse, int live_total, const int32_t* __restrict__ live,
                                                   const int32_t* __restrict__ lscan, const int2* __restrict__ child, const uint8_t* __restrict__ state,
                                                   const int32_t* __restrict__ first, const int32_t* __restrict__ meta, const float4* __restrict__ nlo,
                                                   const float4* __restrict__ nhi, int node_base, int tri_base, float4* __restrict__ nodes2) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= id_count || !live[t]) return;
    int idx = reverse ? live_total - 1 - lscan[t] : lscan[t];
    int2 ch = child[id_base + t];
    int c[2] = {ch.x, ch.y};
    int32_t ref[2];
    float4 lo[2], hi[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        lo[s] = nlo[c[s]]; hi[s] = nhi[c[s]];
        if (state[c[s]] == 1) ref[s] = encode_leaf(tri_base + first[c[s]], meta[c[s]] & kCntMask);
        else { int ls = lscan[c[s] - id_base]; ref[s] = node_base + (reverse ? live_total - 1 - ls : ls); }
    }
    float4* o = nodes2 + 4 * (size_t)idx;
    o[0] = make_float4(lo[0].x, hi[0].x, lo[0].y, hi[0].y);
    o[1] = make_float4(lo[1].x, hi[1].x, lo[1].y, hi[1].y);
    o[2] = make_float4(lo[0].z, hi[0].z, lo[1].z, hi[1].z);
    o[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.f, 0.f);
}

// =====================================================================================================================
// Device binned SAH (RT_BVH_DEVICE_SAH): the host builder's algorithm (bvh_build.cpp: 16 bins x 3 axes over the centroid
// bounds, same costs, same leaf rule) run level-synchronously over ALL nodes of a level at once.  The primitive array is
// never partitioned: each primitive carries the index of the node it currently belongs to (`pnode`) and is re-assigned
// after its node's split is known; leaf ranges are made contiguous at the end by one stable sort on the leaf's DFS offset.
// Per level: S1 centroid bounds per node (float atomics), S2 binning (atomics into the node's 3 x 16 bins), S3 split
// selection per node (the host's sweep, same operation order), scan, S4 child creation, S5 re-assignment.
constexpr int kSahBins = 16;
constexpr int kSahMaxLeaf = 4;
constexpr int kSahMaxDepth = 36;         // deeper levels fall back to Morton-position midpoint splits (the host falls back to medians)
constexpr int kBinWords = 8;             // lo.xyz, hi.xyz, count, pad

struct SahNodes {                        // build nodes, SoA; index = creation order, root = 0, children of a node are adjacent
    float4 *lo, *hi;                     // primitive bounds
    float4 *clo, *chi;                   // centroid bounds (S1)
    int2* child;                         // (left, left + 1), or (-1, -1)
    int32_t *parent, *meta;              // meta: count | kLeafBit
    int32_t *pos_min, *pos_max;          // range of Morton positions of its primitives (fallback split)
    int32_t *split;                      // S3 decision: -1 leaf, else axis | bin << 2 | fallback << 8
    int32_t *pos_mid;
};

__global__ void __launch_bounds__(kTB) k_sah_init_nodes(SahNodes nd, int begin, int end) {
    int b = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= end) return;
    nd.clo[b] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f); nd.chi[b] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f);
    nd.pos_min[b] = 0x7fffffff; nd.pos_max[b] = -1;
    nd.child[b] = make_int2(-1, -1);
}

__global__ void __launch_bounds__(kTB) k_sah_root(SahNodes nd, int n, const float4* __restrict__ plo, const float4* __restrict__ phi, int32_t* __restrict__ pnode) {
    // root bounds: warp-reduced float atomics into node 0 (lo/hi initialised by the caller)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = make_float3(3.0e38f, 3.0e38f, 3.0e38f), hi = make_float3(-3.0e38f, -3.0e38f, -3.0e38f);
    if (i < n) { lo = f3(plo[i]); hi = f3(phi[i]); pnode[i] = 0; }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(full, lo.x, s)); lo.y = fminf(lo.y, __shfl_xor_sync(full, lo.y, s)); lo.z = fminf(lo.z, __shfl_xor_sync(full, lo.z, s));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(full, hi.x, s)); hi.y = fmaxf(hi.y, __shfl_xor_sync(full, hi.y, s)); hi.z = fmaxf(hi.z, __shfl_xor_sync(full, hi.z, s));
    }
    if ((threadIdx.x & 31) == 0 && lo.x <= hi.x) {
        float* l = reinterpret_cast<float*>(nd.lo), *h = reinterpret_cast<float*>(nd.hi);
        atomic_min_f(l + 0, lo.x); atomic_min_f(l + 1, lo.y); atomic_min_f(l + 2, lo.z);
        atomic_max_f(h + 0, hi.x); atomic_max_f(h + 1, hi.y); atomic_max_f(h + 2, hi.z);
    }
}

__device__ __forceinline__ float3 centre_of(float4 lo, float4 hi) { return make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)); }

// S1.  i = Morton position; order[i] = primitive.  Primitives whose node is older than this level sit in finished leaves.
// Lanes of a warp that share a node (neighbours on the Morton curve: all of them near the top of the tree) reduce among
// themselves first and issue one set of atomics per group instead of one per primitive.
__global__ void __launch_bounds__(kTB) k_sah_cbounds(SahNodes nd, int n, int lvl_begin, const uint32_t* __restrict__ order, const float4* __restrict__ plo,
                                                     const float4* __restrict__ phi, const int32_t* __restrict__ pnode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int b = i < n ? pnode[i] : -1;
    bool act = b >= lvl_begin;
    unsigned am = __ballot_sync(0xffffffffu, act);
    if (!act) return;
    uint32_t p = order[i];
    float3 c = centre_of(plo[p], phi[p]);
    unsigned g = __match_any_sync(am, b);
    int lx = __reduce_min_sync(g, f2ord(c.x)), ly = __reduce_min_sync(g, f2ord(c.y)), lz = __reduce_min_sync(g, f2ord(c.z));
    int hx = __reduce_max_sync(g, f2ord(c.x)), hy = __reduce_max_sync(g, f2ord(c.y)), hz = __reduce_max_sync(g, f2ord(c.z));
    int pmin = __reduce_min_sync(g, i), pmax = __reduce_max_sync(g, i);
    if ((int)(threadIdx.x & 31) != __ffs(g) - 1) return;
    float* l = reinterpret_cast<float*>(nd.clo + b), *h = reinterpret_cast<float*>(nd.chi + b);
    atomic_min_f(l + 0, ord2f(lx)); atomic_min_f(l + 1, ord2f(ly)); atomic_min_f(l + 2, ord2f(lz));
    atomic_max_f(h + 0, ord2f(hx)); atomic_max_f(h + 1, ord2f(hy)); atomic_max_f(h + 2, ord2f(hz));
    atomicMin(nd.pos_min + b, pmin); atomicMax(nd.pos_max + b, pmax);
}

__global__ void __launch_bounds__(kTB) k_sah_init_bins(float* bins, long long nbins) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbins) return;
    float* w = bins + t * kBinWords;
    w[0] = w[1] = w[2] = 3.0e38f; w[3] = w[4] = w[5] = -3.0e38f; w[6] = 0.f; w[7] = 0.f;   // w[6] holds the count as an int (bit pattern 0)
}

__device__ __forceinline__ int sah_bin(float c, float lo, float ext) {
    float scale = (float)kSahBins / ext;                                 // as bvh_build.cpp Split(): scale = kBins / ext
    int b = (int)((c - lo) * scale);
    return b < 0 ? 0 : (b > kSahBins - 1 ? kSahBins - 1 : b);
}

// S2.  Same grouping as S1, per axis, on (node, bin).
__global__ void __launch_bounds__(kTB) k_sah_bin(SahNodes nd, int n, int chunk_begin, int chunk_end, const uint32_t* __restrict__ order, const float4* __restrict__ plo,
                                                 const float4* __restrict__ phi, const int32_t* __restrict__ pnode, float* bins) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int b = i < n ? pnode[i] : -1;
    bool act = b >= chunk_begin && b < chunk_end;
    if (act && (nd.meta[b] & kCntMask) <= 1) act = false;               // single-primitive node: a leaf, no split to evaluate
    unsigned am = __ballot_sync(0xffffffffu, act);
    if (!act) return;
    uint32_t p = order[i];
    float4 lo = plo[p], hi = phi[p];
    float3 c = centre_of(lo, hi);
    float4 cl = nd.clo[b], chh = nd.chi[b];
    float cc[3] = {c.x, c.y, c.z}, cmin[3] = {cl.x, cl.y, cl.z}, cmax[3] = {chh.x, chh.y, chh.z};
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float ext = cmax[a] - cmin[a];
        bool ok = ext > 0.f;                                             // uniform over the lanes of a node
        unsigned oka = __ballot_sync(am, ok);
        if (!ok) continue;
        int k = sah_bin(cc[a], cmin[a], ext);
        unsigned g = __match_any_sync(oka, (b - chunk_begin) * kSahBins + k);
        int lx = __reduce_min_sync(g, f2ord(lo.x)), ly = __reduce_min_sync(g, f2ord(lo.y)), lz = __reduce_min_sync(g, f2ord(lo.z));
        int hx = __reduce_max_sync(g, f2ord(hi.x)), hy = __reduce_max_sync(g, f2ord(hi.y)), hz = __reduce_max_sync(g, f2ord(hi.z));
        if (lane != __ffs(g) - 1) continue;
        float* w = bins + (((long long)(b - chunk_begin) * 3 + a) * kSahBins + k) * kBinWords;
        atomic_min_f(w + 0, ord2f(lx)); atomic_min_f(w + 1, ord2f(ly)); atomic_min_f(w + 2, ord2f(lz));
        atomic_max_f(w + 3, ord2f(hx)); atomic_max_f(w + 4, ord2f(hy)); atomic_max_f(w + 5, ord2f(hz));
        atomicAdd(reinterpret_cast<int*>(w + 6), __popc(g));
    }
}

struct Box3 { float lo[3], hi[3]; };
__device__ __forceinline__ void box_reset(Box3& b) { for (int k = 0; k < 3; k++) { b.lo[k] = 3.0e38f; b.hi[k] = -3.0e38f; } }
__device__ __forceinline__ void box_grow(Box3& b, const float* w) { for (int k = 0; k < 3; k++) { if (w[k] < b.lo[k]) b.lo[k] = w[k]; if (w[3 + k] > b.hi[k]) b.hi[k] = w[3 + k]; } }
__device__ __forceinline__ float box_half_area(const Box3& b) {          // Aabb::HalfArea, without FMA contraction: the host's rounding
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.f;
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dz, dx));
}

// S3: the host's Split() on the bins of one node.  Writes the decision and the split flag (scanned into child positions).
struct SahLevelTmp { float4 *llo, *lhi, *rlo, *rhi; int32_t *lc, *rc; };   // per node of the level: the chosen split's children
__global__ void __launch_bounds__(128) k_sah_split(SahNodes nd, int lvl_begin, int chunk_begin, int chunk_end, int depth, const float* __restrict__ bins,
                                                   int32_t* __restrict__ flag, SahLevelTmp tmp) {
    int b = chunk_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= chunk_end) return;
    int count = nd.meta[b] & kCntMask;
    int decision = -1;
    if (count > 1) {
        float best_cost = 3.402823466e38f;
        int best_axis = -1, best_split = -1;
        if (depth < kSahMaxDepth) {
            float4 cl = nd.clo[b], chh = nd.chi[b];
            float cmin[3] = {cl.x, cl.y, cl.z}, cmax[3] = {chh.x, chh.y, chh.z};
            for (int a = 0; a < 3; a++) {
                if (!(cmax[a] - cmin[a] > 0.f)) continue;
                const float* w = bins + ((long long)(b - chunk_begin) * 3 + a) * kSahBins * kBinWords;
                float ra[kSahBins]; int rc[kSahBins];
                Box3 acc; box_reset(acc); int c = 0;
                for (int k = kSahBins - 1; k >= 1; k--) { box_grow(acc, w + k * kBinWords); c += __float_as_int(w[k * kBinWords + 6]); ra[k] = box_half_area(acc); rc[k] = c; }
                box_reset(acc); c = 0;
                for (int k = 0; k < kSahBins - 1; k++) {
                    box_grow(acc, w + k * kBinWords); c += __float_as_int(w[k * kBinWords + 6]);
                    if (c == 0 || rc[k + 1] == 0) continue;
                    float cost = __fadd_rn(__fmul_rn(box_half_area(acc), (float)c), __fmul_rn(ra[k + 1], (float)rc[k + 1]));
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = k; }
                }
            }
        }
        float4 l = nd.lo[b], h = nd.hi[b];
        Box3 nb; nb.lo[0] = l.x; nb.lo[1] = l.y; nb.lo[2] = l.z; nb.hi[0] = h.x; nb.hi[1] = h.y; nb.hi[2] = h.z;
        float area = box_half_area(nb);
        float leaf_cost = (float)count;                                  // kIntersectCost * count
        float split_cost = best_axis >= 0 && area > 0 ? __fadd_rn(kSahTraversalCost, __fdiv_rn(best_cost, area)) : 3.402823466e38f;
        if (count <= kSahMaxLeaf && leaf_cost <= split_cost) decision = -1;
        else if (best_axis >= 0) decision = best_axis | (best_split << 2);
        else if (count <= kSahMaxLeaf) decision = -1;
        else {                                                           // centroids coincide or depth cap: split the Morton range in the middle
            int lo_p = nd.pos_min[b], hi_p = nd.pos_max[b];
            nd.pos_mid[b] = lo_p + (hi_p - lo_p) / 2;
            decision = 1 << 8;
        }
    }
    nd.split[b] = decision;
    flag[b - lvl_begin] = decision >= 0 ? 1 : 0;
    if (decision < 0) nd.meta[b] = count | kLeafBit;
    if (decision >= 0 && !(decision & (1 << 8))) {
        int axis = decision & 3, split = decision >> 2;
        const float* w = bins + ((long long)(b - chunk_begin) * 3 + axis) * kSahBins * kBinWords;
        Box3 lb, rb; box_reset(lb); box_reset(rb);
        int lc = 0, rc = 0;
        for (int k = 0; k < kSahBins; k++) {
            int c = __float_as_int(w[k * kBinWords + 6]);
            if (c == 0) continue;
            if (k <= split) { box_grow(lb, w + k * kBinWords); lc += c; } else { box_grow(rb, w + k * kBinWords); rc += c; }
        }
        int t = b - lvl_begin;
        tmp.llo[t] = make_float4(lb.lo[0], lb.lo[1], lb.lo[2], 0.f); tmp.lhi[t] = make_float4(lb.hi[0], lb.hi[1], lb.hi[2], 0.f);
        tmp.rlo[t] = make_float4(rb.lo[0], rb.lo[1], rb.lo[2], 0.f); tmp.rhi[t] = make_float4(rb.hi[0], rb.hi[1], rb.hi[2], 0.f);
        tmp.lc[t] = lc; tmp.rc[t] = rc;
    }
}

// S4: children of the nodes that split, at next_begin + 2 * scan.  A fallback split keeps the parent's box for both children
// (conservative); their counts come from S5's atomics.
__global__ void __launch_bounds__(kTB) k_sah_children(SahNodes nd, int lvl_begin, int lvl_end, int next_begin, const int32_t* __restrict__ flag,
                                                      const int32_t* __restrict__ scan, SahLevelTmp tmp) {
    int b = lvl_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= lvl_end || !flag[b - lvl_begin]) return;
    int t = b - lvl_begin;
    int left = next_begin + 2 * scan[t];
    int d = nd.split[b];
    nd.child[b] = make_int2(left, left + 1);
    nd.parent[left] = b; nd.parent[left + 1] = b;
    if (d & (1 << 8)) {
        nd.lo[left] = nd.lo[b]; nd.hi[left] = nd.hi[b]; nd.lo[left + 1] = nd.lo[b]; nd.hi[left + 1] = nd.hi[b];
        nd.meta[left] = 0; nd.meta[left + 1] = 0;
        return;
    }
    nd.lo[left] = tmp.llo[t]; nd.hi[left] = tmp.lhi[t]; nd.lo[left + 1] = tmp.rlo[t]; nd.hi[left + 1] = tmp.rhi[t];
    nd.meta[left] = tmp.lc[t]; nd.meta[left + 1] = tmp.rc[t];
}

// S5
__global__ void __launch_bounds__(kTB) k_sah_assign(SahNodes nd, int n, int lvl_begin, const uint32_t* __restrict__ order, const float4* __restrict__ plo,
                                                    const float4* __restrict__ phi, int32_t* __restrict__ pnode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = pnode[i];
    if (b < lvl_begin) return;
    int d = nd.split[b];
    if (d < 0) return;                                                   // its node became a leaf: the primitive stays
    int left = nd.child[b].x;
    int side;
    if (d & (1 << 8)) {
        side = i <= nd.pos_mid[b] ? 0 : 1;
        atomicAdd(nd.meta + left + side, 1);
    } else {
        int axis = d & 3, split = d >> 2;
        uint32_t p = order[i];
        float3 c = centre_of(plo[p], phi[p]);
        float4 cl = nd.clo[b], chh = nd.chi[b];
        float cc = axis == 0 ? c.x : (axis == 1 ? c.y : c.z), cmin = axis == 0 ? cl.x : (axis == 1 ? cl.y : cl.z), cmax = axis == 0 ? chh.x : (axis == 1 ? chh.y : chh.z);
        side = sah_bin(cc, cmin, cmax - cmin) <= split ? 0 : 1;
    }
    pnode[i] = left + side;
}

// DFS offset of every build node (one walk to the root), state for the emission stage, key of every primitive for the final sort
__global__ void __launch_bounds__(kTB) k_sah_layout(SahNodes nd, int nb, int32_t* __restrict__ first, uint8_t* __restrict__ state) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int off = 0, cur = b;
    for (int p = nd.parent[cur]; p >= 0; p = nd.parent[cur]) {
        int left = nd.child[p].x;
        if (cur != left) off += nd.meta[left] & kCntMask;
        cur = p;
    }
    first[b] = off;
    state[b] = (nd.meta[b] & kLeafBit) ? 1 : 0;
}
__global__ void __launch_bounds__(kTB) k_sah_keys(int n, const int32_t* __restrict__ pnode, const int32_t* __restrict__ first, uint32_t* __restrict__ keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (uint32_t)first[pnode[i]];
}
__global__ void __launch_bounds__(kTB) k_emit_tris_ordered(int n, const uint32_t* __restrict__ prims, const float* __restrict__ v, const int32_t* __restrict__ ix,
                                                           int first_tri, float4* __restrict__ tris) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t p = prims[k];
    const int32_t* t = ix + 3 * (size_t)p;
    const float *a = v + 3 * (size_t)t[0], *b = v + 3 * (size_t)t[1], *c = v + 3 * (size_t)t[2];
    float4* o = tris + 3 * (size_t)k;
    o[0] = make_float4(a[0], a[1], a[2], __int_as_float(first_tri + (int)p));
    o[1] = make_float4(b[0] - a[0], b[1] - a[1], b[2] - a[2], 0.f);
    o[2] = make_float4(c[0] - a[0], c[1] - a[1], c[2] - a[2], 0.f);
}

// ---- B6: one level of the BVH2 -> BVH4 collapse.  Item = (binary node id, 4-wide node index, stack entries above it).
__global__ void __launch_bounds__(kTB) k_collapse4(const int4* __restrict__ items, int n_items, int4* __restrict__ items_out, int32_t* ctr,
                                                   const int2* __restrict__ child, const uint8_t* __restrict__ state, const int32_t* __restrict__ first,
                                                   const int32_t* __restrict__ meta, const float4* __restrict__ nlo, const float4* __restrict__ nhi,
                                                   int node4_base, int tri_base, float4* __restrict__ nodes4) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    int4 it = items[t];
    int c[4];
    int nc = 2;
    { int2 ch = child[it.x]; c[0] = ch.x; c[1] = ch.y; }
    while (nc < 4) {                                                  // open the largest inner child until four (as CollapseToBvh4)
        int best = -1;
        float barea = -1.f;
        for (int k = 0; k < nc; k++) {
            if (state[c[k]] != 0) continue;
            float a = half_area(f3(nlo[c[k]]), f3(nhi[c[k]]));
            if (a > barea) { barea = a; best = k; }
        }
        if (best < 0) break;
        int2 sub = child[c[best]];
        for (int k = best; k < nc - 1; k++) c[k] = c[k + 1];
        c[nc - 1] = sub.x; c[nc] = sub.y;
        nc++;
    }
    int stack_here = it.z + nc - 1;
    atomicMax(&ctr[CTR_STACK], stack_here);
    float q[32];
#pragma unroll
    for (int k = 0; k < 32; k++) q[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int32_t ref = kNoChild;
        if (k < nc) {
            float4 lo = nlo[c[k]], hi = nhi[c[k]];
            q[0 + k] = lo.x; q[4 + k] = hi.x; q[8 + k] = lo.y; q[12 + k] = hi.y; q[16 + k] = lo.z; q[20 + k] = hi.z;
            if (state[c[k]] == 1) {
                ref = encode_leaf(tri_base + first[c[k]], meta[c[k]] & kCntMask);
            } else {
                int idx4 = atomicAdd(&ctr[CTR_ALLOC], 1);
                int slot = atomicAdd(&ctr[CTR_QOUT], 1);
                items_out[slot] = make_int4(c[k], idx4, stack_here, 0);
                ref = node4_base + idx4;
            }
        }
        q[24 + k] = __int_as_float(ref);
    }
    float4* o = nodes4 + 8 * (size_t)it.y;
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = make_float4(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
}

// ---- B7
__global__ void __launch_bounds__(kTB) k_quantise4(const float4* __restrict__ nodes4, int n4, uint4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float q[32];
    uint32_t w[16];
    const float4* in = nodes4 + 8 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 8; k++) { float4 v = in[k]; q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w; }
    QuantiseNode4(q, w, kNoChild);
    uint4* o = out + 4 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}

}  // namespace

void FreeDeviceBlas(DeviceBlas* b) {
    if (!b) return;
    if (b->nodes2) cudaFree(b->nodes2);
    if (b->nodes4) cudaFree(b->nodes4);
    if (b->nodes4q) cudaFree(b->nodes4q);
    if (b->tris) cudaFree(b->tris);
    *b = DeviceBlas();
}

bool BuildBlasOnDevice(const float* d_vertices, const int32_t* d_tri_idx, int first_tri, int tri_count, float int_test_eps, int algorithm,
                       int node_base, int node4_base, int tri_base, int max_depth2, int max_stack4, cudaStream_t st, DeviceBlas* out) {
    *out = DeviceBlas();
    const int n = tri_count;
    if (n < 8) throw std::runtime_error("bvh_device: tri_count must be >= 8");
    const bool sah = algorithm == kDeviceAlgoSah;
    const int total = 2 * n - 1;                                         // PLOC id space; the SAH build uses at most 2n build nodes
    const int ids = 2 * n;
    const int32_t* ix = d_tri_idx + 3 * (size_t)first_tri;

    // ---- temporaries
    float4 *plo, *phi, *clo[2], *chi[2], *nlo, *nhi;
    unsigned long long *keys[2], *flags, *scan;
    uint32_t *vals[2], *totals, *skeys[2], *sprims;
    int32_t *nn, *parent, *meta, *first, *live, *lscan, *ctr, *pnode;
    float *cst, *cb, *bins;
    int2* child;
    uint8_t* state;
    int4* items[2];
    SahNodes nd{};
    SahLevelTmp lt{};
    void* cub_tmp;
    const int bin_chunk = std::min(n, 1 << 17);                          // nodes binned per pass of a level (1536 B of bins each)
    size_t sort_bytes = 0, sort32_bytes = 0, scan64_bytes = 0, scan32_bytes = 0;
    BCK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 63, st));
    BCK(cub::DeviceRadixSort::SortPairs(nullptr, sort32_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 32, st));
    BCK(cub::DeviceScan::ExclusiveSum(nullptr, scan64_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, n, st));
    BCK(cub::DeviceScan::ExclusiveSum(nullptr, scan32_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, ids, st));
    size_t cub_bytes = std::max(std::max(sort_bytes, sort32_bytes), std::max(scan64_bytes, scan32_bytes));
    Arena ar;
    ar.Reserve(&plo, n); ar.Reserve(&phi, n);
    for (int k = 0; k < 2; k++) { ar.Reserve(&keys[k], n); ar.Reserve(&vals[k], n); ar.Reserve(&items[k], n); }
    ar.Reserve(&nlo, ids); ar.Reserve(&nhi, ids); ar.Reserve(&child, ids); ar.Reserve(&parent, ids); ar.Reserve(&meta, ids);
    ar.Reserve(&first, ids); ar.Reserve(&state, ids); ar.Reserve(&live, ids); ar.Reserve(&lscan, ids);
    ar.Reserve(&ctr, CTR_N); ar.Reserve(&cb, 6); ar.Reserve(&totals, 2);
    if (sah) {
        ar.Reserve(&nd.clo, ids); ar.Reserve(&nd.chi, ids); ar.Reserve(&nd.pos_min, ids); ar.Reserve(&nd.pos_max, ids); ar.Reserve(&nd.split, ids); ar.Reserve(&nd.pos_mid, ids);
        ar.Reserve(&lt.llo, n); ar.Reserve(&lt.lhi, n); ar.Reserve(&lt.rlo, n); ar.Reserve(&lt.rhi, n); ar.Reserve(&lt.lc, n); ar.Reserve(&lt.rc, n);
        ar.Reserve(&pnode, n); ar.Reserve(&skeys[0], n); ar.Reserve(&skeys[1], n); ar.Reserve(&sprims, n);
        ar.Reserve(&bins, (size_t)bin_chunk * 3 * kSahBins * kBinWords);
    } else {
        for (int k = 0; k < 2; k++) { ar.Reserve(&clo[k], n); ar.Reserve(&chi[k], n); }
        ar.Reserve(&cst, ids); ar.Reserve(&nn, n); ar.Reserve(&flags, n); ar.Reserve(&scan, n);
    }
    ar.Reserve((char**)&cub_tmp, cub_bytes);
    ar.Commit();
    nd.lo = nlo; nd.hi = nhi; nd.child = child; nd.parent = parent; nd.meta = meta;

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    struct Ev { cudaEvent_t& a; cudaEvent_t& b; ~Ev() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } ev_guard{e0, e1};
    BCK(cudaEventCreate(&e0)); BCK(cudaEventCreate(&e1));
    BCK(cudaEventRecord(e0, st));
    auto run = [&]() -> bool {                                           // false: refused (caller falls back to the host builder)
        // B1, B2
        const float cb_init[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
        BCK(cudaMemcpyAsync(cb, cb_init, sizeof cb_init, cudaMemcpyHostToDevice, st));
        k_prim_bounds<<<Grid(n), kTB, 0, st>>>(d_vertices, ix, n, int_test_eps, plo, phi, cb);
        k_morton<<<Grid(n), kTB, 0, st>>>(plo, phi, n, cb, keys[0], vals[0]);
        BCK(cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, keys[0], keys[1], vals[0], vals[1], n, 0, 63, st));
        const uint32_t* order = vals[1];

        // the finished binary tree, as the emission stage sees it
        int id_base = 0, id_count = 0, reverse = 0, root = 0;
        BCK(cudaMemsetAsync(ctr, 0, CTR_N * sizeof(int32_t), st));
        BCK(cudaMalloc((void**)&out->tris, (size_t)n * 3 * sizeof(float4)));

        if (sah) {
            // ---- S: level-synchronous binned SAH
            const float4 root_lo = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 0.f), root_hi = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.f);
            const int32_t root_meta = n, root_parent = -1;
            BCK(cudaMemcpyAsync(nd.lo, &root_lo, 16, cudaMemcpyHostToDevice, st));
            BCK(cudaMemcpyAsync(nd.hi, &root_hi, 16, cudaMemcpyHostToDevice, st));
            BCK(cudaMemcpyAsync(nd.meta, &root_meta, 4, cudaMemcpyHostToDevice, st));
            BCK(cudaMemcpyAsync(nd.parent, &root_parent, 4, cudaMemcpyHostToDevice, st));
            k_sah_root<<<Grid(n), kTB, 0, st>>>(nd, n, plo, phi, pnode);
            int lvl_begin = 0, lvl_end = 1, depth = 0;
            while (lvl_end > lvl_begin) {
                if (depth > kSahMaxDepth + 40) return false;
                const int width = lvl_end - lvl_begin;
                k_sah_init_nodes<<<Grid(width), kTB, 0, st>>>(nd, lvl_begin, lvl_end);
                k_sah_cbounds<<<Grid(n), kTB, 0, st>>>(nd, n, lvl_begin, order, plo, phi, pnode);
                for (int cb0 = lvl_begin; cb0 < lvl_end; cb0 += bin_chunk) {
                    int cb1 = std::min(lvl_end, cb0 + bin_chunk);
                    long long nbins = (long long)(cb1 - cb0) * 3 * kSahBins;
                    k_sah_init_bins<<<Grid(nbins), kTB, 0, st>>>(bins, nbins);
                    k_sah_bin<<<Grid(n), kTB, 0, st>>>(nd, n, cb0, cb1, order, plo, phi, pnode, bins);
                    k_sah_split<<<(cb1 - cb0 + 127) / 128, 128, 0, st>>>(nd, lvl_begin, cb0, cb1, depth, bins, live, lt);
                }
                BCK(cub::DeviceScan::ExclusiveSum(cub_tmp, scan32_bytes, live, lscan, width, st));
                k_sah_children<<<Grid(width), kTB, 0, st>>>(nd, lvl_begin, lvl_end, lvl_end, live, lscan, lt);
                int32_t h_last[2] = {0, 0};
                BCK(cudaMemcpyAsync(&h_last[0], live + (width - 1), 4, cudaMemcpyDeviceToHost, st));
                BCK(cudaMemcpyAsync(&h_last[1], lscan + (width - 1), 4, cudaMemcpyDeviceToHost, st));
                k_sah_assign<<<Grid(n), kTB, 0, st>>>(nd, n, lvl_begin, order, plo, phi, pnode);
                BCK(cudaStreamSynchronize(st));
                const int splits = h_last[0] + h_last[1];
                if (lvl_end + 2 * splits > ids) throw std::runtime_error("bvh_device: SAH build node overflow");
                lvl_begin = lvl_end; lvl_end += 2 * splits;
                if (splits) depth++;
            }
            const int nb = lvl_end;
            out->ploc_iterations = depth;
            out->depth2 = depth;
            k_sah_layout<<<Grid(nb), kTB, 0, st>>>(nd, nb, first, state);
            k_sah_keys<<<Grid(n), kTB, 0, st>>>(n, pnode, first, skeys[0]);
            BCK(cub::DeviceRadixSort::SortPairs(cub_tmp, sort32_bytes, skeys[0], skeys[1], order, sprims, n, 0, 32, st));   // stable: Morton order inside a leaf
            k_emit_tris_ordered<<<Grid(n), kTB, 0, st>>>(n, sprims, d_vertices, ix, first_tri, out->tris);
            id_base = 0; id_count = nb; reverse = 0; root = 0;
        } else {
            k_init_clusters<<<Grid(n), kTB, 0, st>>>(order, plo, phi, n, clo[0], chi[0], nlo, nhi, parent, meta, cst);
            // ---- B3
            int top = kTopClusters;
            if (const char* e = getenv("RT795_PLOC_TOP")) top = std::max(1, atoi(e));
            int m = n, next_id = n, cur = 0, iters = 0;
            int radius = kPlocRadius;
            if (const char* e = getenv("RT795_PLOC_RADIUS")) radius = std::min(kMaxPlocRadius, std::max(1, atoi(e)));
            while (m > top) {
                if (++iters > kMaxPlocIterations) return false;
                k_nearest<<<Grid(m), kTB, 0, st>>>(clo[cur], chi[cur], m, radius, nn);
                k_merge_flags<<<Grid(m), kTB, 0, st>>>(nn, m, flags);
                BCK(cub::DeviceScan::ExclusiveSum(cub_tmp, scan64_bytes, flags, scan, m, st));
                k_merge_emit<<<Grid(m), kTB, 0, st>>>(clo[cur], chi[cur], nn, flags, scan, m, next_id, clo[cur ^ 1], chi[cur ^ 1], nlo, nhi, child, parent, meta, cst, totals);
                uint32_t h_tot[2] = {0, 0};
                BCK(cudaMemcpyAsync(h_tot, totals, sizeof h_tot, cudaMemcpyDeviceToHost, st));
                BCK(cudaStreamSynchronize(st));
                if ((int)h_tot[0] >= m || h_tot[1] == 0) return false;     // no progress (NaN boxes): leave it to the host builder
                m = (int)h_tot[0]; next_id += (int)h_tot[1]; cur ^= 1;
            }
            out->ploc_iterations = iters;
            if (m > 1) {
                // B3b: binned SAH over the remaining clusters on the host; node ids [next_id, total) = reversed flatten order (root last)
                std::vector<float4> hlo(m), hhi(m);
                std::vector<int32_t> hmeta(m);
                std::vector<float> hcst(m);
                int32_t* g_meta = live;                                      // scratch not in use yet
                float* g_cst = reinterpret_cast<float*>(lscan);
                k_gather_clusters<<<Grid(m), kTB, 0, st>>>(clo[cur], m, meta, cst, g_meta, g_cst);
                BCK(cudaMemcpyAsync(hlo.data(), clo[cur], (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, st));
                BCK(cudaMemcpyAsync(hhi.data(), chi[cur], (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, st));
                BCK(cudaMemcpyAsync(hmeta.data(), g_meta, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
                BCK(cudaMemcpyAsync(hcst.data(), g_cst, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
                BCK(cudaStreamSynchronize(st));
                std::vector<Aabb> pb(m);
                for (int i = 0; i < m; i++) { pb[i].lo[0] = hlo[i].x; pb[i].lo[1] = hlo[i].y; pb[i].lo[2] = hlo[i].z; pb[i].hi[0] = hhi[i].x; pb[i].hi[1] = hhi[i].y; pb[i].hi[2] = hhi[i].z; }
                BuiltBvh tb;
                BuildSahBvh(pb.data(), m, 1, &tb);
                const int inner = (int)tb.nodes.size();
                if (inner != m - 1) throw std::runtime_error("bvh_device: top tree is not a full binary tree over the clusters");
                std::vector<float4> tlo(inner), thi(inner);
                std::vector<int2> tch(inner);
                std::vector<int32_t> tmeta(inner);
                std::vector<float> tcst(inner);
                auto slot_of = [&](int i) { return inner - 1 - i; };          // flatten index -> slot in [next_id, total): root (0) -> last
                for (int i = inner - 1; i >= 0; i--) {                        // children come after their parent in the flatten order
                    const BvhNode64& nd = tb.nodes[i];
                    int ids[2], cnt[2]; float cs[2], area[2];
                    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
                    for (int c = 0; c < 2; c++) {
                        int32_t ref; memcpy(&ref, &nd.q[12 + c], 4);
                        float blo[3] = {nd.q[c * 4 + 0], nd.q[c * 4 + 2], nd.q[8 + c * 2 + 0]}, bhi[3] = {nd.q[c * 4 + 1], nd.q[c * 4 + 3], nd.q[8 + c * 2 + 1]};
                        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], blo[k]); hi[k] = std::max(hi[k], bhi[k]); }
                        float dx = bhi[0] - blo[0], dy = bhi[1] - blo[1], dz = bhi[2] - blo[2];
                        area[c] = dx * dy + dy * dz + dz * dx;
                        if (ref >= 0) { int sl = slot_of(ref); ids[c] = next_id + sl; cnt[c] = tmeta[sl] & kCntMask; cs[c] = tcst[sl]; }
                        else { int ci = tb.order[(~ref) >> 3]; int32_t nid; memcpy(&nid, &hlo[ci].w, 4); ids[c] = nid; cnt[c] = hmeta[ci] & kCntMask; cs[c] = hcst[ci]; }
                    }
                    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
                    float a = dx * dy + dy * dz + dz * dx;
                    int c = cnt[0] + cnt[1];
                    float split = a > 0.f ? kSahTraversalCost + (area[0] * cs[0] + area[1] * cs[1]) / a : 3.0e38f;
                    if (!(split <= 3.0e38f)) split = 3.0e38f;
                    bool leaf = c <= 4 && (float)c <= split;
                    int sl = slot_of(i);
                    int32_t idbits = next_id + sl; float idf; memcpy(&idf, &idbits, 4);
                    tlo[sl] = make_float4(lo[0], lo[1], lo[2], idf); thi[sl] = make_float4(hi[0], hi[1], hi[2], 0.f);
                    tch[sl] = make_int2(ids[0], ids[1]);
                    tmeta[sl] = c | (leaf ? kLeafBit : 0); tcst[sl] = leaf ? (float)c : split;
                }
                BCK(cudaMemcpyAsync(nlo + next_id, tlo.data(), (size_t)inner * sizeof(float4), cudaMemcpyHostToDevice, st));
                BCK(cudaMemcpyAsync(nhi + next_id, thi.data(), (size_t)inner * sizeof(float4), cudaMemcpyHostToDevice, st));
                BCK(cudaMemcpyAsync(child + next_id, tch.data(), (size_t)inner * sizeof(int2), cudaMemcpyHostToDevice, st));
                BCK(cudaMemcpyAsync(meta + next_id, tmeta.data(), (size_t)inner * 4, cudaMemcpyHostToDevice, st));
                BCK(cudaMemcpyAsync(cst + next_id, tcst.data(), (size_t)inner * 4, cudaMemcpyHostToDevice, st));
                k_set_parents<<<Grid(inner), kTB, 0, st>>>(next_id, inner, child, parent, next_id + inner - 1);
                BCK(cudaStreamSynchronize(st));                               // the host vectors above go out of scope
                next_id += inner;
            }
            if (next_id != total) throw std::runtime_error("bvh_device: node count mismatch after clustering");
            // B4, B5
            k_layout<<<Grid(total), kTB, 0, st>>>(total, parent, child, meta, first, state, ctr);
            k_emit_tris<<<Grid(n), kTB, 0, st>>>(n, first, order, d_vertices, ix, first_tri, out->tris);
            id_base = n; id_count = n - 1; reverse = 1; root = total - 1;
        }

        // ---- emission: binary nodes, 4-wide collapse, quantised twin
        k_live_flags<<<Grid(id_count), kTB, 0, st>>>(id_base, id_count, state, live);
        BCK(cub::DeviceScan::ExclusiveSum(cub_tmp, scan32_bytes, live, lscan, id_count, st));
        int32_t h_last[2] = {0, 0}, h_ctr[CTR_N];
        BCK(cudaMemcpyAsync(&h_last[0], live + (id_count - 1), 4, cudaMemcpyDeviceToHost, st));
        BCK(cudaMemcpyAsync(&h_last[1], lscan + (id_count - 1), 4, cudaMemcpyDeviceToHost, st));
        BCK(cudaMemcpyAsync(h_ctr, ctr, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
        float h_root[8];
        BCK(cudaMemcpyAsync(h_root, nlo + root, 16, cudaMemcpyDeviceToHost, st));
        BCK(cudaMemcpyAsync(h_root + 4, nhi + root, 16, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        const int live_total = h_last[0] + h_last[1];
        if (!sah) out->depth2 = h_ctr[CTR_DEPTH];
        for (int k = 0; k < 3; k++) { out->bounds.lo[k] = h_root[k]; out->bounds.hi[k] = h_root[4 + k]; }
        if (live_total < 1) throw std::runtime_error("bvh_device: root collapsed into a leaf");
        if (out->depth2 > max_depth2) return false;
        out->n2 = live_total;
        BCK(cudaMalloc((void**)&out->nodes2, (size_t)live_total * 4 * sizeof(float4)));
        k_emit_bvh2<<<Grid(id_count), kTB, 0, st>>>(id_base, id_count, reverse, live_total, live, lscan, child, state, first, meta, nlo, nhi, node_base, tri_base, out->nodes2);

        // B6: every 4-wide node is a distinct live binary node, so live_total bounds their number
        BCK(cudaMalloc((void**)&out->nodes4, (size_t)live_total * 8 * sizeof(float4)));
        {
            int4 root_item = make_int4(root, 0, 0, 0);
            BCK(cudaMemcpyAsync(items[0], &root_item, sizeof root_item, cudaMemcpyHostToDevice, st));
            int32_t init_ctr[CTR_N] = {0, 1, 0, out->depth2};
            BCK(cudaMemcpyAsync(ctr, init_ctr, sizeof init_ctr, cudaMemcpyHostToDevice, st));
            int n_items = 1, q = 0;
            while (n_items > 0) {
                k_collapse4<<<Grid(n_items), kTB, 0, st>>>(items[q], n_items, items[q ^ 1], ctr, child, state, first, meta, nlo, nhi, node4_base, tri_base, out->nodes4);
                BCK(cudaMemcpyAsync(h_ctr, ctr, sizeof h_ctr, cudaMemcpyDeviceToHost, st));
                BCK(cudaMemsetAsync(ctr + CTR_QOUT, 0, sizeof(int32_t), st));
                BCK(cudaStreamSynchronize(st));
                n_items = h_ctr[CTR_QOUT]; q ^= 1;
            }
            out->n4 = h_ctr[CTR_ALLOC];
            out->max_stack4 = h_ctr[CTR_STACK];
            if (out->n4 > live_total) throw std::runtime_error("bvh_device: 4-wide node count exceeds its bound");
            if (out->max_stack4 > max_stack4) return false;
        }
        // B7
        BCK(cudaMalloc((void**)&out->nodes4q, (size_t)out->n4 * 4 * sizeof(uint4)));
        k_quantise4<<<Grid(out->n4), kTB, 0, st>>>(out->nodes4, out->n4, out->nodes4q);
        BCK(cudaEventRecord(e1, st));
        BCK(cudaStreamSynchronize(st));
        BCK(cudaGetLastError());
        BCK(cudaEventElapsedTime(&out->ms, e0, e1));
        out->tri_count = n;
        return true;
    };
    bool ok = false;
    try {
        ok = run();
    } catch (...) {
        cudaStreamSynchronize(st);
        FreeDeviceBlas(out);
        throw;
    }
    if (!ok) {
        cudaStreamSynchronize(st);
        int iters = out->ploc_iterations, d2 = out->depth2, s4 = out->max_stack4;
        FreeDeviceBlas(out);
        out->ploc_iterations = iters; out->depth2 = d2; out->max_stack4 = s4;   // why it was refused, for the caller's log
    }
    return ok;
}

}  // namespace rt795
