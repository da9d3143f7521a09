// D1: device BLAS builder - the whole per-mesh acceleration-structure build on the GPU.
//
// Replaces, for large meshes, the host binned-SAH build of bvh_build.cpp (itself the replacement of the reference's
// per-object median-split BVH, src/BVH.cpp:64-135, whose FindMedian re-sorts the primitives at every node and takes
// seconds on 871 k triangles).  At GPU render speeds (C4: 62 ms/frame) a 0.3-0.7 s host build dominates a one-frame run.
//
// Pipeline (all kernels in bvh_device.cu, one stream, no host work besides loop control):
//   B1 k_prim_bounds   padded triangle boxes (same padding rule as the host builder) + centroid bounds
//   B2 k_morton        63-bit Morton keys of the box centres; sorted with cub::DeviceRadixSort (library sort)
//   B3 PLOC            parallel locally-ordered clustering (Meister & Bittner 2018): every cluster finds its nearest
//                      neighbour (smallest merged surface area) within +-kPlocRadius positions of the Morton-ordered
//                      cluster array, mutual pairs merge, the array is compacted; repeat until one cluster is left.
//                      SAH leaf decision (<= 4 triangles, leaf cost <= split cost) is taken bottom-up while merging.
//   B4 k_layout        every node walks to the root once: DFS leaf offset (so any subtree owns a contiguous triangle
//                      range), "inside a collapsed leaf" flag, depth
//   B5 k_emit_tris / k_emit_bvh2   48-byte {v0, e1, e2} triangle records in leaf order; 64-byte binary nodes
//   B6 k_collapse4     level-synchronous BVH2 -> 4-wide collapse (largest-area child opened first, as the host collapse)
//   B7 k_quantise4     8-bit quantised twin of the 4-wide nodes (same function as the host, bvh_quant.h)
// Output formats are exactly those of bvh_build.h, child references already rebased into the scene-wide arrays.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "bvh_build.h"

namespace rt795 {

struct DeviceBlas {
    // device buffers (cudaMalloc; released by FreeDeviceBlas)
    float4* nodes2 = nullptr;   // n2 x BvhNode64
    float4* nodes4 = nullptr;   // n4 x BvhNode128
    uint4* nodes4q = nullptr;   // n4 x BvhNode64Q
    float4* tris = nullptr;     // tri_count x 3 float4
    int n2 = 0, n4 = 0, tri_count = 0;
    Aabb bounds;
    int depth2 = 0;             // as BuiltBvh::max_depth
    int max_stack4 = 0;         // as BuiltBvh4::max_stack
    int ploc_iterations = 0;
    float ms = 0.f;             // device time of the whole build (CUDA events)
};

// Builds the BLAS of triangles [first_tri, first_tri + tri_count) of the scene-wide index array.  `d_vertices` /
// `d_tri_idx` are the already uploaded scene arrays.  Child references are emitted with node_base / node4_base /
// tri_base added (positions of this mesh's segments in the scene-wide arrays).  Returns false (and frees everything)
// when the build does not converge or the tree is deeper than max_depth2 / needs more than max_stack4 stack entries - the
// caller then uses the host builder.  Throws std::runtime_error on CUDA errors.  tri_count must be >= 8.
constexpr int kDeviceAlgoPloc = 0, kDeviceAlgoSah = 1;
bool BuildBlasOnDevice(const float* d_vertices, const int32_t* d_tri_idx, int first_tri, int tri_count, float int_test_eps, int algorithm,
                       int node_base, int node4_base, int tri_base, int max_depth2, int max_stack4, cudaStream_t st, DeviceBlas* out);
void FreeDeviceBlas(DeviceBlas* b);

}  // namespace rt795
