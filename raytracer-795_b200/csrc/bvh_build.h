// H1: host binned-SAH BVH builder + flattener.
//
// Replaces the reference's per-object, pointer-based, median-split BVH (BVH::ConstructionHelper / FindMedian /
// ComputeBoundingBox, src/BVH.cpp:64-135, 268-283) and the flat top-level loop over objects and instances
// (BVHMethods::FindIntersection, src/Helper.cpp:32-73) with a two-level BVH2 in one linear node array:
//   - one BLAS per mesh (shared by every instance of it), leaves of <= 4 triangles, triangles re-ordered into
//     leaf order as 48-byte {v0, e1, e2} records;
//   - one TLAS over all top-level entries (world AABB of the transformed BLAS / sphere, swept by MotionBlur*[0,1]).
// Any conservative BVH returns the same closest hit as the reference's; only tie order differs.
//
// Node format (64 bytes, 16-byte aligned, 4 x float4) - the two children's boxes live in the parent so one
// 64-byte fetch decides both:
//   q0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   q1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   q2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   q3 = (bits child0, bits child1, 0, 0)
// child >= 0: index of an inner node; child < 0: leaf, ~child = (first << 3) | (count - 1), `first` indexing the
// traversal-triangle array (BLAS) or being the top-level entry index (TLAS, count 1).  An absent child (only in the
// synthetic root of a BVH whose real root is a leaf) is kNoChild and is never entered.
#pragma once
#include <cstdint>
#include <vector>

namespace rt795 {

struct Aabb {
    float lo[3], hi[3];
    void Reset() { for (int k = 0; k < 3; k++) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; } }
    void Grow(const float* p) { for (int k = 0; k < 3; k++) { if (p[k] < lo[k]) lo[k] = p[k]; if (p[k] > hi[k]) hi[k] = p[k]; } }
    void Grow(const Aabb& b) { for (int k = 0; k < 3; k++) { if (b.lo[k] < lo[k]) lo[k] = b.lo[k]; if (b.hi[k] > hi[k]) hi[k] = b.hi[k]; } }
    float HalfArea() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct BvhNode64 { float q[16]; };   // layout above

struct BuiltBvh {
    std::vector<BvhNode64> nodes;    // nodes[0] is the root (always an inner node)
    std::vector<int32_t> order;      // leaf-ordered primitive indices (into the input array)
    Aabb bounds;
    int max_depth = 0;
};

// Builds a BVH2 over `n` primitives given their boxes.  max_leaf in [1,8].  Child leaf payloads index `order`.
void BuildSahBvh(const Aabb* prim_bounds, int n, int max_leaf, BuiltBvh* out);

constexpr int32_t kNoChild = 0x7ffffffd;
// SAH leaf decision: cost of one inner-node visit in units of one triangle test (leaf cost = #triangles).  Measured on C4 with the
// traversal kernels of round 2 (profiles/bench/r02_ab_sah_c4.txt): 1.0 -> 51.3 ms / 29.4 MB of nodes, 2.0 -> 50.7 ms / 19.0 MB (a node
// visit costs ~100 issued instructions against ~35 for a triangle test), 4.0 -> 51.2 ms / 9.9 MB.  The device builders use the same value.
constexpr float kSahTraversalCost = 2.0f;
inline int32_t EncodeLeaf(int32_t first, int32_t count) { return ~((first << 3) | (count - 1)); }

// 4-wide node (128 bytes = one cache line, 8 x float4): SoA boxes of up to four children
//   q0 = lo.x[0..3]  q1 = hi.x[0..3]  q2 = lo.y[0..3]  q3 = hi.y[0..3]  q4 = lo.z[0..3]  q5 = hi.z[0..3]
//   q6 = bits child[0..3] (same encoding as BvhNode64; absent = kNoChild)   q7 = unused
// Built by collapsing the SAH BVH2: a node's children are expanded, largest surface area first, until four.
// Halves the number of dependent node fetches per ray and doubles the independent box tests per fetch.
struct BvhNode128 { float q[32]; };

struct BuiltBvh4 {
    std::vector<BvhNode128> nodes;   // nodes[0] is the root
    int max_stack = 0;               // worst-case traversal stack entries (sum over a root-to-leaf path of children-1)
};

void CollapseToBvh4(const BuiltBvh& b2, BuiltBvh4* out);

// Quantised 4-wide node (64 bytes, 4 x float4): child boxes as 8-bit offsets on the node's own grid
//   q0 = (origin.x, origin.y, origin.z, scale.z)          q1 = bits child[0..3]
//   q2 = (lo.x[4], lo.y[4], lo.z[4], hi.x[4]) one byte per child, packed little-endian into a word per plane set
//   q3 = (hi.y[4], hi.z[4], scale.x, scale.y)
// plane = origin + byte * scale, rounded outwards so the decoded box always contains the fp32 one.  An absent child has the
// inverted range lo = 255, hi = 0.  Four 16-byte loads per node instead of seven: the L1 LSU wavefront rate is the busiest
// unit of the occlusion kernel (77 % of peak, ncu r01e), every lane's 16-byte load of a divergent node being its own wavefront.
struct BvhNode64Q { uint32_t w[16]; };
void QuantiseBvh4(const BuiltBvh4& b4, std::vector<BvhNode64Q>* out);

}  // namespace rt795
