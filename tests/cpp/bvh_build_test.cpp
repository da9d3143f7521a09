// CPU-only check of the host SAH builder (raytracer-795_b200/csrc/bvh_build.cpp): structural invariants of the binary
// BVH and of its 4-wide collapse, and that the multi-threaded build returns exactly the single-threaded tree.
// Built and run by tests/test_bvh_build.py.
#include "bvh_build.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

using namespace rt795;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)

static Aabb ChildBox2(const BvhNode64& n, int s) {
    Aabb b;
    b.lo[0] = n.q[s * 4 + 0]; b.hi[0] = n.q[s * 4 + 1]; b.lo[1] = n.q[s * 4 + 2]; b.hi[1] = n.q[s * 4 + 3]; b.lo[2] = n.q[8 + s * 2]; b.hi[2] = n.q[8 + s * 2 + 1];
    return b;
}
static Aabb ChildBox4(const BvhNode128& n, int k) {
    Aabb b;
    b.lo[0] = n.q[k]; b.hi[0] = n.q[4 + k]; b.lo[1] = n.q[8 + k]; b.hi[1] = n.q[12 + k]; b.lo[2] = n.q[16 + k]; b.hi[2] = n.q[20 + k];
    return b;
}
static bool Contains(const Aabb& o, const Aabb& i) {
    for (int k = 0; k < 3; k++) if (i.lo[k] < o.lo[k] || i.hi[k] > o.hi[k]) return false;
    return true;
}

static void CheckTree(const std::vector<Aabb>& pb, int max_leaf) {
    int n = (int)pb.size();
    BuiltBvh b2;
    BuildSahBvh(pb.data(), n, max_leaf, &b2);
    CHECK((int)b2.order.size() == n);
    std::vector<int> seen(n, 0);
    // binary tree: each leaf range inside its child box, each primitive exactly once
    std::vector<std::pair<int, Aabb>> st;
    Aabb all; all.Reset(); for (auto& b : pb) all.Grow(b);
    if (n) st.push_back({0, all});
    while (!st.empty()) {
        auto [idx, bound] = st.back(); st.pop_back();
        const BvhNode64& nd = b2.nodes[idx];
        for (int s = 0; s < 2; s++) {
            int32_t c; memcpy(&c, &nd.q[12 + s], 4);
            if (c == kNoChild) continue;
            Aabb cb = ChildBox2(nd, s);
            CHECK(Contains(bound, cb));
            if (c >= 0) st.push_back({c, cb});
            else {
                int payload = ~c, first = payload >> 3, count = (payload & 7) + 1;
                CHECK(count <= max_leaf);
                for (int i = first; i < first + count; i++) { CHECK(i < n); seen[b2.order[i]]++; CHECK(Contains(cb, pb[b2.order[i]])); }
            }
        }
    }
    for (int i = 0; i < n; i++) CHECK(seen[i] == 1);
    // 4-wide collapse: same leaves, boxes still conservative, stack bound honest
    BuiltBvh4 b4;
    CollapseToBvh4(b2, &b4);
    std::vector<int> seen4(n, 0);
    struct It { int idx; Aabb bound; int stack; };
    std::vector<It> st4;
    if (n) st4.push_back({0, all, 0});
    int worst = 0;
    while (!st4.empty()) {
        It it = st4.back(); st4.pop_back();
        const BvhNode128& nd = b4.nodes[it.idx];
        int nchild = 0;
        for (int k = 0; k < 4; k++) { int32_t c; memcpy(&c, &nd.q[24 + k], 4); if (c != kNoChild) nchild++; }
        int here = it.stack + nchild - 1;
        if (here > worst) worst = here;
        for (int k = 0; k < 4; k++) {
            int32_t c; memcpy(&c, &nd.q[24 + k], 4);
            if (c == kNoChild) continue;
            Aabb cb = ChildBox4(nd, k);
            CHECK(Contains(it.bound, cb));
            if (c >= 0) st4.push_back({c, cb, here});
            else {
                int payload = ~c, first = payload >> 3, count = (payload & 7) + 1;
                for (int i = first; i < first + count; i++) { seen4[b2.order[i]]++; CHECK(Contains(cb, pb[b2.order[i]])); }
            }
        }
    }
    for (int i = 0; i < n; i++) CHECK(seen4[i] == 1);
    CHECK(worst == b4.max_stack);
    // 8-bit quantised copy: same refs, decoded boxes contain the fp32 boxes, absent children have an inverted range
    std::vector<BvhNode64Q> q4;
    QuantiseBvh4(b4, &q4);
    CHECK(q4.size() == b4.nodes.size());
    for (size_t i = 0; i < q4.size(); i++) {
        const BvhNode64Q& q = q4[i];
        float origin[3], scale[3];
        memcpy(&origin[0], &q.w[0], 4); memcpy(&origin[1], &q.w[1], 4); memcpy(&origin[2], &q.w[2], 4);
        memcpy(&scale[2], &q.w[3], 4); memcpy(&scale[0], &q.w[14], 4); memcpy(&scale[1], &q.w[15], 4);
        const uint32_t lo_w[3] = {q.w[8], q.w[9], q.w[10]}, hi_w[3] = {q.w[11], q.w[12], q.w[13]};
        for (int k = 0; k < 4; k++) {
            int32_t c4, cq; memcpy(&c4, &b4.nodes[i].q[24 + k], 4); memcpy(&cq, &q.w[4 + k], 4);
            CHECK(c4 == cq);
            for (int a = 0; a < 3; a++) {
                unsigned bl = (lo_w[a] >> (8 * k)) & 255u, bh = (hi_w[a] >> (8 * k)) & 255u;
                if (c4 == kNoChild) { CHECK(bl > bh); continue; }
                float dl = origin[a] + (float)bl * scale[a], dh = origin[a] + (float)bh * scale[a];
                Aabb cb = ChildBox4(b4.nodes[i], k);
                CHECK(dl <= cb.lo[a] && dh >= cb.hi[a]);
                CHECK(dh - dl <= (cb.hi[a] - cb.lo[a]) + 2.5f * scale[a]);      // and not looser than two grid steps
            }
        }
    }
    CHECK(b4.nodes.size() <= b2.nodes.size());
}

int main() {
    std::mt19937 g(795);
    std::uniform_real_distribution<float> u(-1, 1);
    for (int n : {0, 1, 2, 3, 5, 17, 1000, 60000}) {
        std::vector<Aabb> pb(n);
        for (auto& b : pb) { float c[3] = {u(g), u(g), u(g)}; for (int k = 0; k < 3; k++) { b.lo[k] = c[k] - 0.01f * std::fabs(u(g)); b.hi[k] = c[k] + 0.01f * std::fabs(u(g)); } }
        CheckTree(pb, 4);
        CheckTree(pb, 1);
    }
    {   // all centroids identical: the builder must fall back to median splits and still terminate
        std::vector<Aabb> pb(100);
        for (auto& b : pb) for (int k = 0; k < 3; k++) { b.lo[k] = -1; b.hi[k] = 1; }
        CheckTree(pb, 4);
    }
    {   // threads do not change the tree
        int n = 200000;
        std::vector<Aabb> pb(n);
        for (auto& b : pb) { float c[3] = {u(g), u(g), u(g)}; for (int k = 0; k < 3; k++) { b.lo[k] = c[k] - 0.002f; b.hi[k] = c[k] + 0.002f; } }
        BuiltBvh a, b;
        setenv("RT795_BVH_PAR", "0", 1); BuildSahBvh(pb.data(), n, 4, &a);
        setenv("RT795_BVH_PAR", "3", 1); BuildSahBvh(pb.data(), n, 4, &b);
        unsetenv("RT795_BVH_PAR");
        CHECK(a.order == b.order);
        CHECK(a.nodes.size() == b.nodes.size());
        CHECK(a.max_depth == b.max_depth);
        // same tree up to node numbering
        std::vector<std::pair<int, int>> st{{0, 0}};
        while (!st.empty() && !fails) {
            auto [x, y] = st.back(); st.pop_back();
            CHECK(memcmp(a.nodes[x].q, b.nodes[y].q, 48) == 0);
            for (int s = 0; s < 2; s++) {
                int32_t ca, cb; memcpy(&ca, &a.nodes[x].q[12 + s], 4); memcpy(&cb, &b.nodes[y].q[12 + s], 4);
                if (ca >= 0 && ca != kNoChild) { CHECK(cb >= 0 && cb != kNoChild); st.push_back({ca, cb}); } else CHECK(ca == cb);
            }
        }
    }
    printf(fails ? "bvh_build_test: %d failure(s)\n" : "bvh_build_test: ok\n", fails);
    return fails ? 1 : 0;
}
