#include "bvh_build.h"
#include "bvh_quant.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>

namespace rt795 {
namespace {

constexpr int kBins = 16;
constexpr float kIntersectCost = 1.0f;
// cost of one inner-node visit relative to one triangle test in the SAH leaf decision.  RT795_SAH_TRAVERSAL_COST / RT795_SAH_MAX_LEAF:
// tuning overrides for the HOST builder only (A/B of the leaf size; the device builder mirrors the defaults)
static float TraversalCost() { static float v = [] { const char* e = getenv("RT795_SAH_TRAVERSAL_COST"); float x = e ? (float)atof(e) : kSahTraversalCost; return x > 0.f ? x : kSahTraversalCost; }(); return v; }
constexpr int kMaxDepth = 36;   // deeper subtrees fall back to median splits (depth <= 36 + log2 n); core.cu checks TLAS+BLAS depth against the traversal stack

struct Tmp {
    Aabb box;
    int left = -1, right = -1;    // Tmp indices, or -1
    int first = 0, count = 0;     // leaf range in `order`
    bool leaf = false;
};

struct Builder {
    const Aabb* pb;
    std::vector<int32_t>& order;
    std::vector<float> cen_own;            // centroids, 3 per primitive (owned by the root builder)
    const std::vector<float>* cen_ref = nullptr;
    const std::vector<float>& cen;         // centroids as seen by this (sub-)builder
    std::vector<Tmp> tmp;
    int max_leaf;
    int max_depth = 0;

    Builder(const Aabb* b, int n, int ml, std::vector<int32_t>& ord) : pb(b), order(ord), cen(cen_own), max_leaf(ml) {
        cen_own.resize((size_t)n * 3);
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) cen_own[(size_t)i * 3 + k] = 0.5f * (b[i].lo[k] + b[i].hi[k]);
        cen_ref = &cen_own;
    }

    // Decides what to do with primitives [first, first+count): returns -1 for "make a leaf", else the split position
    // (the range is partitioned in place).  Fills the node box.
    int Split(int first, int count, int depth, Aabb* box_out) {
        Aabb box, cbox;
        box.Reset(); cbox.Reset();
        for (int i = first; i < first + count; i++) { box.Grow(pb[order[i]]); cbox.Grow(&cen[(size_t)order[i] * 3]); }
        *box_out = box;
        if (count <= 1) return -1;
        int mid = -1;
        if (depth < kMaxDepth) {
            // binned SAH over the three axes
            float best_cost = std::numeric_limits<float>::max();
            int best_axis = -1, best_split = -1;
            for (int axis = 0; axis < 3; axis++) {
                float lo = cbox.lo[axis], ext = cbox.hi[axis] - lo;
                if (!(ext > 0)) continue;
                float scale = kBins / ext;
                Aabb bb[kBins];
                int bc[kBins];
                for (int b = 0; b < kBins; b++) { bb[b].Reset(); bc[b] = 0; }
                for (int i = first; i < first + count; i++) {
                    int p = order[i];
                    int b = std::min(kBins - 1, std::max(0, (int)((cen[(size_t)p * 3 + axis] - lo) * scale)));
                    bb[b].Grow(pb[p]); bc[b]++;
                }
                float ra[kBins]; int rc[kBins];
                Aabb acc; acc.Reset(); int c = 0;
                for (int b = kBins - 1; b >= 1; b--) { acc.Grow(bb[b]); c += bc[b]; ra[b] = acc.HalfArea(); rc[b] = c; }
                acc.Reset(); c = 0;
                for (int b = 0; b < kBins - 1; b++) {
                    acc.Grow(bb[b]); c += bc[b];
                    if (c == 0 || rc[b + 1] == 0) continue;
                    float cost = acc.HalfArea() * c + ra[b + 1] * rc[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
                }
            }
            float area = box.HalfArea();
            float leaf_cost = kIntersectCost * count;
            float split_cost = best_axis >= 0 && area > 0 ? TraversalCost() + kIntersectCost * best_cost / area : std::numeric_limits<float>::max();
            if (count <= max_leaf && leaf_cost <= split_cost) return -1;
            if (best_axis >= 0) {
                float lo = cbox.lo[best_axis], scale = kBins / (cbox.hi[best_axis] - lo);
                auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int32_t p) {
                    int b = std::min(kBins - 1, std::max(0, (int)((cen[(size_t)p * 3 + best_axis] - lo) * scale)));
                    return b <= best_split;
                });
                mid = (int)(it - order.begin());
                if (mid == first || mid == first + count) mid = -1;
            }
        }
        if (mid < 0) {
            if (count <= max_leaf) return -1;
            // centroids coincide (or depth cap reached): median split along the widest box axis
            int axis = 0;
            float e0 = box.hi[0] - box.lo[0], e1 = box.hi[1] - box.lo[1], e2 = box.hi[2] - box.lo[2];
            if (e1 > e0 && e1 >= e2) axis = 1; else if (e2 > e0 && e2 > e1) axis = 2;
            mid = first + count / 2;
            std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                             [&](int32_t a, int32_t b) { return cen[(size_t)a * 3 + axis] < cen[(size_t)b * 3 + axis]; });
        }
        return mid;
    }

    // `par` > 0: the two halves of a large range are built by separate threads into private node vectors and stitched in
    // (the tree does not depend on the thread count: every split is a function of its range only).
    int Build(int first, int count, int depth, int par = 0) {
        max_depth = std::max(max_depth, depth);
        int me = (int)tmp.size();
        tmp.emplace_back();
        Aabb box;
        int mid = Split(first, count, depth, &box);
        tmp[me].box = box;
        if (mid < 0) { tmp[me].leaf = true; tmp[me].first = first; tmp[me].count = count; return me; }
        if (par > 0 && count >= 32768) {
            Builder lb(*this, Shared()), rb(*this, Shared());
            int lroot = 0, rroot = 0;
            std::thread tl([&] { lroot = lb.Build(first, mid - first, depth + 1, par - 1); });
            rroot = rb.Build(mid, first + count - mid, depth + 1, par - 1);
            tl.join();
            auto stitch = [&](Builder& sb, int root) {
                int off = (int)tmp.size();
                for (Tmp t : sb.tmp) { if (t.left >= 0) t.left += off; if (t.right >= 0) t.right += off; tmp.push_back(t); }
                max_depth = std::max(max_depth, sb.max_depth);
                return root + off;
            };
            int l = stitch(lb, lroot), r = stitch(rb, rroot);
            tmp[me].left = l; tmp[me].right = r;
            return me;
        }
        int l = Build(first, mid - first, depth + 1, 0);
        int r = Build(mid, first + count - mid, depth + 1, 0);
        tmp[me].left = l; tmp[me].right = r;
        return me;
    }
    struct Shared {};
    Builder(const Builder& o, Shared) : pb(o.pb), order(o.order), cen_ref(o.cen_ref), cen(o.cen_ref ? *o.cen_ref : o.cen_own), max_leaf(o.max_leaf) {}
};

void PutChild(BvhNode64& n, int slot, const Aabb* box, int32_t child) {
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    if (box) { memcpy(lo, box->lo, 12); memcpy(hi, box->hi, 12); }
    float* q = n.q;
    q[slot * 4 + 0] = lo[0]; q[slot * 4 + 1] = hi[0]; q[slot * 4 + 2] = lo[1]; q[slot * 4 + 3] = hi[1];
    q[8 + slot * 2 + 0] = lo[2]; q[8 + slot * 2 + 1] = hi[2];
    memcpy(&q[12 + slot], &child, 4);
}

}  // namespace

void BuildSahBvh(const Aabb* prim_bounds, int n, int max_leaf, BuiltBvh* out) {
    out->nodes.clear();
    out->order.resize(n);
    for (int i = 0; i < n; i++) out->order[i] = i;
    out->bounds.Reset();
    if (n == 0) {
        BvhNode64 root; memset(&root, 0, sizeof root);
        PutChild(root, 0, nullptr, kNoChild); PutChild(root, 1, nullptr, kNoChild);
        out->nodes.push_back(root);
        return;
    }
    if (const char* e = getenv("RT795_SAH_MAX_LEAF")) { int v = atoi(e); if (v >= 1 && max_leaf > 1) max_leaf = v; }   // TLAS leaves (max_leaf 1) hold exactly one entry
    Builder b(prim_bounds, n, std::min(8, std::max(1, max_leaf)), out->order);
    b.tmp.reserve((size_t)n);
    unsigned hw = std::thread::hardware_concurrency();
    int par = hw >= 16 ? 4 : hw >= 8 ? 3 : hw >= 4 ? 2 : hw >= 2 ? 1 : 0;   // 2^par concurrent subtree builds
    if (const char* e = getenv("RT795_BVH_PAR")) par = atoi(e);             // 0 = single-threaded (same tree either way)
    int root = b.Build(0, n, 0, par);
    out->bounds = b.tmp[root].box;
    out->max_depth = b.max_depth;

    // flatten: inner Tmp nodes get consecutive indices in DFS order (children after parents, near each other)
    std::vector<int> idx(b.tmp.size(), -1);
    int inner = 0;
    if (b.tmp[root].leaf) inner = 1;   // synthetic root
    {
        std::vector<int> st{root};
        while (!st.empty()) {
            int t = st.back(); st.pop_back();
            if (b.tmp[t].leaf) continue;
            idx[t] = inner++;
            st.push_back(b.tmp[t].right); st.push_back(b.tmp[t].left);
        }
    }
    out->nodes.assign(inner, BvhNode64());
    for (auto& nd : out->nodes) memset(&nd, 0, sizeof nd);
    auto child_ref = [&](int t) -> int32_t { return b.tmp[t].leaf ? EncodeLeaf(b.tmp[t].first, b.tmp[t].count) : idx[t]; };
    if (b.tmp[root].leaf) {
        PutChild(out->nodes[0], 0, &b.tmp[root].box, child_ref(root));
        PutChild(out->nodes[0], 1, nullptr, kNoChild);
    }
    for (size_t t = 0; t < b.tmp.size(); t++) {
        if (b.tmp[t].leaf) continue;
        BvhNode64& nd = out->nodes[idx[t]];
        PutChild(nd, 0, &b.tmp[b.tmp[t].left].box, child_ref(b.tmp[t].left));
        PutChild(nd, 1, &b.tmp[b.tmp[t].right].box, child_ref(b.tmp[t].right));
    }
}


// ---------------------------------------------------------------------------------------------- BVH2 -> BVH4 collapse
namespace {
struct Child4 { Aabb box; int32_t ref; };

void ReadChildren2(const BvhNode64& n, std::vector<Child4>* out) {
    for (int slot = 0; slot < 2; slot++) {
        Child4 c;
        memcpy(&c.ref, &n.q[12 + slot], 4);
        if (c.ref == kNoChild) continue;
        c.box.lo[0] = n.q[slot * 4 + 0]; c.box.hi[0] = n.q[slot * 4 + 1]; c.box.lo[1] = n.q[slot * 4 + 2]; c.box.hi[1] = n.q[slot * 4 + 3];
        c.box.lo[2] = n.q[8 + slot * 2 + 0]; c.box.hi[2] = n.q[8 + slot * 2 + 1];
        out->push_back(c);
    }
}
}  // namespace

void CollapseToBvh4(const BuiltBvh& b2, BuiltBvh4* out) {
    out->nodes.clear();
    out->max_stack = 0;
    struct Item { int node2; int node4; int stack_before; };
    std::vector<Item> work;
    out->nodes.emplace_back();
    work.push_back({0, 0, 0});
    while (!work.empty()) {
        Item it = work.back(); work.pop_back();
        std::vector<Child4> ch;
        ReadChildren2(b2.nodes[it.node2], &ch);
        while (ch.size() < 4) {
            int best = -1; float best_area = -1.f;
            for (size_t i = 0; i < ch.size(); i++)
                if (ch[i].ref >= 0 && ch[i].box.HalfArea() > best_area) { best_area = ch[i].box.HalfArea(); best = (int)i; }
            if (best < 0) break;
            std::vector<Child4> sub;
            ReadChildren2(b2.nodes[ch[best].ref], &sub);
            if (ch.size() - 1 + sub.size() > 4) break;
            ch.erase(ch.begin() + best);
            ch.insert(ch.end(), sub.begin(), sub.end());
        }
        BvhNode128 nd;
        memset(&nd, 0, sizeof nd);
        int stack_here = it.stack_before + (int)ch.size() - 1;
        out->max_stack = std::max(out->max_stack, stack_here);
        for (int k = 0; k < 4; k++) {
            int32_t ref = kNoChild;
            if (k < (int)ch.size()) {
                const Child4& c = ch[k];
                nd.q[0 + k] = c.box.lo[0]; nd.q[4 + k] = c.box.hi[0]; nd.q[8 + k] = c.box.lo[1]; nd.q[12 + k] = c.box.hi[1];
                nd.q[16 + k] = c.box.lo[2]; nd.q[20 + k] = c.box.hi[2];
                ref = c.ref;
                if (ref >= 0) {
                    int idx4 = (int)out->nodes.size();
                    out->nodes.emplace_back();
                    work.push_back({c.ref, idx4, stack_here});
                    ref = idx4;
                }
            }
            memcpy(&nd.q[24 + k], &ref, 4);
        }
        out->nodes[it.node4] = nd;
    }
}


void QuantiseBvh4(const BuiltBvh4& b4, std::vector<BvhNode64Q>* out) {
    out->assign(b4.nodes.size(), BvhNode64Q());
    for (size_t i = 0; i < b4.nodes.size(); i++) QuantiseNode4(b4.nodes[i].q, (*out)[i].w, kNoChild);
}

}  // namespace rt795
