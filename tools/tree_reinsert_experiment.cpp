// Experiment: insertion-based optimisation (Bittner, Hapala, Havran 2013) of the reference's binary BVH before the
// wide collapse.  Leaves (and therefore the index-buffer permutation) are kept; only the inner topology changes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <queue>
#include <vector>
#include "../include/rpt_shared_structs.h"

namespace {
struct Box { float lo[3], hi[3]; };
inline Box merge(const Box& a, const Box& b) {
    Box r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); }
    return r;
}
inline double area(const Box& b) {
    const double x = (double)b.hi[0] - b.lo[0], y = (double)b.hi[1] - b.lo[1], z = (double)b.hi[2] - b.lo[2];
    return x * y + y * z + z * x;
}
struct Tree {
    std::vector<Box> box;
    std::vector<int> parent, left, right;  // leaves: left = right = -1
    std::vector<uint32_t> first, count;    // leaves
    int root = 0;
    bool leaf(int n) const { return left[n] < 0; }
    void refit_up(int n) {
        while (n >= 0) {
            const Box b = merge(box[left[n]], box[right[n]]);
            if (std::memcmp(&b, &box[n], sizeof b) == 0) break;
            box[n] = b;
            n = parent[n];
        }
    }
    double sah() const {
        double c = 0;
        for (size_t n = 0; n < box.size(); ++n) if (!leaf((int)n) && alive[n]) c += area(box[n]);
        return c / area(box[root]);
    }
    std::vector<char> alive;
};
}  // namespace

extern "C" int bvh_reinsert(const RptBVHNode* in, uint32_t nnodes, int passes, double fraction, RptBVHNode* out, double* stats) {
    Tree t;
    t.box.resize(nnodes); t.parent.assign(nnodes, -1); t.left.assign(nnodes, -1); t.right.assign(nnodes, -1);
    t.first.assign(nnodes, 0); t.count.assign(nnodes, 0); t.alive.assign(nnodes, 0);
    // only nodes reachable from the root are alive (the reference allocates 2n-1 and may leave a tail unused)
    std::vector<int> todo{0};
    while (!todo.empty()) {
        const int n = todo.back(); todo.pop_back();
        t.alive[n] = 1;
        std::memcpy(t.box[n].lo, in[n].aabb_min, 12); std::memcpy(t.box[n].hi, in[n].aabb_max, 12);
        if (in[n].triangle_count > 0) { t.first[n] = in[n].left_or_first; t.count[n] = in[n].triangle_count; continue; }
        const int l = (int)in[n].left_or_first;
        t.left[n] = l; t.right[n] = l + 1; t.parent[l] = n; t.parent[l + 1] = n;
        todo.push_back(l); todo.push_back(l + 1);
    }
    stats[0] = t.sah();
    std::vector<int> order;
    struct Cand { double bound; int node; double induced; bool operator<(const Cand& o) const { return bound > o.bound; } };
    for (int pass = 0; pass < passes; ++pass) {
        order.clear();
        for (uint32_t n = 0; n < nnodes; ++n)
            if (t.alive[n] && (int)n != t.root && t.parent[n] != t.root) order.push_back((int)n);
        // the nodes with the largest boxes first: that is where overlap costs most
        const size_t take = std::max<size_t>(1, (size_t)(order.size() * fraction));
        std::partial_sort(order.begin(), order.begin() + take, order.end(), [&](int a, int b) { return area(t.box[a]) > area(t.box[b]); });
        order.resize(take);
        size_t moved = 0;
        for (int n : order) {
            const int p = t.parent[n];
            if (p < 0 || p == t.root) continue;
            const int g = t.parent[p];
            const int s = t.left[p] == n ? t.right[p] : t.left[p];
            // ---- remove n and its parent p: the sibling takes p's place
            if (t.left[g] == p) t.left[g] = s; else t.right[g] = s;
            t.parent[s] = g;
            t.refit_up(g);
            // ---- best place for n: branch and bound over the insertion cost
            const Box nb = t.box[n];
            const double na = area(nb);
            double best = 1e300; int best_x = -1;
            std::priority_queue<Cand> q;
            q.push({0.0, t.root, 0.0});
            while (!q.empty()) {
                const Cand c = q.top(); q.pop();
                if (c.bound + na >= best) break;
                const int x = c.node;
                const double direct = area(merge(t.box[x], nb));
                const double total = c.induced + direct;
                if (total < best) { best = total; best_x = x; }
                if (!t.leaf(x)) {
                    const double induced = total - area(t.box[x]);
                    if (induced + na < best) { q.push({induced, t.left[x], induced}); q.push({induced, t.right[x], induced}); }
                }
            }
            // ---- insert: p becomes the parent of (best_x, n) where best_x was
            const int x = best_x;
            const int xp = t.parent[x];
            if (x != s || xp != g) ++moved;
            t.parent[p] = xp;
            if (xp < 0) t.root = p; else if (t.left[xp] == x) t.left[xp] = p; else t.right[xp] = p;
            t.left[p] = x; t.right[p] = n; t.parent[x] = p; t.parent[n] = p;
            t.box[p] = merge(t.box[x], nb);
            if (xp >= 0) t.refit_up(xp);
        }
        stats[1 + pass] = t.sah();
        std::fprintf(stderr, "pass %d: %zu of %zu nodes moved, SAH %.4f\n", pass, moved, order.size(), t.sah());
    }
    // ---- back to the reference layout: children adjacent, depth-first, root at 0
    std::vector<std::pair<int, uint32_t>> stack{{t.root, 0u}};
    uint32_t used = 1;
    while (!stack.empty()) {
        const auto [n, at] = stack.back(); stack.pop_back();
        std::memcpy(out[at].aabb_min, t.box[n].lo, 12); std::memcpy(out[at].aabb_max, t.box[n].hi, 12);
        if (t.leaf(n)) { out[at].triangle_count = t.count[n]; out[at].left_or_first = t.first[n]; continue; }
        out[at].triangle_count = 0; out[at].left_or_first = used;
        stack.push_back({t.right[n], used + 1}); stack.push_back({t.left[n], used});
        used += 2;
    }
    return (int)used;
}
