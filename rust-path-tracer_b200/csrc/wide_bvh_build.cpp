// wide_bvh_build.cpp — collapse the reference's binary BVH (src/bvh.rs output, BVHNode[] +
// permuted index buffer) into the backend's 8-wide compressed layout described in wide_bvh.h.
//
// The binary tree is treated as given: the wide tree holds exactly the same triangles under
// conservative (never smaller) boxes, so a nearest-hit query returns the same triangle as the
// reference traversal except for exact-t ties, which depend on visiting order.
#include "wide_bvh.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>

namespace rpt {
namespace {

struct Box3 {
    float lo[3] = {INFINITY, INFINITY, INFINITY};
    float hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    void grow(const float* p) {
        for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(lo[k], p[k]); hi[k] = std::fmax(hi[k], p[k]); }
    }
    void grow(const Box3& b) {
        for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(lo[k], b.lo[k]); hi[k] = std::fmax(hi[k], b.hi[k]); }
    }
    double half_area() const {
        const double e[3] = {(double)hi[0] - lo[0], (double)hi[1] - lo[1], (double)hi[2] - lo[2]};
        return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
    }
};

// A subtree of the binary BVH, or a run of triangles of an over-full binary leaf.
struct Item {
    Box3 box;
    bool is_range;           // true: triangles [first, first+count); false: binary node `first`
    uint32_t first, count;
};

class Collapser {
  public:
    Collapser(const RptBVHNode* nodes, uint32_t nnodes, const uint32_t* tris, uint32_t ntris, const RptPerVertexData* verts,
              uint32_t nverts, WideBvh& out)
        : nodes_(nodes), nnodes_(nnodes), tris_(tris), ntris_(ntris), verts_(verts), nverts_(nverts), out_(out) {}

    // Triangle range covered by every binary subtree (the reference builder partitions the index
    // buffer in place, so a subtree owns one contiguous run); count 0 marks "not contiguous".
    bool subtree_ranges() {
        sub_first_.assign(nnodes_, 0u);
        sub_count_.assign(nnodes_, 0u);
        std::vector<uint32_t>& order = order_;
        order.clear();
        std::vector<uint32_t> todo{0u};
        order.reserve(nnodes_);
        std::vector<uint8_t> seen(nnodes_, 0);
        while (!todo.empty()) {
            const uint32_t ni = todo.back();
            todo.pop_back();
            if (ni >= nnodes_ || seen[ni]) return false;
            seen[ni] = 1;
            order.push_back(ni);
            if (nodes_[ni].triangle_count == 0) {
                if ((uint64_t)nodes_[ni].left_or_first + 1 >= nnodes_) return false;
                todo.push_back(nodes_[ni].left_or_first);
                todo.push_back(nodes_[ni].left_or_first + 1);
            }
        }
        for (size_t k = order.size(); k-- > 0;) {  // children appear after their parent in `order`
            const uint32_t ni = order[k];
            const RptBVHNode& n = nodes_[ni];
            if (n.triangle_count > 0) {
                sub_first_[ni] = n.left_or_first;
                sub_count_[ni] = n.triangle_count;
            } else {
                const uint32_t l = n.left_or_first, r = l + 1;
                if (sub_count_[l] && sub_count_[r] && sub_first_[l] + sub_count_[l] == sub_first_[r]) {
                    sub_first_[ni] = sub_first_[l];
                    sub_count_[ni] = sub_count_[l] + sub_count_[r];
                }
            }
        }
        return true;
    }

    // Which binary nodes become wide nodes?  Every wide node a ray enters costs one 80-byte fetch and eight
    // box tests whether its slots are full or not, so the collapse minimises the summed surface area of the
    // wide nodes (the expected number of node visits of a random ray) instead of opening the largest child
    // greedily — the dynamic programme of Ylitie, Karras & Laine 2017 (section 3.1) with all leaf terms
    // constant (the reference's leaves are kept as they are):
    //   forest(n, j) = cheapest way to present subtree n as at most j children of some wide node
    //   forest(n, 1) = inner(n) = area(n) + min_k forest(left, k) + forest(right, 8 - k)   (n becomes a wide node)
    //   forest(n, j) = min(forest(n, j-1), min_k forest(left, k) + forest(right, j - k))
    // and a leaf costs nothing however many slots it is offered.
    void plan_collapse(const std::vector<uint32_t>& order) {
        plan_cost_.assign((size_t)nnodes_ * 8, 0.0f);   // [n*8 + j], j = 1..7
        plan_split_.assign((size_t)nnodes_ * 8, 0);     // left share k for forest(n, j); 0 = "same as j-1" / "be a wide node"
        plan_root_split_.assign(nnodes_, 4);
        for (size_t idx = order.size(); idx-- > 0;) {
            const uint32_t n = order[idx];
            if (is_leaf_item(n)) continue;  // cost 0 for every j
            const uint32_t l = nodes_[n].left_or_first, r = l + 1;
            const float* cl = &plan_cost_[(size_t)l * 8];
            const float* cr = &plan_cost_[(size_t)r * 8];
            float* cn = &plan_cost_[(size_t)n * 8];
            uint8_t* sn = &plan_split_[(size_t)n * 8];
            Box3 b;
            std::memcpy(b.lo, nodes_[n].aabb_min, 12);
            std::memcpy(b.hi, nodes_[n].aabb_max, 12);
            float best = INFINITY;
            int best_k = 4;
            for (int k = 1; k <= 7; ++k) {
                const float c = cl[k] + cr[8 - k];
                if (c < best) { best = c; best_k = k; }
            }
            plan_root_split_[n] = (uint8_t)best_k;
            cn[1] = (float)b.half_area() + best;
            sn[1] = 0;
            for (int j = 2; j <= 7; ++j) {
                cn[j] = cn[j - 1];
                sn[j] = 0;
                for (int k = 1; k < j; ++k) {
                    const float c = cl[k] + cr[j - k];
                    if (c < cn[j]) { cn[j] = c; sn[j] = (uint8_t)k; }
                }
            }
        }
    }
    bool is_leaf_item(uint32_t n) const { return nodes_[n].triangle_count > 0 || (sub_count_[n] != 0 && sub_count_[n] <= leaf_merge_); }
    // Children that present binary subtree n as at most j slots, following the plan.
    bool collect(uint32_t n, int j, std::vector<Item>& out) const {
        if (n >= nnodes_ || j < 1) return false;
        while (!is_leaf_item(n) && j > 1 && plan_split_[(size_t)n * 8 + j] == 0) --j;  // "same as j-1"
        Item it;
        if (is_leaf_item(n) || j == 1) {
            if (!node_item(n, it)) return false;
            out.push_back(it);
            return true;
        }
        const int k = plan_split_[(size_t)n * 8 + j];
        const uint32_t l = nodes_[n].left_or_first;
        return collect(l, k, out) && collect(l + 1, j - k, out);
    }

    bool run(const char** error) {
        if (!subtree_ranges()) { *error = "malformed BVH: child index out of range or node referenced twice"; return false; }
        if (use_dp_) plan_collapse(order_);
        out_.nodes.clear();
        out_.tri_pos.clear();
        out_.orig_index.clear();
        out_.wide_index.assign(ntris_, 0xFFFFFFFFu);
        out_.tri_pos.reserve((size_t)ntris_ * 12);
        out_.orig_index.reserve(ntris_);

        Item root;
        if (!node_item(0, root)) { *error = "malformed BVH: bad root node"; return false; }
        struct Work { uint32_t wide; Item item; uint32_t depth; };
        std::deque<Work> queue;
        out_.nodes.emplace_back();
        queue.push_back({0u, root, 0u});
        size_t visited_budget = (size_t)nnodes_ * 2 + (size_t)ntris_ * 2 + 16;
        while (!queue.empty()) {
            Work w = queue.front();
            queue.pop_front();
            out_.max_depth = std::max(out_.max_depth, w.depth);

            // ---- gather up to 8 children
            std::vector<Item> kids;
            if (use_dp_ && !w.item.is_range && !plan_cost_.empty()) {
                // the cut of this binary subtree chosen by the surface-area dynamic programme (plan_collapse)
                const uint32_t l = nodes_[w.item.first].left_or_first;
                const int kl = plan_root_split_[w.item.first];
                if (!collect(l, kl, kids) || !collect(l + 1, 8 - kl, kids)) { *error = "malformed BVH: child index or triangle range out of bounds"; return false; }
            } else if (splittable(w.item)) {
                Item a, b;
                if (!split(w.item, a, b)) { *error = "malformed BVH: child index or triangle range out of bounds"; return false; }
                kids = {a, b};
            } else {
                kids = {w.item};
            }
            // top up greedily (opens over-full leaf runs; the whole job when the DP is off): largest box first
            while (kids.size() < 8) {
                int best = -1;
                double best_area = -1.0;
                for (size_t i = 0; i < kids.size(); ++i) {
                    const bool open = use_dp_ && !plan_cost_.empty() ? (kids[i].is_range && kids[i].count > 3) : splittable(kids[i]);
                    if (open && kids[i].box.half_area() > best_area) { best = (int)i; best_area = kids[i].box.half_area(); }
                }
                if (best < 0) break;
                Item a, b;
                if (!split(kids[best], a, b)) { *error = "malformed BVH: child index or triangle range out of bounds"; return false; }
                kids[best] = a;
                kids.push_back(b);
                if (visited_budget-- == 0) { *error = "malformed BVH: cycle"; return false; }
            }

            // ---- octant-ordered slots: slot s prefers the child lying farthest against direction ds(s)
            Box3 nb;
            for (const Item& k : kids) nb.grow(k.box);
            int slot_of[8], kid_in_slot[8];
            assign_slots(kids, nb, slot_of, kid_in_slot);

            // ---- allocate children: inner nodes contiguous in slot order, triangles in one block
            const uint32_t child_base = (uint32_t)out_.nodes.size();
            const uint32_t tri_base = (uint32_t)out_.orig_index.size();
            uint32_t imask = 0, valid = 0;
            for (int s = 0; s < 8; ++s) {
                const int k = kid_in_slot[s];
                if (k < 0) continue;
                const Item& it = kids[k];
                if (splittable(it)) {
                    imask |= 1u << s;
                    out_.nodes.emplace_back();
                    queue.push_back({(uint32_t)out_.nodes.size() - 1u, it, w.depth + 1u});
                    out_.inner_children++;
                } else {
                    valid |= (it.count == 1 ? 1u : (it.count == 2 ? 3u : 7u)) << (3 * s);  // slot s owns triangle bits 3s..3s+2
                    for (uint32_t t = 0; t < it.count; ++t) emit_triangle(it.first + t);
                    out_.leaf_children++;
                }
            }
            encode(out_.nodes[w.wide], nb, kids, kid_in_slot, child_base, tri_base, imask, valid);
        }
        for (uint32_t t = 0; t < ntris_; ++t)
            if (out_.wide_index[t] == 0xFFFFFFFFu) { *error = "malformed BVH: a triangle is not referenced by any leaf"; return false; }
        return true;
    }

  private:
    bool node_item(uint32_t ni, Item& it) const {
        if (ni >= nnodes_) return false;
        const RptBVHNode& n = nodes_[ni];
        std::memcpy(it.box.lo, n.aabb_min, 12);
        std::memcpy(it.box.hi, n.aabb_max, 12);
        if (n.triangle_count > 0) {
            if ((uint64_t)n.left_or_first + n.triangle_count > ntris_) return false;
            it.is_range = true;
            it.first = n.left_or_first;
            it.count = n.triangle_count;
        } else if (sub_count_[ni] != 0 && sub_count_[ni] <= leaf_merge_) {
            // a whole binary subtree of at most three triangles becomes ONE leaf slot: fewer wide nodes
            // to visit at the price of testing its triangles together
            if ((uint64_t)sub_first_[ni] + sub_count_[ni] > ntris_) return false;
            it.is_range = true;
            it.first = sub_first_[ni];
            it.count = sub_count_[ni];
        } else {
            if ((uint64_t)n.left_or_first + 1 >= nnodes_) return false;
            it.is_range = false;
            it.first = ni;
            it.count = 0;
        }
        return true;
    }
    static bool splittable(const Item& it) { return !it.is_range || it.count > 3; }
    Box3 range_box(uint32_t first, uint32_t count) const {
        Box3 b;
        for (uint32_t t = first; t < first + count; ++t)
            for (int k = 0; k < 3; ++k) b.grow(verts_[tris_[4 * (size_t)t + k]].vertex);
        return b;
    }
    bool split(const Item& it, Item& a, Item& b) const {
        if (!it.is_range) {
            const uint32_t l = nodes_[it.first].left_or_first;
            return node_item(l, a) && node_item(l + 1, b);
        }
        const uint32_t half = it.count / 2;
        a = {range_box(it.first, half), true, it.first, half};
        b = {range_box(it.first + half, it.count - half), true, it.first + half, it.count - half};
        return true;
    }

    static void assign_slots(const std::vector<Item>& kids, const Box3& nb, int* slot_of, int* kid_in_slot) {
        const int n = (int)kids.size();
        double cost[8][8];
        const double nc[3] = {0.5 * ((double)nb.lo[0] + nb.hi[0]), 0.5 * ((double)nb.lo[1] + nb.hi[1]), 0.5 * ((double)nb.lo[2] + nb.hi[2])};
        for (int s = 0; s < 8; ++s) {
            const double ds[3] = {(s & 4) ? -1.0 : 1.0, (s & 2) ? -1.0 : 1.0, (s & 1) ? -1.0 : 1.0};
            for (int i = 0; i < n; ++i) {
                double c = 0.0;
                for (int k = 0; k < 3; ++k) c += (0.5 * ((double)kids[i].box.lo[k] + kids[i].box.hi[k]) - nc[k]) * ds[k];
                cost[s][i] = c;
            }
        }
        for (int s = 0; s < 8; ++s) kid_in_slot[s] = -1;
        for (int i = 0; i < 8; ++i) slot_of[i] = -1;
        for (int round = 0; round < n; ++round) {  // greedy: cheapest remaining (slot, child) pair
            double best = DBL_MAX;
            int bs = -1, bi = -1;
            for (int s = 0; s < 8; ++s) {
                if (kid_in_slot[s] >= 0) continue;
                for (int i = 0; i < n; ++i)
                    if (slot_of[i] < 0 && cost[s][i] < best) { best = cost[s][i]; bs = s; bi = i; }
            }
            kid_in_slot[bs] = bi;
            slot_of[bi] = bs;
        }
    }

    void emit_triangle(uint32_t t) {
        const uint32_t* tri = tris_ + 4 * (size_t)t;
        const float* a = verts_[tri[0]].vertex;
        const float* b = verts_[tri[1]].vertex;
        const float* c = verts_[tri[2]].vertex;
        float rec[12];
        uint32_t bits;
        rec[0] = a[0]; rec[1] = a[1]; rec[2] = a[2];
        bits = t; std::memcpy(&rec[3], &bits, 4);
        rec[4] = b[0] - a[0]; rec[5] = b[1] - a[1]; rec[6] = b[2] - a[2];
        bits = tri[3]; std::memcpy(&rec[7], &bits, 4);
        rec[8] = c[0] - a[0]; rec[9] = c[1] - a[1]; rec[10] = c[2] - a[2];
        rec[11] = 0.0f;
        out_.wide_index[t] = (uint32_t)out_.orig_index.size();
        out_.orig_index.push_back(t);
        out_.tri_pos.insert(out_.tri_pos.end(), rec, rec + 12);
    }

    static void encode(WideNode& node, const Box3& nb, const std::vector<Item>& kids, const int* kid_in_slot, uint32_t child_base,
                       uint32_t tri_base, uint32_t imask, uint32_t valid) {
        uint8_t e[3];
        double cell[3];
        for (int k = 0; k < 3; ++k) {
            const double extent = (double)nb.hi[k] - (double)nb.lo[k];
            int ex = -126;
            if (extent > 0.0) {
                ex = (int)std::ceil(std::log2(extent / 255.0));
                while (std::ceil(extent / std::ldexp(1.0, ex)) > 255.0) ++ex;  // log2 rounding guard
            }
            ex = std::min(std::max(ex, -126), 127);
            e[k] = (uint8_t)(ex + 127);
            cell[k] = std::ldexp(1.0, ex);
        }
        uint8_t q[6][8];
        std::memset(q, 0, sizeof(q));
        for (int s = 0; s < 8; ++s) {
            const int ki = kid_in_slot[s];
            if (ki < 0) continue;
            const Box3& b = kids[ki].box;
            for (int k = 0; k < 3; ++k) {
                const double lo = std::floor(((double)b.lo[k] - (double)nb.lo[k]) / cell[k]);
                const double hi = std::ceil(((double)b.hi[k] - (double)nb.lo[k]) / cell[k]);
                q[k][s] = (uint8_t)std::min(std::max(lo, 0.0), 255.0);
                q[3 + k][s] = (uint8_t)std::min(std::max(hi, 0.0), 255.0);
            }
        }
        auto pack4 = [](const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
        uint32_t* w = node.w;
        std::memcpy(&w[0], &nb.lo[0], 4);
        std::memcpy(&w[1], &nb.lo[1], 4);
        std::memcpy(&w[2], &nb.lo[2], 4);
        w[3] = (uint32_t)e[0] << 23;  // the cell sizes as float bits: x whole, y and z as their upper halves
        w[4] = child_base; w[5] = tri_base; w[6] = valid | (imask << 24); w[7] = ((uint32_t)e[1] << 7) | ((uint32_t)e[2] << 23);
        w[8] = pack4(q[0]); w[9] = pack4(q[0] + 4); w[10] = pack4(q[1]); w[11] = pack4(q[1] + 4);
        w[12] = pack4(q[2]); w[13] = pack4(q[2] + 4); w[14] = pack4(q[3]); w[15] = pack4(q[3] + 4);
        w[16] = pack4(q[4]); w[17] = pack4(q[4] + 4); w[18] = pack4(q[5]); w[19] = pack4(q[5] + 4);
    }

    const RptBVHNode* nodes_;
    uint32_t nnodes_;
    const uint32_t* tris_;
    uint32_t ntris_;
    const RptPerVertexData* verts_;
    uint32_t nverts_;
    WideBvh& out_;
    std::vector<uint32_t> sub_first_, sub_count_, order_;
    std::vector<float> plan_cost_;
    std::vector<uint8_t> plan_split_, plan_root_split_;

  public:
    uint32_t leaf_merge_ = 1;  // largest binary subtree (in triangles) folded into one leaf slot; 1 = off
    bool use_dp_ = true;       // surface-area dynamic programme (plan_collapse) vs. greedy largest-box-first
};

}  // namespace

bool build_wide_bvh(const RptBVHNode* nodes, uint32_t nnodes, const uint32_t* triangles, uint32_t ntriangles,
                    const RptPerVertexData* vertices, uint32_t nvertices, WideBvh& out, const char** error) {
    static const char* none = "";
    const char* dummy = none;
    if (!error) error = &dummy;
    *error = none;
    if (!nodes || !triangles || !vertices || nnodes == 0 || ntriangles == 0) { *error = "empty scene"; return false; }
    for (size_t t = 0; t < (size_t)ntriangles; ++t)
        for (int k = 0; k < 3; ++k)
            if (triangles[4 * t + k] >= nvertices) { *error = "triangle references a vertex out of range"; return false; }
    out = WideBvh{};
    Collapser c(nodes, nnodes, triangles, ntriangles, vertices, nvertices, out);
    if (const char* v = std::getenv("RPT_COLLAPSE")) c.use_dp_ = std::strcmp(v, "greedy") != 0;
    if (const char* v = std::getenv("RPT_LEAF_MERGE")) c.leaf_merge_ = (uint32_t)std::min(3, std::max(1, std::atoi(v)));
    return c.run(error);
}

}  // namespace rpt
