// wide_bvh_build.cpp — collapse the reference's binary BVH (src/bvh.rs output, BVHNode[] +
// permuted index buffer) into the backend's 8-wide compressed layout described in wide_bvh.h.
//
// The binary tree is treated as given: the wide tree holds exactly the same triangles under
// conservative (never smaller) boxes, so a nearest-hit query returns the same triangle as the
// reference traversal except for exact-t ties, which depend on visiting order.
#include "wide_bvh.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace rpt {
namespace {

// fminf / fmaxf written out (the compiler calls libm for them otherwise)
inline float min_f(float x, float y) { return x < y ? x : (x > y ? y : (y != y ? x : y)); }
inline float max_f(float x, float y) { return x > y ? x : (x < y ? y : (y != y ? x : y)); }

// (No default member initialisers: arrays of boxes are allocated for whole tree levels and filled by the build threads.)
struct Box3 {
    float lo[3], hi[3];
    static Box3 empty() { return {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; }
    void grow(const float* p) {
        for (int k = 0; k < 3; ++k) { lo[k] = min_f(lo[k], p[k]); hi[k] = max_f(hi[k], p[k]); }
    }
    void grow(const Box3& b) {
        for (int k = 0; k < 3; ++k) { lo[k] = min_f(lo[k], b.lo[k]); hi[k] = max_f(hi[k], b.hi[k]); }
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

// The (at most eight) children of one wide node while it is being assembled.
struct Kids {
    Item v[8];
    int n;
    bool push(const Item& it) { if (n >= 8) return false; v[n++] = it; return true; }
};

class Collapser {
  public:
    Collapser(const RptBVHNode* nodes, uint32_t nnodes, const uint32_t* tris, uint32_t ntris, const RptPerVertexData* verts,
              uint32_t nverts, WideBvh& out)
        : nodes_(nodes), nnodes_(nnodes), tris_(tris), ntris_(ntris), verts_(verts), nverts_(nverts), out_(out) {}

    // Bottom-up pass over the binary tree, children before parents:
    //  * the triangle range covered by every subtree (the reference builder partitions the index buffer in place, so a
    //    subtree owns one contiguous run); count 0 marks "not contiguous";
    //  * which binary nodes become wide nodes.  Every wide node a ray enters costs one 80-byte fetch and eight box
    //    tests whether its slots are full or not, so the collapse minimises the summed surface area of the wide nodes
    //    (the expected number of node visits of a random ray) instead of opening the largest child greedily — the
    //    dynamic programme of Ylitie, Karras & Laine 2017 (section 3.1) with all leaf terms constant (the reference's
    //    leaves are kept as they are):
    //      forest(n, j) = cheapest way to present subtree n as at most j children of some wide node
    //      forest(n, 1) = inner(n) = area(n) + min_k forest(left, k) + forest(right, 8 - k)   (n becomes a wide node)
    //      forest(n, j) = min(forest(n, j-1), min_k forest(left, k) + forest(right, j - k))
    //    and a leaf costs nothing however many slots it is offered.
    // Disjoint subtrees are independent: the tree is cut a few levels below the root, the host threads take the
    // subtrees under the cut (each in depth-first order, which keeps a subtree's tables in cache), then the few nodes
    // above the cut are finished.  Returns false for a malformed tree (child out of range, node referenced twice).
    bool plan_bottom_up() {
        sub_first_.assign(nnodes_, 0u);
        sub_count_.assign(nnodes_, 0u);
        if (use_dp_) {  // written for inner nodes only, by the thread that owns the subtree (no serial zero fill of 40 bytes per node)
            plan_cost_.reset(new float[(size_t)nnodes_ * 8]);         // [n*8 + j], j = 1..7
            plan_split_.reset(new uint8_t[(size_t)nnodes_ * 8]);      // left share k for forest(n, j); 0 = "same as j-1" / "be a wide node"
            plan_root_split_.reset(new uint8_t[nnodes_]);
        }
        std::vector<std::atomic<uint8_t>> seen(nnodes_);
        for (auto& f : seen) f.store(0, std::memory_order_relaxed);
        std::atomic<bool> malformed{false};
        // children of n, claimed for the caller; false if there are none (leaf) or the tree is malformed
        auto claim_children = [&](uint32_t n, uint32_t& l) {
            if (nodes_[n].triangle_count != 0) return false;
            l = nodes_[n].left_or_first;
            if ((uint64_t)l + 1 >= nnodes_ || seen[l].exchange(1, std::memory_order_relaxed) || seen[l + 1].exchange(1, std::memory_order_relaxed)) {
                malformed.store(true, std::memory_order_relaxed);
                return false;
            }
            return true;
        };

        // ---- the cut: breadth first from the root until there are enough subtrees to share out
        seen[0].store(1, std::memory_order_relaxed);
        std::vector<uint32_t> above, cut{0u};
        const size_t wanted = threads_ > 1 && nnodes_ >= 8192 ? (size_t)threads_ * 8 : 1;
        for (int level = 0; level < 24 && cut.size() < wanted; ++level) {
            std::vector<uint32_t> below;
            for (uint32_t n : cut) {
                uint32_t l;
                if (claim_children(n, l)) { above.push_back(n); below.push_back(l); below.push_back(l + 1); }
                else if (nodes_[n].triangle_count != 0) finish_node(n);
            }
            cut.swap(below);
        }
        if (malformed.load()) return false;

        // ---- subtrees under the cut, handed out one at a time
        std::atomic<size_t> next{0};
        auto worker = [&] {
            std::vector<uint32_t> order, todo;
            for (size_t i; (i = next.fetch_add(1)) < cut.size() && !malformed.load(std::memory_order_relaxed);) {
                order.clear();
                todo.assign(1, cut[i]);
                while (!todo.empty()) {
                    const uint32_t n = todo.back();
                    todo.pop_back();
                    order.push_back(n);
                    uint32_t l;
                    if (claim_children(n, l)) { todo.push_back(l); todo.push_back(l + 1); }
                }
                for (size_t k = order.size(); k-- > 0;) finish_node(order[k]);  // children appear after their parent in `order`
            }
        };
        const size_t workers = std::min<size_t>(threads_, cut.size());
        if (workers <= 1) {
            worker();
        } else {
            std::vector<std::thread> pool;
            for (size_t w = 0; w < workers; ++w) pool.emplace_back(worker);
            for (std::thread& t : pool) t.join();
        }
        if (malformed.load()) return false;
        for (size_t k = above.size(); k-- > 0;) finish_node(above[k]);  // breadth-first order: parents first
        return true;
    }

    // One node of the bottom-up pass; both children are finished.
    void finish_node(uint32_t n) {
        const RptBVHNode& node = nodes_[n];
        if (node.triangle_count > 0) {
            sub_first_[n] = node.left_or_first;
            sub_count_[n] = node.triangle_count;
            return;
        }
        const uint32_t l = node.left_or_first, r = l + 1;
        if (sub_count_[l] && sub_count_[r] && sub_first_[l] + sub_count_[l] == sub_first_[r]) {
            sub_first_[n] = sub_first_[l];
            sub_count_[n] = sub_count_[l] + sub_count_[r];
        }
        if (!use_dp_ || is_leaf_item(n)) return;  // a (merged) leaf costs 0 for every j
        static const float leaf_cost[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        const float* cl = is_leaf_item(l) ? leaf_cost : &plan_cost_[(size_t)l * 8];
        const float* cr = is_leaf_item(r) ? leaf_cost : &plan_cost_[(size_t)r * 8];
        float* cn = &plan_cost_[(size_t)n * 8];
        uint8_t* sn = &plan_split_[(size_t)n * 8];
        Box3 b;
        std::memcpy(b.lo, node.aabb_min, 12);
        std::memcpy(b.hi, node.aabb_max, 12);
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
    bool is_leaf_item(uint32_t n) const { return nodes_[n].triangle_count > 0 || (sub_count_[n] != 0 && sub_count_[n] <= leaf_merge_); }
    // Children that present binary subtree n as at most j slots, following the plan.
    bool collect(uint32_t n, int j, Kids& out) const {
        if (n >= nnodes_ || j < 1) return false;
        while (!is_leaf_item(n) && j > 1 && plan_split_[(size_t)n * 8 + j] == 0) --j;  // "same as j-1"
        Item it;
        if (is_leaf_item(n) || j == 1) return node_item(n, it) && out.push(it);
        const int k = plan_split_[(size_t)n * 8 + j];
        const uint32_t l = nodes_[n].left_or_first;
        return collect(l, k, out) && collect(l + 1, j - k, out);
    }

    // The wide tree is laid out breadth first: the nodes of one level are independent of each other, so every level is
    // built by all host threads — (1) choose each node's children and slots, (2) a prefix sum over the level hands
    // out the child-node and triangle ranges in level order, (3) encode the nodes and write the triangle records into
    // their ranges.  The result is the same array, bit for bit, for any thread count (tests/test_wide_bvh_cpu.py).
    bool run(const char** error) {
        if (!plan_bottom_up()) { *error = "malformed BVH: child index out of range or node referenced twice"; return false; }
        out_.nodes.assign(1, WideNode{});
        out_.tri_pos.resize((size_t)ntris_ * 12);  // (left uninitialised: every record is written below, which the final check verifies)
        out_.orig_index.resize(ntris_);
        out_.wide_index.assign(ntris_, 0xFFFFFFFFu);

        Item root;
        if (!node_item(0, root)) { *error = "malformed BVH: bad root node"; return false; }
        struct Pending { Item item; uint32_t wide; };
        struct Built {
            Kids kids;
            Box3 box;
            int kid_in_slot[8];
            uint32_t inner, triangles, child_base, tri_base;
            bool ok;
        };
        std::vector<Pending> level{{root, 0u}}, next;
        std::unique_ptr<Built[]> built;  // (plain data, allocated uninitialised: 400 bytes per node of the widest level)
        size_t built_capacity = 0;
        uint32_t emitted = 0;
        for (uint32_t depth = 0; !level.empty(); ++depth) {
            out_.max_depth = depth;
            if (level.size() > built_capacity) {
                built_capacity = level.size() + level.size() / 2;
                built.reset(new Built[built_capacity]);
            }
            parallel_for(level.size(), [&](size_t i) {
                Built& b = built[i];
                b.ok = choose_children(level[i].item, b.kids);
                b.inner = b.triangles = 0;
                for (int k = 0; b.ok && k < b.kids.n; ++k) {
                    if (splittable(b.kids.v[k])) b.inner++;
                    else b.triangles += b.kids.v[k].count;
                }
            });
            uint32_t node_cursor = (uint32_t)out_.nodes.size(), tri_cursor = emitted;
            for (size_t i = 0; i < level.size(); ++i) {
                Built& b = built[i];
                if (!b.ok) { *error = "malformed BVH: child index or triangle range out of bounds"; return false; }
                b.child_base = node_cursor;
                b.tri_base = tri_cursor;
                node_cursor += b.inner;
                if ((uint64_t)tri_cursor + b.triangles > ntris_) { *error = "malformed BVH: a triangle is referenced by two leaves"; return false; }
                tri_cursor += b.triangles;
                out_.inner_children += b.inner;
                out_.leaf_children += (uint32_t)b.kids.n - b.inner;
            }
            out_.nodes.resize(node_cursor);
            next.resize(node_cursor - built[0].child_base);
            const uint32_t level_child_base = built[0].child_base;
            parallel_for(level.size(), [&](size_t i) {
                Built& b = built[i];
                // ---- octant-ordered slots: slot s prefers the child lying farthest against direction ds(s)
                b.box = Box3::empty();
                for (int k = 0; k < b.kids.n; ++k) b.box.grow(b.kids.v[k].box);
                int slot_of[8];
                assign_slots(b.kids, b.box, slot_of, b.kid_in_slot);
                // ---- children: inner nodes contiguous in slot order, triangles in one block
                uint32_t imask = 0, valid = 0, child = b.child_base, tri = b.tri_base;
                for (int s = 0; s < 8; ++s) {
                    const int k = b.kid_in_slot[s];
                    if (k < 0) continue;
                    const Item& it = b.kids.v[k];
                    if (splittable(it)) {
                        imask |= 1u << s;
                        next[child - level_child_base] = {it, child};
                        ++child;
                    } else {
                        valid |= (it.count == 1 ? 1u : (it.count == 2 ? 3u : 7u)) << (3 * s);  // slot s owns triangle bits 3s..3s+2
                        for (uint32_t t = 0; t < it.count; ++t) emit_triangle(it.first + t, tri++);
                    }
                }
                encode(out_.nodes[level[i].wide], b.box, b.kids, b.kid_in_slot, b.child_base, b.tri_base, imask, valid);
            });
            emitted = tri_cursor;
            level.swap(next);
            next.clear();
        }
        if (emitted != ntris_) { *error = "malformed BVH: a triangle is not referenced by any leaf"; return false; }
        for (uint32_t t = 0; t < ntris_; ++t)
            if (out_.wide_index[t] == 0xFFFFFFFFu) { *error = "malformed BVH: a triangle is not referenced by any leaf"; return false; }
        return true;
    }

  private:
    // Up to eight children for the wide node that stands for `item`.
    bool choose_children(const Item& item, Kids& kids) const {
        kids.n = 0;
        if (use_dp_ && !item.is_range) {
            // the cut of this binary subtree chosen by the surface-area dynamic programme (plan_collapse)
            const uint32_t l = nodes_[item.first].left_or_first;
            const int kl = plan_root_split_[item.first];
            if (!collect(l, kl, kids) || !collect(l + 1, 8 - kl, kids)) return false;
        } else if (splittable(item)) {
            Item a, b;
            if (!split(item, a, b)) return false;
            kids.push(a);
            kids.push(b);
        } else {
            kids.push(item);
        }
        // top up greedily (opens over-full leaf runs; the whole job when the DP is off): largest box first
        while (kids.n < 8) {
            int best = -1;
            double best_area = -1.0;
            for (int i = 0; i < kids.n; ++i) {
                const bool open = use_dp_ ? (kids.v[i].is_range && kids.v[i].count > 3) : splittable(kids.v[i]);
                if (open && kids.v[i].box.half_area() > best_area) { best = i; best_area = kids.v[i].box.half_area(); }
            }
            if (best < 0) break;
            Item a, b;
            if (!split(kids.v[best], a, b)) return false;
            kids.v[best] = a;
            kids.push(b);
        }
        return true;
    }

    template <class Fn>
    void parallel_for(size_t n, Fn&& fn) const {
        const size_t workers = std::min<size_t>(threads_, n / 512);  // small levels are not worth a thread launch
        if (workers <= 1) {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        std::vector<std::thread> pool;
        pool.reserve(workers);
        for (size_t w = 0; w < workers; ++w)
            pool.emplace_back([&, w] {
                for (size_t i = n * w / workers, end = n * (w + 1) / workers; i < end; ++i) fn(i);
            });
        for (std::thread& t : pool) t.join();
    }

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
        Box3 b = Box3::empty();
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

    static void assign_slots(const Kids& all, const Box3& nb, int* slot_of, int* kid_in_slot) {
        const int n = all.n;
        const Item* kids = all.v;
        double cost[8][8];
        const double nc[3] = {0.5 * ((double)nb.lo[0] + nb.hi[0]), 0.5 * ((double)nb.lo[1] + nb.hi[1]), 0.5 * ((double)nb.lo[2] + nb.hi[2])};
        for (int s = 0; s < 8; ++s) {
            const double ds[3] = {(s & 4) ? -1.0 : 1.0, (s & 2) ? -1.0 : 1.0, (s & 1) ? -1.0 : 1.0};
            for (int i = 0; i < n; ++i) {
                double c = 0.0;
                for (int k = 0; k < 3; ++k) c += (0.5 * ((double)kids[i].box.lo[k] + kids[i].box.hi[k]) - nc[k]) * ds[k];
                cost[s][i] = std::isfinite(c) ? c : 0.0;  // (an empty box — triangles with NaN vertices only — has no centre)
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

    void emit_triangle(uint32_t t, uint32_t position) {
        const uint32_t* tri = tris_ + 4 * (size_t)t;
        const float* a = verts_[tri[0]].vertex;
        const float* b = verts_[tri[1]].vertex;
        const float* c = verts_[tri[2]].vertex;
        float* rec = &out_.tri_pos[(size_t)position * 12];
        uint32_t bits;
        rec[0] = a[0]; rec[1] = a[1]; rec[2] = a[2];
        bits = t; std::memcpy(&rec[3], &bits, 4);
        rec[4] = b[0] - a[0]; rec[5] = b[1] - a[1]; rec[6] = b[2] - a[2];
        bits = tri[3]; std::memcpy(&rec[7], &bits, 4);
        rec[8] = c[0] - a[0]; rec[9] = c[1] - a[1]; rec[10] = c[2] - a[2];
        rec[11] = 0.0f;
        out_.wide_index[t] = position;
        out_.orig_index[position] = t;
    }

    static void encode(WideNode& node, const Box3& nb, const Kids& all, const int* kid_in_slot, uint32_t child_base,
                       uint32_t tri_base, uint32_t imask, uint32_t valid) {
        uint8_t e[3];
        double cell[3];
        for (int k = 0; k < 3; ++k) {
            const double extent = (double)nb.hi[k] - (double)nb.lo[k];
            int ex = -126;
            if (extent > 0.0 && extent < 1e37) {
                ex = (int)std::ceil(std::log2(extent / 255.0));
                while (std::ceil(extent / std::ldexp(1.0, ex)) > 255.0) ++ex;  // log2 rounding guard
            } else if (extent > 0.0) {
                ex = 127;  // (not reachable through build_wide_bvh, which rejects such coordinates)
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
            const Box3& b = all.v[ki].box;
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
    std::vector<uint32_t> sub_first_, sub_count_;
    std::unique_ptr<float[]> plan_cost_;
    std::unique_ptr<uint8_t[]> plan_split_, plan_root_split_;

  public:
    unsigned threads_ = 1;     // host threads building one level together
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
        for (int k = 0; k < 3; ++k) {
            if (triangles[4 * t + k] >= nvertices) { *error = "triangle references a vertex out of range"; return false; }
            // NaN coordinates are fine (min / max ignore them and such a triangle is never hit, as in the reference);
            // infinite or astronomically large ones would overflow the quantised-box arithmetic into NaN and hide
            // whole subtrees, so they are refused instead of traced wrongly.
            const float* p = vertices[triangles[4 * t + k]].vertex;
            if (std::fabs(p[0]) > 1e15f || std::fabs(p[1]) > 1e15f || std::fabs(p[2]) > 1e15f) { *error = "vertex coordinate infinite or beyond 1e15"; return false; }
        }
    out = WideBvh{};
    Collapser c(nodes, nnodes, triangles, ntriangles, vertices, nvertices, out);
    c.threads_ = std::max(1u, std::thread::hardware_concurrency());
    if (const char* v = std::getenv("RPT_BUILD_THREADS")) c.threads_ = (unsigned)std::max(1, std::atoi(v));
    if (const char* v = std::getenv("RPT_COLLAPSE")) c.use_dp_ = std::strcmp(v, "greedy") != 0;
    if (const char* v = std::getenv("RPT_LEAF_MERGE")) c.leaf_merge_ = (uint32_t)std::min(3, std::max(1, std::atoi(v)));
    return c.run(error);
}

}  // namespace rpt
