// world_build.cpp — host-side input producers of the tracing hot path (C ABI in include/rpt_host.h).
//
// These run once per scene load on the CPU in the reference too (src/asset.rs:195-203); they are
// restated here in C++ because the image has no Rust toolchain.  All arithmetic is IEEE fp32 in
// the reference's evaluation order (compile with -ffp-contract=off) so the node array, the
// permuted index buffer and the light table come out the way the Rust builder would emit them.
//
//   build_bvh         follows src/bvh.rs:58-324  (128-bin SAH over 3 axes, in-place partition)
//   light pick table  follows src/light_pick.rs:5-122 (power-weighted two-outcome bins)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/rpt_errors.h"
#include "../../include/rpt_host.h"

namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();

struct F3 {
    float x, y, z;
    float operator[](int a) const { return a == 0 ? x : (a == 1 ? y : z); }
};
inline F3 f3(const float* p) { return {p[0], p[1], p[2]}; }
// fminf / fmaxf (NaN-ignoring, like Rust's f32::min / max) written out: without -ffast-math the compiler calls into
// libm for every one of them, and the builder does a few hundred million.
inline float min_f(float x, float y) { return x < y ? x : (x > y ? y : (y != y ? x : y)); }
inline float max_f(float x, float y) { return x > y ? x : (x < y ? y : (y != y ? x : y)); }
// The same when the first operand is a running bound (+-inf or a value, never NaN): one compare and a select.
inline float acc_min(float acc, float y) { return y <= acc ? y : acc; }
inline float acc_max(float acc, float y) { return y >= acc ? y : acc; }
inline F3 vmin(F3 a, F3 b) { return {min_f(a.x, b.x), min_f(a.y, b.y), min_f(a.z, b.z)}; }
inline F3 vmax(F3 a, F3 b) { return {max_f(a.x, b.x), max_f(a.y, b.y), max_f(a.z, b.z)}; }
inline F3 sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float length(F3 a) { return std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }

// An axis-aligned box that starts empty (+inf, -inf), like `BVHNode::default()`.
struct Box {
    F3 lo{kInf, kInf, kInf};
    F3 hi{-kInf, -kInf, -kInf};
    void grow(F3 l, F3 h) {  // by a box whose bounds may hold NaN (ignored)
        lo = {acc_min(lo.x, l.x), acc_min(lo.y, l.y), acc_min(lo.z, l.z)};
        hi = {acc_max(hi.x, h.x), acc_max(hi.y, h.y), acc_max(hi.z, h.z)};
    }
    void grow(const Box& b) {
        if (b.lo.x == kInf) return;  // src/bvh.rs:21-27: empty bins are skipped
        grow(b.lo, b.hi);
    }
    // src/bvh.rs:29-32 — half the surface area; an empty box evaluates to +inf
    float half_area() const {
        F3 e = sub(hi, lo);
        return (e.x * e.y + e.y * e.z) + e.z * e.x;
    }
};

struct Tri { uint32_t i0, i1, i2, mat; };

// The reference builds the tree with one explicit stack (src/bvh.rs:257-323): a node that is split takes the
// next two free indices for its children, and its left subtree is finished before its right one is started, so
// the array is laid out as  [X, l, r, descendants(l)..., descendants(r)...]  recursively.  A subtree therefore
// depends on nothing but its own triangle range, and its nodes are contiguous once the size of everything before
// it is known: the big subtrees near the root are built by concurrent tasks into local arrays (indices relative to
// the subtree's root) and copied to their places afterwards — same nodes, same order, same permutation of the index
// buffer as the sequential build, in a fraction of the time (1 M triangles: 11 s for the literal loop -> 1.5 s on one core, 0.3 s on eight).
class SahBuilder {
  public:
    SahBuilder(const float* verts, Tri* tris, uint32_t ntris, uint32_t bins) : verts_(verts), tris_(tris), ntris_(ntris), bins_(bins), recs_(ntris) {
        unsigned threads = std::thread::hardware_concurrency();
        if (const char* v = std::getenv("RPT_BUILD_THREADS")) threads = (unsigned)std::max(1, std::atoi(v));
        run_chunks(ntris, ntris < kSharedNodeMinTriangles ? 1u : std::max(1u, threads), [&](unsigned, uint32_t begin, uint32_t end) {
        for (uint32_t t = begin; t < end; ++t) {  // src/bvh.rs:60-68
            F3 a = pos(tris[t].i0), b = pos(tris[t].i1), c = pos(tris[t].i2);
            recs_[t].centroid = {((a.x + b.x) + c.x) / 3.0f, ((a.y + b.y) + c.y) / 3.0f, ((a.z + b.z) + c.z) / 3.0f};
            // The reference grows boxes vertex by vertex; min and max are exact, so growing by the triangle's own box
            // gives the same bits and spares three random vertex fetches per triangle, axis and tree level.
            recs_[t].lo = vmin(vmin(a, b), c);
            recs_[t].hi = vmax(vmax(a, b), c);
        }
        });
    }

    // Nodes of the whole tree in the reference's order; nodes_out has room for 2 * ntris - 1.
    uint32_t run(RptBVHNode* nodes_out) {
        unsigned threads = std::thread::hardware_concurrency();
        if (const char* v = std::getenv("RPT_BUILD_THREADS")) threads = (unsigned)std::max(1, std::atoi(v));
        int fork_levels = 0;
        while ((1u << fork_levels) < std::max(1u, threads)) ++fork_levels;
        const std::unique_ptr<Piece> tree = build_subtree(0, ntris_, threads <= 1 ? 0 : fork_levels + 1, threads);  // one thread: the reference's loop as is
        emit(*tree, nodes_out, 0, 1);
        return (uint32_t)tree->size;
    }

  private:
    // Centroid and box of a triangle, permuted together with the index buffer.
    struct TriRec { F3 centroid, lo, hi; };

    struct Scratch {  // bins of all three axes ([axis * bins + bin]) and the sweep tables of one axis
        std::vector<Box> seg_box;
        std::vector<uint32_t> seg_count, left_count, right_count, touched, tri_bin;
        std::vector<float> left_area, right_area;
        bool dirty = false;
        explicit Scratch(uint32_t bins)
            : seg_box(3 * (size_t)bins), seg_count(3 * (size_t)bins, 0u), left_count(bins), right_count(bins), left_area(bins), right_area(bins) {
            touched.reserve(bins);
            tri_bin.resize(3 * (size_t)bins);
        }
    };
    static constexpr uint32_t kForkMinTriangles = 8192;  // below this a task costs more than it saves
    static constexpr uint32_t kSharedNodeMinTriangles = 1u << 16;  // a node this big is binned by several workers

    // fn(worker, begin, end) over `workers` consecutive runs of [0, count); worker 0 is the caller.
    template <class Fn>
    static void run_chunks(uint32_t count, unsigned workers, Fn&& fn) {
        if (workers <= 1) { fn(0u, 0u, count); return; }
        std::vector<std::thread> pool;
        pool.reserve(workers - 1);
        auto bound = [&](unsigned w) { return (uint32_t)((uint64_t)count * w / workers); };
        for (unsigned w = 1; w < workers; ++w) pool.emplace_back([&fn, &bound, w] { fn(w, bound(w), bound(w + 1)); });
        fn(0u, 0u, bound(1));
        for (std::thread& t : pool) t.join();
    }

    F3 pos(uint32_t v) const { return f3(verts_ + 4 * (size_t)v); }
    // `f as usize` (saturating: negative / NaN -> 0) followed by `.min(bins - 1)` (src/bvh.rs:203-204)
    static uint32_t bin_of(float f, uint32_t bins) {
        if (bins > (1u << 24)) {  // (float)bins would round: the literal form
            const size_t b = f > 0.0f ? (f >= 1.8446744e19f ? SIZE_MAX : (size_t)f) : 0;
            return (uint32_t)std::min<size_t>(b, bins - 1u);
        }
        return f > 0.0f ? (f >= (float)bins ? bins - 1u : (uint32_t)(int32_t)f) : 0u;
    }
    static float node_half_area(const RptBVHNode& n) {
        Box b;
        b.lo = f3(n.aabb_min);
        b.hi = f3(n.aabb_max);
        return b.half_area();
    }
    static RptBVHNode leaf(uint32_t first, uint32_t count) { return RptBVHNode{{kInf, kInf, kInf}, count, {-kInf, -kInf, -kInf}, first}; }

    // A subtree as the concurrent build leaves it: either one array built by the reference's loop (child indices
    // relative to the subtree's root), or a root whose two halves were built by separate tasks.
    struct Piece {
        std::vector<RptBVHNode> flat;
        RptBVHNode root;
        std::unique_ptr<Piece> left, right;
        size_t size = 0;
    };

    // `workers`: host threads this subtree may use for its own root (its two halves get half of them each).
    std::unique_ptr<Piece> build_subtree(uint32_t first, uint32_t count, int fork_levels, unsigned workers) {
        auto piece = std::make_unique<Piece>();
        if (fork_levels <= 0 || count < kForkMinTriangles) {
            piece->flat = build_sequential(first, count);
            piece->size = piece->flat.size();
            return piece;
        }
        Scratch scratch(bins_);
        piece->root = leaf(first, count);
        fit(piece->root, workers);
        uint32_t left_n = 0;
        if (!split(piece->root, scratch, left_n, workers)) {
            piece->flat = {piece->root};
            piece->size = 1;
            return piece;
        }
        const unsigned half = std::max(1u, workers / 2);
        auto left_task = std::async(std::launch::async, [=] { return build_subtree(first, left_n, fork_levels - 1, half); });
        piece->right = build_subtree(first + left_n, count - left_n, fork_levels - 1, half);
        piece->left = left_task.get();
        piece->size = 1 + piece->left->size + piece->right->size;
        return piece;
    }

    // Writes a piece where the reference's single loop would have put it: the subtree's root at `root_at`, then
    // [l, r, descendants(l)..., descendants(r)...] from `rest_at` on (every node is copied once, the big pieces by
    // tasks of their own).
    void emit(const Piece& piece, RptBVHNode* out, size_t root_at, size_t rest_at) const {
        if (!piece.left) {
            const std::vector<RptBVHNode>& flat = piece.flat;
            for (size_t j = 0; j < flat.size(); ++j) {
                RptBVHNode n = flat[j];
                if (n.triangle_count == 0) n.left_or_first += (uint32_t)rest_at - 1u;  // relative child index c >= 1 -> rest_at + c - 1
                out[j == 0 ? root_at : rest_at + j - 1] = n;
            }
            return;
        }
        RptBVHNode root = piece.root;
        root.triangle_count = 0;
        root.left_or_first = (uint32_t)rest_at;
        out[root_at] = root;
        const size_t left_rest = rest_at + 2, right_rest = left_rest + (piece.left->size - 1);
        auto left_task = std::async(std::launch::async, [=, &piece] { emit(*piece.left, out, rest_at, left_rest); });
        emit(*piece.right, out, rest_at + 1, right_rest);
        left_task.get();
    }

    // The reference's loop, src/bvh.rs:257-323, on one subtree.
    std::vector<RptBVHNode> build_sequential(uint32_t first, uint32_t count) {
        Scratch scratch(bins_);
        std::vector<RptBVHNode> nodes;
        nodes.reserve(2 * (size_t)count);
        nodes.push_back(leaf(first, count));
        fit(nodes[0]);
        std::vector<uint32_t> todo{0};
        while (!todo.empty()) {
            const uint32_t ni = todo.back();
            todo.pop_back();
            const RptBVHNode node = nodes[ni];
            uint32_t left_n = 0;
            if (!split(node, scratch, left_n)) continue;
            const uint32_t l = (uint32_t)nodes.size(), r = l + 1;
            nodes[ni].left_or_first = l;
            nodes[ni].triangle_count = 0;
            nodes.push_back(leaf(node.left_or_first, left_n));
            nodes.push_back(leaf(node.left_or_first + left_n, node.triangle_count - left_n));
            fit(nodes[l]);
            fit(nodes[r]);
            todo.push_back(r);
            todo.push_back(l);
        }
        return nodes;
    }

    // Decides whether `node` (a leaf over its triangle range) is split and, if so, partitions the range in place.
    bool split(const RptBVHNode& node, Scratch& scratch, uint32_t& left_n, unsigned workers = 1) {
        int axis;
        float plane, cost;
        best_split(node, scratch, workers, axis, plane, cost);
        const float keep_cost = node_half_area(node) * (float)node.triangle_count;
        if (keep_cost <= cost) return false;
        // in-place partition of [first, first+count) around the plane; 64-bit cursors so the
        // `b -= 1` at b == 0 cannot wrap (the Rust code would panic there)
        int64_t a = node.left_or_first;
        int64_t b = (int64_t)node.left_or_first + node.triangle_count - 1;
        while (a <= b) {
            if (recs_[a].centroid[axis] < plane) {
                ++a;
            } else {
                std::swap(tris_[a], tris_[b]);
                std::swap(recs_[a], recs_[b]);
                --b;
            }
        }
        left_n = (uint32_t)(a - node.left_or_first);
        return left_n != 0 && left_n != node.triangle_count;
    }

    void fit(RptBVHNode& n, unsigned workers = 1) const {  // src/bvh.rs:91-110
        if (n.triangle_count < kSharedNodeMinTriangles) workers = 1;
        std::vector<Box> part(workers);
        run_chunks(n.triangle_count, workers, [&](unsigned w, uint32_t begin, uint32_t end) {
            Box b;
            for (uint32_t k = begin; k < end; ++k) {
                const TriRec& r = recs_[n.left_or_first + k];
                b.grow(r.lo, r.hi);
            }
            part[w] = b;
        });
        Box box;
        for (const Box& b : part) box.grow(b.lo, b.hi);
        n.aabb_min[0] = box.lo.x; n.aabb_min[1] = box.lo.y; n.aabb_min[2] = box.lo.z;
        n.aabb_max[0] = box.hi.x; n.aabb_max[1] = box.hi.y; n.aabb_max[2] = box.hi.z;
    }

    // src/bvh.rs:178-255 — binned sweep; candidate planes sit between adjacent bins.  The reference walks the
    // triangles once per axis and sweeps all bins of every node; here the three axes are binned in one pass and a node
    // with fewer triangles than bins sweeps only the bins it filled.  Same result, bit for bit: a plane that follows an
    // empty bin has the counts and boxes — hence the cost — of the plane before it, which the strict `<` already
    // prefers, and a plane with nothing on one side costs 0 * inf = NaN, which never wins.
    void best_split(const RptBVHNode& node, Scratch& sc, unsigned workers, int& best_axis, float& best_plane, float& best_cost) const {
        best_axis = 0;
        best_plane = 0.0f;
        best_cost = kInf;
        const uint32_t first = node.left_or_first, count = node.triangle_count;
        const uint32_t nb = bins_;
        const TriRec* recs = recs_.data() + first;

        // Big nodes near the root are shared by the workers this subtree owns: each takes a run of the triangles, and
        // the partial bounds, boxes and counts are merged in run order — min / max / + give the sequential result exactly
        // (a tie between +0 and -0 goes to the later triangle either way).
        if (workers > 1 && count < kSharedNodeMinTriangles) workers = 1;
        float lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
        {
            std::vector<float> part(6 * (size_t)workers);
            run_chunks(count, workers, [&](unsigned w, uint32_t begin, uint32_t end) {
                float l[3] = {kInf, kInf, kInf}, h[3] = {-kInf, -kInf, -kInf};
                for (uint32_t k = begin; k < end; ++k) {
                    const F3 c = recs[k].centroid;
                    l[0] = acc_min(l[0], c.x); h[0] = acc_max(h[0], c.x);
                    l[1] = acc_min(l[1], c.y); h[1] = acc_max(h[1], c.y);
                    l[2] = acc_min(l[2], c.z); h[2] = acc_max(h[2], c.z);
                }
                for (int a = 0; a < 3; ++a) { part[6 * (size_t)w + a] = l[a]; part[6 * (size_t)w + 3 + a] = h[a]; }
            });
            for (unsigned w = 0; w < workers; ++w)
                for (int a = 0; a < 3; ++a) { lo[a] = acc_min(lo[a], part[6 * (size_t)w + a]); hi[a] = acc_max(hi[a], part[6 * (size_t)w + 3 + a]); }
        }
        float to_bin[3];
        bool live[3];
        for (int axis = 0; axis < 3; ++axis) {
            live[axis] = !(lo[axis] == hi[axis]);
            to_bin[axis] = (float)nb / (hi[axis] - lo[axis]);
        }
        const bool sparse = count < nb;  // a sparse node empties the bins it filled; a dense one leaves them for the next node to clear
        if (!sparse || sc.dirty) {
            std::fill(sc.seg_box.begin(), sc.seg_box.end(), Box{});
            std::fill(sc.seg_count.begin(), sc.seg_count.end(), 0u);
        }
        sc.dirty = !sparse;
        {
            std::vector<Box> part_box((size_t)(workers - 1) * 3 * nb);
            std::vector<uint32_t> part_count((size_t)(workers - 1) * 3 * nb, 0u);
            run_chunks(count, workers, [&](unsigned w, uint32_t begin, uint32_t end) {
                Box* seg_box = w == 0 ? sc.seg_box.data() : part_box.data() + (size_t)(w - 1) * 3 * nb;
                uint32_t* seg_count = w == 0 ? sc.seg_count.data() : part_count.data() + (size_t)(w - 1) * 3 * nb;
                for (uint32_t k = begin; k < end; ++k) {
                    const TriRec& r = recs[k];
                    for (int axis = 0; axis < 3; ++axis) {
                        if (!live[axis]) continue;
                        const uint32_t b = bin_of((r.centroid[axis] - lo[axis]) * to_bin[axis], nb);
                        if (sparse) sc.tri_bin[(size_t)axis * nb + k] = b;
                        const size_t bin = b + (size_t)axis * nb;
                        seg_box[bin].grow(r.lo, r.hi);
                        seg_count[bin] += 1;
                    }
                }
            });
            for (unsigned w = 1; w < workers; ++w)
                for (size_t bin = 0; bin < 3 * (size_t)nb; ++bin) {
                    const size_t at = (size_t)(w - 1) * 3 * nb + bin;
                    sc.seg_box[bin].grow(part_box[at].lo, part_box[at].hi);
                    sc.seg_count[bin] += part_count[at];
                }
        }

        for (int axis = 0; axis < 3; ++axis) {
            if (!live[axis]) continue;
            const Box* seg_box = sc.seg_box.data() + (size_t)axis * nb;
            const uint32_t* seg_count = sc.seg_count.data() + (size_t)axis * nb;
            const float step = (hi[axis] - lo[axis]) / (float)nb;
            if (!sparse) {
                Box lbox, rbox;
                uint32_t lsum = 0, rsum = 0;
                for (uint32_t i = 0; i + 1 < nb; ++i) {
                    lsum += seg_count[i];
                    sc.left_count[i] = lsum;
                    lbox.grow(seg_box[i]);
                    sc.left_area[i] = lbox.half_area();
                    rsum += seg_count[nb - 1 - i];
                    sc.right_count[nb - 2 - i] = rsum;
                    rbox.grow(seg_box[nb - 1 - i]);
                    sc.right_area[nb - 2 - i] = rbox.half_area();
                }
                for (uint32_t i = 0; i + 1 < nb; ++i) {
                    const float c = (float)sc.left_count[i] * sc.left_area[i] + (float)sc.right_count[i] * sc.right_area[i];
                    if (c < best_cost) {
                        best_axis = axis;
                        best_plane = lo[axis] + step * (float)(i + 1);
                        best_cost = c;
                    }
                }
                continue;
            }
            // the filled bins in ascending order (at least two: the smallest centroid is in bin 0, the largest in the last)
            sc.touched.clear();
            if (count <= 16) {  // a handful of triangles: sort their bins (insertion, dropping repeats)
                const uint32_t* tb = sc.tri_bin.data() + (size_t)axis * nb;
                for (uint32_t k = 0; k < count; ++k) {
                    size_t at = sc.touched.size();
                    while (at > 0 && sc.touched[at - 1] > tb[k]) --at;
                    if (at > 0 && sc.touched[at - 1] == tb[k]) continue;
                    sc.touched.insert(sc.touched.begin() + at, tb[k]);
                }
            } else {
                for (uint32_t i = 0; i < nb; ++i)
                    if (seg_count[i] != 0) sc.touched.push_back(i);
            }
            const uint32_t m = (uint32_t)sc.touched.size();
            Box lbox, rbox;
            uint32_t lsum = 0, rsum = 0;
            for (uint32_t j = 0; j + 1 < m; ++j) {  // entry j: the plane right after filled bin j
                lsum += seg_count[sc.touched[j]];
                sc.left_count[j] = lsum;
                lbox.grow(seg_box[sc.touched[j]]);
                sc.left_area[j] = lbox.half_area();
                rsum += seg_count[sc.touched[m - 1 - j]];
                sc.right_count[m - 2 - j] = rsum;
                rbox.grow(seg_box[sc.touched[m - 1 - j]]);
                sc.right_area[m - 2 - j] = rbox.half_area();
            }
            for (uint32_t j = 0; j + 1 < m; ++j) {
                const float c = (float)sc.left_count[j] * sc.left_area[j] + (float)sc.right_count[j] * sc.right_area[j];
                if (c < best_cost) {
                    best_axis = axis;
                    best_plane = lo[axis] + step * (float)(sc.touched[j] + 1);
                    best_cost = c;
                }
            }
            for (uint32_t bin : sc.touched) {  // leave the bins clean
                sc.seg_box[(size_t)axis * nb + bin] = Box{};
                sc.seg_count[(size_t)axis * nb + bin] = 0u;
            }
        }
    }

    const float* verts_;
    Tri* tris_;
    uint32_t ntris_;
    uint32_t bins_;
    std::vector<TriRec> recs_;
};

// src/light_pick.rs:5-11 — Heron's formula
float heron_area(F3 a, F3 b, F3 c) {
    const float la = length(sub(b, a)), lb = length(sub(c, b)), lc = length(sub(a, c));
    const float s = ((la + lb) + lc) / 2.0f;
    return std::sqrt(((s * (s - la)) * (s - lb)) * (s - lc));
}

}  // namespace

extern "C" int rpt_build_bvh(const float* vertices, uint32_t nverts, uint32_t* indices, uint32_t ntris,
                             uint32_t sah_samples, RptBVHNode* nodes_out, uint32_t* nnodes_out) {
    if (!vertices || !indices || !nodes_out || !nnodes_out) return RPT_ERR_INVALID_ARGUMENT;
    if (ntris == 0 || sah_samples < 2) return RPT_ERR_INVALID_ARGUMENT;
    for (size_t i = 0; i < (size_t)ntris; ++i)
        for (int k = 0; k < 3; ++k)
            if (indices[4 * i + k] >= nverts) return RPT_ERR_INVALID_ARGUMENT;
    SahBuilder builder(vertices, reinterpret_cast<Tri*>(indices), ntris, sah_samples);
    *nnodes_out = builder.run(nodes_out);
    return RPT_OK;
}

extern "C" int rpt_build_light_pick_table(const float* vertices, uint32_t nverts, const uint32_t* indices, uint32_t ntris,
                                          const RptMaterialData* materials, uint32_t nmaterials,
                                          RptLightPickEntry* table_out, uint32_t* nentries_out) {
    if (!vertices || !indices || !materials || !table_out || !nentries_out) return RPT_ERR_INVALID_ARGUMENT;
    const Tri* tris = reinterpret_cast<const Tri*>(indices);
    std::vector<float> area(ntris, 0.0f), power(ntris, 0.0f), prob(ntris, 0.0f);
    float total_power = 0.0f;
    uint32_t emitters = 0;
    for (uint32_t t = 0; t < ntris; ++t) {
        if (tris[t].mat >= nmaterials || tris[t].i0 >= nverts || tris[t].i1 >= nverts || tris[t].i2 >= nverts)
            return RPT_ERR_INVALID_ARGUMENT;
        const float* e = materials[tris[t].mat].emissive;
        if (e[0] == 0.0f && e[1] == 0.0f && e[2] == 0.0f) continue;  // compute_emissive_mask, :13-21
        ++emitters;
        area[t] = heron_area(f3(vertices + 4 * (size_t)tris[t].i0), f3(vertices + 4 * (size_t)tris[t].i1),
                             f3(vertices + 4 * (size_t)tris[t].i2));
        power[t] = ((e[0] * 1.0f + e[1] * 1.0f) + e[2] * 1.0f) * area[t];
        total_power += power[t];
    }
    if (emitters == 0) {  // :53-59 — wgpu cannot bind an empty buffer, hence the sentinel
        std::memset(table_out, 0, sizeof(RptLightPickEntry));
        table_out[0].ratio = -1.0f;
        *nentries_out = 1;
        return RPT_OK;
    }
    float prob_sum = 0.0f;
    for (uint32_t t = 0; t < ntris; ++t) {
        prob[t] = power[t] / total_power;
        prob_sum += prob[t];
    }
    const float mean_prob = prob_sum / (float)emitters;

    struct Bin { uint32_t a; float pa; uint32_t b; float pb; };
    std::vector<Bin> bins;
    for (uint32_t t = 0; t < ntris; ++t)
        if (prob[t] != 0.0f) bins.push_back({t, prob[t], 0u, 0.0f});
    std::stable_sort(bins.begin(), bins.end(), [](const Bin& l, const Bin& r) { return l.pa < r.pa; });

    // :90-104 — top up the least likely bins from the most likely one
    int64_t donor = (int64_t)bins.size() - 1;
    for (size_t i = 0; i < bins.size() && donor >= 0; ++i) {
        const float needed = mean_prob - bins[i].pa;
        if (needed <= 0.0f) break;
        bins[i].b = bins[donor].a;
        bins[i].pb = needed;
        bins[donor].pa -= needed;
        if (bins[donor].pa <= mean_prob) --donor;
    }
    for (size_t i = 0; i < bins.size(); ++i) {
        RptLightPickEntry& e = table_out[i];
        e.triangle_index_a = bins[i].a;
        e.triangle_area_a = area[bins[i].a];
        e.triangle_pick_pdf_a = prob[bins[i].a];
        e.triangle_index_b = bins[i].b;
        e.triangle_area_b = area[bins[i].b];
        e.triangle_pick_pdf_b = prob[bins[i].b];
        e.ratio = bins[i].pa / (bins[i].pa + bins[i].pb);
    }
    *nentries_out = (uint32_t)bins.size();
    return RPT_OK;
}

extern "C" int rpt_pack_per_vertex(const float* vertices, const float* normals, const float* tangents, const float* uvs,
                                   uint32_t nverts, RptPerVertexData* out) {
    if (!vertices || !out) return RPT_ERR_INVALID_ARGUMENT;
    std::memset(out, 0, sizeof(RptPerVertexData) * (size_t)nverts);
    for (size_t v = 0; v < nverts; ++v) {
        std::memcpy(out[v].vertex, vertices + 4 * v, 16);
        if (normals) std::memcpy(out[v].normal, normals + 4 * v, 16);
        if (tangents) std::memcpy(out[v].tangent, tangents + 4 * v, 16);
        if (uvs) std::memcpy(out[v].uv0, uvs + 2 * v, 8);
    }
    return RPT_OK;
}

extern "C" int rpt_make_rng_seeds(const uint8_t* blue_r8, uint32_t bw, uint32_t bh, uint32_t width, uint32_t height,
                                  uint64_t uniform_seed, uint32_t* seeds_xy_out) {
    if (!seeds_xy_out || (blue_r8 && (bw == 0 || bh == 0))) return RPT_ERR_INVALID_ARGUMENT;
    uint64_t s = uniform_seed;
    for (uint32_t y = 0; y < height; ++y) {
        for (uint32_t x = 0; x < width; ++x) {
            uint32_t* out = seeds_xy_out + 2 * ((size_t)y * width + x);
            if (blue_r8) {
                const float pixel = (float)blue_r8[(size_t)(y % bh) * bw + (x % bw)] / 255.0f;
                const float scaled = pixel * 4294967295.0f;  // the literal rounds to 2^32 in f32
                out[0] = 0;
                out[1] = scaled >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)scaled;  // saturating `as u32`
            } else {
                s += 0x9E3779B97F4A7C15ull;  // splitmix64 (the reference draws from thread_rng here)
                uint64_t z = s;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                out[0] = (uint32_t)((z ^ (z >> 31)) >> 32);
                out[1] = 0;
            }
        }
    }
    return RPT_OK;
}

extern "C" int rpt_tile_partition_pixels(uint32_t width, uint32_t height, uint32_t tile_rank, uint32_t tile_count,
                                         uint32_t* pixels_out, uint32_t* npixels_out) {
    if (!npixels_out || tile_count == 0 || tile_rank >= tile_count || width == 0 || height == 0) return RPT_ERR_INVALID_ARGUMENT;
    const uint32_t tiles_x = (width + 31u) / 32u;
    uint32_t n = 0;
    for (uint32_t ty = 0; ty * 32u < height; ++ty)
        for (uint32_t tx = 0; tx < tiles_x; ++tx) {
            if ((ty * tiles_x + tx) % tile_count != tile_rank) continue;
            for (uint32_t y = ty * 32u; y < std::min(height, ty * 32u + 32u); ++y)
                for (uint32_t x = tx * 32u; x < std::min(width, tx * 32u + 32u); ++x) {
                    if (pixels_out) pixels_out[n] = y * width + x;
                    ++n;
                }
        }
    *npixels_out = n;
    return RPT_OK;
}

extern "C" int rpt_camera_matrix(float rot_x, float rot_y, float* m) {
    if (!m) return RPT_ERR_INVALID_ARGUMENT;
    // glam: from_rotation_y cols (c,0,-s),(0,1,0),(s,0,c); from_rotation_x cols (1,0,0),(0,c,s),(0,-s,c)
    const float sy = std::sin(rot_y), cy = std::cos(rot_y), sx = std::sin(rot_x), cx = std::cos(rot_x);
    const float ry[3][3] = {{cy, 0.0f, -sy}, {0.0f, 1.0f, 0.0f}, {sy, 0.0f, cy}};  // [col][row]
    const float rx[3][3] = {{1.0f, 0.0f, 0.0f}, {0.0f, cx, sx}, {0.0f, -sx, cx}};
    for (int c = 0; c < 3; ++c)      // result col c = Ry * rx[c] = (Ry.col0*v.x + Ry.col1*v.y) + Ry.col2*v.z
        for (int r = 0; r < 3; ++r)
            m[3 * c + r] = (ry[0][r] * rx[c][0] + ry[1][r] * rx[c][1]) + ry[2][r] * rx[c][2];
    return RPT_OK;
}
