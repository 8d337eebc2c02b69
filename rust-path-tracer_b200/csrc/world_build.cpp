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
inline F3 vmin(F3 a, F3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline F3 vmax(F3 a, F3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline F3 sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float length(F3 a) { return std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }

// An axis-aligned box that starts empty (+inf, -inf), like `BVHNode::default()`.
struct Box {
    F3 lo{kInf, kInf, kInf};
    F3 hi{-kInf, -kInf, -kInf};
    void grow(F3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
    void grow(const Box& b) {
        if (b.lo.x == kInf) return;  // src/bvh.rs:21-27: empty bins are skipped
        lo = vmin(lo, b.lo);
        hi = vmax(hi, b.hi);
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
// the subtree's root) and spliced together afterwards — same nodes, same order, same permutation of the index
// buffer as the sequential build, in a fraction of the time (1 M triangles: 11.7 s -> ~1.5 s on 16 cores).
class SahBuilder {
  public:
    SahBuilder(const float* verts, Tri* tris, uint32_t ntris, uint32_t bins) : verts_(verts), tris_(tris), ntris_(ntris), bins_(bins), centroids_(ntris) {
        for (uint32_t t = 0; t < ntris; ++t) {  // src/bvh.rs:60-68
            F3 a = pos(tris[t].i0), b = pos(tris[t].i1), c = pos(tris[t].i2);
            centroids_[t] = {((a.x + b.x) + c.x) / 3.0f, ((a.y + b.y) + c.y) / 3.0f, ((a.z + b.z) + c.z) / 3.0f};
        }
    }

    // Nodes of the whole tree in the reference's order; nodes_out has room for 2 * ntris - 1.
    uint32_t run(RptBVHNode* nodes_out) {
        unsigned threads = std::thread::hardware_concurrency();
        if (const char* v = std::getenv("RPT_BUILD_THREADS")) threads = (unsigned)std::max(1, std::atoi(v));
        int fork_levels = 0;
        while ((1u << fork_levels) < std::max(1u, threads)) ++fork_levels;
        const std::vector<RptBVHNode> nodes = build_subtree(0, ntris_, threads <= 1 ? 0 : fork_levels + 1);  // one thread: the reference's loop as is
        std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(RptBVHNode));
        return (uint32_t)nodes.size();
    }

  private:
    struct Scratch {
        std::vector<Box> seg_box;
        std::vector<uint32_t> seg_count, left_count, right_count;
        std::vector<float> left_area, right_area;
        explicit Scratch(uint32_t bins) : seg_box(bins), seg_count(bins), left_count(bins - 1), right_count(bins - 1), left_area(bins - 1), right_area(bins - 1) {}
    };
    static constexpr uint32_t kForkMinTriangles = 8192;  // below this a task costs more than it saves

    F3 pos(uint32_t v) const { return f3(verts_ + 4 * (size_t)v); }
    static float node_half_area(const RptBVHNode& n) {
        Box b;
        b.lo = f3(n.aabb_min);
        b.hi = f3(n.aabb_max);
        return b.half_area();
    }
    static RptBVHNode leaf(uint32_t first, uint32_t count) { return RptBVHNode{{kInf, kInf, kInf}, count, {-kInf, -kInf, -kInf}, first}; }

    // [root, l, r, descendants(l), descendants(r)] of the subtree over triangles [first, first + count); inner
    // nodes hold the index of their left child RELATIVE to the subtree's root.
    std::vector<RptBVHNode> build_subtree(uint32_t first, uint32_t count, int fork_levels) {
        if (fork_levels <= 0 || count < kForkMinTriangles) return build_sequential(first, count);
        Scratch scratch(bins_);
        RptBVHNode root = leaf(first, count);
        fit(root);
        uint32_t left_n = 0;
        if (!split(root, scratch, left_n)) return {root};
        auto left_task = std::async(std::launch::async, [=] { return build_subtree(first, left_n, fork_levels - 1); });
        const std::vector<RptBVHNode> right = build_subtree(first + left_n, count - left_n, fork_levels - 1);
        const std::vector<RptBVHNode> left = left_task.get();
        // splice: L[0] -> 1, R[0] -> 2, L[j >= 1] -> j + 2, R[j >= 1] -> |L| + 1 + j; child indices (always >= 1) move alike
        std::vector<RptBVHNode> out;
        out.reserve(1 + left.size() + right.size());
        root.triangle_count = 0;
        root.left_or_first = 1;
        out.push_back(root);
        out.push_back(left[0]);
        out.push_back(right[0]);
        out.insert(out.end(), left.begin() + 1, left.end());
        out.insert(out.end(), right.begin() + 1, right.end());
        const uint32_t nl = (uint32_t)left.size();
        auto shift = [&](size_t at, uint32_t by) { if (out[at].triangle_count == 0) out[at].left_or_first += by; };
        shift(1, 2);
        shift(2, nl + 1);
        for (size_t j = 1; j < left.size(); ++j) shift(2 + j, 2);
        for (size_t j = 1; j < right.size(); ++j) shift(1 + nl + j, nl + 1);
        return out;
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
    bool split(const RptBVHNode& node, Scratch& scratch, uint32_t& left_n) {
        int axis;
        float plane, cost;
        best_split(node, scratch, axis, plane, cost);
        const float keep_cost = node_half_area(node) * (float)node.triangle_count;
        if (keep_cost <= cost) return false;
        // in-place partition of [first, first+count) around the plane; 64-bit cursors so the
        // `b -= 1` at b == 0 cannot wrap (the Rust code would panic there)
        int64_t a = node.left_or_first;
        int64_t b = (int64_t)node.left_or_first + node.triangle_count - 1;
        while (a <= b) {
            if (centroids_[a][axis] < plane) {
                ++a;
            } else {
                std::swap(tris_[a], tris_[b]);
                std::swap(centroids_[a], centroids_[b]);
                --b;
            }
        }
        left_n = (uint32_t)(a - node.left_or_first);
        return left_n != 0 && left_n != node.triangle_count;
    }

    void fit(RptBVHNode& n) const {  // src/bvh.rs:91-110
        Box box;
        for (uint32_t k = 0; k < n.triangle_count; ++k) {
            const Tri& t = tris_[n.left_or_first + k];
            F3 a = pos(t.i0), b = pos(t.i1), c = pos(t.i2);
            box.lo = vmin(box.lo, vmin(vmin(a, b), c));
            box.hi = vmax(box.hi, vmax(vmax(a, b), c));
        }
        n.aabb_min[0] = box.lo.x; n.aabb_min[1] = box.lo.y; n.aabb_min[2] = box.lo.z;
        n.aabb_max[0] = box.hi.x; n.aabb_max[1] = box.hi.y; n.aabb_max[2] = box.hi.z;
    }

    // src/bvh.rs:178-255 — binned sweep; candidate planes sit between adjacent bins
    void best_split(const RptBVHNode& node, Scratch& sc, int& best_axis, float& best_plane, float& best_cost) const {
        best_axis = 0;
        best_plane = 0.0f;
        best_cost = kInf;
        const uint32_t first = node.left_or_first, count = node.triangle_count;
        const uint32_t nb = bins_;
        for (int axis = 0; axis < 3; ++axis) {
            float lo = kInf, hi = -kInf;
            for (uint32_t k = 0; k < count; ++k) {
                const float c = centroids_[first + k][axis];
                lo = std::fmin(lo, c);
                hi = std::fmax(hi, c);
            }
            if (lo == hi) continue;

            std::fill(sc.seg_box.begin(), sc.seg_box.end(), Box{});
            std::fill(sc.seg_count.begin(), sc.seg_count.end(), 0u);
            const float to_bin = (float)nb / (hi - lo);
            for (uint32_t k = 0; k < count; ++k) {
                const Tri& t = tris_[first + k];
                const float f = (centroids_[first + k][axis] - lo) * to_bin;
                // Rust `as usize` saturates: negative / NaN -> 0
                size_t bin = (f > 0.0f) ? (f >= 1.8446744e19f ? SIZE_MAX : (size_t)f) : 0;
                bin = std::min<size_t>(bin, nb - 1);
                sc.seg_box[bin].grow(pos(t.i0));
                sc.seg_box[bin].grow(pos(t.i1));
                sc.seg_box[bin].grow(pos(t.i2));
                sc.seg_count[bin] += 1;
            }

            Box lbox, rbox;
            uint32_t lsum = 0, rsum = 0;
            for (uint32_t i = 0; i + 1 < nb; ++i) {
                lsum += sc.seg_count[i];
                sc.left_count[i] = lsum;
                lbox.grow(sc.seg_box[i]);
                sc.left_area[i] = lbox.half_area();
                rsum += sc.seg_count[nb - 1 - i];
                sc.right_count[nb - 2 - i] = rsum;
                rbox.grow(sc.seg_box[nb - 1 - i]);
                sc.right_area[nb - 2 - i] = rbox.half_area();
            }

            const float step = (hi - lo) / (float)nb;
            for (uint32_t i = 0; i + 1 < nb; ++i) {
                const float c = (float)sc.left_count[i] * sc.left_area[i] + (float)sc.right_count[i] * sc.right_area[i];
                if (c < best_cost) {
                    best_axis = axis;
                    best_plane = lo + step * (float)(i + 1);
                    best_cost = c;
                }
            }
        }
    }

    const float* verts_;
    Tri* tris_;
    uint32_t ntris_;
    uint32_t bins_;
    std::vector<F3> centroids_;
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
