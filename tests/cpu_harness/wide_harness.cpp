// wide_harness.cpp — TEST-ONLY host build of the backend's wide-BVH code.
//
// The 8-wide BVH collapse (csrc/wide_bvh_build.cpp) and its traversal (csrc/dev/wide_bvh.cuh)
// are plain C++ apart from a handful of intrinsics, so this harness compiles them with g++
// (-ffp-contract=off) to check, without a GPU, that the re-laid-out tree returns the same
// triangle and the same t bits as the reference traversal restated in oracle/oracle.cpp.
// It is not part of the product: nothing under rust-path-tracer_b200/ builds or loads it.
#include <cstdint>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../rust-path-tracer_b200/csrc/dev/wide_bvh.cuh"
#include "../../rust-path-tracer_b200/csrc/wide_bvh.h"

namespace {
struct LocalStack {
    rpt::uint2 data[rpt::kWideStackCapacity];
    int n = 0;
    int high_water = 0;
    void push(rpt::uint2 v) { if (n < (int)rpt::kWideStackCapacity) data[n] = v; ++n; if (n > high_water) high_water = n; }  // (callers check high_water)
    rpt::uint2 pop() { return data[--n]; }
    bool empty() const { return n == 0; }
    void clear() { n = 0; }
    uint32_t permute(uint32_t oct, uint32_t m) const { return rpt::octant_permute(oct, m); }
};
}  // namespace

extern "C" {

// Traversal statistics of nearest-hit rays (tools / experiments): [0] node visits, [1] visits that hit no child and no
// triangle, [2] triangle tests, [3] rays
int harness_wide_stats(const RptPerVertexData* verts, uint32_t nverts, const uint32_t* tris, uint32_t ntris, const RptBVHNode* nodes, uint32_t nnodes,
                       const float* rays_o_d, uint32_t nrays, uint64_t* out_stats) {
    rpt::WideBvh wide;
    const char* err = "";
    if (!rpt::build_wide_bvh(nodes, nnodes, tris, ntris, verts, nverts, wide, &err)) return -1;
    rpt::WideScene scene{reinterpret_cast<const rpt::uint4*>(wide.nodes.data()), reinterpret_cast<const rpt::float4*>(wide.tri_pos.data()), rpt::kHalf1024Bytes};
    uint64_t visits = 0, empty = 0, tests = 0;
    for (uint32_t i = 0; i < nrays; ++i) {
        const float* r = rays_o_d + 6 * (size_t)i;
        LocalStack st;
        rpt::WideCursor<true> c;
        c.begin(rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f);
        while (c.has_nodes()) {
            c.visit_node(scene, st);
            ++visits;
            if (!c.has_triangles() && c.last_child_hits == 0u) ++empty;
            while (c.has_triangles()) { c.test_triangle(scene); ++tests; }
        }
    }
    out_stats[0] = visits; out_stats[1] = empty; out_stats[2] = tests; out_stats[3] = nrays;
    return 0;
}

// Experiment: the same wide tree traversed with EXACT front-to-back order and per-entry distance culling (every hit
// inner child is pushed with its entry distance, far ones first; a popped entry whose distance is not below the best
// hit is dropped without a visit).  Statistics as harness_wide_stats: [0] node visits, [1] entries culled at pop,
// [2] triangle tests, [3] rays, [4] stack high-water.
int harness_wide_stats_sorted(const RptPerVertexData* verts, uint32_t nverts, const uint32_t* tris, uint32_t ntris, const RptBVHNode* nodes,
                              uint32_t nnodes, const float* rays_o_d, uint32_t nrays, uint64_t* out_stats) {
    rpt::WideBvh wide;
    const char* err = "";
    if (!rpt::build_wide_bvh(nodes, nnodes, tris, ntris, verts, nverts, wide, &err)) return -1;
    uint64_t visits = 0, culled = 0, tests = 0, high = 0;
    struct Entry { uint32_t node; float tn; };
    std::vector<Entry> stack;
    for (uint32_t i = 0; i < nrays; ++i) {
        const float* r = rays_o_d + 6 * (size_t)i;
        const rpt::f3 ro = rpt::mk3(r[0], r[1], r[2]), rd = rpt::mk3(r[3], r[4], r[5]);
        const float idir[3] = {rpt::safe_rcp(rd.x), rpt::safe_rcp(rd.y), rpt::safe_rcp(rd.z)};
        const float o[3] = {ro.x, ro.y, ro.z};
        float best = 1000000.0f;
        stack.clear();
        stack.push_back({0u, 0.0f});
        while (!stack.empty()) {
            const Entry e = stack.back();
            stack.pop_back();
            if (e.tn >= best) { ++culled; continue; }
            ++visits;
            const uint32_t* w = wide.nodes[e.node].w;
            float p[3], cell[3];
            std::memcpy(p, w, 12);
            const uint32_t cb[3] = {w[3], w[7] << 16, w[7] & 0xFFFF0000u};
            std::memcpy(cell, cb, 12);
            const uint32_t imask = w[6] >> 24, valid = w[6] & 0x00FFFFFFu;
            Entry kids[8];
            int nk = 0;
            uint32_t below_inner = 0, tri_rank = 0;
            for (uint32_t s = 0; s < 8; ++s) {
                const bool inner = (imask >> s) & 1u;
                const uint32_t tbits = (valid >> (3 * s)) & 7u;
                float tn = 0.0f, tf = best;
                // plane bytes: qlo_x w[8..9], qlo_y w[10..11], qlo_z w[12..13], qhi_x w[14..15], qhi_y w[16..17], qhi_z w[18..19]
                for (int k = 0; k < 3; ++k) {
                    const uint32_t qlo = (w[8 + 2 * k + s / 4] >> (8 * (s % 4))) & 0xFFu, qhi = (w[14 + 2 * k + s / 4] >> (8 * (s % 4))) & 0xFFu;
                    const float t0 = ((p[k] + (float)qlo * cell[k]) - o[k]) * idir[k], t1 = ((p[k] + (float)qhi * cell[k]) - o[k]) * idir[k];
                    tn = std::fmax(tn, std::fmin(t0, t1));
                    tf = std::fmin(tf, std::fmax(t0, t1));
                }
                const bool hit = (inner || tbits) && tn <= tf;
                if (inner) {
                    if (hit) kids[nk++] = {w[4] + below_inner, tn};
                    ++below_inner;
                } else {
                    const uint32_t cnt = tbits == 7u ? 3u : (tbits == 3u ? 2u : (tbits == 1u ? 1u : 0u));
                    if (hit)
                        for (uint32_t t = 0; t < cnt; ++t) {
                            const float* rec = wide.tri_pos.data() + 12 * (size_t)(w[5] + tri_rank + t);
                            float tt;
                            bool back;
                            ++tests;
                            if (rpt::ray_triangle(ro, rd, rpt::mk3(rec[0], rec[1], rec[2]), rpt::mk3(rec[4], rec[5], rec[6]), rpt::mk3(rec[8], rec[9], rec[10]), tt, back) &&
                                tt > 0.001f && tt < best)
                                best = tt;
                        }
                    tri_rank += cnt;
                }
            }
            std::sort(kids, kids + nk, [](const Entry& a, const Entry& b) { return a.tn > b.tn; });  // far first: the nearest is popped next
            for (int k = 0; k < nk; ++k) stack.push_back(kids[k]);
            if (stack.size() > high) high = stack.size();
        }
    }
    out_stats[0] = visits; out_stats[1] = culled; out_stats[2] = tests; out_stats[3] = nrays; out_stats[4] = high;
    return 0;
}

// texel decoding of the shading kernels (dev/vec.cuh), for the exhaustive check against x / 255.0f
float harness_unorm8(uint32_t x) { return rpt::unorm8(x); }

// out_stats: [0] nodes, [1] max_depth, [2] inner_children, [3] leaf_children, [4] stack high-water
int harness_wide_intersect(const RptPerVertexData* verts, uint32_t nverts, const uint32_t* tris, uint32_t ntris, const RptBVHNode* nodes,
                           uint32_t nnodes, const float* rays_o_d, uint32_t nrays, int any_hit, const float* max_t, uint32_t* out_hit,
                           uint32_t* out_tri, float* out_t, uint32_t* out_backface, uint32_t* out_stats) {
    rpt::WideBvh wide;
    const char* err = "";
    if (!rpt::build_wide_bvh(nodes, nnodes, tris, ntris, verts, nverts, wide, &err)) return -1;
    rpt::WideScene scene{reinterpret_cast<const rpt::uint4*>(wide.nodes.data()), reinterpret_cast<const rpt::float4*>(wide.tri_pos.data()), rpt::kHalf1024Bytes};
    int high = 0;
    for (uint32_t i = 0; i < nrays; ++i) {
        const float* r = rays_o_d + 6 * (size_t)i;
        LocalStack st;
        rpt::WideHit h = any_hit ? rpt::wide_intersect<false>(scene, rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), max_t[i], st)
                                 : rpt::wide_intersect<true>(scene, rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f, st);
        out_hit[i] = h.hit;
        out_tri[i] = h.hit ? wide.orig_index[h.triangle] : 0u;
        out_t[i] = h.t;
        out_backface[i] = h.backface;
        if (st.high_water > high) high = st.high_water;
    }
    if (out_stats) {
        out_stats[0] = (uint32_t)wide.nodes.size();
        out_stats[1] = wide.max_depth;
        out_stats[2] = wide.inner_children;
        out_stats[3] = wide.leaf_children;
        out_stats[4] = (uint32_t)high;
    }
    return 0;
}
// Digest of the collapsed tree (FNV-1a over nodes, triangle records and both index maps) + its counts:
// out[0] digest low, [1] digest high, [2] nodes, [3] max depth, [4] inner children, [5] leaf children.
int harness_wide_digest(const RptPerVertexData* verts, uint32_t nverts, const uint32_t* tris, uint32_t ntris, const RptBVHNode* nodes,
                        uint32_t nnodes, uint32_t* out) {
    rpt::WideBvh wide;
    const char* err = "";
    if (!rpt::build_wide_bvh(nodes, nnodes, tris, ntris, verts, nverts, wide, &err)) return -1;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void* p, size_t bytes) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    mix(wide.nodes.data(), wide.nodes.size() * sizeof(rpt::WideNode));
    mix(wide.tri_pos.data(), wide.tri_pos.size() * sizeof(float));
    mix(wide.orig_index.data(), wide.orig_index.size() * sizeof(uint32_t));
    mix(wide.wide_index.data(), wide.wide_index.size() * sizeof(uint32_t));
    out[0] = (uint32_t)h; out[1] = (uint32_t)(h >> 32);
    out[2] = (uint32_t)wide.nodes.size(); out[3] = wide.max_depth; out[4] = wide.inner_children; out[5] = wide.leaf_children;
    return 0;
}
// Experiment: a 32-lane warp of the extend kernel simulated on the CPU, to estimate warp-level
// instruction counts of traversal policies before building them.  Lanes pull rays from a cursor and are refilled
// when fewer than `refill_below` hold a ray (like wf_trace_kernel); one "node round" is charged whenever at least one
// lane visits a node, one "triangle round" whenever at least one lane tests a triangle.
//   defer_threshold == 0: the kernel as it is — every lane tests its triangles right after its visit.
//   defer_threshold  > 0: a lane parks the triangle groups of its visits (`slots` of them) and keeps traversing; the
//                         warp runs triangle rounds only when that many lanes have a parked group, when a lane has no
//                         slot left for a new group, or when nothing else is left to do.  A lane whose traversal is
//                         over waits for the next flush.
// out: [0] node rounds, [1] lane-visits, [2] triangle rounds, [3] lane-tests, [4] rays, [5] refills,
//      [6] rays whose result differs from the plain per-ray loop (must be 0: parking changes when, not what, is tested)
int harness_warp_sim(const RptPerVertexData* verts, uint32_t nverts, const uint32_t* tris, uint32_t ntris, const RptBVHNode* nodes, uint32_t nnodes,
                     const float* rays_o_d, uint32_t nrays, int refill_below, int defer_threshold, int slots, uint64_t* out) {
    rpt::WideBvh wide;
    const char* err = "";
    if (!rpt::build_wide_bvh(nodes, nnodes, tris, ntris, verts, nverts, wide, &err)) return -1;
    rpt::WideScene scene{reinterpret_cast<const rpt::uint4*>(wide.nodes.data()), reinterpret_cast<const rpt::float4*>(wide.tri_pos.data()), rpt::kHalf1024Bytes};
    struct Group { rpt::uint2 tgroup; uint32_t tvalid; };
    struct Lane {
        rpt::WideCursor<true> c;
        LocalStack st;
        bool busy = false;
        uint32_t ray = 0;
        std::vector<Group> parked;  // triangle groups of earlier visits
    };
    std::vector<Lane> lanes(32);
    uint64_t node_rounds = 0, lane_visits = 0, tri_rounds = 0, lane_tests = 0, refills = 0, wrong = 0;
    uint32_t fetch = 0;
    auto live = [&] { int n = 0; for (const Lane& l : lanes) n += l.busy; return n; };
    auto flush = [&] {  // every lane tests one triangle per round until nothing is parked or pending
        for (;;) {
            int active = 0;
            for (Lane& l : lanes) {
                if (!l.busy) continue;
                if (!l.c.has_triangles()) {
                    if (l.parked.empty()) continue;
                    l.c.tgroup = l.parked.back().tgroup;
                    l.c.tvalid = l.parked.back().tvalid;
                    l.parked.pop_back();
                }
                l.c.test_triangle(scene);
                ++active;
            }
            if (!active) break;
            ++tri_rounds;
            lane_tests += (uint64_t)active;
        }
    };
    for (;;) {
        if (fetch < nrays && live() < refill_below) {
            ++refills;
            for (Lane& l : lanes) {
                if (l.busy || fetch >= nrays) continue;
                l.ray = fetch;
                const float* r = rays_o_d + 6 * (size_t)fetch++;
                l.c.begin(rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f);
                l.st.clear();
                l.parked.clear();
                l.busy = true;
            }
        }
        if (live() == 0) break;
        // ---- node round
        int visiting = 0;
        bool slot_needed = false;
        for (Lane& l : lanes) {
            if (!l.busy || !l.c.has_nodes()) continue;
            if (l.c.has_triangles()) {  // the group of the previous visit moves to a slot (there is one: see below)
                l.parked.push_back(Group{l.c.tgroup, l.c.tvalid});
                l.c.tgroup.y = 0u;
            }
            l.c.visit_node(scene, l.st);
            ++visiting;
            if (l.c.has_triangles() && (int)l.parked.size() >= slots) slot_needed = true;  // no slot for this group at the next visit
        }
        if (visiting) { ++node_rounds; lane_visits += (uint64_t)visiting; }
        // ---- triangle rounds
        int holding = 0;
        for (const Lane& l : lanes) holding += l.busy && (l.c.has_triangles() || !l.parked.empty());
        if (defer_threshold == 0 || slot_needed || holding >= defer_threshold || visiting == 0) flush();
        for (Lane& l : lanes)
            if (l.busy && !l.c.has_nodes() && !l.c.has_triangles() && l.parked.empty()) {
                l.busy = false;
                const float* r = rays_o_d + 6 * (size_t)l.ray;
                LocalStack st;
                const rpt::WideHit want = rpt::wide_intersect<true>(scene, rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f, st);
                const rpt::WideHit got = l.c.result();
                if (got.hit != want.hit || (got.hit && (got.triangle != want.triangle || std::memcmp(&got.t, &want.t, 4) != 0 || got.backface != want.backface))) ++wrong;
            }
    }
    out[6] = wrong;
    out[0] = node_rounds; out[1] = lane_visits; out[2] = tri_rounds; out[3] = lane_tests; out[4] = nrays; out[5] = refills;
    return 0;
}

// Experiment 2 (round 2): deferred triangle tests with a per-lane queue of pending triangles and PARTIAL flushes.
// After every node round the warp counts the lanes that hold a pending triangle; triangle rounds start when at least
// `t_hi` lanes do (or `blocked_at` lanes have finished walking and only wait for their triangles, or a lane's queue
// is full, or no lane can visit a node) and go on while at least `t_lo` lanes still do — the rest stays queued for a
// later flush, so a triangle round never runs narrower than t_lo lanes unless it is forced.  A lane whose traversal is
// over keeps its queue and waits (it counts towards the thresholds).  capacity < 0: LIFO queue, else FIFO.
// out as harness_warp_sim, plus [7] forced triangle rounds.
int harness_warp_sim2(const RptPerVertexData* verts, uint32_t nverts, const uint32_t* tris, uint32_t ntris, const RptBVHNode* nodes, uint32_t nnodes,
                      const float* rays_o_d, uint32_t nrays, int refill_below, int t_hi, int t_lo, int capacity, int blocked_at, uint64_t* out) {
    rpt::WideBvh wide;
    const char* err = "";
    if (!rpt::build_wide_bvh(nodes, nnodes, tris, ntris, verts, nverts, wide, &err)) return -1;
    rpt::WideScene scene{reinterpret_cast<const rpt::uint4*>(wide.nodes.data()), reinterpret_cast<const rpt::float4*>(wide.tri_pos.data()), rpt::kHalf1024Bytes};
    struct Lane {
        rpt::WideCursor<true> c;
        LocalStack st;
        bool busy = false;
        uint32_t ray = 0;
        std::vector<uint32_t> queue;  // pending triangles (wide indices)
    };
    std::vector<Lane> lanes(32);
    uint64_t node_rounds = 0, lane_visits = 0, tri_rounds = 0, lane_tests = 0, refills = 0, wrong = 0, forced = 0;
    uint32_t fetch = 0;
    auto live = [&] { int n = 0; for (const Lane& l : lanes) n += l.busy; return n; };
    auto holding = [&] { int n = 0; for (const Lane& l : lanes) n += l.busy && !l.queue.empty(); return n; };
    auto tri_round = [&] {
        int active = 0;
        for (Lane& l : lanes) {
            if (!l.busy || l.queue.empty()) continue;
            uint32_t ti;
            if (capacity < 0) { ti = l.queue.back(); l.queue.pop_back(); }
            else { ti = l.queue.front(); l.queue.erase(l.queue.begin()); }
            l.c.tgroup = rpt::make_uint2(ti, 1u);  // a one-bit group whose rank-0 triangle is ti
            l.c.tvalid = 1u;
            l.c.test_triangle(scene);
            ++active;
        }
        if (active) { ++tri_rounds; lane_tests += (uint64_t)active; }
        return active;
    };
    for (;;) {
        if (fetch < nrays && live() < refill_below) {
            ++refills;
            for (Lane& l : lanes) {
                if (l.busy || fetch >= nrays) continue;
                l.ray = fetch;
                const float* r = rays_o_d + 6 * (size_t)fetch++;
                l.c.begin(rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f);
                l.st.clear();
                l.queue.clear();
                l.busy = true;
            }
        }
        if (live() == 0) break;
        int visiting = 0;
        bool full = false;
        for (Lane& l : lanes) {
            if (!l.busy || !l.c.has_nodes()) continue;
            if ((int)l.queue.size() + 6 > std::abs(capacity)) { full = true; continue; }
            l.c.visit_node(scene, l.st);
            ++visiting;
            while (l.c.has_triangles()) {
                const int k = rpt::highest_bit(l.c.tgroup.y);
                l.c.tgroup.y &= ~(1u << k);
                l.queue.push_back(l.c.tgroup.x + (uint32_t)rpt::popcount(l.c.tvalid & ~(0xFFFFFFFFu << k)));
            }
        }
        if (visiting) { ++node_rounds; lane_visits += (uint64_t)visiting; }
        int h = holding();
        int blocked = 0;
        for (const Lane& l : lanes) blocked += l.busy && !l.c.has_nodes() && !l.queue.empty();
        if (h >= t_hi) {
            while (h >= t_lo && tri_round()) h = holding();
        } else if (blocked >= blocked_at) {
            do { tri_round(); h = holding(); } while (h >= t_lo);
        } else if ((full || visiting == 0) && h > 0) {
            do { tri_round(); ++forced; h = holding(); } while (h >= t_lo);
        }
        for (Lane& l : lanes)
            if (l.busy && !l.c.has_nodes() && l.queue.empty()) {
                l.busy = false;
                const float* r = rays_o_d + 6 * (size_t)l.ray;
                LocalStack st;
                const rpt::WideHit want = rpt::wide_intersect<true>(scene, rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f, st);
                const rpt::WideHit got = l.c.result();
                if (got.hit != want.hit || (got.hit && (got.triangle != want.triangle || std::memcmp(&got.t, &want.t, 4) != 0 || got.backface != want.backface))) ++wrong;
            }
    }
    out[6] = wrong; out[7] = forced;
    out[0] = node_rounds; out[1] = lane_visits; out[2] = tri_rounds; out[3] = lane_tests; out[4] = nrays; out[5] = refills;
    return 0;
}

// Experiment 3 (round 2): ONE STEP PER LANE PER ROUND.  In every round a lane tests one pending triangle if it has one
// and visits its next node otherwise; the warp runs the triangle block (for the lanes that chose it) and the node
// block (for the others) once each.  A lane never waits for another lane's triangles, and every ray performs exactly
// the per-ray sequence of the plain loop (visit, that visit's triangles, next visit), so results and visit counts are
// unchanged; what changes is that the triangle block runs once per round for every lane that currently holds a
// triangle, instead of max-over-lanes times per node round for the few lanes that found one in THAT round.
//   tri_min: the triangle block only runs when at least that many lanes want it or no lane wants a node (lanes that
//            want a triangle idle through the round otherwise); 1 = always.
// out as harness_warp_sim.
int harness_warp_sim3(const RptPerVertexData* verts, uint32_t nverts, const uint32_t* tris, uint32_t ntris, const RptBVHNode* nodes, uint32_t nnodes,
                      const float* rays_o_d, uint32_t nrays, int refill_below, int tri_min, uint64_t* out) {
    rpt::WideBvh wide;
    const char* err = "";
    if (!rpt::build_wide_bvh(nodes, nnodes, tris, ntris, verts, nverts, wide, &err)) return -1;
    rpt::WideScene scene{reinterpret_cast<const rpt::uint4*>(wide.nodes.data()), reinterpret_cast<const rpt::float4*>(wide.tri_pos.data()), rpt::kHalf1024Bytes};
    struct Lane {
        rpt::WideCursor<true> c;
        LocalStack st;
        bool busy = false;
        uint32_t ray = 0;
    };
    std::vector<Lane> lanes(32);
    uint64_t node_rounds = 0, lane_visits = 0, tri_rounds = 0, lane_tests = 0, refills = 0, wrong = 0;
    uint32_t fetch = 0;
    auto live = [&] { int n = 0; for (const Lane& l : lanes) n += l.busy; return n; };
    for (;;) {
        if (fetch < nrays && live() < refill_below) {
            ++refills;
            for (Lane& l : lanes) {
                if (l.busy || fetch >= nrays) continue;
                l.ray = fetch;
                const float* r = rays_o_d + 6 * (size_t)fetch++;
                l.c.begin(rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f);
                l.st.clear();
                l.busy = true;
            }
        }
        if (live() == 0) break;
        int want_tri = 0, want_node = 0;
        for (const Lane& l : lanes) {
            if (!l.busy) continue;
            if (l.c.has_triangles()) ++want_tri;
            else if (l.c.has_nodes()) ++want_node;
        }
        const bool run_tri = want_tri > 0 && (want_tri >= tri_min || want_node == 0);
        int testing = 0, visiting = 0;
        for (Lane& l : lanes) {
            if (!l.busy) continue;
            if (l.c.has_triangles()) {
                if (run_tri) { l.c.test_triangle(scene); ++testing; }
            } else if (l.c.has_nodes()) {
                l.c.visit_node(scene, l.st);
                ++visiting;
            }
        }
        if (testing) { ++tri_rounds; lane_tests += (uint64_t)testing; }
        if (visiting) { ++node_rounds; lane_visits += (uint64_t)visiting; }
        for (Lane& l : lanes)
            if (l.busy && !l.c.has_nodes() && !l.c.has_triangles()) {
                l.busy = false;
                const float* r = rays_o_d + 6 * (size_t)l.ray;
                LocalStack st;
                const rpt::WideHit want = rpt::wide_intersect<true>(scene, rpt::mk3(r[0], r[1], r[2]), rpt::mk3(r[3], r[4], r[5]), 0.0f, st);
                const rpt::WideHit got = l.c.result();
                if (got.hit != want.hit || (got.hit && (got.triangle != want.triangle || std::memcmp(&got.t, &want.t, 4) != 0 || got.backface != want.backface))) ++wrong;
            }
    }
    out[6] = wrong;
    out[0] = node_rounds; out[1] = lane_visits; out[2] = tri_rounds; out[3] = lane_tests; out[4] = nrays; out[5] = refills;
    return 0;
}
}
