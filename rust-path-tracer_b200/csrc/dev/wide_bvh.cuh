// wide_bvh.cuh — traversal of the 8-wide compressed BVH (layout: ../wide_bvh.h).
//
// One ray per lane.  Each node visit is five 16-byte loads.  The eight child boxes are
// de-quantised TWO CHILDREN PER INSTRUCTION: one PRMT builds the fp16 pair (1024 + q_j, 1024 + q_k)
// from two plane bytes, two HADD2.F32 widen it, and one packed FFMA2 (fma.rn.f32x2, new in
// sm_100) evaluates both slab distances with the -1024 bias folded into the per-node addend.
// A hit ORs a per-child constant into one word: three triangle bits (3 * slot ..) and one
// child bit (24 + slot); the triangle bits are masked by the node's valid-triangle word, the child
// bits by its inner mask and then permuted by the ray octant through a 2 KB lookup table so that
// "highest bit first" is the front-to-back visiting order (slot ^ octant) with no per-node sort.
// The traversal stack holds (child base, hits << 24 | inner mask) pairs and lives wherever
// `Stack` puts it — a per-warp shared-memory slab on the device.
//
// Box tests are CONSERVATIVE (quantised outwards at build time, planes pushed out by a few ulps
// of the coordinates here) and may use FMA; the ray/triangle test is exact.cuh's, which evaluates
// the reference's operations in the reference's order.  Compile with -fmad=false.
#pragma once

#include <string.h>

#include "exact.cuh"

namespace rpt {

#if defined(__CUDACC__)
RPT_D int highest_bit(uint32_t v) { return 31 - __clz((int)v); }
RPT_D void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
RPT_D int popcount(uint32_t v) { return __popc(v); }
RPT_D float as_float(uint32_t v) { return __uint_as_float(v); }
RPT_D uint32_t as_uint(float v) { return __float_as_uint(v); }
#else
inline int highest_bit(uint32_t v) { return 31 - __builtin_clz(v); }
inline void prefetch_l1(const void*) {}
inline int popcount(uint32_t v) { return __builtin_popcount(v); }
inline float as_float(uint32_t v) { float f; memcpy(&f, &v, 4); return f; }
inline uint32_t as_uint(float v) { uint32_t u; memcpy(&u, &v, 4); return u; }
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
template <class T> inline T __ldg(const T* p) { return *p; }
#endif

struct WideScene {
    const uint4* nodes;     // 5 per node
    const float4* tri_pos;  // 3 per triangle
    // kHalf1024Bytes, passed as DATA: PRMT takes one immediate, and it has to be the byte selector — a
    // literal here would be folded into that slot and the selectors rematerialised in registers instead
    uint32_t half_1024_bytes;
};
// Plane de-quantisation, two builds (RPT_PLANES_FP32):
//   0  fp16 pairs: one PRMT builds (1024 + q_j, 1024 + q_k) as two halves (0x64 = high byte of 1024.0h), two conversions
//      widen them — 2 instructions per plane byte, the conversions on the FMA-side pipe;
//   1  fp32 directly: one PRMT per plane byte builds the float 32768 + q (0x47000000 = 32768.0f, the byte lands in
//      mantissa bits 8..15, whose unit is exactly 1 at that exponent) — 1.5 instructions per plane byte, all of the
//      de-quantisation on the 16-lane ALU pipe.  The larger bias rounds the folded addend more coarsely (below).
#ifndef RPT_PLANES_FP32
#define RPT_PLANES_FP32 0
#endif
#if RPT_PLANES_FP32
constexpr uint32_t kHalf1024Bytes = 0x00004700u;  // bytes 0x00 and 0x47 for the PRMT that builds 32768 + q
constexpr float kPlaneBias = 32768.0f;
constexpr float kPadCells = 16640.0f;  // 256 (box extent) + 16384: the addend is rounded at magnitude 32768 |adj|, i.e. by < 2^-9 |adj| = 3255 x 6e-7 cells
#else
constexpr uint32_t kHalf1024Bytes = 0x64646464u;  // 0x64 = high byte of 1024.0 in fp16
constexpr float kPlaneBias = 1024.0f;
constexpr float kPadCells = 768.0f;
#endif

struct WideHit {
    float t;            // 1e6 if no hit (kernels/src/intersection.rs:68)
    uint32_t triangle;  // index in WIDE order
    bool hit, backface;
};

// Octant permutation of an 8-bit child set: bit s of `m` moves to bit s ^ oct.  The device keeps the
// 8 x 256 table in shared memory (Stack::permute); this is the definition, used to fill it and by the host.
RPT_HD uint32_t octant_permute(uint32_t oct, uint32_t m) {
    uint32_t p = 0;
    for (uint32_t s = 0; s < 8u; ++s)
        if ((m >> s) & 1u) p |= 1u << (s ^ oct);
    return p;
}

// Ray constants hoisted out of the node loop.
// RPT_NODE_INDEXED_LOADS (device builds): the near / far plane words of a node are LOADED from the place the ray's
// signs point at (three per-ray byte offsets) instead of loading all six pairs and choosing with twelve SELs — the
// selection moves off the half-rate ALU pipe onto the load unit.
#ifndef RPT_NODE_INDEXED_LOADS
#define RPT_NODE_INDEXED_LOADS 0
#endif
#ifndef RPT_NODE_SHORT_PAD
#define RPT_NODE_SHORT_PAD 1
#endif

struct WideRay {
    f3 o, d;
    f3 idir;        // 1 / d with |d| clamped away from 0
    uint32_t oct_inv;  // dx>=0 ? 4 : 0 | dy>=0 ? 2 : 0 | dz>=0 ? 1 : 0
#if RPT_NODE_INDEXED_LOADS && defined(__CUDACC__)
    uint32_t near_x, near_y, near_z;  // which of a node's ten 8-byte words holds the planes the ray enters first, per axis
#endif
};

// 1 / d with |d| clamped away from 0.  It only feeds the conservative box tests, so the device uses the 1-ulp
// MUFU reciprocal instead of the IEEE division (three divisions per ray at refill time, where few lanes are
// enabled); the extra half ulp is part of the padding budget below.
RPT_D float safe_rcp(float d) {
    const float clamped = fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d);
#if defined(__CUDACC__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(clamped));
    return r;
#else
    return 1.0f / clamped;
#endif
}

RPT_D WideRay make_wide_ray(f3 o, f3 d) {
    WideRay r;
    r.o = o;
    r.d = d;
    r.idir = mk3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    r.oct_inv = (d.x < 0.0f ? 0u : 4u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 1u);
#if RPT_NODE_INDEXED_LOADS && defined(__CUDACC__)
    // a node as ten 8-byte words: qlo_x 4, qlo_y 5, qlo_z 6, qhi_x 7, qhi_y 8, qhi_z 9 (wide_bvh.h)
    r.near_x = d.x < 0.0f ? 7u : 4u;
    r.near_y = d.y < 0.0f ? 8u : 5u;
    r.near_z = d.z < 0.0f ? 9u : 6u;
#endif
    return r;
}

// What a hit of the child in slot J contributes: its three triangle bits and its child bit.
#define RPT_CHILD_HIT_BITS(J) ((7u << (3 * (J))) | (1u << (24 + (J))))

// Slab distances of two children (bytes LO and LO + 1 of the six plane words), each
// (1024 + q) * adj + (org -/+ pad - 1024 * adj), evaluated by one fused multiply-add per plane.
#if defined(__CUDACC__)
RPT_D unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
// fp16 pair (1024 + byte[SEL & 7], 1024 + byte[(SEL >> 8) & 7]) widened to two floats; 0x64 is the high byte of 1024.0h
RPT_D unsigned long long dequant_pair(uint32_t word, uint32_t magic, uint32_t sel) {
    const uint32_t h2 = __byte_perm(word, magic, sel);
    float lo, hi;
    asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(lo), "=f"(hi) : "r"(h2));
    return pack_f32x2(lo, hi);
}
#if RPT_PLANES_FP32
// (32768 + byte LO, 32768 + byte LO + 1) as two floats: result bytes (0x00, word byte, 0x00, 0x47), selector 0x5404 | byte << 4
RPT_D void slab_pair(uint32_t word, uint32_t magic, uint32_t lo_byte, float adj, float org, float& t0, float& t1) {
    const float q0 = __uint_as_float(__byte_perm(word, magic, 0x5404u | (lo_byte << 4)));
    const float q1 = __uint_as_float(__byte_perm(word, magic, 0x5404u | ((lo_byte + 1u) << 4)));
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack_f32x2(q0, q1)), "l"(pack_f32x2(adj, adj)), "l"(pack_f32x2(org, org)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(r));
}
#define RPT_PAIR_SEL(LO) (LO)
#else
RPT_D void slab_pair(uint32_t word, uint32_t magic, uint32_t sel, float adj, float org, float& t0, float& t1) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(dequant_pair(word, magic, sel)), "l"(pack_f32x2(adj, adj)), "l"(pack_f32x2(org, org)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(r));
}
#define RPT_PAIR_SEL(LO) (0x4140u + 0x0202u * ((LO) / 2u))
#endif
#else
inline void slab_pair(uint32_t word, uint32_t, uint32_t lo_byte, float adj, float org, float& t0, float& t1) {
    t0 = fmaf(kPlaneBias + (float)((word >> (8u * lo_byte)) & 0xFFu), adj, org);
    t1 = fmaf(kPlaneBias + (float)((word >> (8u * lo_byte + 8u)) & 0xFFu), adj, org);
}
#define RPT_PAIR_SEL(LO) (LO)
#endif

// Test the four children whose bytes sit in (near_x, near_y, near_z, far_x, far_y, far_z); FIRST is the slot of byte 0.
template <uint32_t FIRST>
RPT_D uint32_t test_four(uint32_t magic, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz, f3 adj, f3 org_near,
                         f3 org_far, float best_t) {
    uint32_t hits = 0;
#define RPT_PAIR(LO)                                                                               \
    {                                                                                              \
        float tnx0, tnx1, tny0, tny1, tnz0, tnz1, tfx0, tfx1, tfy0, tfy1, tfz0, tfz1;              \
        slab_pair(nx, magic, RPT_PAIR_SEL(LO), adj.x, org_near.x, tnx0, tnx1);                            \
        slab_pair(ny, magic, RPT_PAIR_SEL(LO), adj.y, org_near.y, tny0, tny1);                            \
        slab_pair(nz, magic, RPT_PAIR_SEL(LO), adj.z, org_near.z, tnz0, tnz1);                            \
        slab_pair(fx, magic, RPT_PAIR_SEL(LO), adj.x, org_far.x, tfx0, tfx1);                             \
        slab_pair(fy, magic, RPT_PAIR_SEL(LO), adj.y, org_far.y, tfy0, tfy1);                             \
        slab_pair(fz, magic, RPT_PAIR_SEL(LO), adj.z, org_far.z, tfz0, tfz1);                             \
        const float tn0 = fmaxf(fmaxf(tnx0, tny0), fmaxf(tnz0, 0.0f));                             \
        const float tf0 = fminf(fminf(tfx0, tfy0), fminf(tfz0, best_t));                           \
        const float tn1 = fmaxf(fmaxf(tnx1, tny1), fmaxf(tnz1, 0.0f));                             \
        const float tf1 = fminf(fminf(tfx1, tfy1), fminf(tfz1, best_t));                           \
        if (tn0 <= tf0) hits |= RPT_CHILD_HIT_BITS(FIRST + (LO));                                  \
        if (tn1 <= tf1) hits |= RPT_CHILD_HIT_BITS(FIRST + (LO) + 1u);                             \
    }
    RPT_PAIR(0u)
    RPT_PAIR(2u)
#undef RPT_PAIR
    return hits;
}

// ---- traversal as a resumable cursor -------------------------------------------------------
// The per-ray state between steps.  A step is either one node visit (`visit_node`) or one
// ray/triangle test (`test_triangle`).  The device kernels interleave these steps across the lanes
// of a warp (ray refill); the plain loop `wide_intersect` below composes them for one ray.
//   next   = the node to visit next (kNoNode: traversal finished).  It is chosen at the END of a visit — before
//            that node's triangles are tested — so that its cache lines can be prefetched while the triangle
//            records are in flight; the choice depends on the box tests only, never on the triangle results.
//   ngroup = (first child node, permuted hits << 24 | inner-slot mask): siblings of `next` still to visit
//   tgroup = (first triangle of the node, 24-bit hit mask), tvalid = the node's valid-triangle bits:
//            triangle bit k is the node's popc(tvalid below k)-th triangle
constexpr uint32_t kNoNode = 0xFFFFFFFFu;
// Measured on B200 (BreakTime proxy): prefetch.global.L1 of the next node makes extend 1.7x SLOWER — at 32 warps
// per SM the prefetched lines (256 B per lane) overflow what is left of the L1 next to 150 KB of shared memory
// and evict the top of the tree.  Kept as a switch for the record; off.
constexpr bool kPrefetchNextNode = false;

template <bool NEAREST>
struct WideCursor {
    WideRay ray;
    float best_t;      // culling bound: nearest accepted t so far (nearest; 1e6 = none) / min(max_t, 1e6) (any)
    float hit_t;       // t of the accepted hit (== best_t for nearest)
    uint32_t hit_tri;  // kNoNode: no hit; else wide triangle | back-face << 31
    uint32_t next;
    uint2 ngroup, tgroup;
    uint32_t tvalid;
#if !defined(__CUDACC__)
    uint32_t last_child_hits = 0;  // host builds only (tests / statistics): child bits of the last visit
#endif

    RPT_D void begin(f3 ro, f3 rd, float max_t_) {
        ray = make_wide_ray(ro, rd);
        best_t = NEAREST ? 1000000.0f : fminf(max_t_, 1000000.0f);  // TraceResult::default().t, intersection.rs:68
        hit_t = 1000000.0f;
        hit_tri = kNoNode;
        next = 0u;  // the root
        ngroup = make_uint2(0u, 0u);
        tgroup = make_uint2(0u, 0u);
        tvalid = 0u;
        // a ray with a non-finite component hits nothing in the reference either (every slab test
        // compares false); without this the NaN-ignoring min/max would visit every node
        if (!(finite3(ro) && finite3(rd))) next = kNoNode;
    }
    RPT_D bool has_nodes() const { return next != kNoNode; }
    RPT_D WideHit result() const { return WideHit{hit_t, hit_tri & 0x7FFFFFFFu, hit_tri != kNoNode, hit_tri != kNoNode && (hit_tri >> 31) != 0u}; }
    RPT_D bool has_triangles() const { return tgroup.y != 0u; }
    // Drop everything still to be traversed (an any-hit query that is already decided).
    template <class Stack>
    RPT_D void abandon(Stack& stack) {
        next = kNoNode;
        ngroup.y = 0u;
        tgroup.y = 0u;
        stack.clear();
    }

    // Tests the eight children of `next`, then picks (and prefetches) the node to visit after it.
    template <class Stack>
    RPT_D void visit_node(const WideScene& s, Stack& stack) {
        const Visit v = test_children(s);
        select_next(s, stack, v);
    }
    // The two halves of a visit, for callers that want to put something between them (the extend kernel issues the
    // loads of the visit's first triangle there): the box tests, which leave the visit's triangle group in the cursor ...
    struct Visit {
        uint32_t child_base, imask, hits;  // first child node, inner-slot mask, RPT_CHILD_HIT_BITS of the children hit
    };
    RPT_D Visit test_children(const WideScene& s) {
        const uint4* node = s.nodes + 5u * (size_t)next;
#if RPT_NODE_INDEXED_LOADS && defined(__CUDACC__)
        const uint4 n0 = __ldg(node), n1 = __ldg(node + 1);
        const uint2* words = reinterpret_cast<const uint2*>(s.nodes);  // (32-bit word indices: the address arithmetic stays IMAD / IMAD.WIDE)
        const uint32_t w10 = 10u * next;
        const uint2 near_wx = __ldg(words + (w10 + ray.near_x)), far_wx = __ldg(words + (w10 + 11u - ray.near_x));
        const uint2 near_wy = __ldg(words + (w10 + ray.near_y)), far_wy = __ldg(words + (w10 + 13u - ray.near_y));
        const uint2 near_wz = __ldg(words + (w10 + ray.near_z)), far_wz = __ldg(words + (w10 + 15u - ray.near_z));
#else
        const uint4 n0 = __ldg(node), n1 = __ldg(node + 1), n2 = __ldg(node + 2), n3 = __ldg(node + 3), n4 = __ldg(node + 4);
#endif

        const f3 p = mk3(as_float(n0.x), as_float(n0.y), as_float(n0.z));
        const f3 cell = mk3(as_float(n0.w), as_float(n1.w << 16), as_float(n1.w & 0xFFFF0000u));  // powers of two
        const f3 adj = cell * ray.idir;
        // Push the planes out so rounding can never cull a box the exact arithmetic would enter.  p - o, its
        // product with idir and the final FMA are each correctly rounded and idir is within 1 ulp of 1/d, i.e.
        // together off by < 3 * 2^-23 of |p - o| resp. of the box extent (<= 256 cells): pad by 5 * 2^-23 (6e-7) of both.  Folding the -1024 bias of the fp16
        // de-quantisation into the addend rounds it at magnitude 1024 |adj|, i.e. by < 2^-14 |adj|: 512 more
        // cells in the same pad term cover that four times over.  (RPT_PLANES_FP32: bias 32768, rounded by < 2^-9 |adj|: kPadCells.)
        const f3 rel_o = p - ray.o;
        const f3 org = rel_o * ray.idir;
#if RPT_NODE_SHORT_PAD
        // the same pad from the products that exist anyway, (kPadCells |adj| + |org|) * 6e-7 (the two differ by the
        // rounding of org and adj, 2^-24 relative, far below the 5/3 margin of the pad), and the biased addend built once:
        // base -+ pad rounds twice at magnitude <= 1024 |adj| + |org| where the form below rounds once — 2^-13 |adj| of the
        // 2^-11.7 |adj| (512 cells * 6e-7) set aside for it.  Six instructions fewer per visit.
        const f3 apad = mk3(fmaf(kPadCells, fabsf(adj.x), fabsf(org.x)) * 6e-7f, fmaf(kPadCells, fabsf(adj.y), fabsf(org.y)) * 6e-7f,
                            fmaf(kPadCells, fabsf(adj.z), fabsf(org.z)) * 6e-7f);
        const f3 base = mk3(fmaf(adj.x, -kPlaneBias, org.x), fmaf(adj.y, -kPlaneBias, org.y), fmaf(adj.z, -kPlaneBias, org.z));
        const f3 org_near = base - apad;
        const f3 org_far = base + apad;
#else
        const f3 apad = mk3(fabsf(fmaf(kPadCells, cell.x, fabsf(rel_o.x)) * ray.idir.x) * 6e-7f, fabsf(fmaf(kPadCells, cell.y, fabsf(rel_o.y)) * ray.idir.y) * 6e-7f,
                            fabsf(fmaf(kPadCells, cell.z, fabsf(rel_o.z)) * ray.idir.z) * 6e-7f);
        const f3 org_near = mk3(fmaf(adj.x, -kPlaneBias, org.x - apad.x), fmaf(adj.y, -kPlaneBias, org.y - apad.y), fmaf(adj.z, -kPlaneBias, org.z - apad.z));
        const f3 org_far = mk3(fmaf(adj.x, -kPlaneBias, org.x + apad.x), fmaf(adj.y, -kPlaneBias, org.y + apad.y), fmaf(adj.z, -kPlaneBias, org.z + apad.z));
#endif

#if RPT_NODE_INDEXED_LOADS && defined(__CUDACC__)
        uint32_t h = test_four<0u>(s.half_1024_bytes, near_wx.x, near_wy.x, near_wz.x, far_wx.x, far_wy.x, far_wz.x, adj, org_near, org_far, best_t);
        h |= test_four<4u>(s.half_1024_bytes, near_wx.y, near_wy.y, near_wz.y, far_wx.y, far_wy.y, far_wz.y, adj, org_near, org_far, best_t);
#else
        const bool nx = (ray.oct_inv & 4u) == 0u, ny = (ray.oct_inv & 2u) == 0u, nz = (ray.oct_inv & 1u) == 0u;  // d.x < 0, ...
        // children 0..3 and 4..7: near/far byte words per axis depend on the ray's sign
        uint32_t h = test_four<0u>(s.half_1024_bytes, nx ? n3.z : n2.x, ny ? n4.x : n2.z, nz ? n4.z : n3.x, nx ? n2.x : n3.z, ny ? n2.z : n4.x, nz ? n3.x : n4.z, adj,
                                   org_near, org_far, best_t);
        h |= test_four<4u>(s.half_1024_bytes, nx ? n3.w : n2.y, ny ? n4.y : n2.w, nz ? n4.w : n3.y, nx ? n2.y : n3.w, ny ? n2.w : n4.y, nz ? n3.y : n4.w, adj,
                           org_near, org_far, best_t);
#endif
        const uint32_t imask = n1.z >> 24;
        tgroup.x = n1.y;
        tvalid = n1.z & 0x00FFFFFFu;
        tgroup.y = h & tvalid;
        return Visit{n1.x, imask, h};
    }
    // ... and the choice of the node after this one: its nearest hit child, else the nearest pending sibling, else the stack
    template <class Stack>
    RPT_D void select_next(const WideScene& s, Stack& stack, const Visit& v) {
        const uint32_t imask = v.imask, h = v.hits;
        const uint32_t child_hits = stack.permute(ray.oct_inv, (h >> 24) & imask) << 24;
#if !defined(__CUDACC__)
        last_child_hits = child_hits;
#endif
        if (child_hits != 0u) {
            if (ngroup.y > 0x00FFFFFFu) stack.push(ngroup);
            ngroup = make_uint2(v.child_base, child_hits | imask);
        } else if (ngroup.y <= 0x00FFFFFFu && !stack.empty()) {
            ngroup = stack.pop();
        }
        if (ngroup.y > 0x00FFFFFFu) {
            const int bit = highest_bit(ngroup.y);
            const uint32_t slot = ((uint32_t)bit - 24u) ^ ray.oct_inv;
            next = ngroup.x + (uint32_t)popcount(ngroup.y & ~(0xFFFFFFFFu << slot) & 0xFFu);
            ngroup.y &= ~(1u << bit);
            if (kPrefetchNextNode) {
                const char* line = reinterpret_cast<const char*>(s.nodes + 5u * (size_t)next);
                prefetch_l1(line);
                prefetch_l1(line + 64);  // an 80-byte node straddles two 128-byte lines half of the time
            }
        } else {
            next = kNoNode;
        }
    }
    // Tests the next pending triangle; returns true when an any-hit query is decided.
    RPT_D bool test_triangle(const WideScene& s) { return test_next<false>(s); }
    template <bool ORDER_FREE>
    RPT_D bool test_next(const WideScene& s) {
        const int k = highest_bit(tgroup.y);
        tgroup.y &= ~(1u << k);
        return test_one<ORDER_FREE>(s, tgroup.x + (uint32_t)popcount(tvalid & ~(0xFFFFFFFFu << k)));
    }
    // Tests triangle `ti` (wide order); returns true when an any-hit query is decided.
    // ORDER_FREE: among hits with bit-equal t the smallest wide index wins, so the result does not depend on the order
    // in which a ray's candidate triangles are tested (deferred tests; the in-order loop keeps "first found wins").
    template <bool ORDER_FREE>
    RPT_D bool test_one(const WideScene& s, uint32_t ti) {
        const float4* rec = s.tri_pos + 3u * (size_t)ti;
        const float4 a = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2);
        return test_record<ORDER_FREE>(ti, a, e1, e2, true);
    }
    // The first pending triangle of the visit: which one it is (wide index; 0 — some valid record — when there is
    // none), removed from the group.  For callers that load the record themselves and then call test_record.
    RPT_D uint32_t take_first_triangle(bool& any) {
        any = tgroup.y != 0u;
        const int k = highest_bit(tgroup.y | 1u);
        const uint32_t ti = tgroup.x + (uint32_t)popcount(tvalid & ~(0xFFFFFFFFu << k));
        tgroup.y &= ~(1u << k);  // (k = 0 when there is none: clearing bit 0 of an empty group changes nothing)
        return any ? ti : 0u;
    }
    // The test proper on a record that is already in registers; `enabled` = false discards the result.
    template <bool ORDER_FREE>
    RPT_D bool test_record(uint32_t ti, float4 a, float4 e1, float4 e2, bool enabled) {
        float t;
        bool back;
        // accept 0.001 < t < best so far (nearest) resp. t <= max_t and t < 1e6 (any), intersection.rs:195
        if (ray_triangle(ray.o, ray.d, mk3(a.x, a.y, a.z), mk3(e1.x, e1.y, e1.z), mk3(e2.x, e2.y, e2.z), t, back) && enabled && t > 0.001f &&
            (NEAREST ? (t < best_t || (ORDER_FREE && t == best_t && hit_tri != kNoNode && ti < (hit_tri & 0x7FFFFFFFu))) : (t <= best_t && t < 1000000.0f))) {
            hit_t = t;
            hit_tri = ti | (back ? 0x80000000u : 0u);
            if (!NEAREST) return true;
            best_t = t;
        }
        return false;
    }
};

// Nearest hit (NEAREST = true, `t < best`) or any hit with t <= max_t (NEAREST = false), with the
// reference's acceptance window t > 0.001 (intersection.rs:195).  Stack: push(uint2), pop(),
// empty(), clear(), permute(octant, child set).
template <bool NEAREST, class Stack>
RPT_D WideHit wide_intersect(const WideScene& s, f3 ro, f3 rd, float max_t, Stack& stack) {
    WideCursor<NEAREST> c;
    c.begin(ro, rd, max_t);
    while (c.has_nodes()) {
        c.visit_node(s, stack);
        while (c.has_triangles())
            if (c.test_triangle(s)) return c.result();
    }
    return c.result();
}

}  // namespace rpt
