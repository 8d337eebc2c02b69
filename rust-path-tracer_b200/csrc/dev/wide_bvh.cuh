// wide_bvh.cuh — traversal of the 8-wide compressed BVH (layout: ../wide_bvh.h).
//
// One ray per lane.  Each node visit is five 16-byte loads; the eight child boxes are
// de-quantised with one PRMT + FADD + FFMA per plane and tested against the current best t; hits
// are scattered into a 32-bit mask whose bit order (slot ^ ray octant) is the front-to-back
// visiting order, so there is no per-node sort.  The traversal stack holds (base index, hit mask)
// pairs and lives wherever `Stack` puts it — a per-warp shared-memory slab on the device.
//
// Box tests are CONSERVATIVE (quantised outwards at build time, planes pushed out by a few ulps
// of the coordinates here) and may use FMA; the ray/triangle test is exact.cuh's, which evaluates
// the reference's operations in the reference's order.  Compile with -fmad=false.
#pragma once

#include <string.h>

#include "exact.cuh"

namespace rpt {

#if defined(__CUDACC__)
RPT_D uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// 0xFF in every byte whose bit 7 is set.  (__byte_perm masks selector nibbles to 3 bits, so the
// sign-replicating form of PRMT is only reachable through PTX.)
RPT_D uint32_t sign_extend_s8x4(uint32_t v) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0, 0xBA98;" : "=r"(r) : "r"(v));
    return r;
}
RPT_D int highest_bit(uint32_t v) { return 31 - __clz((int)v); }
RPT_D int popcount(uint32_t v) { return __popc(v); }
RPT_D float as_float(uint32_t v) { return __uint_as_float(v); }
RPT_D uint32_t as_uint(float v) { return __float_as_uint(v); }
#else
inline uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t both = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        const uint32_t byte = (uint32_t)(both >> (8 * (sel & 7))) & 0xFF;  // like __byte_perm: 3-bit selectors
        r |= byte << (8 * i);
    }
    return r;
}
inline uint32_t sign_extend_s8x4(uint32_t v) { return ((v >> 7) & 0x01010101u) * 0xFFu; }
inline int highest_bit(uint32_t v) { return 31 - __builtin_clz(v); }
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
};

struct WideHit {
    float t;            // 1e6 if no hit (kernels/src/intersection.rs:68)
    uint32_t triangle;  // index in WIDE order
    bool hit, backface;
};

// Ray constants hoisted out of the node loop.
struct WideRay {
    f3 o, d;
    f3 idir;        // 1 / d with |d| clamped away from 0
    f3 pad_scale;   // |idir| * 2^-21: conservative slack per unit of coordinate magnitude
    uint32_t oct_inv4;  // (dx>=0 ? 4 : 0 | dy>=0 ? 2 : 0 | dz>=0 ? 1 : 0) * 0x01010101
};

RPT_D float safe_rcp(float d) { return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }

RPT_D WideRay make_wide_ray(f3 o, f3 d) {
    WideRay r;
    r.o = o;
    r.d = d;
    r.idir = mk3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    r.pad_scale = mk3(fabsf(r.idir.x) * 4.8e-7f, fabsf(r.idir.y) * 4.8e-7f, fabsf(r.idir.z) * 4.8e-7f);
    const uint32_t oct = (d.x < 0.0f ? 0u : 4u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 1u);
    r.oct_inv4 = oct * 0x01010101u;
    return r;
}

// Test four children whose bytes sit in (near_x, near_y, near_z, far_x, far_y, far_z); returns
// their contribution to the hit mask.
RPT_D uint32_t test_four(uint32_t meta4, uint32_t oct_inv4, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz,
                         f3 adj, f3 org_near, f3 org_far, float best_t) {
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);  // 0xFF per inner byte
    const uint32_t bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    uint32_t mask = 0;
#define RPT_CHILD(J, SEL)                                                                                     \
    {                                                                                                         \
        const float tnx = fmaf(as_float(byte_perm(nx, 0x4B000000u, SEL)) - 8388608.0f, adj.x, org_near.x);    \
        const float tny = fmaf(as_float(byte_perm(ny, 0x4B000000u, SEL)) - 8388608.0f, adj.y, org_near.y);    \
        const float tnz = fmaf(as_float(byte_perm(nz, 0x4B000000u, SEL)) - 8388608.0f, adj.z, org_near.z);    \
        const float tfx = fmaf(as_float(byte_perm(fx, 0x4B000000u, SEL)) - 8388608.0f, adj.x, org_far.x);     \
        const float tfy = fmaf(as_float(byte_perm(fy, 0x4B000000u, SEL)) - 8388608.0f, adj.y, org_far.y);     \
        const float tfz = fmaf(as_float(byte_perm(fz, 0x4B000000u, SEL)) - 8388608.0f, adj.z, org_far.z);     \
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));                                            \
        const float tf = fminf(fminf(tfx, tfy), fminf(tfz, best_t));                                          \
        if (tn <= tf) mask |= ((child_bits4 >> (8 * J)) & 0xFFu) << ((bit_index4 >> (8 * J)) & 0xFFu);        \
    }
    RPT_CHILD(0, 0x7650u)
    RPT_CHILD(1, 0x7651u)
    RPT_CHILD(2, 0x7652u)
    RPT_CHILD(3, 0x7653u)
#undef RPT_CHILD
    return mask;
}

// ---- traversal as a resumable cursor -------------------------------------------------------
// The per-ray state between steps.  A step is either one node visit (`visit_node`) or one
// ray/triangle test (`test_triangle`); `advance` pops the next group.  The device kernels
// interleave these steps across the lanes of a warp (triangle postponing, ray refill); the
// plain loop `wide_intersect` below composes them for one ray.
//   ngroup = (first child node, hits << 24 | inner-slot mask): child nodes still to visit
//   tgroup = (first triangle, 24-bit mask): triangles still to test
// Stack entries are groups of either kind; a popped entry without node hits is a triangle group.
template <bool NEAREST>
struct WideCursor {
    WideRay ray;
    float best_t;   // current culling bound: best hit so far (nearest) or max_t (any)
    float max_t;
    WideHit res;
    uint2 ngroup, tgroup;

    RPT_D void begin(f3 ro, f3 rd, float max_t_) {
        res = WideHit{1000000.0f, 0u, false, false};
        max_t = max_t_;
        ray = make_wide_ray(ro, rd);
        best_t = NEAREST ? res.t : fminf(max_t_, res.t);
        ngroup = make_uint2(0u, 0x80000000u);  // the root, as "child bit 31 of a virtual parent"
        tgroup = make_uint2(0u, 0u);
        // a ray with a non-finite component hits nothing in the reference either (every slab test
        // compares false); without this the NaN-ignoring min/max would visit every node
        if (!(finite3(ro) && finite3(rd))) ngroup.y = 0u;
    }
    RPT_D bool has_nodes() const { return ngroup.y > 0x00FFFFFFu; }
    RPT_D bool has_triangles() const { return tgroup.y != 0u; }

    // Pops the nearest pending child of ngroup and tests its eight children.
    template <class Stack>
    RPT_D void visit_node(const WideScene& s, Stack& stack) {
        const uint32_t oct_inv = ray.oct_inv4 & 7u;
        const uint32_t hits = ngroup.y;
        const int bit = highest_bit(hits);
        ngroup.y &= ~(1u << bit);
        if (ngroup.y > 0x00FFFFFFu) stack.push(ngroup);
        const uint32_t slot = ((uint32_t)bit - 24u) ^ oct_inv;
        const uint32_t rel = (uint32_t)popcount(hits & ~(0xFFFFFFFFu << slot) & 0xFFu);
        const uint4* node = s.nodes + 5u * (size_t)(ngroup.x + rel);
        const uint4 n0 = __ldg(node), n1 = __ldg(node + 1), n2 = __ldg(node + 2), n3 = __ldg(node + 3), n4 = __ldg(node + 4);

        const f3 p = mk3(as_float(n0.x), as_float(n0.y), as_float(n0.z));
        const f3 cell = mk3(as_float((n0.w & 0xFFu) << 23), as_float(((n0.w >> 8) & 0xFFu) << 23), as_float(((n0.w >> 16) & 0xFFu) << 23));
        const f3 adj = cell * ray.idir;
        // Push the planes out by a few ulps so rounding can never cull a box the exact arithmetic
        // would enter: p - o, its product with idir and the FMA below are each correctly rounded,
        // i.e. off by < 2^-23 of |p - o| resp. of the box extent (<= 256 cells); pad by 2^-21 of both.
        const f3 rel_o = p - ray.o;
        const f3 apad = mk3(fmaf(256.0f, cell.x, fabsf(rel_o.x)) * ray.pad_scale.x, fmaf(256.0f, cell.y, fabsf(rel_o.y)) * ray.pad_scale.y,
                            fmaf(256.0f, cell.z, fabsf(rel_o.z)) * ray.pad_scale.z);
        const f3 org = rel_o * ray.idir;
        const f3 org_near = org - apad, org_far = org + apad;

        const bool nx = ray.d.x < 0.0f, ny = ray.d.y < 0.0f, nz = ray.d.z < 0.0f;
        // children 0..3 and 4..7: near/far byte words per axis depend on the ray's sign
        uint32_t hitmask = test_four(n1.z, ray.oct_inv4, nx ? n3.z : n2.x, ny ? n4.x : n2.z, nz ? n4.z : n3.x, nx ? n2.x : n3.z,
                                     ny ? n2.z : n4.x, nz ? n3.x : n4.z, adj, org_near, org_far, best_t);
        hitmask |= test_four(n1.w, ray.oct_inv4, nx ? n3.w : n2.y, ny ? n4.y : n2.w, nz ? n4.w : n3.y, nx ? n2.y : n3.w,
                             ny ? n2.w : n4.y, nz ? n3.y : n4.w, adj, org_near, org_far, best_t);
        ngroup.x = n1.x;
        ngroup.y = (hitmask & 0xFF000000u) | (n0.w >> 24);
        tgroup.x = n1.y;
        tgroup.y = hitmask & 0x00FFFFFFu;
    }
    // ngroup holds no node hits: whatever it holds is a (postponed) triangle group.
    RPT_D void take_triangle_group() {
        tgroup = ngroup;
        ngroup = make_uint2(0u, 0u);
    }
    // Tests the next pending triangle; returns true when an any-hit query is decided.
    RPT_D bool test_triangle(const WideScene& s) {
        const int k = highest_bit(tgroup.y);
        tgroup.y &= ~(1u << k);
        const uint32_t ti = tgroup.x + (uint32_t)k;
        const float4* rec = s.tri_pos + 3u * (size_t)ti;
        const float4 a = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2);
        float t;
        bool back;
        if (ray_triangle(ray.o, ray.d, mk3(a.x, a.y, a.z), mk3(e1.x, e1.y, e1.z), mk3(e2.x, e2.y, e2.z), t, back) && t > 0.001f && t < res.t &&
            (NEAREST || t <= max_t)) {
            res.t = t;
            res.triangle = ti;
            res.hit = true;
            res.backface = back;
            if (!NEAREST) return true;
            best_t = t;
        }
        return false;
    }
    // After the triangles: make sure ngroup holds work, popping the stack; false when the ray is done.
    template <class Stack>
    RPT_D bool advance(Stack& stack) {
        if (ngroup.y <= 0x00FFFFFFu) {
            if (stack.empty()) return false;
            ngroup = stack.pop();
        }
        return true;
    }
};

// Nearest hit (NEAREST = true, `t < best`) or any hit with t <= max_t (NEAREST = false), with the
// reference's acceptance window t > 0.001 (intersection.rs:195).  Stack: push(uint2), pop(),
// empty().
template <bool NEAREST, class Stack>
RPT_D WideHit wide_intersect(const WideScene& s, f3 ro, f3 rd, float max_t, Stack& stack) {
    WideCursor<NEAREST> c;
    c.begin(ro, rd, max_t);
    if (!c.has_nodes()) return c.res;
    for (;;) {
        if (c.has_nodes()) c.visit_node(s, stack);
        else c.take_triangle_group();
        while (c.has_triangles())
            if (c.test_triangle(s)) return c.res;
        if (!c.advance(stack)) break;
    }
    return c.res;
}

}  // namespace rpt
