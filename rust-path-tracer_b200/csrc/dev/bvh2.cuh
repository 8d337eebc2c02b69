// bvh2.cuh — the reference's ordered binary-BVH traversal, restated for the megakernel arm.
//
// Same node order, same slab test with true divisions, same child ordering and the same
// strict `t < best` acceptance as kernels/src/intersection.rs:104-122,169-234, reading the
// reference's own buffers (BVHNode 32 B, UVec4 index, PerVertexData 64 B).  Because nothing is
// re-ordered, exact-t ties resolve to the same triangle as on the CPU path.  Requires -fmad=false.
#pragma once

#include "exact.cuh"

namespace rpt {

struct Hit {
    float t;            // 1e6 when nothing was hit (intersection.rs:68)
    uint32_t triangle;  // index into the index buffer
    bool hit, backface;
};

struct Bvh2Scene {
    const float4* nodes;     // 2 float4 per node: (min.xyz, count bits), (max.xyz, left/first bits)
    const uint4* triangles;  // (i0, i1, i2, material)
    const float4* vertices;  // PerVertexData as 4 float4: vertex, normal, tangent, (uv0, uv1)
};

// intersection.rs:104-122
RPT_D float slab_test(float4 bmin, float4 bmax, f3 ro, f3 rd, float best_t) {
    const float tx1 = (bmin.x - ro.x) / rd.x, tx2 = (bmax.x - ro.x) / rd.x;
    float tmin = fminf(tx1, tx2), tmax = fmaxf(tx1, tx2);
    const float ty1 = (bmin.y - ro.y) / rd.y, ty2 = (bmax.y - ro.y) / rd.y;
    tmin = fmaxf(tmin, fminf(ty1, ty2));
    tmax = fminf(tmax, fmaxf(ty1, ty2));
    const float tz1 = (bmin.z - ro.z) / rd.z, tz2 = (bmax.z - ro.z) / rd.z;
    tmin = fmaxf(tmin, fminf(tz1, tz2));
    tmax = fminf(tmax, fmaxf(tz1, tz2));
    return (tmax >= tmin && tmax > 0.0f && tmin < best_t) ? tmin : INFINITY;
}

template <bool NEAREST>
RPT_D Hit bvh2_intersect(const Bvh2Scene& s, f3 ro, f3 rd, float max_t) {
    uint32_t stack[32];  // FixedVec<usize, 32>, kernels/src/vec.rs
    int sp = 0;
    stack[sp++] = 0;
    Hit res{1000000.0f, 0u, false, false};
    while (sp > 0) {
        const uint32_t ni = stack[--sp];
        const float4 nmin = __ldg(s.nodes + 2 * ni), nmax = __ldg(s.nodes + 2 * ni + 1);
        const uint32_t count = __float_as_uint(nmin.w), first = __float_as_uint(nmax.w);
        if (count > 0) {
            for (uint32_t i = 0; i < count; ++i) {
                const uint32_t ti = first + i;
                const uint4 tri = __ldg(s.triangles + ti);
                const f3 a = xyz(__ldg(s.vertices + 4 * tri.x)), b = xyz(__ldg(s.vertices + 4 * tri.y)), c = xyz(__ldg(s.vertices + 4 * tri.z));
                float t = 0.0f;
                bool back = false;
                if (ray_triangle(ro, rd, a, b - a, c - a, t, back) && t > 0.001f && t < res.t && (NEAREST || t <= max_t)) {
                    res.t = fminf(res.t, t);
                    res.triangle = ti;
                    res.hit = true;
                    res.backface = back;
                    if (!NEAREST) return res;
                }
            }
        } else {
            uint32_t near_i = first, far_i = first + 1;
            float near_d = slab_test(__ldg(s.nodes + 2 * near_i), __ldg(s.nodes + 2 * near_i + 1), ro, rd, res.t);
            float far_d = slab_test(__ldg(s.nodes + 2 * far_i), __ldg(s.nodes + 2 * far_i + 1), ro, rd, res.t);
            if (near_d > far_d) {
                const uint32_t ti = near_i; near_i = far_i; far_i = ti;
                const float td = near_d; near_d = far_d; far_d = td;
            }
            if (isinf(near_d)) continue;
            if (isfinite(far_d) && sp < 31) stack[sp++] = far_i;
            if (sp < 32) stack[sp++] = near_i;
        }
    }
    return res;
}

}  // namespace rpt
