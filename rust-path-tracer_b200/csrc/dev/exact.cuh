// exact.cuh — the ID-critical arithmetic: R-sequence, camera ray, ray/triangle test.
//
// INCLUDE ONLY FROM TRANSLATION UNITS COMPILED WITH -fmad=false (build.py: *_nofma.cu; host
// harness: -ffp-contract=off).  With contraction off these functions evaluate the same IEEE
// fp32 operations in the same order as the reference's CPU path, so primary-hit triangle ids
// and hit distances are bit-identical to it.
#pragma once

#include "vec.cuh"

namespace rpt {

// ---- R-sequence, kernels/src/rng.rs:19-62 ------------------------------------------------
// lds(n, dim, offset) = f32(PRIME[dim] * (n + offset) mod 2^32) * 2^-32 with `dim`
// pre-incremented (the first number drawn uses PRIME[1]).  Random access in (n, dim): a path
// only has to carry its 5-bit dimension cursor.
#if defined(__CUDACC__)
static __constant__ uint32_t c_lds_primes[32] = {
#else
static const uint32_t c_lds_primes[32] = {
#endif
    0x6a09e667u, 0xbb67ae84u, 0x3c6ef372u, 0xa54ff539u, 0x510e527fu, 0x9b05688au, 0x1f83d9abu, 0x5be0cd18u,
    0xcbbb9d5cu, 0x629a2929u, 0x91590159u, 0x452fecd8u, 0x67332667u, 0x8eb44a86u, 0xdb0c2e0bu, 0x47b5481du,
    0xae5f9155u, 0xcf6c85d1u, 0x2f73477du, 0x6d1826cau, 0x8b43d455u, 0xe360b595u, 0x1c456002u, 0x6f196330u,
    0xd94ebeafu, 0x9cc4a611u, 0x261dc1f2u, 0x5815a7bdu, 0x70b7ed67u, 0xa1513c68u, 0x44f93634u, 0x720dcdfcu};

struct Rng {
    uint32_t key;  // n + offset (wrapping): sample index + per-pixel offset
    uint32_t dim;  // last dimension used
    RPT_D float next() {
        dim += 1;
        // `u32 as f32` rounds to nearest even; can produce exactly 1.0 (rng.rs:31)
        return (float)(c_lds_primes[dim & 31u] * key) * (1.0f / 4294967296.0f);
    }
    // One of the numbers this path draws is exactly 1.0: `u32 as f32` rounds 0xFFFFFF80 and above up to 2^32, i.e. the
    // NEGATED product PRIME[d] * (-key) mod 2^32 is at most 128 (0 — the number 0.0 — is let through as well: harmless).
    // All 31 dimensions, drawn or not, so that the loop unrolls into one multiply per dimension (the prime is a
    // constant-bank operand) and half a three-input minimum; ~50 instructions for the lanes that ask.
    RPT_D bool draws_a_one() const {
        const uint32_t negated = 0u - key;
        uint32_t least = 0xFFFFFFFFu;
#if defined(__CUDACC__)
#pragma unroll
#endif
        for (uint32_t d = 1u; d < 32u; ++d) {
            const uint32_t q = c_lds_primes[d] * negated;
            least = q < least ? q : least;
        }
        return least <= 128u;
    }
};

// ---- camera ray, kernels/src/lib.rs:38-51 -------------------------------------------------
// cam_m is RotY(rot.y) * RotX(rot.x) as three columns, computed on the HOST (rpt_camera_matrix)
// so device libm never touches a primary ray.
struct Camera {
    f3 position;
    f3 c0, c1, c2;
    float width, height, aspect;  // f32(width), f32(height), f32(height) / f32(width)
};

RPT_D void camera_ray(const Camera& cam, uint32_t px, uint32_t py, Rng& rng, f3& ro, f3& rd) {
    const float jx = rng.next();
    const float jy = rng.next();
    const float sx = (float)px + jx;
    const float sy = (float)py + jy;
    const float ux = (sx / cam.width) * 2.0f - 1.0f;
    float uy = (1.0f - sy / cam.height) * 2.0f - 1.0f;
    uy = uy * cam.aspect;
    const f3 d = normalize(mk3(ux, uy, 1.0f));
    ro = cam.position;
    rd = (cam.c0 * d.x + cam.c1 * d.y) + cam.c2 * d.z;
}

// ---- Moller-Trumbore, kernels/src/intersection.rs:9-54 ------------------------------------
// Takes the precomputed edges e1 = b - a, e2 = c - a (single IEEE subtractions, so computing
// them once on the host changes no bit).  Returns true with t (>= 0) and the back-face flag;
// the caller applies `t > 0.001 && t < best` (intersection.rs:195).
// RPT_TRI_STRAIGHT (build switch): the same operations without the early returns — every value that decides or is
// returned is computed by the identical instruction sequence, the four rejections are combined at the end (a rejected
// triangle's later values, possibly inf / NaN, are discarded).  Straight-line code keeps all three record loads in flight
// together (with the early returns ptxas sinks the load of `a` below the determinant test) at the price of finishing
// tests that could have stopped early — which a warp only profits from when ALL of its active lanes stop.
#ifndef RPT_TRI_STRAIGHT
#define RPT_TRI_STRAIGHT 1
#endif
RPT_HD bool ray_triangle(f3 ro, f3 rd, f3 a, f3 e1, f3 e2, float& t_out, bool& backface) {
    const f3 pv = cross(rd, e2);
    const float det = dot(e1, pv);
    backface = signbit(det);
#if RPT_TRI_STRAIGHT
    {
        const float inv_det = 1.0f / det;
        const f3 tv = ro - a;
        const float u = dot(tv, pv) * inv_det;
        const f3 qv = cross(tv, e1);
        const float v = dot(rd, qv) * inv_det;
        const float t = dot(e2, qv) * inv_det;
        t_out = t;
        const bool reject = (fabsf(det) < 1e-6f) | (u < 0.0f) | (u > 1.0f) | (v < 0.0f) | (u + v > 1.0f) | (t < 0.0f);
        return !reject;
    }
#endif
    if (fabsf(det) < 1e-6f) return false;
    const float inv_det = 1.0f / det;
    const f3 tv = ro - a;
    const float u = dot(tv, pv) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    const f3 qv = cross(tv, e1);
    const float v = dot(rd, qv) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    const float t = dot(e2, qv) * inv_det;
    if (t < 0.0f) return false;
    t_out = t;
    return true;
}

}  // namespace rpt
