// vec.cuh — small fp32 vector helpers for device code.
//
// Evaluation order follows glam 0.22's scalar Vec3 (dot = (x*x' + y*y') + z*z', cross as below,
// normalize = v * (1/len), lerp = a + (b-a)*s) so that translation units compiled with
// -fmad=false (ray generation, traversal, ray/triangle test) reproduce the reference's CPU path
// bit for bit; in units compiled with FMA contraction the same code is merely ulp-close.
// Functions are RPT_HD so tests/cpu_harness can compile the traversal code on the host.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RPT_HD __host__ __device__ __forceinline__
#define RPT_D __device__ __forceinline__
#else
#define RPT_HD inline
#define RPT_D inline
#endif

namespace rpt {

struct f3 {
    float x, y, z;
};

RPT_HD f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
RPT_HD f3 splat3(float s) { return f3{s, s, s}; }
RPT_HD f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
RPT_HD f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
RPT_HD f3 operator*(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
RPT_HD f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
RPT_HD f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
RPT_HD f3 operator/(f3 a, float s) { return f3{a.x / s, a.y / s, a.z / s}; }
RPT_HD f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
RPT_HD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RPT_HD f3 cross(f3 a, f3 b) { return f3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
RPT_HD float length(f3 a) { return sqrtf(dot(a, a)); }
RPT_HD f3 normalize(f3 a) { return a * (1.0f / length(a)); }
RPT_HD f3 lerp3(f3 a, f3 b, float s) { return a + ((b - a) * s); }
RPT_HD bool finite3(f3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }
RPT_HD bool zero3(f3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
RPT_HD float max_element(f3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
RPT_HD f3 mask_nan(f3 v) { return finite3(v) ? v : splat3(0.0f); }  // kernels/src/util.rs:271-277

#if defined(__CUDACC__)
RPT_HD f3 xyz(float4 v) { return f3{v.x, v.y, v.z}; }
RPT_HD float4 mk4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
#endif

// x / 255 for an integer 0 <= x <= 255, correctly rounded — the very value of the IEEE division the CPU path does
// (src/asset.rs:266-273; all 256 inputs are checked in tests/test_wide_bvh_cpu.py) — without the division's
// special-case machinery: one Newton step on x * (1/255).
RPT_HD float unorm8(uint32_t x) {
    const float xf = (float)x, r = 1.0f / 255.0f;
    const float q = xf * r;
    return fmaf(fmaf(-q, 255.0f, xf), r, q);
}
// size - 1 for a power-of-two image size (coordinates wrap with a mask), else 0
RPT_HD uint32_t pow2_mask(uint32_t size) { return (size & (size - 1u)) == 0u ? size - 1u : 0u; }

constexpr float kPi = 3.14159265358979323846f;
constexpr float kEps = 0.001f;  // kernels/src/util.rs:5

}  // namespace rpt
