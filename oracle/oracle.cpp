// oracle.cpp — CPU restatement of the reference's per-pixel-sample tracing loop.
//
// *** TEST INFRASTRUCTURE ONLY. ***  This file is the checker the CUDA path is compared with.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
// build, load or call it; nothing under rust-path-tracer_b200/ links or imports it.
//
// PARITY PINNING.  The reference (Rust, rust-gpu nightly + assimp + wgpu) cannot be compiled in
// this image, and it ships no golden vectors for this path except ONE known answer:
// tests/correctness_tests.rs:14-33 — FurnaceTest.glb, 128x128, 32 spp, pixel (65,75), every
// channel ^(1/2.2) == 0.8 +- 0.02, with nee = 0 and nee = MIS.  tests/test_oracle_furnace.py
// pins this oracle to that value.  Everything finer (per-pixel images, primary-hit ids) is
// "parity unpinned" at the glam / libm / assimp boundary: the restatement below follows the
// cited lines op for op in IEEE fp32 (build with -O2 -ffp-contract=off; glibc libm is what
// Rust's f32::sin/cos/... call on Linux), and third-party arithmetic is restated from the
// published behaviour of glam 0.22 (Vec3 scalar: dot = (x*x'+y*y')+z*z', normalize = v*(1/len),
// lerp = a+(b-a)*s, Mat3*v = (c0*v.x+c1*v.y)+c2*v.z) and compiler-rt's powi.
// Beyond that one pixel, tests/test_known_answers.py holds this file to answers that do not come from it: closed
// forms and numpy quadrature of the reference's own formulas (diffuse-only and mirror-limit spheres, the lat-long sky
// lookup), and laws its estimators obey by construction (an emitter seen from inside == a constant sky, constant
// textures == constant factors, NEE off / MIS / direct-only agree, Russian roulette is unbiased) — tests/known_answers.py.
//
// Layout of this file (reference file it follows):
//   vector helpers           glam 0.22 semantics
//   rng                      kernels/src/rng.rs:19-62
//   intersection             kernels/src/intersection.rs:9-54, 56-74, 104-122, 169-234; vec.rs
//   shading utilities        kernels/src/util.rs (live subset)
//   texture fetch            shared_structs/src/image_polyfill.rs:32-55
//   PBR bsdf                 kernels/src/bsdf.rs:179-387
//   next-event estimation    kernels/src/light_pick.rs:8-23, 30-87, 89-199
//   sky                      kernels/src/skybox.rs
//   trace_pixel              kernels/src/lib.rs:21-186
//   driver loop              src/trace.rs:273-308 (row-parallel, OpenMP standing in for rayon)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/rpt_shared_structs.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------ glam-style vectors
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 v3(float x, float y, float z) { return {x, y, z}; }
inline V3 splat(float s) { return {s, s, s}; }
inline V3 xyz(const float* p) { return {p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) { return a * (1.0f / length(a)); }
inline V3 lerp(V3 a, V3 b, float s) { return a + ((b - a) * s); }
inline bool is_finite(V3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }
inline bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
inline float max_element(V3 a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }
inline V3 powf3(V3 a, float e) { return {std::pow(a.x, e), std::pow(a.y, e), std::pow(a.z, e)}; }
inline V3 exp3(V3 a) { return {std::exp(a.x), std::exp(a.y), std::exp(a.z)}; }

inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator*(V2 a, V2 b) { return {a.x * b.x, a.y * b.y}; }
inline V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }
inline V2 operator*(float s, V2 a) { return {s * a.x, s * a.y}; }

inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator-(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 lerp(V4 a, V4 b, float s) { return a + ((b - a) * s); }

// 3x3 matrix as columns; M * v = (c0*v.x + c1*v.y) + c2*v.z
struct M3 { V3 c0, c1, c2; };
inline V3 operator*(const M3& m, V3 v) { return (m.c0 * v.x + m.c1 * v.y) + m.c2 * v.z; }
inline M3 rotation_y(float a) { float s = std::sin(a), c = std::cos(a); return {v3(c, 0, -s), v3(0, 1, 0), v3(s, 0, c)}; }
inline M3 rotation_x(float a) { float s = std::sin(a), c = std::cos(a); return {v3(1, 0, 0), v3(0, c, s), v3(0, -s, c)}; }
inline M3 operator*(const M3& a, const M3& b) { return {a * b.c0, a * b.c1, a * b.c2}; }

// compiler-rt __powisf2 / LLVM constant-exponent expansion
inline float powi2(float x) { return x * x; }
inline float powi5(float x) { float x2 = x * x; float x4 = x2 * x2; return x4 * x; }

constexpr float kPi = 3.14159265358979323846f;
constexpr float kEps = 0.001f;  // util.rs:5

// Rust `f32 as usize`: saturating, NaN -> 0
inline size_t f32_as_usize(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)f;
}
// Rust `f32 as i32`: saturating, NaN -> 0
inline int32_t f32_as_i32(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

// ------------------------------------------------------------------ scene view + counters
struct Image {
    const V4* texels;
    uint32_t width, height;
};

struct Counters {
    uint64_t nearest_rays = 0, any_rays = 0, nodes_popped = 0, boxes_tested = 0, tris_tested = 0;
    uint64_t boxes_tested_any = 0, tris_tested_any = 0;  // the share of the two totals spent in intersect_any
    uint64_t stack_overflows = 0, light_index_clamped = 0;
    void add(const Counters& o) {
        nearest_rays += o.nearest_rays; any_rays += o.any_rays; nodes_popped += o.nodes_popped;
        boxes_tested += o.boxes_tested; tris_tested += o.tris_tested;
        boxes_tested_any += o.boxes_tested_any; tris_tested_any += o.tris_tested_any;
        stack_overflows += o.stack_overflows; light_index_clamped += o.light_index_clamped;
    }
};

struct Scene {
    const RptPerVertexData* verts;
    const uint32_t* tris;  // 4 per triangle
    const RptBVHNode* nodes;
    const RptMaterialData* mats;
    const RptLightPickEntry* lights;
    uint32_t nlights;
    Image atlas, sky;
};

// ------------------------------------------------------------------ rng.rs:19-62
const uint32_t kLdsPrimes[32] = {
    0x6a09e667u, 0xbb67ae84u, 0x3c6ef372u, 0xa54ff539u, 0x510e527fu, 0x9b05688au, 0x1f83d9abu, 0x5be0cd18u,
    0xcbbb9d5cu, 0x629a2929u, 0x91590159u, 0x452fecd8u, 0x67332667u, 0x8eb44a86u, 0xdb0c2e0bu, 0x47b5481du,
    0xae5f9155u, 0xcf6c85d1u, 0x2f73477du, 0x6d1826cau, 0x8b43d455u, 0xe360b595u, 0x1c456002u, 0x6f196330u,
    0xd94ebeafu, 0x9cc4a611u, 0x261dc1f2u, 0x5815a7bdu, 0x70b7ed67u, 0xa1513c68u, 0x44f93634u, 0x720dcdfcu};

struct Rng {
    uint32_t n, offset;
    uint32_t dimension = 0;
    bool exhausted = false;
    float r1() {  // rng.rs:51-54 with lds() of :29-32; the dimension is incremented BEFORE use
        dimension += 1;
        if (dimension >= 32) { exhausted = true; return 0.0f; }  // the Rust code panics here
        return (float)(uint32_t)(kLdsPrimes[dimension] * (n + offset)) * (1.0f / 4294967296.0f);
    }
    V2 r2() { float a = r1(); float b = r1(); return {a, b}; }
    V3 r3() { float a = r1(); float b = r1(); float c = r1(); return {a, b, c}; }
    // (test switch only, see oracle_set_retire_dead_paths) one of the numbers still to be drawn is exactly 1.0:
    // `u32 as f32` rounds 0xFFFFFF80 and above up to 2^32
    // (all 31 dimensions, drawn or not — the backend's form of the guard, which is cheaper to evaluate that way)
    bool draws_a_one() const {
        for (uint32_t d = 1; d < 32; ++d)
            if ((uint32_t)(kLdsPrimes[d] * (n + offset)) >= 0xFFFFFF80u) return true;
        return false;
    }
};

// ------------------------------------------------------------------ intersection.rs
struct TraceResult {  // intersection.rs:56-74
    uint32_t tri[4] = {0, 0, 0, 0};
    uint32_t triangle_index = 0;
    float t = 1000000.0f;
    bool hit = false;
    bool backface = false;
};

// intersection.rs:9-54
inline bool muller_trumbore(V3 ro, V3 rd, V3 a, V3 b, V3 c, float& out_t, bool& out_backface) {
    out_t = 0.0f;
    V3 edge1 = b - a;
    V3 edge2 = c - a;
    V3 pv = cross(rd, edge2);
    float det = dot(edge1, pv);
    out_backface = std::signbit(det);
    if (std::fabs(det) < 1e-6f) return false;
    float inv_det = 1.0f / det;
    V3 tv = ro - a;
    float u = dot(tv, pv) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    V3 qv = cross(tv, edge1);
    float v = dot(rd, qv) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = dot(edge2, qv) * inv_det;
    if (t < 0.0f) return false;
    out_t = t;
    return true;
}

// intersection.rs:104-122 — slab test with true divisions; f32::min/max ignore NaN
inline float intersect_aabb(const RptBVHNode& n, V3 ro, V3 rd, float prev_min_t) {
    float tx1 = (n.aabb_min[0] - ro.x) / rd.x;
    float tx2 = (n.aabb_max[0] - ro.x) / rd.x;
    float tmin = std::fmin(tx1, tx2);
    float tmax = std::fmax(tx1, tx2);
    float ty1 = (n.aabb_min[1] - ro.y) / rd.y;
    float ty2 = (n.aabb_max[1] - ro.y) / rd.y;
    tmin = std::fmax(tmin, std::fmin(ty1, ty2));
    tmax = std::fmin(tmax, std::fmax(ty1, ty2));
    float tz1 = (n.aabb_min[2] - ro.z) / rd.z;
    float tz2 = (n.aabb_max[2] - ro.z) / rd.z;
    tmin = std::fmax(tmin, std::fmin(tz1, tz2));
    tmax = std::fmin(tmax, std::fmax(tz1, tz2));
    if (tmax >= tmin && tmax > 0.0f && tmin < prev_min_t) return tmin;
    return std::numeric_limits<float>::infinity();
}

// intersection.rs:169-234 — ordered DFS with a 32-entry stack (vec.rs)
template <bool NEAREST>
TraceResult intersect_front_to_back(const Scene& s, V3 ro, V3 rd, float max_t, Counters& ctr) {
    uint32_t stack[32];
    uint32_t sp = 0;
    stack[sp++] = 0;
    TraceResult result;
    while (sp > 0) {
        const uint32_t ni = stack[--sp];
        const RptBVHNode& node = s.nodes[ni];
        ctr.nodes_popped++;
        if (node.triangle_count > 0) {
            for (uint32_t i = 0; i < node.triangle_count; ++i) {
                const uint32_t ti = node.left_or_first + i;
                const uint32_t* tri = s.tris + 4 * (size_t)ti;
                V3 a = xyz(s.verts[tri[0]].vertex), b = xyz(s.verts[tri[1]].vertex), c = xyz(s.verts[tri[2]].vertex);
                float t = 0.0f;
                bool backface = false;
                ctr.tris_tested++;
                if (!NEAREST) ctr.tris_tested_any++;
                if (muller_trumbore(ro, rd, a, b, c, t, backface) && t > 0.001f && t < result.t && (NEAREST || t <= max_t)) {
                    std::memcpy(result.tri, tri, 16);
                    result.triangle_index = ti;
                    result.t = std::fmin(result.t, t);
                    result.hit = true;
                    result.backface = backface;
                    if (!NEAREST) return result;
                }
            }
        } else {
            uint32_t near_i = node.left_or_first, far_i = node.left_or_first + 1;
            float near_d = intersect_aabb(s.nodes[near_i], ro, rd, result.t);
            float far_d = intersect_aabb(s.nodes[far_i], ro, rd, result.t);
            ctr.boxes_tested += 2;
            if (!NEAREST) ctr.boxes_tested_any += 2;
            if (near_d > far_d) {
                std::swap(near_i, far_i);
                std::swap(near_d, far_d);
            }
            if (std::isinf(near_d)) continue;
            if (std::isfinite(far_d)) {
                if (sp >= 32) { ctr.stack_overflows++; return result; }  // Rust: index panic
                stack[sp++] = far_i;
            }
            if (sp >= 32) { ctr.stack_overflows++; return result; }
            stack[sp++] = near_i;
        }
    }
    return result;
}

// ------------------------------------------------------------------ util.rs (live subset)
inline V3 cosine_sample_hemisphere(float r1, float r2) {  // util.rs:24-32
    float theta = std::acos(std::sqrt(r1));
    float phi = 2.0f * kPi * r2;
    return {std::sin(theta) * std::cos(phi), std::cos(theta), std::sin(theta) * std::sin(phi)};
}
inline void create_cartesian(V3 up, V3& right, V3& forward) {  // util.rs:34-40
    V3 arbitrary = v3(0.1f, 0.5f, 0.9f);
    V3 temp = normalize(cross(up, arbitrary));
    right = normalize(cross(temp, up));
    forward = normalize(cross(up, right));
}
inline V3 reflect(V3 i, V3 n) { return i - n * 2.0f * dot(i, n); }  // util.rs:42-44
inline float ggx_distribution(V3 n, V3 h, float roughness) {  // util.rs:58-64
    float numerator = roughness * roughness;
    float n_dot_h = std::fmax(dot(n, h), 0.0f);
    float denominator = (n_dot_h * n_dot_h) * (numerator - 1.0f) + 1.0f;
    denominator = std::fmax(kPi * (denominator * denominator), kEps);
    return numerator / denominator;
}
inline V3 sample_ggx(float r1, float r2, V3 refl, float roughness) {  // util.rs:67-85
    float a = roughness * roughness;
    float phi = 2.0f * kPi * r1;
    float cos_theta = std::sqrt((1.0f - r2) / (r2 * (a * a - 1.0f) + 1.0f));
    float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    V3 halfway = v3(std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, cos_theta);
    V3 up = std::fabs(refl.z) < 0.999f ? v3(0, 0, 1) : v3(1, 0, 0);
    V3 tangent = normalize(cross(up, refl));
    V3 bitangent = cross(refl, tangent);
    return normalize((tangent * halfway.x + bitangent * halfway.y) + refl * halfway.z);
}
inline float geometry_schlick_ggx(V3 n, V3 v, float roughness) {  // util.rs:211-216
    float numerator = std::fmax(dot(n, v), 0.0f);
    float r = (roughness * roughness) / 8.0f;
    float denominator = numerator * (1.0f - r) + r;
    return numerator / denominator;
}
inline float geometry_smith_schlick_ggx(V3 n, V3 v, V3 l, float roughness) {  // util.rs:219-227
    return geometry_schlick_ggx(n, v, roughness) * geometry_schlick_ggx(n, l, roughness);
}
inline V3 fresnel_schlick(float cos_theta, V3 f0) {  // util.rs:229-231
    return f0 + (splat(1.0f) - f0) * powi5(1.0f - cos_theta);
}
inline float fresnel_schlick_scalar(float in_ior, float out_ior, float cos_theta) {  // util.rs:233-236
    float f0 = powi2((in_ior - out_ior) / (in_ior + out_ior));
    return f0 + (1.0f - f0) * powi5(1.0f - cos_theta);
}
inline V3 barycentric(V3 p, V3 a, V3 b, V3 c) {  // util.rs:238-251
    V3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    float v = (d11 * d20 - d01 * d21) / denom;
    float w = (d00 * d21 - d01 * d20) / denom;
    return {1.0f - v - w, v, w};
}
inline float power_heuristic(float p1, float p2) { float p = p1 * p1; return p / (p + p2 * p2); }  // util.rs:253-256
inline V3 mask_nan(V3 v) { return is_finite(v) ? v : splat(0.0f); }                                   // util.rs:271-277
inline float lerpf(float a, float b, float t) { return a * (1.0f - t) + b * t; }                     // util.rs:279-281

// ------------------------------------------------------------------ image_polyfill.rs:32-55
inline V4 sample_raw(const Image& img, int32_t x, int32_t y) {
    // `coord.x as usize % width`: a negative i32 sign-extends to a huge usize before the modulo
    size_t ux = (size_t)(int64_t)x % (size_t)img.width;
    size_t uy = (size_t)(int64_t)y % (size_t)img.height;
    return img.texels[uy * (size_t)img.width + ux];
}
inline V4 sample_by_lod(const Image& img, V2 coord) {
    V2 scaled = coord * V2{(float)img.width, (float)img.height};
    float fx = scaled.x - std::floor(scaled.x), fy = scaled.y - std::floor(scaled.y);
    int32_t cx = f32_as_i32(std::ceil(scaled.x)), cy = f32_as_i32(std::ceil(scaled.y));
    int32_t lx = f32_as_i32(std::floor(scaled.x)), ly = f32_as_i32(std::floor(scaled.y));
    V4 c00 = sample_raw(img, lx, ly);
    V4 c01 = sample_raw(img, lx, cy);
    V4 c10 = sample_raw(img, cx, ly);
    V4 c11 = sample_raw(img, cx, cy);
    V4 a = lerp(c00, c10, fx);
    V4 b = lerp(c01, c11, fx);
    return lerp(a, b, fy);
}

// ------------------------------------------------------------------ bsdf.rs:179-387
enum Lobe : uint32_t { kDiffuse = 0, kSpecular = 1 };
struct BsdfSample {  // bsdf.rs:20-26; Default = zeros, lobe DiffuseReflection
    float pdf = 0.0f;
    Lobe lobe = kDiffuse;
    V3 spectrum = {0, 0, 0};
    V3 direction = {0, 0, 0};
};
constexpr float kDielectricIor = 1.5f;
constexpr float kDielectricF0Sqrt = (kDielectricIor - 1.0f) / (kDielectricIor + 1.0f);
constexpr float kDielectricF0 = kDielectricF0Sqrt * kDielectricF0Sqrt;

struct Pbr {
    V3 albedo;
    float roughness, metallic;
    V2 clamp;

    V3 diffuse_term(float cos_theta, float sw, V3 ks) const {  // bsdf.rs:187-196
        V3 kd = (splat(1.0f) - ks) * (1.0f - metallic);
        V3 diffuse = kd * albedo / kPi;
        return diffuse * cos_theta / (1.0f - sw);
    }
    V3 specular_term(V3 v, V3 n, V3 l, float cos_theta, float d, float sw, V3 ks) const {  // bsdf.rs:198-213
        float g = geometry_smith_schlick_ggx(n, v, l, roughness);
        V3 numerator = d * g * ks;
        float denominator = 4.0f * std::fmax(dot(n, v), 0.0f) * cos_theta;
        V3 specular = numerator / std::fmax(denominator, kEps);
        return specular * cos_theta / sw;
    }
    static float pdf_diffuse(float cos_theta) { return cos_theta / kPi; }  // bsdf.rs:215-217
    static float pdf_specular(V3 v, V3 n, V3 h, float d) { return (d * dot(n, h)) / (4.0f * dot(v, h)); }  // :219-227
    float specular_weight(V3 v, V3 n) const {  // bsdf.rs:238-242 / 275-280
        float fres = fresnel_schlick_scalar(1.0f, kDielectricIor, std::fmax(dot(n, v), 0.0f));
        float sw = lerpf(fres, 1.0f, metallic);
        if (sw != 0.0f && sw != 1.0f) sw = sw < clamp.x ? clamp.x : (sw > clamp.y ? clamp.y : sw);  // f32::clamp
        return sw;
    }

    V3 evaluate(V3 v, V3 n, V3 l, Lobe lobe) const {  // bsdf.rs:231-270
        float sw = specular_weight(v, n);
        float cos_theta = std::fmax(dot(n, l), 0.0f);
        V3 h = normalize(v + l);
        V3 f0 = lerp(splat(kDielectricF0), albedo, metallic);
        V3 ks = fresnel_schlick(std::fmax(dot(h, v), 0.0f), f0);
        if (lobe == kDiffuse) return diffuse_term(cos_theta, sw, ks);
        float d = ggx_distribution(n, h, roughness);
        return specular_term(v, n, l, cos_theta, d, sw, ks);
    }
    float pdf(V3 v, V3 n, V3 l, Lobe lobe) const {  // bsdf.rs:336-351
        if (lobe == kDiffuse) return pdf_diffuse(std::fmax(dot(n, l), 0.0f));
        V3 h = normalize(v + l);
        return pdf_specular(v, n, h, ggx_distribution(n, h, roughness));
    }
    BsdfSample sample(V3 v, V3 n, Rng& rng) const {  // bsdf.rs:272-334
        V3 r = rng.r3();
        float sw = specular_weight(v, n);
        BsdfSample out;
        if (r.z >= sw) {
            V3 nt, nb;
            create_cartesian(n, nt, nb);
            V3 s = cosine_sample_hemisphere(r.x, r.y);
            out.direction = normalize(v3(s.x * nb.x + s.y * n.x + s.z * nt.x, s.x * nb.y + s.y * n.y + s.z * nt.y,
                                         s.x * nb.z + s.y * n.z + s.z * nt.z));
            out.lobe = kDiffuse;
        } else {
            out.direction = sample_ggx(r.x, r.y, reflect(-v, n), roughness);
            out.lobe = kSpecular;
        }
        float cos_theta = std::fmax(dot(n, out.direction), kEps);
        V3 h = normalize(v + out.direction);
        V3 f0 = lerp(splat(kDielectricF0), albedo, metallic);
        V3 ks = fresnel_schlick(std::fmax(dot(h, v), 0.0f), f0);
        if (out.lobe == kDiffuse) {
            out.pdf = pdf_diffuse(cos_theta);
            out.spectrum = diffuse_term(cos_theta, sw, ks);
        } else {
            float d = ggx_distribution(n, h, roughness);
            out.pdf = pdf_specular(v, n, h, d);
            out.spectrum = specular_term(v, n, out.direction, cos_theta, d, sw, ks);
        }
        return out;
    }
};

inline V2 atlas_uv(const float* rect, V2 uv) { return V2{rect[0], rect[1]} + uv * V2{rect[2], rect[3]}; }

Pbr get_pbr_bsdf(const RptTracingConfig& cfg, const RptMaterialData& m, V2 uv, const Image& atlas) {  // bsdf.rs:354-387
    Pbr b;
    if (m.has_albedo_texture) { V4 t = sample_by_lod(atlas, atlas_uv(m.albedo, uv)); b.albedo = {t.x, t.y, t.z}; }
    else b.albedo = xyz(m.albedo);
    float roughness = m.has_roughness_texture ? sample_by_lod(atlas, atlas_uv(m.roughness, uv)).x : m.roughness[0];
    float metallic = m.has_metallic_texture ? sample_by_lod(atlas, atlas_uv(m.metallic, uv)).x : m.metallic[0];
    b.roughness = std::fmax(roughness, kEps);
    b.metallic = std::fmin(metallic, 1.0f - kEps);
    b.clamp = {cfg.specular_weight_clamp[0], cfg.specular_weight_clamp[1]};
    return b;
}

// ------------------------------------------------------------------ light_pick.rs
struct DirectLightSample {  // light_pick.rs:89-98; Default = zeros
    float light_area = 0.0f;
    V3 light_normal = {0, 0, 0};
    float light_pick_pdf = 0.0f;
    V3 light_emission = {0, 0, 0};
    uint32_t light_triangle_index = 0;
    V3 throughput = {0, 0, 0};
    V3 contribution = {0, 0, 0};
};

// light_pick.rs:30-79 (the code is the last three statements)
inline float calculate_light_pdf(float area, float distance, V3 light_normal, V3 light_direction) {
    float cos_theta = dot(light_normal, -light_direction);
    if (cos_theta <= 0.0f) return 0.0f;
    return powi2(distance) / (area * cos_theta);
}
inline float mis_weight(uint32_t nee, float p1, float p2) {  // light_pick.rs:81-87
    return nee == RPT_NEE_MIS ? power_heuristic(p1, p2) : 1.0f;
}

DirectLightSample sample_direct_lighting(uint32_t nee, const Scene& s, V3 throughput, const Pbr& bsdf, V3 p, V3 n,
                                         V3 ray_direction, Rng& rng, Counters& ctr) {  // light_pick.rs:100-173
    DirectLightSample info;
    if (s.lights[0].ratio < 0.0f) return info;  // sentinel: no rng draw

    // pick_light, light_pick.rs:8-16
    V2 r = rng.r2();
    size_t slot = f32_as_usize(r.x * (float)s.nlights);
    if (slot >= s.nlights) {  // r.x can round to exactly 1.0 (p ~ 2^-25): Rust panics; clamp and count
        slot = s.nlights - 1;
        ctr.light_index_clamped++;
    }
    const RptLightPickEntry& e = s.lights[slot];
    uint32_t light_index;
    float light_area, pick_pdf;
    if (r.y < e.ratio) { light_index = e.triangle_index_a; light_area = e.triangle_area_a; pick_pdf = e.triangle_pick_pdf_a; }
    else { light_index = e.triangle_index_b; light_area = e.triangle_area_b; pick_pdf = e.triangle_pick_pdf_b; }

    const uint32_t* tri = s.tris + 4 * (size_t)light_index;
    V3 va = xyz(s.verts[tri[0]].vertex), vb = xyz(s.verts[tri[1]].vertex), vc = xyz(s.verts[tri[2]].vertex);
    V3 light_normal = ((xyz(s.verts[tri[0]].normal) + xyz(s.verts[tri[1]].normal)) + xyz(s.verts[tri[2]].normal)) / 3.0f;
    V3 emission = xyz(s.mats[tri[3]].emissive);

    // pick_triangle_point, light_pick.rs:19-23
    V2 q = rng.r2();
    float sq = std::sqrt(q.x);
    V3 light_point = ((1.0f - sq) * va + (sq * (1.0f - q.y)) * vb) + (sq * q.y) * vc;
    V3 to_light = light_point - p;
    float distance = length(to_light);
    V3 l = to_light / distance;

    V3 direct = splat(0.0f);
    ctr.any_rays++;
    TraceResult shadow = intersect_front_to_back<false>(s, p + l * kEps, l, distance - kEps * 2.0f, ctr);
    if (!shadow.hit) {
        float light_pdf = calculate_light_pdf(light_area, distance, light_normal, l);
        if (light_pdf > 0.0f) {
            V3 f = bsdf.evaluate(-ray_direction, n, l, kDiffuse);
            float bsdf_pdf = bsdf.pdf(-ray_direction, n, l, kDiffuse);
            if (bsdf_pdf > 0.0f) {
                float w = mis_weight(nee, light_pdf, bsdf_pdf);
                direct = (f * emission * w / light_pdf) / pick_pdf;
            }
        }
    }
    info.light_area = light_area;
    info.light_normal = light_normal;
    info.light_pick_pdf = pick_pdf;
    info.light_emission = emission;
    info.light_triangle_index = light_index;
    info.throughput = throughput;
    info.contribution = throughput * direct;
    return info;
}

// light_pick.rs:179-199
V3 bsdf_mis_contribution(const TraceResult& tr, const BsdfSample& last_bsdf, const DirectLightSample& last_light) {
    if (tr.triangle_index != last_light.light_triangle_index) return splat(0.0f);
    float light_pdf = calculate_light_pdf(last_light.light_area, tr.t, last_light.light_normal, last_bsdf.direction);
    if (light_pdf > 0.0f) {
        float w = power_heuristic(last_bsdf.pdf, light_pdf);
        V3 direct = (last_bsdf.spectrum * last_light.light_emission * w / last_bsdf.pdf) / last_light.light_pick_pdf;
        return last_light.throughput * direct;
    }
    return splat(0.0f);
}

// ------------------------------------------------------------------ skybox.rs
namespace sky {
const V3 kRayCoeff = {58e-7f, 135e-7f, 331e-7f};
const V3 kMieScatter = {2e-5f, 2e-5f, 2e-5f};
const V3 kMieEffective = {2e-5f * 1.1f, 2e-5f * 1.1f, 2e-5f * 1.1f};
constexpr float kEarthRadius = 6360e3f, kAtmosphereRadius = 6380e3f, kHRay = 8e3f, kHMie = 12e2f;
const V3 kCenter = {0.0f, -kEarthRadius, 0.0f};

float escape(V3 p, V3 d, float r) {  // skybox.rs:18-32
    V3 v = p - kCenter;
    float b = dot(v, d);
    float det = b * b - dot(v, v) + r * r;
    if (det < 0.0f) return -1.0f;
    det = std::sqrt(det);
    float t1 = -b - det, t2 = -b + det;
    return t1 >= 0.0f ? t1 : t2;
}
V2 densities_rm(V3 p) {  // skybox.rs:34-39
    float h = std::fmax(length(p - kCenter) - kEarthRadius, 0.0f);
    return {std::exp(-h / kHRay), std::exp(-h / kHMie)};
}
V2 scatter_depth_int(V3 o, V3 d, float l) {  // skybox.rs:41-44
    return densities_rm(o) * (l / 2.0f) + densities_rm(o + d * l) * (l / 2.0f);
}
V3 scatter(const float* sundir4, V3 origin, V3 direction) {  // skybox.rs:46-94
    const V3 sundir = xyz(sundir4);
    const uint32_t steps = 12;
    float depth = escape(origin, direction, kAtmosphereRadius) / (float)steps;
    V3 i_r = splat(0.0f), i_m = splat(0.0f);
    V2 total = {0.0f, 0.0f};
    for (uint32_t i = 0; i < steps; ++i) {
        V3 p = origin + direction * (depth * (float)i);
        V2 d_rm = densities_rm(p) * depth;
        total = total + d_rm;
        V2 sum = total + scatter_depth_int(p, sundir, escape(p, sundir, kAtmosphereRadius));
        V3 a = exp3((-kRayCoeff) * sum.x - kMieEffective * sum.y);
        i_r = i_r + a * d_rm.x;
        i_m = i_m + a * d_rm.y;
    }
    float mu = dot(direction, sundir);
    V3 res = (sundir4[3] * (1.0f + mu * mu)) *
             (i_r * kRayCoeff * 0.0597f + i_m * kMieScatter * 0.0196f / std::pow(1.58f - 1.52f * mu, 1.5f));
    return powf3(mask_nan(v3(std::sqrt(res.x), std::sqrt(res.y), std::sqrt(res.z))), 2.2f);
}
}  // namespace sky

// ------------------------------------------------------------------ kernels/src/lib.rs:21-186
static bool g_retire_dead_paths = false;  // test switch, see oracle_set_retire_dead_paths

struct PixelResult {
    V3 radiance;
    uint32_t primary_triangle;  // diagnostics: triangle_index of bounce 0, 0xFFFFFFFF on miss
    float primary_t;
    bool rng_exhausted;
};

// camera ray of lib.rs:36-51
inline void camera_ray(const RptTracingConfig& cfg, uint32_t px, uint32_t py, Rng& rng, V3& ro, V3& rd) {
    V2 jitter = rng.r2();
    float sx = (float)px + jitter.x, sy = (float)py + jitter.y;
    float ux = (sx / (float)cfg.width) * 2.0f - 1.0f;
    float uy = (1.0f - sy / (float)cfg.height) * 2.0f - 1.0f;
    uy *= (float)cfg.height / (float)cfg.width;
    ro = xyz(cfg.cam_position);
    rd = normalize(v3(ux, uy, 1.0f));
    M3 euler = rotation_y(cfg.cam_rotation[1]) * rotation_x(cfg.cam_rotation[0]);
    rd = euler * rd;
}

PixelResult trace_pixel(uint32_t px, uint32_t py, const RptTracingConfig& cfg, uint32_t seed_x, uint32_t seed_y,
                        const Scene& s, Counters& ctr) {
    const uint32_t nee_mode = cfg.nee <= 2 ? cfg.nee : 0;  // from_u32: unknown -> None
    const bool nee = nee_mode != RPT_NEE_NONE;
    Rng rng{seed_x, seed_y};
    V3 ro, rd;
    camera_ray(cfg, px, py, rng, ro, rd);

    V3 throughput = splat(1.0f), radiance = splat(0.0f);
    BsdfSample last_bsdf;
    DirectLightSample last_light;
    PixelResult out{splat(0.0f), 0xFFFFFFFFu, 0.0f, false};

    for (uint32_t bounce = 0; bounce < cfg.max_bounces; ++bounce) {
        ctr.nearest_rays++;
        TraceResult tr = intersect_front_to_back<true>(s, ro, rd, 0.0f, ctr);
        V3 hit = ro + rd * tr.t;
        if (bounce == 0 && tr.hit) { out.primary_triangle = tr.triangle_index; out.primary_t = tr.t; }

        if (!tr.hit) {
            if (cfg.has_skybox == 0) {
                radiance = radiance + throughput * sky::scatter(cfg.sun_direction, ro, rd);
            } else {  // lib.rs:70-78 — lat-long image, yaw taken from the sun direction
                float rotation = std::atan2(cfg.sun_direction[2], cfg.sun_direction[0]);
                V3 rotated = rotation_y(rotation) * rd;
                float u = 0.5f + std::atan2(rotated.z, rotated.x) / (2.0f * kPi);
                float v = 1.0f - (0.5f + std::asin(rotated.y) / kPi);
                float intensity = cfg.sun_direction[3] * (1.0f / 15.0f);
                V4 texel = sample_by_lod(s.sky, V2{u, v});
                radiance = radiance + throughput * v3(texel.x, texel.y, texel.z) * intensity;
            }
            break;
        }

        const RptMaterialData& material = s.mats[tr.tri[3]];
        if (!is_zero(xyz(material.emissive))) {  // lib.rs:86-109
            if (tr.backface) break;
            if (!nee || bounce == 0 || last_bsdf.lobe != kDiffuse) {
                radiance = radiance + mask_nan(throughput * xyz(material.emissive));
                break;
            }
            if (nee_mode == RPT_NEE_MIS && last_bsdf.lobe == kDiffuse) {
                radiance = radiance + mask_nan(bsdf_mis_contribution(tr, last_bsdf, last_light));
                break;
            }
            // mode 2 after a diffuse bounce: no add, no break — the emitter is shaded as a surface
        }

        // lib.rs:111-129 — attribute interpolation with barycentrics re-derived from the hit point
        const RptPerVertexData& da = s.verts[tr.tri[0]];
        const RptPerVertexData& db = s.verts[tr.tri[1]];
        const RptPerVertexData& dc = s.verts[tr.tri[2]];
        V3 bary = barycentric(hit, xyz(da.vertex), xyz(db.vertex), xyz(dc.vertex));
        V3 normal = (bary.x * xyz(da.normal) + bary.y * xyz(db.normal)) + bary.z * xyz(dc.normal);
        V2 uv = (bary.x * V2{da.uv0[0], da.uv0[1]} + bary.y * V2{db.uv0[0], db.uv0[1]}) + bary.z * V2{dc.uv0[0], dc.uv0[1]};
        {   // `if uv.clamp(0,1) != uv { uv = uv.fract() }`
            float cx = std::fmin(std::fmax(uv.x, 0.0f), 1.0f), cy = std::fmin(std::fmax(uv.y, 0.0f), 1.0f);
            if (cx != uv.x || cy != uv.y) uv = {uv.x - std::floor(uv.x), uv.y - std::floor(uv.y)};
        }
        if (material.has_normal_texture) {  // lib.rs:131-141
            V4 t = sample_by_lod(s.atlas, atlas_uv(material.normals, uv));
            V3 nm = v3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
            V3 tangent = (bary.x * xyz(da.tangent) + bary.y * xyz(db.tangent)) + bary.z * xyz(dc.tangent);
            M3 tbn{tangent, cross(tangent, normal), normal};
            normal = normalize(tbn * nm);
        }

        Pbr bsdf = get_pbr_bsdf(cfg, material, uv, s.atlas);
        BsdfSample bs = bsdf.sample(-rd, normal, rng);
        last_bsdf = bs;

        if (nee && bs.lobe == kDiffuse) {  // lib.rs:149-165
            last_light = sample_direct_lighting(nee_mode, s, throughput, bsdf, hit, normal, rd, rng, ctr);
            radiance = radiance + mask_nan(last_light.contribution);
        }

        throughput = throughput * (bs.spectrum / bs.pdf);
        rd = bs.direction;
        ro = hit + rd * kEps;

        if (bounce > cfg.min_bounces) {  // lib.rs:175-181
            float prob = max_element(throughput);
            if (rng.r1() > prob) break;
            throughput = throughput * (1.0f / prob);
        } else if (g_retire_dead_paths && is_zero(throughput) && !rng.draws_a_one()) {
            break;  // NOT in the reference: see oracle_set_retire_dead_paths
        }
    }
    out.radiance = radiance;
    out.rng_exhausted = rng.exhausted;
    return out;
}

std::vector<V4> to_float_texels(const uint8_t* rgba8, uint32_t w, uint32_t h) {
    // dynamic_image_to_cpu_buffer, src/asset.rs:266-273: into_rgb8 drops alpha, then (r,g,b,255)/255
    std::vector<V4> out((size_t)w * h);
    for (size_t i = 0; i < out.size(); ++i)
        out[i] = {(float)rgba8[4 * i] / 255.0f, (float)rgba8[4 * i + 1] / 255.0f, (float)rgba8[4 * i + 2] / 255.0f, 255.0f / 255.0f};
    return out;
}

}  // namespace

// ====================================================================== C entry points (ctypes)
extern "C" {

// Checker for one optimisation of the CUDA backend (wavefront_shade.cu retires paths whose throughput is exactly zero):
// with the switch on, the restatement stops such paths too, so tests can show on the CPU — at sizes and on scenes the
// GPU tests do not cover — that the accumulator does not change in a single bit.  Off by default: the reference
// walks those paths to their first roulette bounce.  The backend's rule has one guard, restated here with it: a dead
// path that can still draw a random number of exactly 1.0 is NOT retired — a lobe selector of 1.0 meeting a specular
// weight of 1.0 gives the diffuse term 1 / (1 - 1) = inf (bsdf.rs:202-211), 0 x inf = NaN, and the NaN then reaches the
// accumulator through the unmasked sky term (lib.rs:69): found on the BreakTime proxy at sample 4534, where 4 of the
// reference's 20 NaN pixels are such paths (tests/test_dead_paths_cpu.py).
void oracle_set_retire_dead_paths(int on) { g_retire_dead_paths = on != 0; }

struct OracleWorld {
    const RptPerVertexData* verts; uint32_t nverts;
    const uint32_t* tris; uint32_t ntris;
    const RptBVHNode* nodes; uint32_t nnodes;
    const RptMaterialData* mats; uint32_t nmats;
    const RptLightPickEntry* lights; uint32_t nlights;
    const uint8_t* atlas_rgba8; uint32_t atlas_w, atlas_h;   // may be NULL -> 1x1 white
    const float* sky_rgba32f; uint32_t sky_w, sky_h;         // may be NULL -> 2x2 magenta (src/asset.rs:283-290)
    // The atlas as the CPU path reads it: float texels, converted ONCE before the sample loop (src/trace.rs:268-271).
    // Optional: oracle_convert_atlas() output kept by the caller across oracle_trace calls; NULL -> converted per call.
    const float* atlas_rgba32f;
};

struct OracleCounters {
    uint64_t paths, nearest_rays, any_rays, nodes_popped, boxes_tested, tris_tested;
    uint64_t stack_overflows, light_index_clamped, rng_exhausted;
    uint64_t boxes_tested_any, tris_tested_any;
};

int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// RGBA8 -> the float texels of `dynamic_image_to_cpu_buffer` (src/asset.rs:266-273); out holds 4 * w * h floats.
int oracle_convert_atlas(const uint8_t* rgba8, uint32_t w, uint32_t h, float* out) {
    if (!rgba8 || !out) return -1;
    const std::vector<V4> texels = to_float_texels(rgba8, w, h);
    std::memcpy(out, texels.data(), texels.size() * sizeof(V4));
    return 0;
}

// Advance every pixel by `n_samples` sample indices, exactly like n passes of the `while running`
// body in trace_cpu (src/trace.rs:273-298): output[i] += (radiance, 1); rng[i] = (x + 1, y).
// primary_ids (optional, width*height) receives the bounce-0 triangle_index of the LAST pass.
int oracle_trace(const RptTracingConfig* cfg, const OracleWorld* w, uint32_t* rng_xy, float* output_rgba,
                 uint32_t n_samples, int n_threads, uint32_t* primary_ids, OracleCounters* counters_out) {
    if (!cfg || !w || !rng_xy || !output_rgba) return -1;
    std::vector<V4> atlas_f, sky_f;
    const V4 white = {1, 1, 1, 1};
    const V4 magenta[4] = {{1, 0, 1, 1}, {1, 0, 1, 1}, {1, 0, 1, 1}, {1, 0, 1, 1}};
    Scene s{w->verts, w->tris, w->nodes, w->mats, w->lights, w->nlights, {&white, 1, 1}, {magenta, 2, 2}};
    if (w->atlas_rgba8 && w->atlas_rgba32f) {
        s.atlas = {reinterpret_cast<const V4*>(w->atlas_rgba32f), w->atlas_w, w->atlas_h};
    } else if (w->atlas_rgba8) {
        atlas_f = to_float_texels(w->atlas_rgba8, w->atlas_w, w->atlas_h);
        s.atlas = {atlas_f.data(), w->atlas_w, w->atlas_h};
    }
    if (w->sky_rgba32f) s.sky = {reinterpret_cast<const V4*>(w->sky_rgba32f), w->sky_w, w->sky_h};

    const uint32_t W = cfg->width, H = cfg->height;
    Counters total;
    uint64_t exhausted = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    for (uint32_t pass = 0; pass < n_samples; ++pass) {
#pragma omp parallel
        {
            Counters local;
            uint64_t local_exhausted = 0;
#pragma omp for schedule(dynamic, 1)
            for (int64_t y = 0; y < (int64_t)H; ++y) {  // one task per image row, like par_chunks_mut(width)
                for (uint32_t x = 0; x < W; ++x) {
                    const size_t i = (size_t)y * W + x;
                    PixelResult r = trace_pixel(x, (uint32_t)y, *cfg, rng_xy[2 * i], rng_xy[2 * i + 1], s, local);
                    output_rgba[4 * i + 0] += r.radiance.x;
                    output_rgba[4 * i + 1] += r.radiance.y;
                    output_rgba[4 * i + 2] += r.radiance.z;
                    output_rgba[4 * i + 3] += 1.0f;
                    rng_xy[2 * i] += 1;
                    if (primary_ids) primary_ids[i] = r.primary_triangle;
                    local_exhausted += r.rng_exhausted;
                }
            }
#pragma omp critical
            { total.add(local); exhausted += local_exhausted; }
        }
    }
    if (counters_out) {
        counters_out->paths = (uint64_t)W * H * n_samples;
        counters_out->nearest_rays = total.nearest_rays;
        counters_out->any_rays = total.any_rays;
        counters_out->nodes_popped = total.nodes_popped;
        counters_out->boxes_tested = total.boxes_tested;
        counters_out->tris_tested = total.tris_tested;
        counters_out->stack_overflows = total.stack_overflows;
        counters_out->light_index_clamped = total.light_index_clamped;
        counters_out->rng_exhausted = exhausted;
        counters_out->boxes_tested_any = total.boxes_tested_any;
        counters_out->tris_tested_any = total.tris_tested_any;
    }
    return 0;
}

// Single rays through the reference traversal (unit tests of wide-BVH / any-hit parity).
// out per ray: hit (0/1), triangle_index, t bits, backface.
int oracle_intersect(const OracleWorld* w, const float* rays_o_d, uint32_t nrays, int any_hit, const float* max_t,
                     uint32_t* out_hit, uint32_t* out_tri, float* out_t, uint32_t* out_backface) {
    Scene s{w->verts, w->tris, w->nodes, w->mats, w->lights, w->nlights, {nullptr, 1, 1}, {nullptr, 2, 2}};
    Counters c;
    for (uint32_t i = 0; i < nrays; ++i) {
        V3 ro = xyz(rays_o_d + 6 * (size_t)i), rd = xyz(rays_o_d + 6 * (size_t)i + 3);
        TraceResult r = any_hit ? intersect_front_to_back<false>(s, ro, rd, max_t[i], c) : intersect_front_to_back<true>(s, ro, rd, 0.0f, c);
        out_hit[i] = r.hit;
        out_tri[i] = r.triangle_index;
        out_t[i] = r.t;
        out_backface[i] = r.backface;
    }
    return c.stack_overflows ? -2 : 0;
}

// Camera rays for a given sample (used to feed oracle_intersect and the wide-BVH tests).
int oracle_camera_rays(const RptTracingConfig* cfg, const uint32_t* rng_xy, float* rays_o_d) {
    for (uint32_t y = 0; y < cfg->height; ++y)
        for (uint32_t x = 0; x < cfg->width; ++x) {
            size_t i = (size_t)y * cfg->width + x;
            Rng rng{rng_xy[2 * i], rng_xy[2 * i + 1]};
            V3 ro, rd;
            camera_ray(*cfg, x, y, rng, ro, rd);
            float* o = rays_o_d + 6 * i;
            o[0] = ro.x; o[1] = ro.y; o[2] = ro.z; o[3] = rd.x; o[4] = rd.y; o[5] = rd.z;
        }
    return 0;
}

}  // extern "C"
