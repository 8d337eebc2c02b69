// shading.cuh — device-side shading of the tracing loop: texture fetch, the PBR
// metallic/roughness lobe-select BSDF, light picking / MIS terms, and the two sky models.
//
// Semantics follow the reference functions cited on each block; the arithmetic is organised
// for the GPU (sincosf, algebraic sin/cos of acos, reciprocal reuse), so results agree with the
// CPU path to a few ulps per operation rather than bit for bit — the parity bar for radiance is
// a mean absolute error, not identity (only ray generation / traversal / ray-triangle are exact).
#pragma once

#include "../../../include/rpt_shared_structs.h"
#include "exact.cuh"
#include "vec.cuh"

namespace rpt {

struct f2 {
    float x, y;
};

// ---- texture fetch: the CPU polyfill, shared_structs/src/image_polyfill.rs:32-55 ------------
// Manual bilinear, no half-texel shift, wrap by modulo (with the reference's `i32 as usize`
// sign-extension), alpha ignored.  Texels are either RGBA8 (atlas; /255 like
// dynamic_image_to_cpu_buffer, src/asset.rs:266-273) or float4 (sky).
// Rust's `f32 as i32` saturates and maps NaN to 0 — exactly cvt.rzi.s32.f32.
RPT_D int f32_as_i32_sat(float f) { return __float2int_rz(f); }
// `coord as usize % size`: a negative i32 sign-extends to a huge usize first.  Images whose size is a power of
// two (the 4096^2 atlas, lat-long skies) wrap non-negative coordinates — the only ones valid uvs produce — with
// a mask (`mask` = size - 1, or 0 when the size is not a power of two); everything else takes the general path.
static __device__ __noinline__ uint32_t wrap_coord_general(int c, uint32_t size) {
    if (c >= 0) return (uint32_t)c % size;
    return (uint32_t)((unsigned long long)(long long)c % (unsigned long long)size);
}
RPT_D uint32_t wrap_coord(int c, uint32_t size, uint32_t mask) {
    if (c >= 0 && mask != 0u) return (uint32_t)c & mask;
    return wrap_coord_general(c, size);
}

struct TexelRGBA8 {
    const uchar4* texels;
    RPT_D f3 operator()(uint32_t i) const {
        const uchar4 t = __ldg(texels + i);
        return mk3(unorm8(t.x), unorm8(t.y), unorm8(t.z));
    }
};
struct TexelF32 {
    const float4* texels;
    RPT_D f3 operator()(uint32_t i) const { return xyz(__ldg(texels + i)); }
};

// The lookup in two halves: where the four taps are and how they are weighted, then the blend.
struct BilinearTaps {
    uint32_t i00, i10, i01, i11;
    float fx, fy;
};
RPT_D BilinearTaps bilinear_taps(uint32_t width, uint32_t height, uint32_t wmask, uint32_t hmask, float u, float v) {
    const float sx = u * (float)width, sy = v * (float)height;
    const float flx = floorf(sx), fly = floorf(sy);
    const int cx0 = f32_as_i32_sat(flx), cx1 = f32_as_i32_sat(ceilf(sx)), cy0 = f32_as_i32_sat(fly), cy1 = f32_as_i32_sat(ceilf(sy));
    uint32_t x0, x1, y0, y1;
    if ((cx0 | cx1 | cy0 | cy1) >= 0 && wmask != 0u && hmask != 0u) {  // one test for the four coordinates: the usual case
        x0 = (uint32_t)cx0 & wmask; x1 = (uint32_t)cx1 & wmask;
        y0 = (uint32_t)cy0 & hmask; y1 = (uint32_t)cy1 & hmask;
    } else {
        x0 = wrap_coord(cx0, width, wmask); x1 = wrap_coord(cx1, width, wmask);
        y0 = wrap_coord(cy0, height, hmask); y1 = wrap_coord(cy1, height, hmask);
    }
    return BilinearTaps{y0 * width + x0, y0 * width + x1, y1 * width + x0, y1 * width + x1, sx - flx, sy - fly};
}
RPT_D f3 bilinear_blend(f3 c00, f3 c10, f3 c01, f3 c11, float fx, float fy) {
    const f3 a = lerp3(c00, c10, fx), b = lerp3(c01, c11, fx);
    return lerp3(a, b, fy);
}
template <class Fetch>
RPT_D f3 sample_bilinear(const Fetch& fetch, uint32_t width, uint32_t height, uint32_t wmask, uint32_t hmask, float u, float v) {
    const BilinearTaps t = bilinear_taps(width, height, wmask, hmask, u, v);
    const f3 c00 = fetch(t.i00), c10 = fetch(t.i10);
    const f3 c01 = fetch(t.i01), c11 = fetch(t.i11);
    return bilinear_blend(c00, c10, c01, c11, t.fx, t.fy);
}

struct Atlas {
    const uchar4* texels;
    uint32_t width, height, wmask, hmask;  // masks: pow2_mask(size)
    RPT_D f3 sample(const float* rect, f2 uv) const {  // rect = (u0, v0, su, sv), kernels/src/bsdf.rs:356
        return sample_bilinear(TexelRGBA8{texels}, width, height, wmask, hmask, rect[0] + uv.x * rect[2], rect[1] + uv.y * rect[3]);
    }
};

// ---- sampling utilities, kernels/src/util.rs -----------------------------------------------
// The specular lobe keeps IEEE division and sqrt even in the translation unit built with the approximate forms:
// at low roughness D and the lobe pdf are huge and steep, so a 2-ulp error there is amplified into radiance error
// that an HDR sky makes visible (measured on PBRTest + HDR sky: MAE 6.5e-5 with approximate, 8.7e-7 with IEEE).
RPT_D float div_exact(float a, float b) { return __fdiv_rn(a, b); }
RPT_D float sqrt_exact(float a) { return __fsqrt_rn(a); }
RPT_D float powi5(float x) { const float x2 = x * x; return x2 * x2 * x; }

// util.rs:34-40 — frame around `up` from the fixed helper (0.1, 0.5, 0.9)
RPT_D void create_cartesian(f3 up, f3& right, f3& forward) {
    const f3 temp = normalize(cross(up, mk3(0.1f, 0.5f, 0.9f)));
    right = normalize(cross(temp, up));
    forward = normalize(cross(up, right));
}
// util.rs:58-64
RPT_D float ggx_distribution(float n_dot_h_raw, float roughness) {
    const float a = roughness * roughness;
    const float ndh = fmaxf(n_dot_h_raw, 0.0f);
    float d = (ndh * ndh) * (a - 1.0f) + 1.0f;
    d = fmaxf(kPi * (d * d), kEps);
    return div_exact(a, d);
}
// util.rs:211-227
RPT_D float geometry_schlick_ggx(float n_dot_x_raw, float roughness) {
    const float num = fmaxf(n_dot_x_raw, 0.0f);
    const float r = (roughness * roughness) / 8.0f;
    return div_exact(num, num * (1.0f - r) + r);
}
RPT_D f3 fresnel_schlick(float cos_theta, f3 f0) { return f0 + (splat3(1.0f) - f0) * powi5(1.0f - cos_theta); }  // util.rs:229-231
RPT_D float power_heuristic(float p1, float p2) { const float a = p1 * p1; return a / (a + p2 * p2); }           // util.rs:253-256

// ---- PBR bsdf, kernels/src/bsdf.rs:179-387 --------------------------------------------------
enum : uint32_t { kLobeDiffuse = 0, kLobeSpecular = 1 };
constexpr float kDielectricF0 = (0.5f / 2.5f) * (0.5f / 2.5f);  // ((1.5-1)/(1.5+1))^2, bsdf.rs:173-176

struct Pbr {
    f3 albedo;
    float roughness, metallic;
    float clamp_lo, clamp_hi;

    // lobe-select probability, bsdf.rs:275-280 (same code at :238-242)
    RPT_D float specular_weight(float n_dot_v_raw) const {
        const float f0 = kDielectricF0;  // ((1.0-1.5)/(1.0+1.5))^2, util.rs:233-236
        const float fres = f0 + (1.0f - f0) * powi5(1.0f - fmaxf(n_dot_v_raw, 0.0f));
        float sw = fres * (1.0f - metallic) + 1.0f * metallic;  // util::lerp
        if (sw != 0.0f && sw != 1.0f) sw = sw < clamp_lo ? clamp_lo : (sw > clamp_hi ? clamp_hi : sw);
        return sw;
    }
    RPT_D f3 ks(float h_dot_v_raw) const {
        const f3 f0 = lerp3(splat3(kDielectricF0), albedo, metallic);
        return fresnel_schlick(fmaxf(h_dot_v_raw, 0.0f), f0);
    }
    // bsdf.rs:187-196 (already divided by the lobe-select probability)
    RPT_D f3 diffuse_term(float cos_theta, float sw, f3 ks_) const {
        const f3 kd = (splat3(1.0f) - ks_) * (1.0f - metallic);
        return ((kd * albedo) / kPi) * cos_theta / (1.0f - sw);
    }
    // bsdf.rs:198-213
    RPT_D f3 specular_term(float n_dot_v_raw, float n_dot_l_raw, float cos_theta, float d, float sw, f3 ks_) const {
        const float g = geometry_schlick_ggx(n_dot_v_raw, roughness) * geometry_schlick_ggx(n_dot_l_raw, roughness);
        const float denom = fmaxf(4.0f * fmaxf(n_dot_v_raw, 0.0f) * cos_theta, kEps);
        const f3 dgk = (d * g) * ks_;
        return mk3(div_exact(div_exact(dgk.x, denom) * cos_theta, sw), div_exact(div_exact(dgk.y, denom) * cos_theta, sw),
                   div_exact(div_exact(dgk.z, denom) * cos_theta, sw));
    }
};

struct BsdfSample {
    f3 direction;
    f3 spectrum;
    float pdf;
    uint32_t lobe;
};

// PBR::sample, bsdf.rs:272-334.  r = the three numbers of `rng.gen_r3()`.
RPT_D BsdfSample pbr_sample(const Pbr& m, f3 v, f3 n, f3 r) {
    const float ndv = dot(n, v);
    const float sw = m.specular_weight(ndv);
    BsdfSample out;
    if (r.z >= sw) {
        // cosine_sample_hemisphere (util.rs:24-32): theta = acos(sqrt(r1)) => cos = sqrt(r1), sin = sqrt(1 - r1)
        f3 nt, nb;
        create_cartesian(n, nt, nb);
        const float ct = sqrtf(r.x), st = sqrtf(fmaxf(1.0f - r.x, 0.0f));
        float sp, cp;
        sincospif(2.0f * r.y, &sp, &cp);  // phi = 2*pi*r2
        const float sx = st * cp, sy = ct, sz = st * sp;
        out.direction = normalize(mk3(sx * nb.x + sy * n.x + sz * nt.x, sx * nb.y + sy * n.y + sz * nt.y, sx * nb.z + sy * n.z + sz * nt.z));
        out.lobe = kLobeDiffuse;
    } else {
        // sample_ggx around the mirror direction (util.rs:67-85, bsdf.rs:293-301)
        const f3 i = -v;
        const f3 refl = i - n * 2.0f * dot(i, n);
        const float a = m.roughness * m.roughness;
        float sp, cp;
        sincospif(2.0f * r.x, &sp, &cp);  // phi = 2*pi*r1
        const float ct = sqrt_exact(div_exact(1.0f - r.y, r.y * (a * a - 1.0f) + 1.0f));
        const float st = sqrt_exact(1.0f - ct * ct);
        const f3 up = fabsf(refl.z) < 0.999f ? mk3(0.0f, 0.0f, 1.0f) : mk3(1.0f, 0.0f, 0.0f);
        const f3 tangent = normalize(cross(up, refl));
        const f3 bitangent = cross(refl, tangent);
        out.direction = normalize((tangent * (cp * st) + bitangent * (sp * st)) + refl * ct);
        out.lobe = kLobeSpecular;
    }
    const float ndl = dot(n, out.direction);
    const float cos_theta = fmaxf(ndl, kEps);
    const f3 h = normalize(v + out.direction);
    const f3 ks_ = m.ks(dot(h, v));
    if (out.lobe == kLobeDiffuse) {
        out.pdf = cos_theta / kPi;
        out.spectrum = m.diffuse_term(cos_theta, sw, ks_);
    } else {
        const float ndh = dot(n, h);
        const float d = ggx_distribution(ndh, m.roughness);
        out.pdf = div_exact(d * ndh, 4.0f * dot(v, h));
        out.spectrum = m.specular_term(ndv, ndl, cos_theta, d, sw, ks_);
    }
    return out;
}

// PBR::evaluate + PBR::pdf for the DIFFUSE lobe — the only lobe NEE uses (light_pick.rs:153-155)
RPT_D void pbr_eval_diffuse(const Pbr& m, f3 v, f3 n, f3 l, f3& f_out, float& pdf_out) {
    const float sw = m.specular_weight(dot(n, v));
    const float cos_theta = fmaxf(dot(n, l), 0.0f);
    const f3 h = normalize(v + l);
    f_out = m.diffuse_term(cos_theta, sw, m.ks(dot(h, v)));
    pdf_out = cos_theta / kPi;
}

// get_pbr_bsdf, bsdf.rs:354-387 (NB: roughness and metallic textures both read channel x)
RPT_D Pbr make_pbr(const RptMaterialData& mat, f2 uv, const Atlas& atlas, float clamp_lo, float clamp_hi) {
    Pbr m;
    m.albedo = mat.has_albedo_texture ? atlas.sample(mat.albedo, uv) : mk3(mat.albedo[0], mat.albedo[1], mat.albedo[2]);
    const float roughness = mat.has_roughness_texture ? atlas.sample(mat.roughness, uv).x : mat.roughness[0];
    const float metallic = mat.has_metallic_texture ? atlas.sample(mat.metallic, uv).x : mat.metallic[0];
    m.roughness = fmaxf(roughness, kEps);
    m.metallic = fminf(metallic, 1.0f - kEps);
    m.clamp_lo = clamp_lo;
    m.clamp_hi = clamp_hi;
    return m;
}

// ---- light sampling, kernels/src/light_pick.rs ----------------------------------------------
// calculate_light_pdf, light_pick.rs:74-79: area pdf -> solid angle; 0 when the light faces away
RPT_D float light_pdf(float area, float distance, f3 light_normal, f3 light_direction) {
    const float cos_theta = dot(light_normal, -light_direction);
    if (cos_theta <= 0.0f) return 0.0f;
    return (distance * distance) / (area * cos_theta);
}

// barycentric(), util.rs:238-251, with v0 = e1 and v1 = e2 precomputed
RPT_D f3 barycentric(f3 p, f3 a, f3 e1, f3 e2) {
    const f3 v2 = p - a;
    const float d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2), d20 = dot(v2, e1), d21 = dot(v2, e2);
    const float denom = d00 * d11 - d01 * d01;
    const float v = (d11 * d20 - d01 * d21) / denom;
    const float w = (d00 * d21 - d01 * d20) / denom;
    return mk3(1.0f - v - w, v, w);
}

// ---- procedural sky, kernels/src/skybox.rs --------------------------------------------------
namespace sky {
constexpr float kEarthRadius = 6360e3f, kAtmosphereRadius = 6380e3f, kHRay = 8e3f, kHMie = 12e2f;

RPT_D float escape(f3 p, f3 d, float r) {  // skybox.rs:18-32; CENTER = (0, -EARTH_RADIUS, 0)
    const f3 v = mk3(p.x, p.y + kEarthRadius, p.z);
    const float b = dot(v, d);
    float det = b * b - dot(v, v) + r * r;
    if (det < 0.0f) return -1.0f;
    det = sqrtf(det);
    const float t1 = -b - det;
    return t1 >= 0.0f ? t1 : -b + det;
}
RPT_D f2 densities_rm(f3 p) {  // skybox.rs:34-39
    const float h = fmaxf(length(mk3(p.x, p.y + kEarthRadius, p.z)) - kEarthRadius, 0.0f);
    return f2{expf(-h / kHRay), expf(-h / kHMie)};
}
RPT_D f3 scatter(f3 sundir, float sun_intensity, f3 origin, f3 direction) {  // skybox.rs:46-94
    const f3 ray_coeff = mk3(58e-7f, 135e-7f, 331e-7f);
    const float mie_scatter = 2e-5f, mie_effective = 2e-5f * 1.1f;
    const float depth = escape(origin, direction, kAtmosphereRadius) / 12.0f;
    f3 i_r = splat3(0.0f), i_m = splat3(0.0f);
    float total_r = 0.0f, total_m = 0.0f;
#pragma unroll 1
    for (int i = 0; i < 12; ++i) {
        const f3 p = origin + direction * (depth * (float)i);
        const f2 d = densities_rm(p);
        const float dr = d.x * depth, dm = d.y * depth;
        total_r += dr;
        total_m += dm;
        // scatter_depth_int(p, sundir, escape(p, sundir, R_atm)), skybox.rs:41-44
        const float l = escape(p, sundir, kAtmosphereRadius);
        const f2 d1 = densities_rm(p + sundir * l);
        const float sum_r = total_r + (d.x * (l / 2.0f) + d1.x * (l / 2.0f));
        const float sum_m = total_m + (d.y * (l / 2.0f) + d1.y * (l / 2.0f));
        const float m = mie_effective * sum_m;
        const f3 a = mk3(expf(-ray_coeff.x * sum_r - m), expf(-ray_coeff.y * sum_r - m), expf(-ray_coeff.z * sum_r - m));
        i_r = i_r + a * dr;
        i_m = i_m + a * dm;
    }
    const float mu = dot(direction, sundir);
    const float mie_phase = mie_scatter * 0.0196f / powf(1.58f - 1.52f * mu, 1.5f);
    const f3 res = (sun_intensity * (1.0f + mu * mu)) * (i_r * ray_coeff * 0.0597f + i_m * mie_phase);
    const f3 g = mask_nan(mk3(sqrtf(res.x), sqrtf(res.y), sqrtf(res.z)));
    return mk3(powf(g.x, 2.2f), powf(g.y, 2.2f), powf(g.z, 2.2f));
}
}  // namespace sky

// HDR lat-long sky, kernels/src/lib.rs:70-78.  yaw_sin/yaw_cos = sin/cos(atan2(sun.z, sun.x))
// are computed on the host (one value per config).
struct SkyImage {
    const float4* texels;
    uint32_t width, height, wmask, hmask;  // masks: pow2_mask(size)
    float yaw_sin, yaw_cos, intensity;  // intensity = sun.w * (1/15)
    RPT_D f3 lookup(f3 d) const {
        // Mat3::from_rotation_y(yaw) * d, columns (c,0,-s), (0,1,0), (s,0,c)
        const f3 r = mk3(yaw_cos * d.x + yaw_sin * d.z, d.y, -yaw_sin * d.x + yaw_cos * d.z);
        // (IEEE divisions even where the translation unit uses approximate ones: an HDR sky can have gradients of
        // thousands per unit of u, which would amplify a 2-ulp error in the coordinate into visible radiance error)
        const float u = 0.5f + __fdiv_rn(atan2f(r.z, r.x), 2.0f * kPi);
        const float v = 1.0f - (0.5f + __fdiv_rn(asinf(r.y), kPi));
        return sample_bilinear(TexelF32{texels}, width, height, wmask, hmask, u, v) * intensity;
    }
};

}  // namespace rpt
