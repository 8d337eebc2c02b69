// display_nofma.cu — what happens to the accumulator on its way to the screen (display half):
//   * normalise:  framebuffer = output.xyz / samples as packed RGB f32          (src/trace.rs:199-204)
//   * display:    the fragment stage of src/resources/render.wgsl (fs_main, :150-185) — the tonemap operator
//                 selected by `Tonemapping` (src/app.rs:20-28) applied to the normalised colour — and, for the
//                 8-bit variant, the render-target conversion save_render relies on (src/app.rs:759-840:
//                 unorm8 or sRGB-encoded unorm8 attachment, alpha 1, bytes delivered as RGBA).
// The fragment shader reads render_buffer[idx] with idx = row-major pixel of a top-left origin (its `uv.y = 1 - uv.y`
// undoes clip space), so in buffer terms display[i] = tonemap(framebuffer[i]): no flip.
// Built with IEEE division and without FMA contraction: the normalisation is the host's `/` in the reference and is
// compared bit for bit; the tonemap curves run in fp32 like the shader and, evaluated in source order, equal the
// float32 numpy restatement in oracle/display_oracle.py bit for bit (WGSL itself leaves precision open).
#include "device_scene.h"

namespace rpt {

namespace {

struct rgb3 { float r, g, b; };

template <class F>
__device__ __forceinline__ rgb3 per_channel(rgb3 v, F f) { return {f(v.r), f(v.g), f(v.b)}; }

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }  // NaN -> 0 (fmaxf drops it)

// render.wgsl:36-43
__device__ __forceinline__ float aces_narkowicz(float x) {
    return clamp01((x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f));
}

// render.wgsl:46-71 (the matrices are written row by row there: transpose(mat3x3(rows)))
__device__ __forceinline__ rgb3 aces_hill(rgb3 x) {
    rgb3 c = {0.59719f * x.r + 0.35458f * x.g + 0.04823f * x.b,
              0.07600f * x.r + 0.90834f * x.g + 0.01566f * x.b,
              0.02840f * x.r + 0.13383f * x.g + 0.83777f * x.b};
    c = per_channel(c, [](float v) {
        const float a = v * (v + 0.0245786f) - 0.000090537f;
        const float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
        return a / b;
    });
    return {clamp01(1.60475f * c.r - 0.53108f * c.g - 0.07367f * c.b),
            clamp01(-0.10208f * c.r + 1.10813f * c.g - 0.00605f * c.b),
            clamp01(-0.00327f * c.r - 0.07276f * c.g + 1.07602f * c.b)};
}

// render.wgsl:77-79 and :104-112 are the same rational curve with different constants
__device__ __forceinline__ float filmic_curve(float x, float a, float b, float c, float d, float e, float f) {
    return ((x * (a * x + c * b) + d * e) / (x * (a * x + b) + d * f)) - e / f;
}

// render.wgsl:81-102 (white level 5.3, white clip 1)
__device__ __forceinline__ float neutral(float x) {
    const float a = 0.2f, b = 0.29f, c = 0.24f, d = 0.272f, e = 0.02f, f = 0.3f;
    const float white_scale = 1.0f / filmic_curve(5.3f, a, b, c, d, e, f);
    return (filmic_curve(x * white_scale, a, b, c, d, e, f) * white_scale) / 1.0f;
}

// render.wgsl:114-121 (exposure bias 2, white point 11.2)
__device__ __forceinline__ float uncharted(float x) {
    const float a = 0.15f, b = 0.50f, c = 0.10f, d = 0.20f, e = 0.02f, f = 0.30f;
    return filmic_curve(x * 2.0f, a, b, c, d, e, f) * (1.0f / filmic_curve(11.2f, a, b, c, d, e, f));
}

// render.wgsl:162-184: the switch of fs_main; any other value is "no tonemapping"
__device__ __forceinline__ rgb3 tonemap(rgb3 v, uint32_t op) {
    switch (op) {
        case 1u: return per_channel(v, [](float x) { return x / (x + 1.0f); });
        case 2u: return per_channel(v, [](float x) { return aces_narkowicz(x * 0.6f); });
        case 3u: return per_channel(v, [](float x) { return aces_narkowicz(x); });
        case 4u: return aces_hill(v);
        case 5u: return per_channel(v, [](float x) { return neutral(x); });
        case 6u: return per_channel(v, [](float x) { return uncharted(x); });
        default: return v;
    }
}

// Colour-attachment store: clamp, optional linear -> sRGB transfer, round to the nearest of 256 levels.
__device__ __forceinline__ uint32_t to_unorm8(float x, bool srgb) {
    x = clamp01(x);
    if (srgb) x = x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
    return (uint32_t)__float2int_rn(x * 255.0f);
}

__global__ void __launch_bounds__(256) normalize_kernel(const float4* __restrict__ output, float* __restrict__ rgb, uint32_t npixels, float samples) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npixels) return;
    const float4 c = output[i];
    rgb[3 * (size_t)i + 0] = __fdiv_rn(c.x, samples);
    rgb[3 * (size_t)i + 1] = __fdiv_rn(c.y, samples);
    rgb[3 * (size_t)i + 2] = __fdiv_rn(c.z, samples);
}

template <bool BYTES>
__global__ void __launch_bounds__(256) display_kernel(const float4* __restrict__ output, void* __restrict__ out, uint32_t npixels, float samples,
                                                      uint32_t op, bool srgb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npixels) return;
    const float4 c = output[i];
    const rgb3 v = tonemap({__fdiv_rn(c.x, samples), __fdiv_rn(c.y, samples), __fdiv_rn(c.z, samples)}, op);
    if (BYTES) {
        static_cast<uint32_t*>(out)[i] = to_unorm8(v.r, srgb) | (to_unorm8(v.g, srgb) << 8) | (to_unorm8(v.b, srgb) << 16) | 0xFF000000u;
    } else {
        float* rgb = static_cast<float*>(out);
        rgb[3 * (size_t)i + 0] = v.r;
        rgb[3 * (size_t)i + 1] = v.g;
        rgb[3 * (size_t)i + 2] = v.b;
    }
}

}  // namespace

void launch_normalize(const float4* output, float* rgb, uint32_t npixels, float samples, cudaStream_t stream) {
    normalize_kernel<<<(npixels + 255) / 256, 256, 0, stream>>>(output, rgb, npixels, samples);
}
void launch_display(const float4* output, float* rgb, uint32_t npixels, float samples, uint32_t tonemap_op, cudaStream_t stream) {
    display_kernel<false><<<(npixels + 255) / 256, 256, 0, stream>>>(output, rgb, npixels, samples, tonemap_op, false);
}
void launch_display_rgba8(const float4* output, uint32_t* rgba8, uint32_t npixels, float samples, uint32_t tonemap_op, bool srgb, cudaStream_t stream) {
    display_kernel<true><<<(npixels + 255) / 256, 256, 0, stream>>>(output, rgba8, npixels, samples, tonemap_op, srgb);
}

}  // namespace rpt
