// atlas_build.cpp — host-side texture atlas producer: src/atlas.rs and the texture part of
// the material loop in src/asset.rs:135-192, restated in C++ behind rpt_host.h.
//
//   * quadtree split of the atlas until there are more leaves than textures; leaves stable-sorted by descending
//     width and truncated (src/atlas.rs:26-69);
//   * every texture resized to its leaf with a Lanczos3 convolution, flipped vertically, copied in (:71-87);
//   * the rect handed to the kernels is (x/W, y/W, w/W, h/H) — the y offset really is divided by the atlas WIDTH
//     (:16-23; harmless for the square atlas the reference uses);
//   * albedo textures are gamma-2.2 decoded in 8 bits before packing (src/asset.rs:140-147).
//
// The reference resizes through the fast_image_resize crate (2.7.3, not vendored under /root/reference).  Its
// U8x4 convolution is restated here from the published algorithm it ports (Pillow's ImagingResample: separable,
// support 3 * max(scale, 1), coefficients normalised and rounded to 22-bit fixed point, horizontal pass then
// vertical pass, round-half-up and clip to 8 bits).  No golden vectors of the crate exist in the reference, so the
// resampled texels are parity-UNPINNED; a texture that already has its leaf's size is copied bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <atomic>
#include <thread>
#include <cstdlib>
#include <vector>

#include "../../include/rpt_errors.h"
#include "../../include/rpt_host.h"

namespace {

struct Rect {
    uint32_t x, y, w, h;
};

std::vector<Rect> packing_rects(uint32_t ntextures, uint32_t atlas_w, uint32_t atlas_h) {
    std::deque<Rect> queue{Rect{0, 0, atlas_w, atlas_h}};
    while (queue.size() <= ntextures) {
        const Rect n = queue.front();
        queue.pop_front();
        const uint32_t hw = n.w / 2, hh = n.h / 2;
        queue.push_back(Rect{n.x, n.y, hw, hh});
        queue.push_back(Rect{n.x + hw, n.y, hw, hh});
        queue.push_back(Rect{n.x, n.y + hh, hw, hh});
        queue.push_back(Rect{n.x + hw, n.y + hh, hw, hh});
    }
    std::vector<Rect> leaves(queue.begin(), queue.end());
    std::stable_sort(leaves.begin(), leaves.end(), [](const Rect& a, const Rect& b) { return a.w > b.w; });  // slice::sort_by is stable
    leaves.resize(ntextures);
    return leaves;
}

// ---- Lanczos3 convolution, 8-bit RGBA ----------------------------------------------------------------------
constexpr int kPrecisionBits = 32 - 8 - 2;

double lanczos3(double x) {
    if (x == 0.0) return 1.0;
    if (x < -3.0 || x >= 3.0) return 0.0;
    const double a = x * 3.14159265358979323846;
    return 3.0 * std::sin(a) * std::sin(a / 3.0) / (a * a);
}

struct Taps {
    std::vector<int> first, count;
    std::vector<int32_t> weight;  // ksize per output sample
    int ksize = 0;
};

Taps make_taps(uint32_t in_size, uint32_t out_size) {
    Taps t;
    const double scale = (double)in_size / (double)out_size;
    const double filterscale = std::max(scale, 1.0);
    const double support = 3.0 * filterscale;
    t.ksize = (int)std::ceil(support) * 2 + 1;
    t.first.resize(out_size);
    t.count.resize(out_size);
    t.weight.assign((size_t)out_size * t.ksize, 0);
    std::vector<double> k(t.ksize);
    for (uint32_t xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > (int)in_size) xmax = (int)in_size;
        const int n = xmax - xmin;
        double sum = 0.0;
        for (int x = 0; x < n; ++x) {
            k[x] = lanczos3((x + xmin - center + 0.5) / filterscale);
            sum += k[x];
        }
        for (int x = 0; x < n; ++x) {
            const double w = sum != 0.0 ? k[x] / sum : 0.0;
            t.weight[(size_t)xx * t.ksize + x] = (int32_t)(w < 0 ? w * (1 << kPrecisionBits) - 0.5 : w * (1 << kPrecisionBits) + 0.5);
        }
        t.first[xx] = xmin;
        t.count[xx] = n;
    }
    return t;
}

inline uint8_t clip8(int64_t v) {
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

std::vector<uint8_t> resize_rgba8(const uint8_t* src, uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh) {
    // horizontal pass: sw x sh -> dw x sh
    const Taps hx = make_taps(sw, dw);
    std::vector<uint8_t> tmp((size_t)dw * sh * 4);
    for (uint32_t y = 0; y < sh; ++y)
        for (uint32_t x = 0; x < dw; ++x) {
            int64_t acc[4] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
            const int32_t* w = &hx.weight[(size_t)x * hx.ksize];
            for (int k = 0; k < hx.count[x]; ++k) {
                const uint8_t* p = src + ((size_t)y * sw + (size_t)(hx.first[x] + k)) * 4;
                for (int c = 0; c < 4; ++c) acc[c] += (int64_t)p[c] * w[k];
            }
            for (int c = 0; c < 4; ++c) tmp[((size_t)y * dw + x) * 4 + c] = clip8(acc[c]);
        }
    // vertical pass: dw x sh -> dw x dh
    const Taps vy = make_taps(sh, dh);
    std::vector<uint8_t> out((size_t)dw * dh * 4);
    for (uint32_t y = 0; y < dh; ++y)
        for (uint32_t x = 0; x < dw; ++x) {
            int64_t acc[4] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
            const int32_t* w = &vy.weight[(size_t)y * vy.ksize];
            for (int k = 0; k < vy.count[y]; ++k) {
                const uint8_t* p = tmp.data() + ((size_t)(vy.first[y] + k) * dw + x) * 4;
                for (int c = 0; c < 4; ++c) acc[c] += (int64_t)p[c] * w[k];
            }
            for (int c = 0; c < 4; ++c) out[((size_t)y * dw + x) * 4 + c] = clip8(acc[c]);
        }
    return out;
}

}  // namespace

extern "C" int rpt_atlas_rects(uint32_t ntextures, uint32_t atlas_w, uint32_t atlas_h, uint32_t* rects_xywh_out) {
    if ((ntextures && !rects_xywh_out) || atlas_w == 0 || atlas_h == 0) return RPT_ERR_INVALID_ARGUMENT;
    // the split halves the leaves until there are enough of them; a leaf must keep at least one texel
    uint64_t leaves = 1;
    uint32_t w = atlas_w, h = atlas_h;
    while (leaves <= ntextures) { leaves *= 4; w /= 2; h /= 2; }
    if (ntextures && (w == 0 || h == 0)) return RPT_ERR_UNSUPPORTED;
    const std::vector<Rect> rects = packing_rects(ntextures, atlas_w, atlas_h);
    for (uint32_t i = 0; i < ntextures; ++i) {
        rects_xywh_out[4 * i + 0] = rects[i].x; rects_xywh_out[4 * i + 1] = rects[i].y;
        rects_xywh_out[4 * i + 2] = rects[i].w; rects_xywh_out[4 * i + 3] = rects[i].h;
    }
    return RPT_OK;
}

extern "C" int rpt_atlas_pack(const uint8_t* const* textures_rgba8, const uint32_t* widths, const uint32_t* heights, uint32_t ntextures,
                              uint32_t atlas_w, uint32_t atlas_h, uint8_t* atlas_rgba8_out, float* sts_out) {
    if (!atlas_rgba8_out || atlas_w == 0 || atlas_h == 0 || (ntextures && (!textures_rgba8 || !widths || !heights || !sts_out)))
        return RPT_ERR_INVALID_ARGUMENT;
    std::vector<uint32_t> rects((size_t)ntextures * 4);
    const int rc = rpt_atlas_rects(ntextures, atlas_w, atlas_h, rects.data());
    if (rc != RPT_OK) return rc;
    std::memset(atlas_rgba8_out, 0, (size_t)atlas_w * atlas_h * 4);  // DynamicImage::new_rgba8: transparent black
    for (uint32_t i = 0; i < ntextures; ++i)
        if (!textures_rgba8[i] || widths[i] == 0 || heights[i] == 0) return RPT_ERR_INVALID_ARGUMENT;
    // Every texture owns its rectangle of the atlas: the resizes run on the host threads (RPT_BUILD_THREADS, like the
    // BVH builders), textures handed out one at a time.
    unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    if (const char* v = std::getenv("RPT_BUILD_THREADS")) threads = (unsigned)std::max(1, std::atoi(v));
    threads = std::min<unsigned>(threads, std::max(1u, ntextures));
    std::atomic<uint32_t> next{0};
    auto worker = [&] {
        for (uint32_t i; (i = next.fetch_add(1)) < ntextures;) {
            const uint32_t x = rects[4 * i], y = rects[4 * i + 1], w = rects[4 * i + 2], h = rects[4 * i + 3];
            std::vector<uint8_t> resized;
            const uint8_t* texels = textures_rgba8[i];
            if (widths[i] != w || heights[i] != h) {
                resized = resize_rgba8(texels, widths[i], heights[i], w, h);
                texels = resized.data();
            }
            for (uint32_t row = 0; row < h; ++row)  // flipv, then copy_from at (x, y)
                std::memcpy(atlas_rgba8_out + ((size_t)(y + row) * atlas_w + x) * 4, texels + (size_t)(h - 1 - row) * w * 4, (size_t)w * 4);
            sts_out[4 * i + 0] = (float)x / (float)atlas_w;
            sts_out[4 * i + 1] = (float)y / (float)atlas_w;  // sic: the reference divides the y offset by the width
            sts_out[4 * i + 2] = (float)w / (float)atlas_w;
            sts_out[4 * i + 3] = (float)h / (float)atlas_h;
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (std::thread& t : pool) t.join();
    return RPT_OK;
}

extern "C" int rpt_decode_albedo_gamma(const uint8_t* rgba8_in, size_t npixels, uint8_t* rgba8_out) {
    if ((npixels && (!rgba8_in || !rgba8_out))) return RPT_ERR_INVALID_ARGUMENT;
    uint8_t lut[256];
    for (int v = 0; v < 256; ++v) {  // `((p as f32 / 255.0).powf(2.2) * 255.0) as u8`, src/asset.rs:143-146
        const float lin = std::pow((float)v / 255.0f, 2.2f) * 255.0f;
        lut[v] = (uint8_t)(lin < 0.0f ? 0.0f : (lin > 255.0f ? 255.0f : lin));
    }
    for (size_t i = 0; i < npixels; ++i) {  // into_rgb8 drops alpha; back to RGBA it is opaque
        rgba8_out[4 * i + 0] = lut[rgba8_in[4 * i + 0]];
        rgba8_out[4 * i + 1] = lut[rgba8_in[4 * i + 1]];
        rgba8_out[4 * i + 2] = lut[rgba8_in[4 * i + 2]];
        rgba8_out[4 * i + 3] = 255;
    }
    return RPT_OK;
}
