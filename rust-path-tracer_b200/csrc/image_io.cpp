// image_io.cpp — host-side image conversion either side of the sky lookup:
//
//   rpt_decode_hdr   <- load_dynamic_image's `.hdr` branch, src/asset.rs:238-254: image::codecs::hdr::HdrDecoder
//                       (`image` 0.24.6, not vendored in the reference checkout — its published behaviour is restated
//                       here: strict "#?RADIANCE" / FORMAT=32-bit_rle_rgbe header, "-Y h +X w" orientation only,
//                       new-style per-component RLE, old-style run markers, flat scanlines; RGBE -> float as
//                       c * 2^(e - 136) with e == 0 -> 0, no exposure applied — parity unpinned, no golden file exists)
//   rpt_sky_texels   <- the two texel conventions the reference feeds the sky lookup (kernels/src/lib.rs:70-78):
//                       GPU path `into_rgba32f()` = (r, g, b, 1) as decoded (src/asset.rs:257-264), CPU path
//                       `into_rgb8()` then (r, g, b, 255) / 255 (src/asset.rs:266-273: every texel clamped to [0, 1]
//                       and quantised to 8 bits — an HDR sky loses its range on the reference's CPU path)
//
// Pure CPU code; nothing throws across the ABI.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rpt_errors.h"
#include "../../include/rpt_host.h"

namespace {

struct Reader {
    const uint8_t* p;
    size_t n, at = 0;
    bool line(std::string& out) {  // up to '\n' (excluded); false at end of input
        if (at >= n) return false;
        out.clear();
        while (at < n && p[at] != '\n') out.push_back((char)p[at++]);
        if (at < n) ++at;
        return true;
    }
    bool bytes(uint8_t* dst, size_t count) {
        if (n - at < count) return false;
        std::memcpy(dst, p + at, count);
        at += count;
        return true;
    }
    bool byte(uint8_t& b) { return bytes(&b, 1); }
};

inline void rgbe_to_float(const uint8_t* px, float* out) {
    if (px[3] == 0) { out[0] = out[1] = out[2] = 0.0f; return; }
    const float scale = std::ldexp(1.0f, (int)px[3] - (128 + 8));
    out[0] = scale * (float)px[0];
    out[1] = scale * (float)px[1];
    out[2] = scale * (float)px[2];
}

// One component of a new-style scanline: runs (count > 128: count - 128 copies of the next byte) and literals.
bool read_component(Reader& r, uint8_t* dst, uint32_t width) {
    uint32_t x = 0;
    while (x < width) {
        uint8_t count;
        if (!r.byte(count)) return false;
        if (count > 128) {
            count -= 128;
            uint8_t value;
            if (count == 0 || x + count > width || !r.byte(value)) return false;
            std::memset(dst + x, value, count);
        } else {
            if (count == 0 || x + count > width || !r.bytes(dst + x, count)) return false;
        }
        x += count;
    }
    return true;
}

// Old-style scanline whose first pixel `first` has been read already: pixels (1, 1, 1, n) repeat the previous pixel
// n << shift times, with the shift growing by 8 for consecutive markers.
bool read_old_scanline(Reader& r, const uint8_t* first, uint8_t* rgbe, uint32_t width) {
    uint32_t x = 0;
    int shift = 0;
    uint8_t px[4];
    std::memcpy(px, first, 4);
    bool have = true;
    while (x < width) {
        if (!have && !r.bytes(px, 4)) return false;
        have = false;
        if (px[0] == 1 && px[1] == 1 && px[2] == 1) {
            if (x == 0) return false;  // nothing to repeat
            uint64_t run = (uint64_t)px[3] << shift;
            if (run > width - x) return false;
            for (uint64_t k = 0; k < run; ++k, ++x) std::memcpy(rgbe + 4 * (size_t)x, rgbe + 4 * (size_t)(x - 1), 4);
            shift += 8;
        } else {
            std::memcpy(rgbe + 4 * (size_t)x, px, 4);
            ++x;
            shift = 0;
        }
    }
    return true;
}

}  // namespace

extern "C" int rpt_decode_hdr(const uint8_t* bytes, size_t nbytes, float* rgb_out, uint32_t* width_out, uint32_t* height_out) {
    if (!bytes || !width_out || !height_out) return RPT_ERR_INVALID_ARGUMENT;
    Reader r{bytes, nbytes};
    std::string line;
    if (!r.line(line) || (line.rfind("#?RADIANCE", 0) != 0 && line.rfind("#?RGBE", 0) != 0)) return RPT_ERR_INVALID_ARGUMENT;
    bool format_ok = false;
    for (;;) {
        if (!r.line(line)) return RPT_ERR_INVALID_ARGUMENT;
        if (line.empty() || line == "\r") break;
        if (line.rfind("FORMAT=", 0) == 0) {
            if (line.rfind("FORMAT=32-bit_rle_rgbe", 0) != 0) return RPT_ERR_UNSUPPORTED;  // (XYZE files are rejected by the reference's decoder too)
            format_ok = true;
        }
    }
    if (!format_ok) return RPT_ERR_INVALID_ARGUMENT;
    if (!r.line(line)) return RPT_ERR_INVALID_ARGUMENT;
    unsigned long h = 0, w = 0;
    char tail = 0;
    if (std::sscanf(line.c_str(), "-Y %lu +X %lu%c", &h, &w, &tail) < 2 || (tail != 0 && tail != '\r')) return RPT_ERR_UNSUPPORTED;  // other orientations
    if (w == 0 || h == 0 || w > 0x7FFFFFFFul / h) return RPT_ERR_INVALID_ARGUMENT;
    *width_out = (uint32_t)w;
    *height_out = (uint32_t)h;
    if (!rgb_out) return RPT_OK;  // size query

    std::vector<uint8_t> scan((size_t)w * 4), comp(w);
    for (unsigned long y = 0; y < h; ++y) {
        uint8_t first[4];
        if (!r.bytes(first, 4)) return RPT_ERR_INVALID_ARGUMENT;
        const bool new_rle = w >= 8 && w < 32768 && first[0] == 2 && first[1] == 2 && first[2] < 128;
        if (new_rle) {
            if ((((unsigned long)first[2] << 8) | first[3]) != w) return RPT_ERR_INVALID_ARGUMENT;
            for (int c = 0; c < 4; ++c) {
                if (!read_component(r, comp.data(), (uint32_t)w)) return RPT_ERR_INVALID_ARGUMENT;
                for (unsigned long x = 0; x < w; ++x) scan[4 * x + c] = comp[x];
            }
        } else if (!read_old_scanline(r, first, scan.data(), (uint32_t)w)) {
            return RPT_ERR_INVALID_ARGUMENT;
        }
        float* row = rgb_out + 3 * (size_t)y * w;
        for (unsigned long x = 0; x < w; ++x) rgbe_to_float(&scan[4 * x], row + 3 * x);
    }
    return RPT_OK;
}

extern "C" int rpt_sky_texels(const float* rgb, uint32_t width, uint32_t height, int cpu_path_rgb8, float* rgba_out) {
    if (!rgb || !rgba_out || width == 0 || height == 0) return RPT_ERR_INVALID_ARGUMENT;
    const size_t n = (size_t)width * height;
    for (size_t i = 0; i < n; ++i) {
        for (int c = 0; c < 3; ++c) {
            float v = rgb[3 * i + c];
            if (cpu_path_rgb8) {
                // image 0.24 f32 -> u8: (clamp(v, 0, 1) * 255).round() (half away from zero); then `as f32 / 255.0`.
                // (A NaN texel makes the reference panic; it becomes 0 here.)
                const float clamped = v != v ? 0.0f : std::fmin(std::fmax(v, 0.0f), 1.0f);
                v = (float)(uint8_t)std::round(clamped * 255.0f) / 255.0f;
            }
            rgba_out[4 * i + c] = v;
        }
        rgba_out[4 * i + 3] = cpu_path_rgb8 ? 255.0f / 255.0f : 1.0f;
    }
    return RPT_OK;
}
