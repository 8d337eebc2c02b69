// trace.cpp — see trace.hpp.  Loop structure of src/trace.rs:136-224 over the C ABI.
#include "trace.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

namespace rpt {

namespace {

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return false;
    const std::streamsize n = f.tellg();
    f.seekg(0);
    out.resize((size_t)n);
    return (bool)f.read(reinterpret_cast<char*>(out.data()), n);
}

// Minimal .npy reader (little-endian, C order): returns the payload offset and the header text.
bool npy_payload(const std::vector<uint8_t>& buf, size_t& offset, std::string& header) {
    if (buf.size() < 12 || std::memcmp(buf.data(), "\x93NUMPY", 6) != 0) return false;
    size_t hlen, start;
    if (buf[6] == 1) { hlen = buf[8] | (buf[9] << 8); start = 10; }
    else { hlen = buf[8] | (buf[9] << 8) | (buf[10] << 16) | ((size_t)buf[11] << 24); start = 12; }
    if (start + hlen > buf.size()) return false;
    header.assign(reinterpret_cast<const char*>(buf.data() + start), hlen);
    offset = start + hlen;
    return true;
}

std::string module_dir() {  // .../rust-path-tracer_b200/host/ -> resources live one level up
    const std::string f = __FILE__;
    return f.substr(0, f.find_last_of('/'));
}

bool load_blue_noise(std::vector<uint8_t>& tile) {
    std::vector<uint8_t> buf;
    size_t off;
    std::string hdr;
    const char* env = std::getenv("RPT_RESOURCES");
    const std::string path = (env ? std::string(env) : module_dir() + "/../resources") + "/bluenoise_r8.npy";
    if (!read_file(path, buf) || !npy_payload(buf, off, hdr) || buf.size() - off < 256 * 256) return false;
    tile.assign(buf.begin() + off, buf.begin() + off + 256 * 256);
    return true;
}

// (H, W, 3|4) float32 .npy -> RGBA32F texels, the role of load_dynamic_image + into_rgba32f (src/asset.rs:238-264)
bool load_skybox(const char* path, std::vector<float>& texels, uint32_t& w, uint32_t& h) {
    if (!path) return false;
    std::vector<uint8_t> buf;
    size_t off;
    std::string hdr;
    const std::string name(path);
    if (name.size() > 4 && name.compare(name.size() - 4, 4, ".hdr") == 0) {  // load_dynamic_image's HDR branch, src/asset.rs:240-253
        if (!read_file(path, buf) || rpt_decode_hdr(buf.data(), buf.size(), nullptr, &w, &h) != RPT_OK) return false;
        std::vector<float> rgb((size_t)w * h * 3);
        if (rpt_decode_hdr(buf.data(), buf.size(), rgb.data(), &w, &h) != RPT_OK) return false;
        texels.resize((size_t)w * h * 4);
        return rpt_sky_texels(rgb.data(), w, h, /*cpu_path_rgb8=*/0, texels.data()) == RPT_OK;  // Rgba32Float upload, src/asset.rs:257-264
    }
    if (!read_file(path, buf) || !npy_payload(buf, off, hdr) || hdr.find("<f4") == std::string::npos) return false;
    unsigned long sh, sw, sc;
    const size_t p = hdr.find("'shape': (");
    if (p == std::string::npos || std::sscanf(hdr.c_str() + p, "'shape': (%lu, %lu, %lu", &sh, &sw, &sc) != 3 || (sc != 3 && sc != 4)) return false;
    if (buf.size() - off < (size_t)sh * sw * sc * 4) return false;
    const float* src = reinterpret_cast<const float*>(buf.data() + off);
    texels.resize((size_t)sh * sw * 4);
    for (size_t i = 0; i < (size_t)sh * sw; ++i) {
        for (unsigned c = 0; c < 3; ++c) texels[4 * i + c] = src[sc * i + c];
        texels[4 * i + 3] = sc == 4 ? src[4 * i + 3] : 1.0f;
    }
    w = (uint32_t)sw;
    h = (uint32_t)sh;
    return true;
}

}  // namespace

RptTracingConfig default_config() {
    RptTracingConfig c{};
    c.cam_position[0] = 0.0f; c.cam_position[1] = 1.0f; c.cam_position[2] = -5.0f; c.cam_position[3] = 0.0f;
    c.width = 1280; c.height = 720; c.min_bounces = 3; c.max_bounces = 4;
    const float sx = 0.5f, sy = 1.3f, sz = 1.0f;
    const float inv = 1.0f / std::sqrt((sx * sx + sy * sy) + sz * sz);  // glam normalize: v * (1 / length)
    c.sun_direction[0] = sx * inv; c.sun_direction[1] = sy * inv; c.sun_direction[2] = sz * inv; c.sun_direction[3] = 15.0f;
    c.nee = 0; c.has_skybox = 0;
    c.specular_weight_clamp[0] = 0.1f; c.specular_weight_clamp[1] = 0.9f;
    return c;
}

TracingState::TracingState(uint32_t width, uint32_t height) : framebuffer((size_t)width * height * 3, 0.0f), config(default_config()) {
    config.width = width;
    config.height = height;
}

std::shared_ptr<TracingState> setup_trace(uint32_t width, uint32_t height, uint32_t samples) {
    auto state = std::make_shared<TracingState>(width, height);
    state->running.store(true, std::memory_order_relaxed);
    std::thread([state, samples] {
        while (state->samples.load(std::memory_order_relaxed) < samples) std::this_thread::yield();
        state->running.store(false, std::memory_order_relaxed);
    }).detach();
    return state;
}

// `.rptw`: "RPTW0001", u32 nverts, ntris, nmats, atlas_w, atlas_h, then vertices f32x4, normals f32x4,
// tangents f32x4, uvs f32x2, indices u32x4, materials 96 B each, atlas RGBA8 (see glb.py: save_rptw).
std::optional<World> World::from_path(const std::string& path) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf) || buf.size() < 28 || std::memcmp(buf.data(), "RPTW0001", 8) != 0) return std::nullopt;
    uint32_t hdr[5];
    std::memcpy(hdr, buf.data() + 8, 20);
    const size_t nv = hdr[0], nt = hdr[1], nm = hdr[2], aw = hdr[3], ah = hdr[4];
    const size_t need = 28 + nv * (16 + 16 + 16 + 8) + nt * 16 + nm * 96 + aw * ah * 4;
    if (buf.size() < need || nv == 0 || nt == 0 || nm == 0) return std::nullopt;
    const uint8_t* p = buf.data() + 28;
    const float* vertices = reinterpret_cast<const float*>(p); p += nv * 16;
    const float* normals = reinterpret_cast<const float*>(p); p += nv * 16;
    const float* tangents = reinterpret_cast<const float*>(p); p += nv * 16;
    const float* uvs = reinterpret_cast<const float*>(p); p += nv * 8;
    World w;
    w.index_buffer.assign(reinterpret_cast<const uint32_t*>(p), reinterpret_cast<const uint32_t*>(p) + nt * 4); p += nt * 16;
    w.material_data_buffer.resize(nm);
    std::memcpy(w.material_data_buffer.data(), p, nm * 96); p += nm * 96;
    if (aw * ah) { w.atlas.assign(p, p + aw * ah * 4); w.atlas_w = (uint32_t)aw; w.atlas_h = (uint32_t)ah; }

    // BVH building permutes the index buffer (src/asset.rs:195-196), then the light table reads it (:201-202)
    w.nodes.resize(2 * nt - 1);
    uint32_t nnodes = 0, nlights = 0;
    if (rpt_build_bvh(vertices, (uint32_t)nv, w.index_buffer.data(), (uint32_t)nt, 128, w.nodes.data(), &nnodes) != RPT_OK) return std::nullopt;
    w.nodes.resize(nnodes);
    w.light_pick_buffer.resize(nt);
    if (rpt_build_light_pick_table(vertices, (uint32_t)nv, w.index_buffer.data(), (uint32_t)nt, w.material_data_buffer.data(), (uint32_t)nm,
                                   w.light_pick_buffer.data(), &nlights) != RPT_OK) return std::nullopt;
    w.light_pick_buffer.resize(nlights);
    w.per_vertex_buffer.resize(nv);
    rpt_pack_per_vertex(vertices, normals, tangents, uvs, (uint32_t)nv, w.per_vertex_buffer.data());
    return w;
}

int trace_gpu(const std::string& scene_path, const char* skybox_path, std::shared_ptr<TracingState> state, int device) {
    auto world = World::from_path(scene_path);
    if (!world) return RPT_OK;  // `let Some(world) = ... else { return; }`
    std::vector<float> sky;
    uint32_t sky_w = 0, sky_h = 0;
    const bool has_sky = load_skybox(skybox_path, sky, sky_w, sky_h);

    const RptTracingConfig cfg0 = state->read_config();
    const uint32_t width = cfg0.width, height = cfg0.height;
    const size_t pixel_count = (size_t)width * height;
    std::vector<uint8_t> blue;
    std::vector<uint32_t> rng_blue(pixel_count * 2), rng_uniform(pixel_count * 2);
    if (!load_blue_noise(blue)) { std::fprintf(stderr, "trace_gpu: blue-noise tile not found\n"); return RPT_ERR_INVALID_ARGUMENT; }
    rpt_make_rng_seeds(blue.data(), 256, 256, width, height, 0, rng_blue.data());
    rpt_make_rng_seeds(nullptr, 0, 0, width, height, 0x5eed, rng_uniform.data());
    auto seeds = [&]() -> const uint32_t* { return state->use_blue_noise.load(std::memory_order_relaxed) ? rng_blue.data() : rng_uniform.data(); };

    rpt_context* ctx = nullptr;
    int rc = rpt_create(device, &ctx);
    auto fail = [&](const char* what) {
        std::fprintf(stderr, "trace_gpu: %s failed (%d): %s\n", what, rc, rpt_last_error(ctx));
        if (ctx) rpt_destroy(ctx);
        return rc;
    };
    if (rc != RPT_OK) return fail("rpt_create");
    rc = rpt_upload_world(ctx, world->per_vertex_buffer.data(), (uint32_t)world->per_vertex_buffer.size(), world->index_buffer.data(),
                          (uint32_t)(world->index_buffer.size() / 4), world->nodes.data(), (uint32_t)world->nodes.size(),
                          world->material_data_buffer.data(), (uint32_t)world->material_data_buffer.size(), world->light_pick_buffer.data(),
                          (uint32_t)world->light_pick_buffer.size(), world->atlas.empty() ? nullptr : world->atlas.data(), world->atlas_w,
                          world->atlas_h, has_sky ? sky.data() : nullptr, sky_w, sky_h);
    if (rc != RPT_OK) return fail("rpt_upload_world");
    if ((rc = rpt_set_config(ctx, &cfg0)) != RPT_OK) return fail("rpt_set_config");
    if ((rc = rpt_write_rng(ctx, seeds(), pixel_count)) != RPT_OK) return fail("rpt_write_rng");

    // restore previous state, if there is any (src/trace.rs:162-164)
    const float samples_init = (float)state->samples.load(std::memory_order_relaxed);
    if (samples_init > 0.0f) {
        std::vector<float> init(pixel_count * 4);
        std::shared_lock<std::shared_mutex> l(state->framebuffer_lock);
        for (size_t i = 0; i < pixel_count; ++i) {
            for (int c = 0; c < 3; ++c) init[4 * i + c] = state->framebuffer[3 * i + c] * samples_init;
            init[4 * i + 3] = samples_init;
        }
        if ((rc = rpt_write_output(ctx, init.data(), pixel_count)) != RPT_OK) return fail("rpt_write_output");
    }

    // the per-batch readback lands in page-locked staging memory (the reference reads through a mapped staging buffer too)
    float* image_buffer = nullptr;
    if ((rc = rpt_host_alloc(pixel_count * 3 * sizeof(float), reinterpret_cast<void**>(&image_buffer))) != RPT_OK) return fail("rpt_host_alloc");
    auto fail_loop = [&](const char* what) {
        rpt_host_free(image_buffer);
        return fail(what);
    };
    while (state->running.load(std::memory_order_relaxed)) {
        const uint32_t sync_rate = state->sync_rate.load(std::memory_order_relaxed);
        // The reference reads its control flags after EVERY sample of a batch (src/trace.rs:182-193): it leaves the
        // batch when `interacting | dirty` is set and abandons it when `running` is cleared.  Same here, sample by
        // sample, while somebody may be steering (a frame that already flushes takes one sample, as in the reference);
        // an undisturbed batch is one uninterrupted enqueue.
        bool flush = state->interacting.load(std::memory_order_relaxed) || state->dirty.load(std::memory_order_relaxed);
        uint32_t batch = 0;
        if (flush || state->poll_every_sample.load(std::memory_order_relaxed)) {
            while (batch < sync_rate) {
                if ((rc = rpt_enqueue(ctx, 1)) != RPT_OK) return fail_loop("rpt_enqueue");
                if ((rc = rpt_sync(ctx)) != RPT_OK) return fail_loop("rpt_sync");
                ++batch;
                flush |= state->interacting.load(std::memory_order_relaxed) || state->dirty.load(std::memory_order_relaxed);
                if (flush) break;
                if (!state->running.load(std::memory_order_relaxed)) { rpt_host_free(image_buffer); rpt_destroy(ctx); return RPT_OK; }
            }
        } else {
            batch = sync_rate;
            if ((rc = rpt_enqueue(ctx, batch)) != RPT_OK) return fail_loop("rpt_enqueue");
            if ((rc = rpt_sync(ctx)) != RPT_OK) return fail_loop("rpt_sync");
        }
        state->samples.fetch_add(batch, std::memory_order_relaxed);

        const float sample_count = (float)state->samples.load(std::memory_order_relaxed);
        if ((rc = rpt_read_framebuffer(ctx, image_buffer, pixel_count, sample_count)) != RPT_OK) return fail_loop("rpt_read_framebuffer");
        {
            std::unique_lock<std::shared_mutex> l(state->framebuffer_lock);
            state->framebuffer.assign(image_buffer, image_buffer + pixel_count * 3);
        }
        if (flush) {  // src/trace.rs:216-222
            state->dirty.store(false, std::memory_order_relaxed);
            state->samples.store(0, std::memory_order_relaxed);
            const RptTracingConfig cfg = state->read_config();
            if ((rc = rpt_set_config(ctx, &cfg)) != RPT_OK) return fail_loop("rpt_set_config");
            if ((rc = rpt_write_output(ctx, nullptr, pixel_count)) != RPT_OK) return fail_loop("rpt_write_output");
            if ((rc = rpt_write_rng(ctx, seeds(), pixel_count)) != RPT_OK) return fail_loop("rpt_write_rng");
        }
    }
    rpt_host_free(image_buffer);
    rpt_destroy(ctx);
    return RPT_OK;
}

}  // namespace rpt
