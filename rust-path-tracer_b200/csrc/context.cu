// context.cu — implementation of the C ABI in include/rpt_b200.h: one context = one B200.
//
// Owns every device allocation (the reference-layout scene copy used by the megakernel arm, the
// private wide-BVH / triangle-stream / light-record layouts used by the wavefront pipeline, the
// rng and accumulation buffers, and the path-state arrays of a wave), and drives the wavefront
// loop: per wave  generate -> { extend -> miss | shade -> shadow-connect } x max_bounces ->
// accumulate, all on one stream with queue lengths kept on the device (no host round-trips).
// Every failure becomes a status code + message; nothing throws or aborts across the ABI.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <thread>
#include <vector>

#include "../../include/rpt_b200.h"
#include "../../include/rpt_host.h"
#include "device_build.h"
#include "device_scene.h"
#include "wide_bvh.h"

using namespace rpt;

namespace {

thread_local std::string g_create_error;

// Path slots per wave.  Every stage is a persistent kernel whose last blocks drain alone, so long waves amortise
// the launch tails (measured: 4 Mi -> 16 Mi slots = +4% Mpaths/s on the 1080p workloads); 16 Mi slots are 2.6 GB of
// path state, nothing next to 180 GB of HBM.  kShadedSlotMask caps a wave below 2^30 slots.
constexpr uint32_t kDefaultWaveSlots = 1u << 24, kMaxWaveSlots = (1u << 30) - 1u;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* host, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

// ---- NCCL, resolved at run time so the library loads (and host-only tests run) without it ----
struct NcclUniqueId {  // ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
    char internal[128];
};
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (handle) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(handle, "ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(handle, "ncclCommInitRank"));
        Reduce = reinterpret_cast<decltype(Reduce)>(dlsym(handle, "ncclReduce"));
        Send = reinterpret_cast<decltype(Send)>(dlsym(handle, "ncclSend"));
        Recv = reinterpret_cast<decltype(Recv)>(dlsym(handle, "ncclRecv"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(handle, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(handle, "ncclGroupEnd"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !Reduce || !CommDestroy || !Send || !Recv || !GroupStart || !GroupEnd) { err = "libnccl is missing expected symbols"; return false; }
        return true;
    }
};
NcclApi g_nccl;

}  // namespace

struct rpt_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string error;
    int pipeline = RPT_PIPELINE_WAVEFRONT;
    uint32_t wave_slots = kDefaultWaveSlots;
    // trace-kernel tunables (defaults chosen on B200; RPT_* env vars override for sweeps)
    bool log_queues = false;  // RPT_LOG_QUEUES=1: print every bounce's queue lengths (syncs; for reading ncu captures)
    int trace_blocks_per_sm = 9;  // = the __launch_bounds__ of wf_trace_kernel: 56 registers, 36 warps per SM
    int refill_below = 20;
    // deferred triangle tests (wf_trace_deferred_kernel): RPT_DEFER_EXTEND / RPT_DEFER_SHADOW / RPT_FLUSH_AT / RPT_FLUSH_KEEP
    bool l2_persist = false;     // RPT_L2_PERSIST=1: access-policy window over the scene (measured: extend -0.8 %, shade +16 % — the set-aside is L2 the shading stage loses)
    size_t l2_window_bytes = 0;
    bool defer_extend = false, defer_shadow = false;  // measured: -2 % extend time on the 1M-triangle proxy, +3.5 % on DarkCornell — off
    int flush_at = 12, flush_keep = 6;
    // Paths whose throughput is exactly zero are retired instead of traced to their first roulette (RPT_KEEP_DEAD_PATHS=1
    // keeps them).  Exact whenever nothing such a path can still meet is non-finite — 0 x inf would be a NaN the
    // reference adds to the pixel: sky texels and sun parameters are checked (sources_finite, frame_params); what is
    // not covered is a dead path that later samples a NaN direction on degenerate geometry and escapes to an HDR sky,
    // and the 2^-32 roulette draw of exactly 0 that would turn a dead path's throughput into NaN.
    bool retire_dead_paths = true;
    bool sources_finite = true;  // set by rpt_upload_world: sky texels finite, scene bounds finite and below 1e15

    // scene, reference layouts (megakernel arm)
    DevBuf<RptPerVertexData> d_vertices;
    DevBuf<uint4> d_triangles;
    DevBuf<RptBVHNode> d_nodes;
    DevBuf<RptMaterialData> d_materials;
    DevBuf<RptLightPickEntry> d_lights;
    DevBuf<uchar4> d_atlas;
    DevBuf<float4> d_sky;
    uint32_t atlas_w = 1, atlas_h = 1, sky_w = 2, sky_h = 2, nlights = 0, nmaterials = 0;
    // scene, private layouts (wavefront arm)
    DeviceBuildResult tree;  // wide nodes, triangle position stream, shading records, leaf-order maps, refit scratch (owned)
    uint32_t shade_stride = kShadeStridePlain;
    uint32_t ntriangles = 0, nvertices = 0;
    bool device_built = false;  // the tree came from device_build_wide_bvh (no reference BVH: the megakernel arm cannot run)
    // host copies a refit needs to rebuild the light records (small next to the scene: 16 B per triangle)
    std::vector<uint32_t> h_triangles, h_wide_index;
    std::vector<RptMaterialData> h_materials;
    std::vector<RptLightPickEntry> h_lights;
    DevBuf<LightBin> d_light_bins;
    DevBuf<LightRecord> d_light_records;
    uint32_t nbins = 0;
    bool deep_tree = false;  // the wide tree needs more stack entries than the shared slab holds
    bool has_world = false;

    // render state
    RptTracingConfig config{};
    bool has_config = false;
    Camera camera{};
    float sky_yaw_sin = 0.0f, sky_yaw_cos = 1.0f;
    DevBuf<uint2> d_rng;
    DevBuf<float4> d_output;
    DevBuf<float> d_rgb;
    DevBuf<uint32_t> d_rgba8;
    DevBuf<uint32_t> d_ids;
    bool rng_written = false;
    uint32_t tile_rank = 0, tile_count = 1;
    DevBuf<uint32_t> d_pixel_map;
    uint32_t pixel_map_len = 0;

    // wave state
    uint32_t wave_capacity = 0;
    DevBuf<float4> w_ray_o, w_ray_d, w_thr, w_rad, w_sh_o, w_sh_d, w_sh_c;
    DevBuf<uint2> w_hit, w_stack_overflow;
    DevBuf<uint32_t> w_q0, w_q1, w_qhit, w_qmiss, w_qshaded, w_qshadow;
    DevBuf<WaveCtl> w_ctl;
    DevBuf<unsigned long long> d_counters;

    uint64_t kernel_launches = 0;
    uint64_t mega_paths = 0;

    // CUDA graphs of whole waves.  A wave is a fixed sequence of ~30 launches whose arguments only depend on the
    // scene, the config, the buffers and the wave's shape, so it is captured once and replayed: one launch per
    // wave instead of thirty, which is what bounds small frames (a 128x128 x 32 spp batch is ~150 us of GPU work).
    // `graph_epoch` moves whenever something baked into the captured arguments changes.
    // A shape is captured the second time it runs in an epoch: a host that changes the config before every batch
    // (camera motion: `interacting` flushes after each sample, src/trace.rs:187-193) never pays for a capture.
    struct WaveGraph { WaveDesc desc; uint64_t epoch; cudaGraphExec_t exec; uint64_t launches; };
    std::vector<WaveGraph> wave_graphs;
    std::vector<WaveDesc> shapes_run_once;
    uint64_t graph_epoch = 0;
    bool use_graphs = true;  // RPT_GRAPHS=0 turns them off
    void drop_graphs() {
        for (auto& g : wave_graphs) cudaGraphExecDestroy(g.exec);
        wave_graphs.clear();
        shapes_run_once.clear();
        ++graph_epoch;
    }
    bool stage_timing = false;
    bool trace_statistics = false;  // rpt_set_trace_statistics: launch the counting build of the trace kernels
    struct StageEvent { int stage; cudaEvent_t e0, e1; };
    std::vector<StageEvent> stage_events;
    RptStageTiming stage_totals{};

    // Run `launch` as one counted kernel launch of `stage`, bracketed by events when stage timing is on.
    template <class Fn>
    void launch(int stage, Fn&& fn) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        const bool timed_launch = stage_timing && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
        if (timed_launch) cudaEventRecord(e0, stream);
        fn();
        if (timed_launch) { cudaEventRecord(e1, stream); stage_events.push_back({stage, e0, e1}); }
        kernel_launches++;
    }
    void drain_stage_events() {
        for (auto& ev : stage_events) {
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, ev.e0, ev.e1) == cudaSuccess) { stage_totals.ms[ev.stage] += t; stage_totals.launches[ev.stage]++; }
            cudaEventDestroy(ev.e0);
            cudaEventDestroy(ev.e1);
        }
        stage_events.clear();
    }
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
    float device_ms = 0.0f;
    cudaEvent_t region_start = nullptr, region_stop = nullptr;  // rpt_timer_start / rpt_timer_stop
    // asynchronous readback (rpt_read_framebuffer_async): two device frames, filled on the render stream and copied out
    // on `copy_stream`, so the next batch renders while the previous frame crosses PCIe
    cudaStream_t copy_stream = nullptr;
    DevBuf<float> d_rgb_async[2];
    cudaEvent_t ev_frame_ready[2] = {nullptr, nullptr}, ev_frame_copied[2] = {nullptr, nullptr};
    int async_next = 0;
    // post-normalise hook (the reference's denoiser slot, src/trace.rs:207-210)
    rpt_frame_hook frame_hook = nullptr;
    void* frame_hook_user = nullptr;

    void* nccl_comm = nullptr;
    int nccl_rank = 0, nccl_nranks = 1;
    // The combine never touches an accumulator: it snapshots this rank's contribution on the render stream, moves it
    // over NVLink on `comm_stream` — overlapped with whatever the render stream does next — and leaves the combined frame
    // in `d_combined` on the root, which the read calls then return (until rpt_write_output / a resize drops it).
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_snapshot = nullptr, ev_combined = nullptr;
    DevBuf<float4> d_snapshot, d_combined, d_gather;  // d_gather (root, tile mode): every rank's packed tiles, rank after rank
    bool combined_valid = false;
    struct RankTiles { uint32_t count; DevBuf<uint32_t> map; };
    std::vector<RankTiles> gather_maps;  // root, tile mode: the pixel map of every rank
    uint32_t gather_w = 0, gather_h = 0;
    void drop_gather_maps() {
        for (auto& t : gather_maps) t.map.release();
        gather_maps.clear();
    }
    // the frame the read calls return: the combined one after a combine (waits for it), else this context's accumulator
    const float4* frame_for_read() {
        if (!combined_valid) return d_output.p;
        cudaStreamWaitEvent(stream, ev_combined, 0);
        return d_combined.p;
    }

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        error = buf;
        return code;
    }
    int cuda(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return RPT_OK;
        return fail(RPT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    uint32_t npixels() const { return config.width * config.height; }
};

#define RPT_TRY(expr)                        \
    do {                                     \
        const int rpt_status_ = (expr);      \
        if (rpt_status_ != RPT_OK) return rpt_status_; \
    } while (0)
#define RPT_CUDA(ctx, expr) RPT_TRY((ctx)->cuda((expr), #expr))

namespace {

int bind_device(rpt_context* c) { return c->cuda(cudaSetDevice(c->device), "cudaSetDevice"); }

FrameParams frame_params(const rpt_context* c) {
    FrameParams f{};
    f.camera = c->camera;
    f.width = c->config.width;
    f.height = c->config.height;
    f.min_bounces = c->config.min_bounces;
    f.max_bounces = c->config.max_bounces;
    f.nee = c->config.nee <= 2 ? c->config.nee : 0;  // NextEventEstimation::from_u32
    f.has_skybox = c->config.has_skybox;
    f.clamp_lo = c->config.specular_weight_clamp[0];
    f.clamp_hi = c->config.specular_weight_clamp[1];
    f.sun_dir = mk3(c->config.sun_direction[0], c->config.sun_direction[1], c->config.sun_direction[2]);
    f.sun_intensity = c->config.sun_direction[3];
    f.sky = SkyImage{c->d_sky.p, c->sky_w, c->sky_h, pow2_mask(c->sky_w), pow2_mask(c->sky_h), c->sky_yaw_sin, c->sky_yaw_cos, c->config.sun_direction[3] * (1.0f / 15.0f)};
    f.atlas = Atlas{c->d_atlas.p, c->atlas_w, c->atlas_h, pow2_mask(c->atlas_w), pow2_mask(c->atlas_h)};
    f.tile_rank = c->tile_rank;
    f.tile_count = c->tile_count;
    f.retire_dead_paths = c->retire_dead_paths && c->sources_finite && std::isfinite(c->config.sun_direction[0]) && std::isfinite(c->config.sun_direction[1]) &&
                          std::isfinite(c->config.sun_direction[2]) && std::isfinite(c->config.sun_direction[3]);
    return f;
}

MegaParams mega_params(const rpt_context* c) {
    const FrameParams f = frame_params(c);
    MegaParams p{};
    p.camera = f.camera;
    p.width = f.width; p.height = f.height; p.min_bounces = f.min_bounces; p.max_bounces = f.max_bounces;
    p.nee = f.nee; p.has_skybox = f.has_skybox; p.clamp_lo = f.clamp_lo; p.clamp_hi = f.clamp_hi;
    p.sun_dir = f.sun_dir; p.sun_intensity = f.sun_intensity; p.sky = f.sky;
    p.atlas = c->d_atlas.p; p.atlas_w = c->atlas_w; p.atlas_h = c->atlas_h;
    p.nodes = reinterpret_cast<const float4*>(c->d_nodes.p);
    p.triangles = c->d_triangles.p;
    p.vertices = reinterpret_cast<const float4*>(c->d_vertices.p);
    p.materials = c->d_materials.p;
    p.lights = c->d_lights.p;
    p.nlights = c->nlights;
    p.rng = c->d_rng.p;
    p.output = c->d_output.p;
    p.counters = c->d_counters.p;
    p.tile_rank = c->tile_rank;
    p.tile_count = c->tile_count;
    return p;
}

WideWorld wide_world(const rpt_context* c) {
    WideWorld w{};
    w.bvh = WideScene{c->tree.nodes, c->tree.tri_pos, kHalf1024Bytes};
    w.tri_shade = c->tree.tri_shade;
    w.shade_stride = c->shade_stride;
    w.materials = c->d_materials.p;
    w.nmaterials = c->nmaterials;
    w.light_bins = c->d_light_bins.p;
    w.nbins = c->nbins;
    w.lights = c->d_light_records.p;
    w.nlights = (uint32_t)c->d_light_records.n;
    return w;
}

WaveState wave_state(rpt_context* c) {
    WaveState s{};
    s.ray_o = c->w_ray_o.p; s.ray_d = c->w_ray_d.p; s.thr = c->w_thr.p; s.rad = c->w_rad.p;
    s.hit = c->w_hit.p;
    s.sh_o = c->w_sh_o.p; s.sh_d = c->w_sh_d.p; s.sh_c = c->w_sh_c.p;
    s.q_ext[0] = c->w_q0.p; s.q_ext[1] = c->w_q1.p; s.q_hit = c->w_qhit.p; s.q_miss = c->w_qmiss.p;
    s.q_shaded = c->w_qshaded.p; s.q_shadow = c->w_qshadow.p;
    s.ctl = c->w_ctl.p;
    s.counters = c->d_counters.p;
    return s;
}

void release_wave(rpt_context* c) {
    c->w_ray_o.release(); c->w_ray_d.release(); c->w_thr.release(); c->w_rad.release();
    c->w_sh_o.release(); c->w_sh_d.release(); c->w_sh_c.release(); c->w_hit.release();
    c->w_q0.release(); c->w_q1.release(); c->w_qhit.release(); c->w_qmiss.release(); c->w_qshaded.release(); c->w_qshadow.release();
    c->wave_capacity = 0;
}

int ensure_wave(rpt_context* c, uint32_t slots) {
    if (slots <= c->wave_capacity) return RPT_OK;
    c->drop_graphs();  // the path-state buffers move
    // All or nothing: a failed allocation leaves NO wave buffers (capacity 0), so the next call allocates again
    // instead of launching on a half-resized set.
    cudaError_t e = cudaSuccess;
    auto grow = [&](auto& buf) { if (e == cudaSuccess) e = buf.alloc(slots); };
    grow(c->w_ray_o); grow(c->w_ray_d); grow(c->w_thr); grow(c->w_rad);
    grow(c->w_sh_o); grow(c->w_sh_d); grow(c->w_sh_c); grow(c->w_hit);
    grow(c->w_q0); grow(c->w_q1); grow(c->w_qhit); grow(c->w_qmiss); grow(c->w_qshaded); grow(c->w_qshadow);
    if (e == cudaSuccess && !c->w_ctl.p) e = c->w_ctl.alloc(1);
    // (only for trees deeper than the shared-memory part of the traversal stack)
    if (e == cudaSuccess && c->deep_tree && !c->w_stack_overflow.p)
        e = c->w_stack_overflow.alloc(trace_stack_overflow_entries(c->sm_count * c->trace_blocks_per_sm));
    if (e != cudaSuccess) {
        release_wave(c);
        cudaGetLastError();
        return c->fail(RPT_ERR_CUDA, "path-state allocation for %u slots: %s", slots, cudaGetErrorString(e));
    }
    c->wave_capacity = slots;
    return RPT_OK;
}

// Worst-case R-sequence dimensions a path can consume (kernels/src/rng.rs:19-32 holds 32 primes,
// index 0 unused): 2 for the camera, then per bounce 3 (BSDF) + 4 (NEE) + 1 (Russian roulette).
uint32_t rng_dimension_budget(const RptTracingConfig& cfg) {
    uint64_t dims = 2;
    for (uint32_t b = 0; b < cfg.max_bounces && dims < 1000; ++b) dims += 3u + (cfg.nee != 0 && cfg.nee <= 2 ? 4u : 0u) + (b > cfg.min_bounces ? 1u : 0u);
    return (uint32_t)std::min<uint64_t>(dims, 1000);
}

int rebuild_pixel_map(rpt_context* c) {
    c->d_pixel_map.release();
    c->pixel_map_len = 0;
    if (c->tile_count <= 1 || !c->has_config) return RPT_OK;
    uint32_t count = 0;
    rpt_tile_partition_pixels(c->config.width, c->config.height, c->tile_rank, c->tile_count, nullptr, &count);
    std::vector<uint32_t> map(count);
    if (count) rpt_tile_partition_pixels(c->config.width, c->config.height, c->tile_rank, c->tile_count, map.data(), &count);
    c->pixel_map_len = (uint32_t)map.size();
    if (!map.empty()) {
        RPT_CUDA(c, c->d_pixel_map.upload(map.data(), map.size(), c->stream));
        RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return RPT_OK;
}

int check_ready(rpt_context* c) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_world) return c->fail(RPT_ERR_NOT_READY, "no world uploaded (rpt_upload_world)");
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "no config set (rpt_set_config)");
    if (!c->rng_written) return c->fail(RPT_ERR_NOT_READY, "rng seeds not written (rpt_write_rng)");
    return bind_device(c);
}

// One wave of the wavefront pipeline; `primary_only` stops after the first extend (diagnostics).
int run_wave(rpt_context* c, const WaveDesc& d, bool primary_only, uint32_t* ids_out) {
    const FrameParams f = frame_params(c);
    const WideWorld w = wide_world(c);
    const WaveState s = wave_state(c);
    const WaveLaunch l{c->sm_count, c->stream, c->trace_blocks_per_sm, c->refill_below, c->deep_tree ? c->w_stack_overflow.p : nullptr,
                       c->defer_extend && !c->trace_statistics, c->defer_shadow && !c->trace_statistics, c->flush_at, std::max(1, c->flush_keep),
                       c->trace_statistics};
    const uint32_t nslots = d.npix * d.k_samples;
    c->launch(RPT_STAGE_OTHER, [&] { launch_wf_reset(l, s, 1, true); });
    c->launch(RPT_STAGE_GENERATE, [&] { launch_wf_generate(l, f, s, d, c->d_rng.p); });
    const uint32_t bounces = primary_only ? 1u : f.max_bounces;
    for (uint32_t b = 0; b < bounces; ++b) {
        const int cur = (int)(b & 1u), nxt = cur ^ 1;
        if (b > 0) c->launch(RPT_STAGE_OTHER, [&] { launch_wf_reset(l, s, nxt, false); });
        c->launch(RPT_STAGE_EXTEND, [&] { launch_wf_extend(l, w.bvh, s, cur, b == 0, nslots); });
        c->kernel_launches++;  // extend = trace + queue compaction
        if (primary_only) {
            c->launch(RPT_STAGE_OTHER, [&] { launch_wf_export_primary(l, w.bvh, s, d, ids_out); });
            c->kernel_launches++;  // the export is two kernels
            break;
        }
        c->launch(RPT_STAGE_MISS, [&] { launch_wf_miss(l, f, s); });
        c->launch(RPT_STAGE_SHADE, [&] { launch_wf_shade(l, f, w, s, d, c->d_rng.p, b); launch_wf_compact_shaded(l, s, nxt); });
        c->kernel_launches++;  // shade = shading + queue compaction
        if (f.nee != RPT_NEE_NONE && c->nbins > 0) c->launch(RPT_STAGE_SHADOW, [&] { launch_wf_shadow(l, w.bvh, s); });
        if (c->log_queues) {
            WaveCtl ctl{};
            cudaMemcpyAsync(&ctl, c->w_ctl.p, sizeof ctl, cudaMemcpyDeviceToHost, c->stream);
            cudaStreamSynchronize(c->stream);
            std::fprintf(stderr, "[rpt] bounce %u: extend traced %u rays -> %u hits, %u misses; %u shadow rays; %u paths go on\n", b,
                         b == 0 ? nslots : ctl.n_ext[cur], ctl.n_hit, ctl.n_miss, ctl.n_shadow, ctl.n_ext[nxt]);
        }
    }
    if (!primary_only) c->launch(RPT_STAGE_ACCUMULATE, [&] { launch_wf_accumulate(l, s, d, c->d_rng.p, c->d_output.p); });
    return c->cuda(cudaGetLastError(), "wavefront launch");
}

// run_wave through a captured graph (captured on first use of this wave shape).
int run_wave_graphed(rpt_context* c, const WaveDesc& d) {
    if (!c->use_graphs || c->stage_timing || c->log_queues) return run_wave(c, d, false, nullptr);
    for (auto& g : c->wave_graphs)
        if (g.epoch == c->graph_epoch && g.desc.pix_base == d.pix_base && g.desc.npix == d.npix && g.desc.k_samples == d.k_samples &&
            g.desc.pixel_map == d.pixel_map) {
            c->kernel_launches += g.launches;
            return c->cuda(cudaGraphLaunch(g.exec, c->stream), "cudaGraphLaunch");
        }
    auto same = [&](const WaveDesc& o) { return o.pix_base == d.pix_base && o.npix == d.npix && o.k_samples == d.k_samples && o.pixel_map == d.pixel_map; };
    if (std::none_of(c->shapes_run_once.begin(), c->shapes_run_once.end(), same)) {
        if (c->shapes_run_once.size() >= 64) c->shapes_run_once.clear();
        c->shapes_run_once.push_back(d);
        return run_wave(c, d, false, nullptr);
    }
    if (c->wave_graphs.size() >= 16) c->drop_graphs();  // (a frame is a handful of wave shapes)
    const uint64_t before = c->kernel_launches;
    RPT_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int status = run_wave(c, d, false, nullptr);
    cudaGraph_t graph = nullptr;
    const cudaError_t end = cudaStreamEndCapture(c->stream, &graph);
    if (status != RPT_OK || end != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        c->kernel_launches = before;
        return status != RPT_OK ? status : c->cuda(end != cudaSuccess ? end : cudaErrorUnknown, "wave graph capture");
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) { c->kernel_launches = before; return c->cuda(inst, "cudaGraphInstantiate"); }
    c->wave_graphs.push_back({d, c->graph_epoch, exec, c->kernel_launches - before});
    return c->cuda(cudaGraphLaunch(exec, c->stream), "cudaGraphLaunch");
}

template <class Fn>
int for_each_wave(rpt_context* c, uint32_t n_samples, Fn&& fn) {
    const uint32_t total = c->tile_count > 1 ? c->pixel_map_len : c->npixels();
    if (total == 0) return RPT_OK;
    const uint32_t slots = std::min(std::max<uint32_t>(c->wave_slots, 1024u), kMaxWaveSlots);
    const uint32_t chunk = std::min(total, slots);
    for (uint32_t base = 0; base < total; base += chunk) {
        const uint32_t np = std::min(chunk, total - base);
        const uint32_t kmax = std::max(1u, slots / np);
        for (uint32_t s0 = 0; s0 < n_samples; s0 += kmax) {
            WaveDesc d{base, np, std::min(kmax, n_samples - s0), c->tile_count > 1 ? c->d_pixel_map.p : nullptr};
            RPT_TRY(ensure_wave(c, d.npix * d.k_samples));
            RPT_TRY(fn(d));
        }
    }
    return RPT_OK;
}

}  // namespace

// ============================================================================ lifetime
extern "C" int rpt_create(int device_id, rpt_context** out_ctx) {
    if (!out_ctx) return RPT_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                         "); this backend has no CPU fallback";
        cudaGetLastError();
        return RPT_ERR_NO_DEVICE;
    }
    if (device_id < 0 || device_id >= count) {
        g_create_error = "device id out of range";
        return RPT_ERR_INVALID_ARGUMENT;
    }
    rpt_context* c = new (std::nothrow) rpt_context();
    if (!c) return RPT_ERR_CUDA;
    c->device = device_id;
    cudaDeviceProp prop{};
    if ((e = cudaSetDevice(device_id)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess || (e = c->d_counters.alloc(8)) != cudaSuccess ||
        (e = cudaMemsetAsync(c->d_counters.p, 0, 8 * sizeof(unsigned long long), c->stream)) != cudaSuccess) {
        g_create_error = std::string("device initialisation failed: ") + cudaGetErrorString(e);
        delete c;
        return RPT_ERR_CUDA;
    }
    if (prop.major < 10) {
        g_create_error = std::string("device '") + prop.name + "' is not sm_100-class; this library ships sm_100a code only";
        cudaStreamDestroy(c->stream);
        delete c;
        return RPT_ERR_NO_DEVICE;
    }
    c->sm_count = prop.multiProcessorCount;
    if (const char* v = getenv("RPT_TRACE_BLOCKS_PER_SM")) c->trace_blocks_per_sm = std::max(1, atoi(v));
    if (const char* v = getenv("RPT_REFILL_BELOW")) c->refill_below = atoi(v);
    if (const char* v = getenv("RPT_LOG_QUEUES")) c->log_queues = atoi(v) != 0;
    if (const char* v = getenv("RPT_L2_PERSIST")) c->l2_persist = atoi(v) != 0;
    if (const char* v = getenv("RPT_DEFER_EXTEND")) c->defer_extend = atoi(v) != 0;
    if (const char* v = getenv("RPT_DEFER_SHADOW")) c->defer_shadow = atoi(v) != 0;
    if (const char* v = getenv("RPT_FLUSH_AT")) c->flush_at = atoi(v);
    if (const char* v = getenv("RPT_FLUSH_KEEP")) c->flush_keep = atoi(v);
    if (const char* v = getenv("RPT_GRAPHS")) c->use_graphs = atoi(v) != 0;
    if (const char* v = getenv("RPT_KEEP_DEAD_PATHS")) c->retire_dead_paths = atoi(v) == 0;
    if (const char* v = getenv("RPT_WAVE_SLOTS")) c->wave_slots = (uint32_t)std::max(1024, atoi(v));
    *out_ctx = c;
    return RPT_OK;
}

extern "C" int rpt_destroy(rpt_context* c) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->drop_graphs();
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    if (c->ev_snapshot) cudaEventDestroy(c->ev_snapshot);
    if (c->ev_combined) cudaEventDestroy(c->ev_combined);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    c->drop_gather_maps();
    c->d_snapshot.release(); c->d_combined.release(); c->d_gather.release();
    for (auto& ev : c->timed) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    if (c->region_start) { cudaEventDestroy(c->region_start); cudaEventDestroy(c->region_stop); }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(c->ev_frame_ready[k]); cudaEventDestroy(c->ev_frame_copied[k]); c->d_rgb_async[k].release(); }
        cudaStreamDestroy(c->copy_stream);
    }
    c->drain_stage_events();
    c->tree.release();
    for (auto* b : {&c->w_ray_o, &c->w_ray_d, &c->w_thr, &c->w_rad, &c->w_sh_o, &c->w_sh_d, &c->w_sh_c, &c->d_sky, &c->d_output})
        b->release();
    c->w_hit.release(); c->w_stack_overflow.release(); c->w_q0.release(); c->w_q1.release(); c->w_qhit.release(); c->w_qmiss.release(); c->w_qshaded.release();
    c->w_qshadow.release(); c->w_ctl.release();
    c->d_counters.release(); c->d_vertices.release(); c->d_triangles.release(); c->d_nodes.release(); c->d_materials.release();
    c->d_lights.release(); c->d_atlas.release(); c->d_light_bins.release(); c->d_light_records.release();
    c->d_rng.release(); c->d_rgb.release(); c->d_rgba8.release(); c->d_ids.release(); c->d_pixel_map.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RPT_OK;
}

extern "C" const char* rpt_last_error(const rpt_context* c) { return c ? c->error.c_str() : g_create_error.c_str(); }

extern "C" int rpt_set_pipeline(rpt_context* c, int pipeline) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (pipeline != RPT_PIPELINE_WAVEFRONT && pipeline != RPT_PIPELINE_MEGAKERNEL) return c->fail(RPT_ERR_INVALID_ARGUMENT, "unknown pipeline %d", pipeline);
    c->pipeline = pipeline;
    return RPT_OK;
}

extern "C" int rpt_set_wave_slots(rpt_context* c, uint32_t slots) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    c->wave_slots = slots == 0 ? kDefaultWaveSlots : slots;
    return RPT_OK;
}

// ============================================================================ scene
namespace {
// fn(begin, end) over consecutive runs of [0, n) on the host threads (RPT_BUILD_THREADS, like the BVH builders).
template <class Fn>
void host_parallel_for(uint32_t n, Fn&& fn) {
    unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    if (const char* v = getenv("RPT_BUILD_THREADS")) threads = (unsigned)std::max(1, atoi(v));
    threads = std::min<unsigned>(threads, n / 65536u);  // small scenes: not worth a thread launch
    if (threads <= 1) { fn(0u, n); return; }
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < threads; ++w)
        pool.emplace_back([&fn, n, w, threads] { fn((uint32_t)((uint64_t)n * w / threads), (uint32_t)((uint64_t)n * (w + 1) / threads)); });
    fn(0u, (uint32_t)((uint64_t)n / threads));
    for (std::thread& t : pool) t.join();
}
}  // namespace

namespace {

// light-pick table -> bins over compact light records (a one-entry table with ratio < 0 is the sentinel: no bins).
// `wide_index`: caller's triangle index -> leaf order.
int build_light_records(rpt_context* c, const RptPerVertexData* vertices, const uint32_t* triangles, uint32_t ntriangles, const RptMaterialData* materials,
                        const RptLightPickEntry* lights, uint32_t nlights, const uint32_t* wide_index, std::vector<LightBin>& bins, std::vector<LightRecord>& records) {
    bins.clear();
    records.clear();
    if (nlights == 1 && lights[0].ratio < 0.0f) return RPT_OK;
    std::unordered_map<uint32_t, uint32_t> record_of;
    int status = RPT_OK;
    auto record = [&](uint32_t tri_index, float area, float pdf) -> uint32_t {
        auto it = record_of.find(tri_index);
        if (it != record_of.end()) return it->second;
        if (tri_index >= ntriangles) { status = RPT_ERR_INVALID_ARGUMENT; return 0; }
        const uint32_t* tri = triangles + 4 * (size_t)tri_index;
        const RptPerVertexData &a = vertices[tri[0]], &b = vertices[tri[1]], &cc = vertices[tri[2]];
        LightRecord r{};
        r.a_area = make_float4(a.vertex[0], a.vertex[1], a.vertex[2], area);
        r.e1_pdf = make_float4(b.vertex[0] - a.vertex[0], b.vertex[1] - a.vertex[1], b.vertex[2] - a.vertex[2], pdf);
        const uint32_t wi = wide_index[tri_index];
        float wbits;
        std::memcpy(&wbits, &wi, 4);
        r.e2_tri = make_float4(cc.vertex[0] - a.vertex[0], cc.vertex[1] - a.vertex[1], cc.vertex[2] - a.vertex[2], wbits);
        r.normal = make_float4(((a.normal[0] + b.normal[0]) + cc.normal[0]) / 3.0f, ((a.normal[1] + b.normal[1]) + cc.normal[1]) / 3.0f,
                               ((a.normal[2] + b.normal[2]) + cc.normal[2]) / 3.0f, 0.0f);
        const float* em = materials[tri[3]].emissive;
        r.emission = make_float4(em[0], em[1], em[2], 0.0f);
        const uint32_t id = (uint32_t)records.size();
        records.push_back(r);
        record_of.emplace(tri_index, id);
        return id;
    };
    bins.reserve(nlights);
    for (uint32_t i = 0; i < nlights; ++i) {
        const RptLightPickEntry& e = lights[i];
        LightBin bin{};
        bin.light_a = record(e.triangle_index_a, e.triangle_area_a, e.triangle_pick_pdf_a);
        // entries that were never topped up keep index_b = 0 with probability_b = 0: ratio == 1, never picked
        bin.light_b = e.ratio >= 1.0f ? bin.light_a : record(e.triangle_index_b, e.triangle_area_b, e.triangle_pick_pdf_b);
        bin.ratio = e.ratio;
        bins.push_back(bin);
    }
    if (status != RPT_OK) return c->fail(status, "light-pick table references a triangle out of range");
    if (records.size() >= (1u << 23)) return c->fail(RPT_ERR_UNSUPPORTED, "more than 2^23 emissive triangles (the path state keeps the sampled light in 23 bits)");
    return RPT_OK;
}

// The trace kernels' scene data (wide nodes + triangle positions: 63 MB for a million triangles) fits the 126 MB L2, but
// the shading stage streams gigabytes of path state, shading records and texels through it between two extend
// launches.  An access-policy window marks the scene range as PERSISTING in the L2 set-aside, so node and triangle
// fetches that miss the L1 keep hitting the L2.  Best effort: a device without the feature just runs without it.
void pin_scene_in_l2(rpt_context* c, const void* base, size_t bytes) {
    c->l2_window_bytes = 0;
    if (!c->l2_persist || !base || bytes == 0) return;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) { cudaGetLastError(); return; }
    const size_t window = std::min(bytes, (size_t)prop.accessPolicyMaxWindowSize);
    const size_t carve = std::min(window, (size_t)prop.persistingL2CacheMaxSize);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    attr.accessPolicyWindow.num_bytes = window;
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)window);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) { cudaGetLastError(); return; }
    c->l2_window_bytes = window;
}

int upload_light_records(rpt_context* c, const std::vector<LightBin>& bins, const std::vector<LightRecord>& records) {
    c->d_light_bins.release();
    c->d_light_records.release();
    if (!bins.empty()) {
        RPT_CUDA(c, c->d_light_bins.upload(bins.data(), bins.size(), c->stream));
        RPT_CUDA(c, c->d_light_records.upload(records.data(), records.size(), c->stream));
        RPT_CUDA(c, cudaStreamSynchronize(c->stream));  // (the vectors are the caller's locals)
    }
    c->nbins = (uint32_t)bins.size();
    return RPT_OK;
}

}  // namespace

// nodes == NULL (and nnodes == 0): no reference BVH is supplied — the wide tree is built on the device
// (device_build.cu); triangle ids still refer to the caller's index buffer.
extern "C" int rpt_upload_world(rpt_context* c, const RptPerVertexData* vertices, uint32_t nvertices, const uint32_t* triangles,
                                uint32_t ntriangles, const RptBVHNode* nodes, uint32_t nnodes, const RptMaterialData* materials,
                                uint32_t nmaterials, const RptLightPickEntry* lights, uint32_t nlights, const uint8_t* atlas_rgba8,
                                uint32_t atlas_w, uint32_t atlas_h, const float* sky_rgba32f, uint32_t sky_w, uint32_t sky_h) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    const bool build_on_device = nodes == nullptr && nnodes == 0;
    if (!vertices || !triangles || (!nodes && !build_on_device) || !materials || !lights || nvertices == 0 || ntriangles == 0 || (nnodes == 0 && !build_on_device) ||
        nmaterials == 0 || nlights == 0)
        return c->fail(RPT_ERR_INVALID_ARGUMENT, "rpt_upload_world: null or empty buffer");
    if (ntriangles >= 0x80000000u) return c->fail(RPT_ERR_UNSUPPORTED, "more than 2^31 triangles");
    if ((atlas_rgba8 && (atlas_w == 0 || atlas_h == 0)) || (sky_rgba32f && (sky_w == 0 || sky_h == 0)))
        return c->fail(RPT_ERR_INVALID_ARGUMENT, "rpt_upload_world: image with a zero dimension");
    for (size_t t = 0; t < (size_t)ntriangles; ++t)
        if (triangles[4 * t + 3] >= nmaterials) return c->fail(RPT_ERR_INVALID_ARGUMENT, "triangle %zu references material %u of %u", t, triangles[4 * t + 3], nmaterials);
    RPT_TRY(bind_device(c));

    bool any_normal_map = false;
    for (uint32_t m = 0; m < nmaterials; ++m) any_normal_map |= materials[m].has_normal_texture != 0;
    const bool textured = [&] {
        for (uint32_t m = 0; m < nmaterials; ++m)
            if (materials[m].has_albedo_texture || materials[m].has_metallic_texture || materials[m].has_roughness_texture || materials[m].has_normal_texture) return true;
        return false;
    }();
    if (textured && !atlas_rgba8) return c->fail(RPT_ERR_INVALID_ARGUMENT, "a material is textured but no atlas was supplied");
    const uint32_t shade_stride = any_normal_map ? kShadeStrideTangents : kShadeStridePlain;

    // ---- private re-layout.  Nothing of the context is touched until the input has been validated and every
    // host-side layout built: a rejected scene leaves the previous world in place.
    WideBvh wide;
    UninitVector<float> shade;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
    if (!build_on_device) {
        const char* err = "";
        if (!build_wide_bvh(nodes, nnodes, triangles, ntriangles, vertices, nvertices, wide, &err)) return c->fail(RPT_ERR_INVALID_ARGUMENT, "rpt_upload_world: %s", err);
        if (wide.max_depth + 1 > kWideStackCapacity)
            return c->fail(RPT_ERR_UNSUPPORTED, "wide BVH depth %u exceeds the traversal stack (%u)", wide.max_depth, kWideStackCapacity);
        // per-triangle shading records in wide order (layout: device_scene.h; three scattered vertex reads per triangle,
        // shared out over the host threads).  a / e1 / e2 are copied from the traversal stream, so shading sees the bits
        // the ray/triangle test saw.
        shade.resize((size_t)ntriangles * shade_stride * 4);
        host_parallel_for(ntriangles, [&](uint32_t begin, uint32_t end) {
            for (uint32_t wi = begin; wi < end; ++wi) {
                const uint32_t* tri = triangles + 4 * (size_t)wide.orig_index[wi];
                const RptPerVertexData &a = vertices[tri[0]], &b = vertices[tri[1]], &cc = vertices[tri[2]];
                const float* pos = wide.tri_pos.data() + 12 * (size_t)wi;
                float* o = shade.data() + (size_t)shade_stride * 4 * wi;
                o[0] = pos[0]; o[1] = pos[1]; o[2] = pos[2]; o[3] = pos[7];  // a, bits(material)
                o[4] = pos[4]; o[5] = pos[5]; o[6] = pos[6]; o[7] = a.uv0[0];
                o[8] = pos[8]; o[9] = pos[9]; o[10] = pos[10]; o[11] = a.uv0[1];
                o[12] = a.normal[0]; o[13] = a.normal[1]; o[14] = a.normal[2]; o[15] = b.uv0[0];
                o[16] = b.normal[0]; o[17] = b.normal[1]; o[18] = b.normal[2]; o[19] = b.uv0[1];
                o[20] = cc.normal[0]; o[21] = cc.normal[1]; o[22] = cc.normal[2]; o[23] = cc.uv0[0];
                o[24] = cc.uv0[1];
                if (any_normal_map) {
                    for (int k = 0; k < 3; ++k) { o[25 + k] = a.tangent[k]; o[28 + k] = b.tangent[k]; o[31 + k] = cc.tangent[k]; }
                    for (int k = 34; k < 40; ++k) o[k] = 0.0f;
                } else {
                    for (int k = 25; k < 32; ++k) o[k] = 0.0f;
                }
            }
        });
        for (int k = 0; k < 3; ++k) { scene_lo[k] = nodes[0].aabb_min[k]; scene_hi[k] = nodes[0].aabb_max[k]; }  // (the binary root holds the scene bounds)
    } else {
        // what build_wide_bvh checks on its way: indices in range, no infinite / astronomically large coordinate
        std::atomic<int> bad{0};
        host_parallel_for(ntriangles, [&](uint32_t begin, uint32_t end) {
            for (uint32_t t = begin; t < end; ++t)
                for (int k = 0; k < 3; ++k) {
                    const uint32_t v = triangles[4 * (size_t)t + k];
                    if (v >= nvertices) { bad.store(1); return; }
                    const float* p = vertices[v].vertex;
                    if (std::fabs(p[0]) > 1e15f || std::fabs(p[1]) > 1e15f || std::fabs(p[2]) > 1e15f) { bad.store(2); return; }
                }
        });
        if (bad.load() == 1) return c->fail(RPT_ERR_INVALID_ARGUMENT, "rpt_upload_world: triangle references a vertex out of range");
        if (bad.load() == 2) return c->fail(RPT_ERR_INVALID_ARGUMENT, "rpt_upload_world: vertex coordinate infinite or beyond 1e15");
    }

    // ---- uploads (from here on the previous world is gone, whatever happens)
    c->has_world = false;
    c->drop_graphs();
    cudaStream_t s = c->stream;
    RPT_CUDA(c, c->d_vertices.upload(vertices, nvertices, s));
    RPT_CUDA(c, c->d_triangles.upload(reinterpret_cast<const uint4*>(triangles), ntriangles, s));
    RPT_CUDA(c, c->d_materials.upload(materials, nmaterials, s));
    RPT_CUDA(c, c->d_lights.upload(lights, nlights, s));
    c->tree.release();
    uint32_t max_depth = 0;
    std::vector<uint32_t> wide_index;
    if (!build_on_device) {
        RPT_CUDA(c, c->d_nodes.upload(nodes, nnodes, s));
        // nodes and triangle positions — everything the trace kernels read of the scene — in ONE allocation, so that a
        // single L2 access-policy window can cover them (pin_scene_in_l2)
        const size_t node_bytes = (wide.nodes.size() * sizeof(WideNode) + 255) & ~(size_t)255;
        RPT_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->tree.nodes), node_bytes + (size_t)ntriangles * 48));
        c->tree.tri_pos = reinterpret_cast<float4*>(reinterpret_cast<char*>(c->tree.nodes) + node_bytes);
        c->tree.tri_pos_in_nodes_block = true;
        RPT_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->tree.tri_shade), (size_t)ntriangles * shade_stride * 16));
        RPT_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->tree.orig_index), (size_t)ntriangles * 4));
        RPT_CUDA(c, cudaMemcpyAsync(c->tree.nodes, wide.nodes.data(), wide.nodes.size() * sizeof(WideNode), cudaMemcpyHostToDevice, s));
        RPT_CUDA(c, cudaMemcpyAsync(c->tree.tri_pos, wide.tri_pos.data(), (size_t)ntriangles * 48, cudaMemcpyHostToDevice, s));
        RPT_CUDA(c, cudaMemcpyAsync(c->tree.tri_shade, shade.data(), (size_t)ntriangles * shade_stride * 16, cudaMemcpyHostToDevice, s));
        RPT_CUDA(c, cudaMemcpyAsync(c->tree.orig_index, wide.orig_index.data(), (size_t)ntriangles * 4, cudaMemcpyHostToDevice, s));
        c->tree.nnodes = (uint32_t)wide.nodes.size();
        c->tree.max_depth = max_depth = wide.max_depth;
        wide_index = std::move(wide.wide_index);
    } else {
        c->d_nodes.release();
        const cudaError_t e = device_build_wide_bvh(c->d_vertices.p, c->d_triangles.p, ntriangles, shade_stride, c->tree, s);
        if (e != cudaSuccess) return c->fail(RPT_ERR_CUDA, "device BVH build: %s", cudaGetErrorString(e));
        max_depth = c->tree.max_depth;
        if (max_depth + 1 > kWideStackCapacity) {
            c->tree.release();
            return c->fail(RPT_ERR_UNSUPPORTED, "device-built BVH depth %u exceeds the traversal stack (%u)", max_depth, kWideStackCapacity);
        }
        wide_index.resize(ntriangles);
        RPT_CUDA(c, cudaMemcpyAsync(wide_index.data(), c->tree.wide_index, (size_t)ntriangles * 4, cudaMemcpyDeviceToHost, s));
        float box[6];
        RPT_CUDA(c, cudaMemcpyAsync(box, c->tree.node_box, sizeof box, cudaMemcpyDeviceToHost, s));  // node 0: the scene bounds
        RPT_CUDA(c, cudaStreamSynchronize(s));
        for (int k = 0; k < 3; ++k) { scene_lo[k] = box[k]; scene_hi[k] = box[3 + k]; }
    }
    c->device_built = build_on_device;
    pin_scene_in_l2(c, c->tree.nodes, build_on_device ? (size_t)c->tree.nnodes * sizeof(WideNode)
                                                      : (size_t)(reinterpret_cast<char*>(c->tree.tri_pos) - reinterpret_cast<char*>(c->tree.nodes)) + (size_t)ntriangles * 48);
    c->shade_stride = shade_stride;
    c->ntriangles = ntriangles;
    c->nvertices = nvertices;

    std::vector<LightBin> bins;
    std::vector<LightRecord> records;
    RPT_TRY(build_light_records(c, vertices, triangles, ntriangles, materials, lights, nlights, wide_index.data(), bins, records));
    RPT_TRY(upload_light_records(c, bins, records));
    c->h_triangles.assign(triangles, triangles + 4 * (size_t)ntriangles);
    c->h_materials.assign(materials, materials + nmaterials);
    c->h_lights.assign(lights, lights + nlights);
    c->h_wide_index = std::move(wide_index);

    const uchar4 white = make_uchar4(255, 255, 255, 255);
    if (atlas_rgba8) {
        RPT_CUDA(c, c->d_atlas.upload(reinterpret_cast<const uchar4*>(atlas_rgba8), (size_t)atlas_w * atlas_h, s));
        c->atlas_w = atlas_w; c->atlas_h = atlas_h;
    } else {
        RPT_CUDA(c, c->d_atlas.upload(&white, 1, s));
        c->atlas_w = c->atlas_h = 1;
    }
    const float magenta[16] = {1, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1};  // fallback_gpu_image, src/asset.rs:275-281
    c->sources_finite = true;
    for (int k = 0; k < 3; ++k)
        if (!(std::fabs(scene_lo[k]) < 1e15f) || !(std::fabs(scene_hi[k]) < 1e15f)) c->sources_finite = false;
    if (sky_rgba32f) {
        std::atomic<bool> finite{true};
        host_parallel_for((uint32_t)std::min<uint64_t>((uint64_t)sky_w * sky_h, 0xFFFFFFFFull), [&](uint32_t begin, uint32_t end) {
            bool ok = true;
            for (size_t i = (size_t)begin * 4; i < (size_t)end * 4; ++i) ok &= std::isfinite(sky_rgba32f[i]);
            if (!ok) finite.store(false, std::memory_order_relaxed);
        });
        if (!finite.load()) c->sources_finite = false;
        RPT_CUDA(c, c->d_sky.upload(reinterpret_cast<const float4*>(sky_rgba32f), (size_t)sky_w * sky_h, s));
        c->sky_w = sky_w; c->sky_h = sky_h;
    } else {
        RPT_CUDA(c, c->d_sky.upload(reinterpret_cast<const float4*>(magenta), 4, s));
        c->sky_w = c->sky_h = 2;
    }
    RPT_CUDA(c, cudaStreamSynchronize(s));  // host staging vectors die at return
    c->nlights = nlights;
    c->nmaterials = nmaterials;
    c->deep_tree = max_depth + 1 > kWideStackShared;
    if (c->deep_tree && c->wave_capacity != 0 && !c->w_stack_overflow.p)
        RPT_CUDA(c, c->w_stack_overflow.alloc(trace_stack_overflow_entries(c->sm_count * c->trace_blocks_per_sm)));
    c->has_world = true;
    return RPT_OK;
}

// Vertices moved, topology unchanged (same triangles, same materials): new vertex records are uploaded, the triangle
// streams are rewritten and every box of the tree is refitted bottom-up on the device (device_build.cu).  `lights` may
// carry a new light-pick table (areas and pick pdfs change when emitters deform); NULL keeps the table and only
// refreshes the emitters' geometry.
extern "C" int rpt_refit_world(rpt_context* c, const RptPerVertexData* vertices, uint32_t nvertices, const RptLightPickEntry* lights, uint32_t nlights) {
    if (!c || !vertices) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_world) return c->fail(RPT_ERR_NOT_READY, "rpt_refit_world before rpt_upload_world");
    if (nvertices != c->nvertices) return c->fail(RPT_ERR_SIZE_MISMATCH, "world has %u vertices, refit got %u", c->nvertices, nvertices);
    if (lights && nlights == 0) return c->fail(RPT_ERR_INVALID_ARGUMENT, "empty light table");
    RPT_TRY(bind_device(c));
    std::atomic<int> bad{0};
    host_parallel_for(nvertices, [&](uint32_t begin, uint32_t end) {
        for (uint32_t v = begin; v < end; ++v) {
            const float* p = vertices[v].vertex;
            if (std::fabs(p[0]) > 1e15f || std::fabs(p[1]) > 1e15f || std::fabs(p[2]) > 1e15f) { bad.store(1); return; }
        }
    });
    if (bad.load()) return c->fail(RPT_ERR_INVALID_ARGUMENT, "rpt_refit_world: vertex coordinate infinite or beyond 1e15");
    std::vector<LightBin> bins;
    std::vector<LightRecord> records;
    const RptLightPickEntry* table = lights ? lights : c->h_lights.data();
    const uint32_t ntable = lights ? nlights : (uint32_t)c->h_lights.size();
    RPT_TRY(build_light_records(c, vertices, c->h_triangles.data(), c->ntriangles, c->h_materials.data(), table, ntable, c->h_wide_index.data(), bins, records));
    const bool had_lights = c->nbins > 0;
    if (had_lights != !bins.empty() || bins.size() != c->d_light_bins.n || records.size() != c->d_light_records.n) c->drop_graphs();  // light buffers move
    RPT_CUDA(c, cudaMemcpyAsync(c->d_vertices.p, vertices, (size_t)nvertices * sizeof(RptPerVertexData), cudaMemcpyHostToDevice, c->stream));
    const cudaError_t e = device_refit_wide_bvh(c->d_vertices.p, c->d_triangles.p, c->ntriangles, c->shade_stride, c->tree, c->stream);
    if (e != cudaSuccess) return c->fail(RPT_ERR_CUDA, "device BVH refit: %s", cudaGetErrorString(e));
    if (bins.size() == c->d_light_bins.n && records.size() == c->d_light_records.n && !bins.empty()) {  // same shape: in place, captured graphs stay valid
        RPT_CUDA(c, cudaMemcpyAsync(c->d_light_bins.p, bins.data(), bins.size() * sizeof(LightBin), cudaMemcpyHostToDevice, c->stream));
        RPT_CUDA(c, cudaMemcpyAsync(c->d_light_records.p, records.data(), records.size() * sizeof(LightRecord), cudaMemcpyHostToDevice, c->stream));
    } else {
        RPT_TRY(upload_light_records(c, bins, records));
    }
    if (lights) {
        RPT_CUDA(c, c->d_lights.upload(lights, nlights, c->stream));
        c->h_lights.assign(lights, lights + nlights);
        c->nlights = nlights;
    }
    // the reference-layout node copy (megakernel arm) is NOT refitted: that arm needs a fresh rpt_upload_world
    c->d_nodes.release();
    c->device_built = true;
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_refit_world");
}

// ============================================================================ render state
extern "C" int rpt_set_config(rpt_context* c, const RptTracingConfig* cfg) {
    if (!c || !cfg) return RPT_ERR_INVALID_ARGUMENT;
    if (cfg->width == 0 || cfg->height == 0 || (uint64_t)cfg->width * cfg->height > 0x7FFFFFFFull)
        return c->fail(RPT_ERR_INVALID_ARGUMENT, "bad frame size %ux%u", cfg->width, cfg->height);
    const uint32_t dims = rng_dimension_budget(*cfg);
    if (dims > 31)
        return c->fail(RPT_ERR_RNG_DIMENSIONS, "config can consume %u R-sequence dimensions per path; the reference's table has 31 (kernels/src/rng.rs:19-27)", dims);
    RPT_TRY(bind_device(c));
    const bool resized = !c->has_config || cfg->width != c->config.width || cfg->height != c->config.height;
    if (!c->has_config || std::memcmp(&c->config, cfg, sizeof *cfg) != 0) c->drop_graphs();  // the config is baked into the launches
    c->config = *cfg;
    c->has_config = true;
    float m[9];
    rpt_camera_matrix(cfg->cam_rotation[0], cfg->cam_rotation[1], m);
    c->camera.position = mk3(cfg->cam_position[0], cfg->cam_position[1], cfg->cam_position[2]);
    c->camera.c0 = mk3(m[0], m[1], m[2]);
    c->camera.c1 = mk3(m[3], m[4], m[5]);
    c->camera.c2 = mk3(m[6], m[7], m[8]);
    c->camera.width = (float)cfg->width;
    c->camera.height = (float)cfg->height;
    c->camera.aspect = (float)cfg->height / (float)cfg->width;
    const float yaw = std::atan2(cfg->sun_direction[2], cfg->sun_direction[0]);  // lib.rs:71
    c->sky_yaw_sin = std::sin(yaw);
    c->sky_yaw_cos = std::cos(yaw);
    if (resized) {
        const size_t n = (size_t)cfg->width * cfg->height;
        if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
        c->combined_valid = false;
        c->d_snapshot.release(); c->d_combined.release(); c->d_gather.release();
        c->drop_gather_maps();
        RPT_CUDA(c, c->d_rng.alloc(n));
        RPT_CUDA(c, c->d_output.alloc(n));
        RPT_CUDA(c, cudaMemsetAsync(c->d_output.p, 0, n * sizeof(float4), c->stream));
        if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
        c->d_rgb_async[0].release(); c->d_rgb_async[1].release();
        c->d_rgb.release();
        c->d_rgba8.release();
        c->d_ids.release();
        c->rng_written = false;
        RPT_TRY(rebuild_pixel_map(c));
    }
    return RPT_OK;
}

extern "C" int rpt_write_rng(rpt_context* c, const uint32_t* seeds, size_t npixels) {
    if (!c || !seeds) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "rpt_write_rng before rpt_set_config");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "rng has %zu entries, frame has %u pixels", npixels, c->npixels());
    RPT_TRY(bind_device(c));
    RPT_CUDA(c, cudaMemcpyAsync(c->d_rng.p, seeds, npixels * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->rng_written = true;
    return RPT_OK;
}

extern "C" int rpt_read_rng(rpt_context* c, uint32_t* seeds, size_t npixels) {
    if (!c || !seeds) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config || !c->rng_written) return c->fail(RPT_ERR_NOT_READY, "rng not written yet");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "rng has %u entries, caller asked for %zu", c->npixels(), npixels);
    RPT_TRY(bind_device(c));
    RPT_CUDA(c, cudaMemcpyAsync(seeds, c->d_rng.p, npixels * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream));
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_read_rng");
}

extern "C" int rpt_write_output(rpt_context* c, const float* rgba, size_t npixels) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "rpt_write_output before rpt_set_config");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "output has %zu entries, frame has %u pixels", npixels, c->npixels());
    RPT_TRY(bind_device(c));
    c->combined_valid = false;  // a reset / resume starts a new frame: reads return this context's accumulator again
    if (rgba) RPT_CUDA(c, cudaMemcpyAsync(c->d_output.p, rgba, npixels * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    else RPT_CUDA(c, cudaMemsetAsync(c->d_output.p, 0, npixels * sizeof(float4), c->stream));
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_write_output");
}

extern "C" int rpt_set_tile_partition(rpt_context* c, uint32_t tile_rank, uint32_t tile_count) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (tile_count == 0 || tile_rank >= tile_count) return c->fail(RPT_ERR_INVALID_ARGUMENT, "tile rank %u of %u", tile_rank, tile_count);
    RPT_TRY(bind_device(c));
    c->tile_rank = tile_rank;
    c->tile_count = tile_count;
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    c->drop_gather_maps();
    c->drop_graphs();
    return rebuild_pixel_map(c);
}

// ============================================================================ host staging memory
extern "C" int rpt_host_alloc(size_t bytes, void** out_ptr) {
    if (!out_ptr) return RPT_ERR_INVALID_ARGUMENT;
    *out_ptr = nullptr;
    if (bytes == 0) return RPT_ERR_INVALID_ARGUMENT;
    const cudaError_t e = cudaHostAlloc(out_ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaHostAlloc: ") + cudaGetErrorString(e);
        cudaGetLastError();
        *out_ptr = nullptr;
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? RPT_ERR_NO_DEVICE : RPT_ERR_CUDA;
    }
    return RPT_OK;
}

extern "C" int rpt_host_free(void* ptr) {
    if (!ptr) return RPT_ERR_INVALID_ARGUMENT;
    const cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaFreeHost: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return RPT_ERR_CUDA;
    }
    return RPT_OK;
}

// ============================================================================ run
extern "C" int rpt_enqueue(rpt_context* c, uint32_t n_samples) {
    RPT_TRY(check_ready(c));
    if (n_samples == 0) return RPT_OK;
    // Device time of every batch (rpt_get_device_ms).  Pairs whose batch has completed are folded into the running
    // total here, so a host that never asks — an interactive session of millions of batches — holds a handful of events.
    while (c->timed.size() > 8 && cudaEventQuery(c->timed.front().second) == cudaSuccess) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, c->timed.front().first, c->timed.front().second) == cudaSuccess) c->device_ms += t;
        cudaEventDestroy(c->timed.front().first);
        cudaEventDestroy(c->timed.front().second);
        c->timed.erase(c->timed.begin());
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    RPT_CUDA(c, cudaEventCreate(&e0));
    if (const cudaError_t e = cudaEventCreate(&e1); e != cudaSuccess) {
        cudaEventDestroy(e0);
        return c->cuda(e, "cudaEventCreate");
    }
    if (const cudaError_t e = cudaEventRecord(e0, c->stream); e != cudaSuccess) {
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return c->cuda(e, "cudaEventRecord");
    }
    int status = RPT_OK;
    if (c->pipeline == RPT_PIPELINE_MEGAKERNEL) {
        if (!c->d_nodes.p) return c->fail(RPT_ERR_NOT_READY, "the megakernel arm traverses the reference BVH, and this world has none (device-built or refitted tree)");
        c->launch(RPT_STAGE_MEGAKERNEL, [&] { launch_mega_trace(mega_params(c), n_samples, c->stream); });
        status = c->cuda(cudaGetLastError(), "megakernel launch");
        // the megakernel counts its rays on the device; finished paths are counted here
        if (status == RPT_OK) c->mega_paths += (uint64_t)(c->tile_count > 1 ? c->pixel_map_len : c->npixels()) * n_samples;
    } else {
        status = for_each_wave(c, n_samples, [&](const WaveDesc& d) { return run_wave_graphed(c, d); });
    }
    const cudaError_t rec = cudaEventRecord(e1, c->stream);
    if (rec != cudaSuccess) {
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return status != RPT_OK ? status : c->cuda(rec, "cudaEventRecord");
    }
    c->timed.emplace_back(e0, e1);
    return status;
}

// The reference's dispatch loop looks at its control flags after EVERY sample (src/trace.rs:182-193) and leaves the
// batch early when the camera moves or the render is stopped.  Here the batch is cut into groups of `poll_samples`
// samples; after each group the stream is drained and `*stop_flag` (host memory another thread may set) is read.
extern "C" int rpt_enqueue_interruptible(rpt_context* c, uint32_t n_samples, const volatile uint32_t* stop_flag, uint32_t poll_samples,
                                         uint32_t* finished_out) {
    if (!c || !finished_out) return RPT_ERR_INVALID_ARGUMENT;
    *finished_out = 0;
    if (poll_samples == 0) poll_samples = 1;
    for (uint32_t done = 0; done < n_samples;) {
        const uint32_t group = std::min(poll_samples, n_samples - done);
        RPT_TRY(rpt_enqueue(c, group));
        RPT_CUDA(c, cudaStreamSynchronize(c->stream));
        done += group;
        *finished_out = done;
        if (stop_flag && *stop_flag != 0u) break;
    }
    return RPT_OK;
}

extern "C" int rpt_sync(rpt_context* c) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    RPT_TRY(bind_device(c));
    if (c->comm_stream) RPT_CUDA(c, cudaStreamSynchronize(c->comm_stream));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->copy_stream) RPT_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    return RPT_OK;
}

// ============================================================================ readback
extern "C" int rpt_read_output(rpt_context* c, float* rgba, size_t npixels) {
    if (!c || !rgba) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "rpt_read_output before rpt_set_config");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "frame has %u pixels, caller asked for %zu", c->npixels(), npixels);
    RPT_TRY(bind_device(c));
    RPT_CUDA(c, cudaMemcpyAsync(rgba, c->frame_for_read(), npixels * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_read_output");
}

extern "C" int rpt_read_framebuffer(rpt_context* c, float* rgb, size_t npixels, float samples) {
    if (!c || !rgb) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "rpt_read_framebuffer before rpt_set_config");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "frame has %u pixels, caller asked for %zu", c->npixels(), npixels);
    RPT_TRY(bind_device(c));
    if (c->d_rgb.n != npixels * 3) RPT_CUDA(c, c->d_rgb.alloc(npixels * 3));
    launch_normalize(c->frame_for_read(), c->d_rgb.p, (uint32_t)npixels, samples, c->stream);
    c->kernel_launches++;
    RPT_CUDA(c, cudaGetLastError());
    if (c->frame_hook) c->frame_hook(c->d_rgb.p, c->config.width, c->config.height, c->stream, c->frame_hook_user);
    RPT_CUDA(c, cudaMemcpyAsync(rgb, c->d_rgb.p, npixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_read_framebuffer");
}

extern "C" int rpt_set_frame_hook(rpt_context* c, rpt_frame_hook hook, void* user) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    c->frame_hook = hook;
    c->frame_hook_user = user;
    return RPT_OK;
}

// rpt_read_framebuffer without the wait: the frame is normalised on the render stream into one of two device
// buffers and copied to `rgb` (page-locked: rpt_host_alloc) on a copy stream; the caller goes on enqueueing and calls
// rpt_readback_wait before it reads `rgb`.
extern "C" int rpt_read_framebuffer_async(rpt_context* c, float* rgb, size_t npixels, float samples) {
    if (!c || !rgb) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "rpt_read_framebuffer_async before rpt_set_config");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "frame has %u pixels, caller asked for %zu", c->npixels(), npixels);
    RPT_TRY(bind_device(c));
    if (!c->copy_stream) {
        RPT_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            RPT_CUDA(c, cudaEventCreateWithFlags(&c->ev_frame_ready[k], cudaEventDisableTiming));
            RPT_CUDA(c, cudaEventCreateWithFlags(&c->ev_frame_copied[k], cudaEventDisableTiming));
        }
    }
    const int k = c->async_next;
    c->async_next ^= 1;
    if (c->d_rgb_async[k].n != npixels * 3) RPT_CUDA(c, c->d_rgb_async[k].alloc(npixels * 3));
    RPT_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_frame_copied[k], 0));  // the copy that last used this buffer
    launch_normalize(c->frame_for_read(), c->d_rgb_async[k].p, (uint32_t)npixels, samples, c->stream);
    c->kernel_launches++;
    RPT_CUDA(c, cudaGetLastError());
    if (c->frame_hook) c->frame_hook(c->d_rgb_async[k].p, c->config.width, c->config.height, c->stream, c->frame_hook_user);
    RPT_CUDA(c, cudaEventRecord(c->ev_frame_ready[k], c->stream));
    RPT_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_frame_ready[k], 0));
    RPT_CUDA(c, cudaMemcpyAsync(rgb, c->d_rgb_async[k].p, npixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
    return c->cuda(cudaEventRecord(c->ev_frame_copied[k], c->copy_stream), "rpt_read_framebuffer_async");
}

extern "C" int rpt_readback_wait(rpt_context* c) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->copy_stream) return RPT_OK;
    RPT_TRY(bind_device(c));
    return c->cuda(cudaStreamSynchronize(c->copy_stream), "rpt_readback_wait");
}

extern "C" int rpt_read_display(rpt_context* c, float* rgb, size_t npixels, float samples, uint32_t tonemap) {
    if (!c || !rgb) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "rpt_read_display before rpt_set_config");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "frame has %u pixels, caller asked for %zu", c->npixels(), npixels);
    RPT_TRY(bind_device(c));
    if (c->d_rgb.n != npixels * 3) RPT_CUDA(c, c->d_rgb.alloc(npixels * 3));
    launch_display(c->frame_for_read(), c->d_rgb.p, (uint32_t)npixels, samples, tonemap, c->stream);
    c->kernel_launches++;
    RPT_CUDA(c, cudaGetLastError());
    RPT_CUDA(c, cudaMemcpyAsync(rgb, c->d_rgb.p, npixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_read_display");
}

extern "C" int rpt_read_display_rgba8(rpt_context* c, uint8_t* rgba, size_t npixels, float samples, uint32_t tonemap, uint32_t srgb_encode) {
    if (!c || !rgba) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "rpt_read_display_rgba8 before rpt_set_config");
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "frame has %u pixels, caller asked for %zu", c->npixels(), npixels);
    RPT_TRY(bind_device(c));
    if (c->d_rgba8.n != npixels) RPT_CUDA(c, c->d_rgba8.alloc(npixels));
    launch_display_rgba8(c->frame_for_read(), c->d_rgba8.p, (uint32_t)npixels, samples, tonemap, srgb_encode != 0, c->stream);
    c->kernel_launches++;
    RPT_CUDA(c, cudaGetLastError());
    RPT_CUDA(c, cudaMemcpyAsync(rgba, c->d_rgba8.p, npixels * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_read_display_rgba8");
}

extern "C" int rpt_read_primary_ids(rpt_context* c, uint32_t* ids, size_t npixels) {
    RPT_TRY(check_ready(c));
    if (!ids) return RPT_ERR_INVALID_ARGUMENT;
    if (npixels != c->npixels()) return c->fail(RPT_ERR_SIZE_MISMATCH, "frame has %u pixels, caller asked for %zu", c->npixels(), npixels);
    if (c->d_ids.n != npixels) RPT_CUDA(c, c->d_ids.alloc(npixels));
    RPT_CUDA(c, cudaMemsetAsync(c->d_ids.p, 0xFF, npixels * sizeof(uint32_t), c->stream));
    // counters must not move for a diagnostic pass: save and restore them around it
    unsigned long long saved[8];
    RPT_CUDA(c, cudaMemcpyAsync(saved, c->d_counters.p, sizeof saved, cudaMemcpyDeviceToHost, c->stream));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    int status = RPT_OK;
    if (c->pipeline == RPT_PIPELINE_MEGAKERNEL) {
        if (!c->d_nodes.p) return c->fail(RPT_ERR_NOT_READY, "the megakernel arm traverses the reference BVH, and this world has none (device-built or refitted tree)");
        launch_mega_primary(mega_params(c), c->d_ids.p, c->stream);
        c->kernel_launches++;
        status = c->cuda(cudaGetLastError(), "primary-id launch");
    } else {
        status = for_each_wave(c, 1, [&](const WaveDesc& d) {
            WaveDesc one = d;
            one.k_samples = 1;
            return run_wave(c, one, true, c->d_ids.p);
        });
    }
    RPT_TRY(status);
    RPT_CUDA(c, cudaMemcpyAsync(c->d_counters.p, saved, sizeof saved, cudaMemcpyHostToDevice, c->stream));
    RPT_CUDA(c, cudaMemcpyAsync(ids, c->d_ids.p, npixels * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    return c->cuda(cudaStreamSynchronize(c->stream), "rpt_read_primary_ids");
}

extern "C" int rpt_get_counters(rpt_context* c, RptCounters* out) {
    if (!c || !out) return RPT_ERR_INVALID_ARGUMENT;
    RPT_TRY(bind_device(c));
    unsigned long long host[4] = {0, 0, 0, 0};
    RPT_CUDA(c, cudaMemcpyAsync(host, c->d_counters.p, sizeof host, cudaMemcpyDeviceToHost, c->stream));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    out->paths = host[0] + c->mega_paths;
    out->nearest_rays = host[1];
    out->any_rays = host[2];
    out->kernel_launches = c->kernel_launches;
    return RPT_OK;
}

extern "C" int rpt_reset_counters(rpt_context* c) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    RPT_TRY(bind_device(c));
    RPT_CUDA(c, cudaMemsetAsync(c->d_counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& ev : c->timed) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    c->timed.clear();
    c->device_ms = 0.0f;
    c->kernel_launches = 0;
    c->mega_paths = 0;
    return RPT_OK;
}

extern "C" int rpt_get_device_ms(rpt_context* c, float* ms) {
    if (!c || !ms) return RPT_ERR_INVALID_ARGUMENT;
    RPT_TRY(bind_device(c));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& ev : c->timed) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, ev.first, ev.second) == cudaSuccess) c->device_ms += t;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    c->timed.clear();
    *ms = c->device_ms;
    return RPT_OK;
}

extern "C" int rpt_get_sm_count(rpt_context* c, int* sm_count) {
    if (!c || !sm_count) return RPT_ERR_INVALID_ARGUMENT;
    *sm_count = c->sm_count;
    return RPT_OK;
}

extern "C" int rpt_set_trace_statistics(rpt_context* c, int enable) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (c->trace_statistics != (enable != 0)) c->drop_graphs();  // the captured waves launch the other build
    c->trace_statistics = enable != 0;
    return RPT_OK;
}

extern "C" int rpt_get_trace_statistics(rpt_context* c, RptTraceStatistics* out) {
    if (!c || !out) return RPT_ERR_INVALID_ARGUMENT;
    RPT_TRY(bind_device(c));
    unsigned long long host[8] = {0};
    RPT_CUDA(c, cudaMemcpyAsync(host, c->d_counters.p, sizeof host, cudaMemcpyDeviceToHost, c->stream));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    out->nearest_rays = host[1];
    out->any_rays = host[2];
    out->nearest_node_visits = host[4];
    out->nearest_triangle_tests = host[5];
    out->any_node_visits = host[6];
    out->any_triangle_tests = host[7];
    out->shaded_hits = host[3];
    out->node_bytes = sizeof(WideNode);
    out->triangle_bytes = 48;
    return RPT_OK;
}

// Device time of a region of calls — enqueues AND the combine on the side stream — from CUDA events on the streams the
// work is launched on (bench.py times multi-GPU steps with this, never by wall clock).
extern "C" int rpt_timer_start(rpt_context* c) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    RPT_TRY(bind_device(c));
    if (!c->region_start) {
        RPT_CUDA(c, cudaEventCreate(&c->region_start));
        RPT_CUDA(c, cudaEventCreate(&c->region_stop));
    }
    if (c->comm_stream) RPT_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_combined, 0));  // an earlier combine is not part of the region
    return c->cuda(cudaEventRecord(c->region_start, c->stream), "rpt_timer_start");
}

extern "C" int rpt_timer_stop(rpt_context* c, float* ms) {
    if (!c || !ms) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->region_start) return c->fail(RPT_ERR_NOT_READY, "rpt_timer_stop before rpt_timer_start");
    RPT_TRY(bind_device(c));
    if (c->comm_stream) RPT_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_combined, 0));  // the region ends when the combine has landed
    RPT_CUDA(c, cudaEventRecord(c->region_stop, c->stream));
    RPT_CUDA(c, cudaEventSynchronize(c->region_stop));
    return c->cuda(cudaEventElapsedTime(ms, c->region_start, c->region_stop), "rpt_timer_stop");
}

extern "C" int rpt_set_stage_timing(rpt_context* c, int enable) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    c->stage_timing = enable != 0;
    return RPT_OK;
}

extern "C" int rpt_get_stage_timing(rpt_context* c, RptStageTiming* out) {
    if (!c || !out) return RPT_ERR_INVALID_ARGUMENT;
    RPT_TRY(bind_device(c));
    RPT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->drain_stage_events();
    *out = c->stage_totals;
    c->stage_totals = RptStageTiming{};
    return RPT_OK;
}

// ============================================================================ multi-GPU combine
extern "C" int rpt_comm_unique_id(uint8_t* id_bytes_128) {
    if (!id_bytes_128) return RPT_ERR_INVALID_ARGUMENT;
    if (!g_nccl.load(g_create_error)) return RPT_ERR_NCCL;
    NcclUniqueId id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r != 0) {
        g_create_error = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        return RPT_ERR_NCCL;
    }
    std::memcpy(id_bytes_128, id.internal, 128);
    return RPT_OK;
}

extern "C" int rpt_comm_init(rpt_context* c, const uint8_t* id_bytes_128, int rank, int nranks) {
    if (!c || !id_bytes_128 || nranks < 1 || rank < 0 || rank >= nranks) return RPT_ERR_INVALID_ARGUMENT;
    if (!g_nccl.load(c->error)) return RPT_ERR_NCCL;
    RPT_TRY(bind_device(c));
    if (c->nccl_comm) { g_nccl.CommDestroy(c->nccl_comm); c->nccl_comm = nullptr; }
    NcclUniqueId id;
    std::memcpy(id.internal, id_bytes_128, 128);
    const int r = g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (r != 0) return c->fail(RPT_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    c->nccl_rank = rank;
    c->nccl_nranks = nranks;
    return RPT_OK;
}

namespace {
int nccl_fail(rpt_context* c, int r, const char* what) {
    return c->fail(RPT_ERR_NCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
}

int ensure_comm_stream(rpt_context* c) {
    if (c->comm_stream) return RPT_OK;
    RPT_CUDA(c, cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    RPT_CUDA(c, cudaEventCreateWithFlags(&c->ev_snapshot, cudaEventDisableTiming));
    RPT_CUDA(c, cudaEventCreateWithFlags(&c->ev_combined, cudaEventDisableTiming));
    return RPT_OK;
}
}  // namespace

// Combine the per-rank accumulators on `root`.  No accumulator is modified: each rank snapshots its contribution on
// the render stream, the exchange runs on a side stream (so the next rpt_enqueue overlaps it), and the result is
// the root's "combined frame", which its read calls return from then on (rpt_write_output or a resize drops it).
//   * whole-frame ranks (sample-index split): snapshot = the accumulator; ncclReduce(sum) over NVLink.
//   * tile ranks (rpt_set_tile_partition(rank, nranks)): snapshot = the pixels this rank owns, packed; every rank
//     sends its pack to the root (grouped ncclSend / ncclRecv), which scatters them: (nranks - 1) / nranks of ONE
//     frame crosses NVLink instead of a reduce over nranks full frames of mostly zeros.  Bit-identical to one GPU.
extern "C" int rpt_comm_reduce_output(rpt_context* c, int root) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (!c->nccl_comm) return c->fail(RPT_ERR_NOT_READY, "rpt_comm_reduce_output before rpt_comm_init");
    if (!c->has_config) return c->fail(RPT_ERR_NOT_READY, "no frame to reduce");
    if (root < 0 || root >= c->nccl_nranks) return c->fail(RPT_ERR_INVALID_ARGUMENT, "root %d of %d ranks", root, c->nccl_nranks);
    RPT_TRY(bind_device(c));
    RPT_TRY(ensure_comm_stream(c));
    const uint32_t npix = c->npixels();
    const bool is_root = c->nccl_rank == root;
    const bool tiles = c->tile_count > 1;
    if (tiles && ((int)c->tile_count != c->nccl_nranks || (int)c->tile_rank != c->nccl_rank))
        return c->fail(RPT_ERR_INVALID_ARGUMENT, "tile partition %u of %u does not match rank %d of %d", c->tile_rank, c->tile_count, c->nccl_rank, c->nccl_nranks);
    // the previous combine must be off the wire before its buffers are reused
    RPT_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_combined, 0));
    if (is_root && c->d_combined.n != npix) RPT_CUDA(c, c->d_combined.alloc(npix));

    if (!tiles) {
        if (c->d_snapshot.n != npix) RPT_CUDA(c, c->d_snapshot.alloc(npix));
        RPT_CUDA(c, cudaMemcpyAsync(c->d_snapshot.p, c->d_output.p, (size_t)npix * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        RPT_CUDA(c, cudaEventRecord(c->ev_snapshot, c->stream));
        RPT_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_snapshot, 0));
        const int r = g_nccl.Reduce(c->d_snapshot.p, is_root ? c->d_combined.p : nullptr, (size_t)npix * 4, /*ncclFloat32*/ 7, /*ncclSum*/ 0, root, c->nccl_comm, c->comm_stream);
        if (r != 0) return nccl_fail(c, r, "ncclReduce");
    } else {
        // pixel maps: this rank's own (d_pixel_map) and, on the root, everybody's
        const uint32_t mine = c->pixel_map_len;
        if (c->d_snapshot.n < std::max(mine, 1u)) RPT_CUDA(c, c->d_snapshot.alloc(std::max(mine, 1u)));
        if (is_root && (c->gather_maps.size() != (size_t)c->nccl_nranks || c->gather_w != c->config.width || c->gather_h != c->config.height)) {
            c->drop_gather_maps();
            c->gather_maps.resize((size_t)c->nccl_nranks);
            size_t total = 0;
            for (int r = 0; r < c->nccl_nranks; ++r) {
                uint32_t count = 0;
                rpt_tile_partition_pixels(c->config.width, c->config.height, (uint32_t)r, c->tile_count, nullptr, &count);
                std::vector<uint32_t> map(count);
                if (count) rpt_tile_partition_pixels(c->config.width, c->config.height, (uint32_t)r, c->tile_count, map.data(), &count);
                c->gather_maps[(size_t)r].count = count;
                if (count) RPT_CUDA(c, c->gather_maps[(size_t)r].map.upload(map.data(), count, c->stream));
                RPT_CUDA(c, cudaStreamSynchronize(c->stream));  // `map` dies here
                total += count;
            }
            if (total != npix) return c->fail(RPT_ERR_CUDA, "tile partitions cover %zu of %u pixels", total, npix);
            RPT_CUDA(c, c->d_gather.alloc(npix));
            c->gather_w = c->config.width; c->gather_h = c->config.height;
        }
        launch_pack_pixels(c->d_output.p, c->d_pixel_map.p, c->d_snapshot.p, mine, c->sm_count, c->stream);
        c->kernel_launches++;
        RPT_CUDA(c, cudaGetLastError());
        RPT_CUDA(c, cudaEventRecord(c->ev_snapshot, c->stream));
        RPT_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->ev_snapshot, 0));
        int r = g_nccl.GroupStart();
        if (r != 0) return nccl_fail(c, r, "ncclGroupStart");
        if (!is_root) {
            if (mine) r = g_nccl.Send(c->d_snapshot.p, (size_t)mine * 4, 7, root, c->nccl_comm, c->comm_stream);
        } else {
            size_t offset = 0;
            for (int k = 0; k < c->nccl_nranks && r == 0; ++k) {
                const uint32_t count = c->gather_maps[(size_t)k].count;
                if (k != root && count) r = g_nccl.Recv(c->d_gather.p + offset, (size_t)count * 4, 7, k, c->nccl_comm, c->comm_stream);
                offset += count;
            }
        }
        const int e = g_nccl.GroupEnd();
        if (r != 0 || e != 0) return nccl_fail(c, r != 0 ? r : e, "tile gather (ncclSend / ncclRecv)");
        if (is_root) {
            size_t offset = 0;
            for (int k = 0; k < c->nccl_nranks; ++k) {
                const rpt_context::RankTiles& t = c->gather_maps[(size_t)k];
                launch_unpack_pixels(k == root ? c->d_snapshot.p : c->d_gather.p + offset, t.map.p, c->d_combined.p, t.count, c->sm_count, c->comm_stream);
                c->kernel_launches++;
                offset += t.count;
            }
            RPT_CUDA(c, cudaGetLastError());
        }
    }
    RPT_CUDA(c, cudaEventRecord(c->ev_combined, c->comm_stream));
    if (is_root) c->combined_valid = true;
    return RPT_OK;
}

extern "C" int rpt_comm_destroy(rpt_context* c) {
    if (!c) return RPT_ERR_INVALID_ARGUMENT;
    if (c->nccl_comm) {
        bind_device(c);
        if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    return RPT_OK;
}
