// device_scene.h — parameter blocks passed by value to the kernels, and the launch wrappers
// the context (context.cu) calls.  All pointers are device pointers owned by the context.
#pragma once

#include <cuda_runtime.h>

#include "dev/shading.cuh"
#include "dev/wide_bvh.cuh"

namespace rpt {

// Per-config constants (rpt_set_config) + read-only images.
struct FrameParams {
    Camera camera;
    uint32_t width, height, min_bounces, max_bounces, nee, has_skybox;
    float clamp_lo, clamp_hi;
    f3 sun_dir;
    float sun_intensity;
    SkyImage sky;
    Atlas atlas;
    uint32_t tile_rank, tile_count;  // 32x32-tile round-robin partition (tile_count <= 1: whole frame)
    uint32_t retire_dead_paths;      // a path whose throughput became exactly (0,0,0) is not continued (context.cu: when that is exact)
};

// ---- megakernel arm: reads the reference's own layouts ---------------------------------------
struct MegaParams {
    Camera camera;
    uint32_t width, height, min_bounces, max_bounces, nee, has_skybox;
    float clamp_lo, clamp_hi;
    f3 sun_dir;
    float sun_intensity;
    SkyImage sky;
    const uchar4* atlas;
    uint32_t atlas_w, atlas_h;
    const float4* nodes;
    const uint4* triangles;
    const float4* vertices;
    const RptMaterialData* materials;
    const RptLightPickEntry* lights;
    uint32_t nlights;
    uint2* rng;
    float4* output;
    unsigned long long* counters;  // [0] paths, [1] nearest rays, [2] any rays
    uint32_t tile_rank, tile_count;
};

void launch_mega_trace(const MegaParams& p, uint32_t n_samples, cudaStream_t stream);
void launch_mega_primary(const MegaParams& p, uint32_t* ids, cudaStream_t stream);

// ---- wavefront arm: private layouts ----------------------------------------------------------
// One emissive triangle, in the order the light-pick table refers to it.
struct LightRecord {
    float4 a_area;        // vertex a, triangle area (LightPickEntry.triangle_area_*)
    float4 e1_pdf;        // edge b-a, pick pdf (LightPickEntry.triangle_pick_pdf_*)
    float4 e2_tri;        // edge c-a, bits(wide triangle index)
    float4 normal;        // mean of the three vertex normals (light_pick.rs:128)
    float4 emission;      // material emissive rgb
};
// LightPickEntry with triangle indices replaced by LightRecord indices.
struct LightBin {
    uint32_t light_a, light_b;
    float ratio;
};

struct WideWorld {
    WideScene bvh;                     // nodes + triangle position stream
    // One shading record per triangle, wide order, `shade_stride` float4 apart: everything wf_shade_kernel gathers for
    // a hit in ONE aligned run of 32-byte sectors (it used to be three gathers from three streams):
    //   [0] a.xyz, bits(material)   [1] e1.xyz, uv_a.x   [2] e2.xyz, uv_a.y
    //   [3] n_a.xyz, uv_b.x         [4] n_b.xyz, uv_b.y  [5] n_c.xyz, uv_c.x    [6] uv_c.y, -, -, -       (7 float4, stride 8:
    //   one 128-byte line) and, when a material is normal-mapped, the tangents in words [6].yzw [7] [8].xy (stride 10: five sectors).
    const float4* tri_shade;
    uint32_t shade_stride;             // 8 or 10 (kShadeStrideTangents: records carry tangents)
    const RptMaterialData* materials;
    uint32_t nmaterials;
    const LightBin* light_bins;        // null / nbins == 0: the "no lights" sentinel
    uint32_t nbins;
    const LightRecord* lights;
    uint32_t nlights;                  // light records
};

constexpr uint32_t kShadeStridePlain = 8, kShadeStrideTangents = 10;
constexpr uint32_t kSmemLightBytes = 4 * 1024;  // light records + bins are staged in shared memory when they fit this

struct WaveCtl {
    uint32_t n_ext[2];  // rays queued for the current / next extend pass
    uint32_t n_hit, n_miss, n_shadow;
    uint32_t fetch_extend, fetch_shadow;  // work cursors of the persistent trace kernels
    uint32_t pad;
};

// Path state of one wave, structure-of-arrays over `slots` path slots.
struct WaveState {
    float4* ray_o;   // origin.xyz, Russian-roulette probability the throughput was last divided by (1 if none)
    float4* ray_d;   // direction.xyz, bits(rng dimension | last lobe << 8 | light record of the last NEE sample << 9)
    float4* thr;     // throughput.xyz, pdf of the last BSDF sample
    float4* rad;     // radiance of this pixel-sample so far
    uint2* hit;      // bits(t), wide triangle | backface << 31
    float4* sh_o;    // shadow rays, indexed like q_hit: origin.xyz, max_t
    float4* sh_d;    //               direction.xyz, bits(slot)
    float4* sh_c;    //               contribution.xyz if unoccluded
    uint32_t* q_ext[2];
    uint32_t* q_hit;
    uint32_t* q_miss;
    uint32_t* q_shaded;  // per hit: slot | kShadedNoNext | kShadedShadow (wf_shade_kernel -> wf_compact_shaded_kernel)
    uint32_t* q_shadow;  // indices of the shadow rays to trace
    WaveCtl* ctl;
    unsigned long long* counters;  // [0] paths, [1] nearest rays, [2] any rays, [3] surface hits shaded; diagnostic build: [4] node visits / [5] triangle tests of nearest rays, [6] / [7] of any rays
};

constexpr uint32_t kShadedShadow = 0x80000000u, kShadedNoNext = 0x40000000u, kShadedSlotMask = 0x3FFFFFFFu;  // waves hold < 2^30 slots

// Which pixel-samples a wave covers: slot = k * npix + j  ->  pixel = map(pix_base + j), sample k.
struct WaveDesc {
    uint32_t pix_base, npix, k_samples;
    const uint32_t* pixel_map;  // null: identity
};

struct WaveLaunch {
    int grid;  // SM count: persistent grids are multiples of it
    cudaStream_t stream;
    int trace_blocks_per_sm;  // resident 128-thread blocks per SM of the trace kernels
    int refill_below;         // refill idle lanes once fewer than this many lanes of a warp hold a ray
    uint2* stack_overflow;    // global overflow area of the traversal stacks (trace_stack_overflow_entries); null when the tree fits the shared slab
    bool defer_extend, defer_shadow;  // trace with deferred triangle tests (wf_trace_deferred_kernel)
    int flush_at, flush_keep;         // triangle rounds start at / go on while this many lanes hold a queued triangle
    bool trace_statistics;            // diagnostic build of the trace kernels: count node visits / triangle tests
};

size_t trace_stack_overflow_entries(int grid_blocks);  // uint2 entries the trace kernels need for a grid of that many blocks
void launch_wf_reset(const WaveLaunch& l, const WaveState& s, int next_queue, bool whole);
void launch_wf_generate(const WaveLaunch& l, const FrameParams& f, const WaveState& s, const WaveDesc& d, const uint2* rng);
void launch_wf_extend(const WaveLaunch& l, const WideScene& bvh, const WaveState& s, int in_queue, bool identity_queue, uint32_t n_identity);
void launch_wf_shadow(const WaveLaunch& l, const WideScene& bvh, const WaveState& s);
void launch_wf_export_primary(const WaveLaunch& l, const WideScene& bvh, const WaveState& s, const WaveDesc& d, uint32_t* ids);
void launch_wf_shade(const WaveLaunch& l, const FrameParams& f, const WideWorld& w, const WaveState& s, const WaveDesc& d, const uint2* rng,
                     uint32_t bounce);
void launch_wf_compact_shaded(const WaveLaunch& l, const WaveState& s, int out_queue);
void launch_wf_miss(const WaveLaunch& l, const FrameParams& f, const WaveState& s);
void launch_wf_accumulate(const WaveLaunch& l, const WaveState& s, const WaveDesc& d, uint2* rng, float4* output);
void launch_pack_pixels(const float4* frame, const uint32_t* map, float4* packed, uint32_t n, int grid, cudaStream_t stream);
void launch_unpack_pixels(const float4* packed, const uint32_t* map, float4* frame, uint32_t n, int grid, cudaStream_t stream);
void launch_normalize(const float4* output, float* rgb, uint32_t npixels, float samples, cudaStream_t stream);  // display_nofma.cu
void launch_display(const float4* output, float* rgb, uint32_t npixels, float samples, uint32_t tonemap_op, cudaStream_t stream);
void launch_display_rgba8(const float4* output, uint32_t* rgba8, uint32_t npixels, float samples, uint32_t tonemap_op, bool srgb, cudaStream_t stream);

}  // namespace rpt
