/*
 * rpt_b200.h — C ABI of the B200 (sm_100a) tracing backend.
 *
 * This is the boundary that replaces the reference's wgpu / gpgpu-rs compute dispatch inside
 * `trace_gpu` (src/trace.rs:136-224).  One context drives one GPU (one process per GPU); every
 * entry point takes plain pointers and sizes, copies what it is given (the caller keeps
 * ownership), returns an RPT_* status (rpt_errors.h) and never unwinds across the boundary.
 * There is NO CPU fallback: without a CUDA device `rpt_create` fails with RPT_ERR_NO_DEVICE.
 *
 * Reference call -> replacement
 *   gpgpu::Framework (lazy_static FW, src/trace.rs:3-5,25-38)            -> rpt_create / rpt_destroy
 *   World::into_gpu: GpuBuffer::from_slice x5 + atlas GpuConstImage
 *     (src/asset.rs:226-235, src/bvh.rs:40-43), skybox GpuConstImage
 *     (src/asset.rs:257-281, src/trace.rs:144)                            -> rpt_upload_world
 *   GpuUniformBuffer::from_slice / config_buffer.write
 *     (src/trace.rs:168,219)                                              -> rpt_set_config
 *   GpuBuffer::from_slice(rng) / rng_buffer.write (src/trace.rs:169,221)  -> rpt_write_rng
 *   GpuBuffer::from_slice(output) / output_buffer.write
 *     (src/trace.rs:170,220)                                              -> rpt_write_output
 *   Shader/DescriptorSet/Program/Kernel::new (src/trace.rs:97-122)        -> (inside rpt_upload_world)
 *   `for _ in 0..sync_rate { kernel.enqueue(w/8,h/8,1); FW.poll_blocking() }`
 *     (src/trace.rs:182-193)                                              -> rpt_enqueue(n) + rpt_sync
 *   output_buffer.read_blocking + `/ sample_count` (src/trace.rs:198-204) -> rpt_read_output /
 *                                                                            rpt_read_framebuffer
 *
 * Semantics kept from the reference kernel (kernels/src/lib.rs:189-227): one "sample" advances
 * EVERY pixel of this context's partition by one sample index — output[i] += (radiance, 1) and
 * rng[i] = (x + 1, y).  rpt_enqueue(n) is n such samples, batched on the device.
 */
#ifndef RPT_B200_H
#define RPT_B200_H

#include "rpt_errors.h"
#include "rpt_shared_structs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rpt_context rpt_context;

/* How rpt_enqueue runs the loop.  WAVEFRONT is the product path; MEGAKERNEL is the 1:1
 * one-thread-per-pixel form of the reference kernel (binary BVH, reference traversal order),
 * kept as the bit-exact comparison arm. */
enum { RPT_PIPELINE_WAVEFRONT = 0, RPT_PIPELINE_MEGAKERNEL = 1 };

/* Device counters accumulated since the last rpt_reset_counters (all in units of rays/paths). */
typedef struct RptCounters {
    uint64_t paths;        /* pixel-samples finished */
    uint64_t nearest_rays; /* intersect_nearest calls (primary + bounce rays) */
    uint64_t any_rays;     /* intersect_any calls (shadow rays) */
    uint64_t kernel_launches;
} RptCounters;

/* ---- lifetime -------------------------------------------------------------------------- */
int rpt_create(int device_id, rpt_context** out_ctx);
int rpt_destroy(rpt_context* ctx);
/* Message of the last failure on this context ("" if none).  ctx == NULL: last rpt_create failure. */
const char* rpt_last_error(const rpt_context* ctx);
int rpt_set_pipeline(rpt_context* ctx, int pipeline);
/* Tunable: path slots kept in flight per wave of the wavefront pipeline (0 = default). */
int rpt_set_wave_slots(rpt_context* ctx, uint32_t slots);

/* ---- scene ----------------------------------------------------------------------------- */
/* Host pointers; the library copies and re-lays-out (wide BVH, triangle streams) privately.
 * atlas_rgba8 may be NULL (no textured material); sky_rgba32f may be NULL (2x2 magenta
 * fallback, src/asset.rs:275-281).  lights must hold >= 1 entry (sentinel allowed). */
int rpt_upload_world(rpt_context* ctx,
                     const RptPerVertexData* vertices, uint32_t nvertices,
                     const uint32_t* triangles_xyzw, uint32_t ntriangles,
                     const RptBVHNode* nodes, uint32_t nnodes,
                     const RptMaterialData* materials, uint32_t nmaterials,
                     const RptLightPickEntry* lights, uint32_t nlights,
                     const uint8_t* atlas_rgba8, uint32_t atlas_w, uint32_t atlas_h,
                     const float* sky_rgba32f, uint32_t sky_w, uint32_t sky_h);
/* Build half of SURVEY §8 f4.  rpt_upload_world with nodes == NULL and nnodes == 0 builds the traversal structure ON THE
 * DEVICE from the vertices and triangles alone (Morton order + bottom-up fit; src/bvh.rs:257-323 is then not needed
 * on the host): triangle ids keep referring to the caller's index buffer, results stay within the same parity bars
 * (nearest hits do not depend on the tree), traversal is slower than over the reference's SAH tree.
 * rpt_refit_world: the vertices moved but the topology did not — new vertex records in, triangle streams rewritten,
 * every box refitted bottom-up on the device; `lights` optionally replaces the light-pick table (NULL keeps it and only
 * refreshes the emitters' geometry).  After a refit the megakernel comparison arm has no reference BVH to traverse. */
int rpt_refit_world(rpt_context* ctx, const RptPerVertexData* vertices, uint32_t nvertices, const RptLightPickEntry* lights, uint32_t nlights);

/* ---- per-render state ------------------------------------------------------------------ */
/* (Re)allocates rng/output when width*height changes; validates the RNG dimension budget. */
int rpt_set_config(rpt_context* ctx, const RptTracingConfig* config);
int rpt_write_rng(rpt_context* ctx, const uint32_t* seeds_xy, size_t npixels);
int rpt_read_rng(rpt_context* ctx, uint32_t* seeds_xy, size_t npixels);
/* rgba == NULL zeroes the accumulator (the flush of src/trace.rs:220). */
int rpt_write_output(rpt_context* ctx, const float* rgba, size_t npixels);

/* Multi-GPU partition of the frame.  Sample-index-range splits need no call:
 * the host offsets seeds_xy[].x per rank.  Tile split: this context renders only pixels whose
 * 32x32 tile index t satisfies t % tile_count == tile_rank; other pixels stay untouched. */
int rpt_set_tile_partition(rpt_context* ctx, uint32_t tile_rank, uint32_t tile_count);

/* ---- host staging memory --------------------------------------------------------------- */
/* Page-locked host memory for the buffers that cross the boundary every batch (seeds in, frame out):
 * the counterpart of gpgpu-rs's mapped staging buffers behind GpuBuffer::write / read_blocking
 * (src/trace.rs:198,219-221).  Any host pointer is accepted by the calls above and below; buffers from
 * rpt_host_alloc are copied at full PCIe/C2C rate instead of through the driver's bounce buffer. */
int rpt_host_alloc(size_t bytes, void** out_ptr);
int rpt_host_free(void* ptr);

/* ---- run ------------------------------------------------------------------------------- */
int rpt_enqueue(rpt_context* ctx, uint32_t n_samples); /* asynchronous */
int rpt_sync(rpt_context* ctx);                        /* FW.poll_blocking() */
/* rpt_enqueue that can be cut short like the reference's dispatch loop, which looks at `interacting | dirty` and
 * `running` after every sample (src/trace.rs:182-193): after every `poll_samples` samples (0 = 1) the batch is drained and
 * *stop_flag — host memory another thread may write, or NULL — is read; non-zero ends the batch.  *finished_out =
 * samples rendered (a multiple of poll_samples unless the batch ended).  Returns with the stream idle. */
int rpt_enqueue_interruptible(rpt_context* ctx, uint32_t n_samples, const volatile uint32_t* stop_flag, uint32_t poll_samples,
                              uint32_t* finished_out);

/* ---- readback -------------------------------------------------------------------------- */
/* Raw accumulator, float[4] per pixel (sum rgb, w = sample count). Implies rpt_sync. */
int rpt_read_output(rpt_context* ctx, float* rgba, size_t npixels);
/* Packed RGB `output.xyz / samples` normalised on the device (src/trace.rs:199-204). */
int rpt_read_framebuffer(rpt_context* ctx, float* rgb, size_t npixels, float samples);
/* The same without the wait: normalise on the render stream, copy on a copy stream into `rgb` (page-locked memory:
 * rpt_host_alloc; it must stay valid until rpt_readback_wait), two frames in flight at most.  The next rpt_enqueue
 * overlaps the copy; rpt_readback_wait (or rpt_sync) returns when `rgb` holds the frame. */
int rpt_read_framebuffer_async(rpt_context* ctx, float* rgb, size_t npixels, float samples);
int rpt_readback_wait(rpt_context* ctx);
/* Post-normalise hook — the slot of the reference's denoiser (src/trace.rs:207-210: `denoise_image(w, h, &mut
 * image_buffer)` between the division by the sample count and the hand-over to the display): called by
 * rpt_read_framebuffer[_async] after the frame has been normalised on the device and before it is copied out, with the
 * DEVICE pointer of the packed RGB frame (width * height * 3 floats, modifiable in place) and the cudaStream_t the
 * library works on — the hook must enqueue its work on that stream (or synchronise itself).  NULL removes it. */
typedef void (*rpt_frame_hook)(float* rgb_device, uint32_t width, uint32_t height, void* cuda_stream, void* user);
int rpt_set_frame_hook(rpt_context* ctx, rpt_frame_hook hook, void* user);
/* Display resolve on the device — the fragment stage of src/resources/render.wgsl:150-185 (fs_main):
 * rgb = tonemap(output.xyz / samples), row-major, top-left origin.  `tonemap` is the reference's `Tonemapping`
 * enum value (src/app.rs:20-28): 0 none, 1 Reinhard, 2 ACES Narkowicz (input x 0.6), 3 ACES Narkowicz
 * "overexposed", 4 ACES Hill, 5 Neutral, 6 Uncharted; any other value means none, like the shader's default arm. */
int rpt_read_display(rpt_context* ctx, float* rgb, size_t npixels, float samples, uint32_t tonemap);
/* The same, stored as the colour attachment save_render reads back (src/app.rs:759-840): 4 bytes per pixel in
 * R,G,B,A order (the reference swizzles its BGRA surface to this before writing the PNG), alpha 255, channels
 * clamped to [0,1] (NaN -> 0) and rounded to the nearest of 256 levels; srgb_encode != 0 applies the linear ->
 * sRGB transfer first, which is what an *Srgb surface format does in hardware.  A third of the readback bytes. */
int rpt_read_display_rgba8(rpt_context* ctx, uint8_t* rgba, size_t npixels, float samples, uint32_t tonemap, uint32_t srgb_encode);
/* Diagnostics: bounce-0 triangle_index per pixel (0xFFFFFFFF = miss) for the NEXT sample index
 * of the current rng state, without advancing any state. */
int rpt_read_primary_ids(rpt_context* ctx, uint32_t* triangle_ids, size_t npixels);
int rpt_get_counters(rpt_context* ctx, RptCounters* out);
int rpt_reset_counters(rpt_context* ctx);
/* Milliseconds the device spent in the kernels of all rpt_enqueue calls since the last
 * rpt_reset_counters (CUDA events on the context's stream). */
int rpt_get_device_ms(rpt_context* ctx, float* ms);

/* Device time of a region of calls (enqueues and the combine) from CUDA events on the library's own streams. */
int rpt_timer_start(rpt_context* ctx);
int rpt_timer_stop(rpt_context* ctx, float* ms); /* waits for the region's work */

/* Per-stage device time, for roofline accounting.  While enabled, every kernel launch of
 * rpt_enqueue is bracketed by CUDA events on the context's stream (this perturbs the pipeline a
 * little, so whole-job throughput is measured with it off). */
enum { RPT_STAGE_GENERATE = 0, RPT_STAGE_EXTEND, RPT_STAGE_MISS, RPT_STAGE_SHADE, RPT_STAGE_SHADOW, RPT_STAGE_ACCUMULATE,
       RPT_STAGE_MEGAKERNEL, RPT_STAGE_OTHER, RPT_STAGE_COUNT };
typedef struct RptStageTiming {
    float ms[RPT_STAGE_COUNT];          /* summed launch durations */
    uint64_t launches[RPT_STAGE_COUNT]; /* launches timed */
} RptStageTiming;
int rpt_set_stage_timing(rpt_context* ctx, int enable);
int rpt_get_stage_timing(rpt_context* ctx, RptStageTiming* out); /* syncs; clears the totals */

/* Traversal statistics, for roofline accounting in the backend's OWN layout.  While enabled, rpt_enqueue launches a
 * counting build of the trace kernels (a little slower; the product build carries no counters): node visits (one
 * 80-byte wide node each) and ray/triangle tests (one 48-byte record each) since the last rpt_reset_counters. */
typedef struct RptTraceStatistics {
    uint64_t nearest_rays, nearest_node_visits, nearest_triangle_tests;
    uint64_t any_rays, any_node_visits, any_triangle_tests;
    uint64_t shaded_hits;               /* surface hits shaded (counted by the product build too) */
    uint32_t node_bytes, triangle_bytes; /* bytes read per visit / per test */
} RptTraceStatistics;
int rpt_get_sm_count(rpt_context* ctx, int* sm_count); /* multiprocessors of the context's device (persistent grids are multiples of it) */
int rpt_set_trace_statistics(rpt_context* ctx, int enable);
int rpt_get_trace_statistics(rpt_context* ctx, RptTraceStatistics* out);

/* ---- multi-GPU combine over NCCL (one rank per GPU) ------------------------------------- */
/* id_bytes: 128-byte ncclUniqueId made by rank 0 and distributed by the host (any channel). */
int rpt_comm_unique_id(uint8_t* id_bytes_128);
int rpt_comm_init(rpt_context* ctx, const uint8_t* id_bytes_128, int rank, int nranks);
/* Combine the per-rank accumulators on `root` over NVLink.  NO accumulator is modified (calling it twice combines the
 * same samples twice into the same result): every rank snapshots its contribution on its render stream, the exchange
 * runs on a side stream — the next rpt_enqueue overlaps it — and the result is root's COMBINED FRAME, which root's
 * rpt_read_output / rpt_read_framebuffer / rpt_read_display* return from then on (they wait for the exchange);
 * rpt_write_output or a resize drops it.  Whole-frame ranks (sample-index split): ncclReduce(sum).  Tile ranks
 * (rpt_set_tile_partition(rank, nranks) on every rank): each rank sends only the pixels it owns (ncclSend / ncclRecv),
 * root scatters them — bit-identical to one GPU. */
int rpt_comm_reduce_output(rpt_context* ctx, int root);
int rpt_comm_destroy(rpt_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RPT_B200_H */
