/*
 * rpt_host.h — C ABI of the host-side input producers either side of the tracing hot path
 * (SURVEY.md §8 f1).  Pure CPU code, as in the reference, where these run once at scene load:
 *
 *   rpt_build_bvh               <- BVHBuilder::new(..).sah_samples(128).build(), src/bvh.rs:58-324
 *                                  (called from src/asset.rs:196; permutes the index buffer in place)
 *   rpt_build_light_pick_table  <- compute_emissive_mask + build_light_pick_table,
 *                                  src/light_pick.rs:13-122 (called from src/asset.rs:201-202)
 *   rpt_pack_per_vertex         <- the PerVertexData packing loop, src/asset.rs:205-215
 *   rpt_make_rng_seeds          <- blue-noise / uniform seed tables, src/trace.rs:149-160, 245-256
 *   rpt_atlas_rects / rpt_atlas_pack / rpt_decode_albedo_gamma
 *                               <- pack_textures, src/atlas.rs:26-92, and the albedo decode of src/asset.rs:140-147
 *   rpt_decode_hdr / rpt_sky_texels
 *                               <- load_dynamic_image's .hdr branch and the GPU / CPU texel conventions of the sky image,
 *                                  src/asset.rs:238-273
 *   rpt_tile_partition_pixels   <- (new) the tile split used by rpt_set_tile_partition
 *   rpt_camera_matrix           <- Mat3::from_rotation_y(ry) * Mat3::from_rotation_x(rx),
 *                                  kernels/src/lib.rs:50 (host libm keeps primary rays bit-exact)
 *
 * All functions return 0 on success or a negative RPT_ERR_* code; nothing throws across the ABI.
 */
#ifndef RPT_HOST_H
#define RPT_HOST_H

#include <stddef.h>

#include "rpt_shared_structs.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Binned-SAH BVH build over `ntris` triangles.
 *   vertices : nverts x float[4] positions (w ignored)
 *   indices  : ntris x uint32[4] (i0,i1,i2,material) — permuted IN PLACE into BVH leaf order
 *   nodes_out: capacity 2*ntris-1 nodes; *nnodes_out receives the used count. */
int rpt_build_bvh(const float* vertices, uint32_t nverts, uint32_t* indices, uint32_t ntris,
                  uint32_t sah_samples, RptBVHNode* nodes_out, uint32_t* nnodes_out);

/* Power-weighted two-outcome light-pick table over the emissive triangles (post-BVH order).
 *   table_out: capacity max(ntris,1) entries; a scene without emitters yields the one-entry
 *   sentinel {ratio = -1}. */
int rpt_build_light_pick_table(const float* vertices, uint32_t nverts, const uint32_t* indices, uint32_t ntris,
                               const RptMaterialData* materials, uint32_t nmaterials,
                               RptLightPickEntry* table_out, uint32_t* nentries_out);

/* Interleave vertex / normal / tangent / uv streams into PerVertexData (missing streams -> 0). */
int rpt_pack_per_vertex(const float* vertices, const float* normals, const float* tangents, const float* uvs,
                        uint32_t nverts, RptPerVertexData* out);

/* Per-pixel rng seeds.  blue != NULL: x = 0, y = u32(R8(x % bw, y % bh) / 255 * 4294967295.0)
 * (blue-noise mode); blue == NULL: x = splitmix-style uniform from `uniform_seed`, y = 0. */
int rpt_make_rng_seeds(const uint8_t* blue_r8, uint32_t bw, uint32_t bh, uint32_t width, uint32_t height,
                       uint64_t uniform_seed, uint32_t* seeds_xy_out);

/* Texture atlas (src/atlas.rs).  Leaf rectangles (x, y, w, h) of `ntextures` textures in packing order. */
int rpt_atlas_rects(uint32_t ntextures, uint32_t atlas_w, uint32_t atlas_h, uint32_t* rects_xywh_out);
/* Pack RGBA8 textures (row-major, widths[i] x heights[i]) into a zeroed atlas_w x atlas_h RGBA8 atlas: Lanczos3
 * resize to the leaf when the size differs, vertical flip, copy.  sts_out: ntextures x float[4] (x/W, y/W, w/W, h/H). */
int rpt_atlas_pack(const uint8_t* const* textures_rgba8, const uint32_t* widths, const uint32_t* heights, uint32_t ntextures,
                   uint32_t atlas_w, uint32_t atlas_h, uint8_t* atlas_rgba8_out, float* sts_out);
/* `((p / 255).powf(2.2) * 255) as u8` on RGB, alpha -> 255 (albedo textures, before packing). */
int rpt_decode_albedo_gamma(const uint8_t* rgba8_in, size_t npixels, uint8_t* rgba8_out);

/* Radiance .hdr (RGBE) file -> packed RGB float (src/asset.rs:238-254; `image` 0.24 HdrDecoder semantics: RGBE value
 * c * 2^(e-136), e == 0 -> 0; "-Y h +X w" orientation only).  rgb_out may be NULL to query the size. */
int rpt_decode_hdr(const uint8_t* bytes, size_t nbytes, float* rgb_out, uint32_t* width_out, uint32_t* height_out);
/* Sky texels as the tracing loop reads them, float[4] per texel.  cpu_path_rgb8 == 0: (r, g, b, 1), the GPU path's
 * Rgba32Float upload (src/asset.rs:257-264); != 0: the CPU path's `into_rgb8()` quantisation, every channel clamped to
 * [0, 1], rounded to 8 bits and divided by 255 (src/asset.rs:266-273). */
int rpt_sky_texels(const float* rgb, uint32_t width, uint32_t height, int cpu_path_rgb8, float* rgba_out);

/* Multi-GPU tile split: pixel indices (row-major y*width+x) of the 32x32 tiles t
 * with t % tile_count == tile_rank, tile by tile.  pixels_out may be NULL to query the count. */
int rpt_tile_partition_pixels(uint32_t width, uint32_t height, uint32_t tile_rank, uint32_t tile_count,
                              uint32_t* pixels_out, uint32_t* npixels_out);

/* 3x3 camera rotation, column-major (col0,col1,col2), computed with the host libm. */
int rpt_camera_matrix(float rot_x, float rot_y, float* m9_out);

#ifdef __cplusplus
}
#endif
#endif /* RPT_HOST_H */
