/*
 * rpt_shared_structs.h — host/device shared record layouts of the tracing hot path.
 *
 * These are the `#[repr(C)]` layouts of the reference's `shared_structs` crate, restated as
 * plain C structs so a Rust host can pass its own `Vec<T>` buffers through the FFI unchanged
 * (`bytemuck::cast_slice`).  Every struct cites the reference definition it mirrors; sizes
 * and offsets are pinned with static asserts below.
 *
 *   RptTracingConfig   <- shared_structs/src/lib.rs:12-25   (80 B uniform)
 *   RptMaterialData    <- shared_structs/src/lib.rs:44-56   (96 B)
 *   RptPerVertexData   <- shared_structs/src/lib.rs:92-100  (64 B)
 *   RptLightPickEntry  <- shared_structs/src/lib.rs:102-112 (28 B)
 *   RptBVHNode         <- shared_structs/src/lib.rs:121-126 (32 B)
 *   index buffer       <- UVec4 (i0,i1,i2,material), src/asset.rs:106
 *   rng buffer         <- UVec2 (x = sample index, y = per-pixel offset), kernels/src/rng.rs:34-49
 *   output buffer      <- Vec4 running sum (rgb, w = sample count), kernels/src/lib.rs:185,225
 */
#ifndef RPT_SHARED_STRUCTS_H
#define RPT_SHARED_STRUCTS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* shared_structs/src/lib.rs:12-25 */
typedef struct RptTracingConfig {
    float cam_position[4];          /* @0  xyz used */
    float cam_rotation[4];          /* @16 x = pitch, y = yaw (radians) */
    uint32_t width;                 /* @32 */
    uint32_t height;                /* @36 */
    uint32_t min_bounces;           /* @40 Russian roulette starts after this bounce */
    uint32_t max_bounces;           /* @44 */
    float sun_direction[4];         /* @48 xyz direction, w intensity */
    uint32_t nee;                   /* @64 0 none, 1 MIS, 2 direct only (lib.rs:193-227) */
    uint32_t has_skybox;            /* @68 0 procedural sky, else lat-long image */
    float specular_weight_clamp[2]; /* @72 */
} RptTracingConfig;

/* shared_structs/src/lib.rs:44-56: each float[4] is a colour OR an atlas rect (u0,v0,su,sv) */
typedef struct RptMaterialData {
    float emissive[4];
    float albedo[4];
    float roughness[4];
    float metallic[4];
    float normals[4];
    uint32_t has_albedo_texture;
    uint32_t has_metallic_texture;
    uint32_t has_roughness_texture;
    uint32_t has_normal_texture;
} RptMaterialData;

/* shared_structs/src/lib.rs:92-100 */
typedef struct RptPerVertexData {
    float vertex[4];
    float normal[4];
    float tangent[4];
    float uv0[2];
    float uv1[2];
} RptPerVertexData;

/* shared_structs/src/lib.rs:102-112; ratio < 0 marks the "no lights" sentinel (:115-119) */
typedef struct RptLightPickEntry {
    uint32_t triangle_index_a;
    float triangle_area_a;
    float triangle_pick_pdf_a;
    uint32_t triangle_index_b;
    float triangle_area_b;
    float triangle_pick_pdf_b;
    float ratio;
} RptLightPickEntry;

/* shared_structs/src/lib.rs:121-126: aabb_min.w = bit-cast u32 triangle_count (>0 => leaf),
 * aabb_max.w = bit-cast u32 left child (inner; right = left + 1) or first triangle (leaf). */
typedef struct RptBVHNode {
    float aabb_min[3];
    uint32_t triangle_count;
    float aabb_max[3];
    uint32_t left_or_first;
} RptBVHNode;

/* NextEventEstimation, shared_structs/src/lib.rs:193-199 */
enum { RPT_NEE_NONE = 0, RPT_NEE_MIS = 1, RPT_NEE_DIRECT = 2 };

#ifdef __cplusplus
}
static_assert(sizeof(RptTracingConfig) == 80, "TracingConfig is an 80-byte uniform");
static_assert(offsetof(RptTracingConfig, width) == 32 && offsetof(RptTracingConfig, sun_direction) == 48 &&
              offsetof(RptTracingConfig, nee) == 64 && offsetof(RptTracingConfig, specular_weight_clamp) == 72,
              "TracingConfig field offsets");
static_assert(sizeof(RptMaterialData) == 96, "MaterialData is 96 bytes");
static_assert(sizeof(RptPerVertexData) == 64, "PerVertexData is 64 bytes");
static_assert(sizeof(RptLightPickEntry) == 28, "LightPickEntry is 28 bytes");
static_assert(sizeof(RptBVHNode) == 32, "BVHNode is 32 bytes");
#endif

#endif /* RPT_SHARED_STRUCTS_H */
