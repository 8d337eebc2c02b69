// trace.hpp — C++ host-side mirror of the reference's trace driver (src/trace.rs) above the C ABI.
//
// The reference host is Rust; this image has no Rust toolchain, so the compiled-language host the
// reference interface maps onto is C++.  Names, fields, argument meaning and loop structure follow
// src/trace.rs so that host/correctness_tests.cpp reads like tests/correctness_tests.rs and
// host/benchmark.cpp like benches/benchmark.rs:
//
//   rpt::TracingState            <- pub struct TracingState          (src/trace.rs:40-92)
//   rpt::setup_trace(w, h, n)    <- pub fn setup_trace               (src/trace.rs:331-344)
//   rpt::trace_gpu(scene, sky, state) <- pub fn trace_gpu            (src/trace.rs:136-224)
//   rpt::World::from_path        <- World::from_path                 (src/asset.rs:55-224)
//
// There is no trace_cpu: the reference's CPU path exists in this repo only as the test oracle.
// Scene files: the reference imports .glb through assimp; here World::from_path reads the baked
// `.rptw` container written by rust-path-tracer_b200/glb.py (BakedScene.save_rptw) — the output of
// that import — and then runs the same steps as the reference: BVH build (permuting the index
// buffer), light-pick table, PerVertexData packing.
#pragma once

#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <optional>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/rpt_b200.h"
#include "../../include/rpt_host.h"

namespace rpt {

RptTracingConfig default_config();  // TracingConfig::default(), shared_structs/src/lib.rs:27-42

struct TracingState {
    std::shared_mutex framebuffer_lock;  // RwLock<Vec<f32>>
    std::vector<float> framebuffer;      // packed RGB, output.xyz / samples
    std::atomic<bool> running{false};
    std::atomic<uint32_t> samples{0};
    std::atomic<bool> denoise{false};
    std::atomic<uint32_t> sync_rate{32};
    std::atomic<bool> use_blue_noise{true};
    std::atomic<bool> interacting{false};
    std::atomic<bool> dirty{false};
    // (not in the reference) an interactive host sets this so that a batch is rendered sample by sample with the control
    // flags read in between, exactly like the reference's dispatch loop; benches and tests leave it off: one enqueue per batch
    std::atomic<bool> poll_every_sample{false};
    std::shared_mutex config_lock;  // RwLock<TracingConfig>
    RptTracingConfig config;

    TracingState(uint32_t width, uint32_t height);
    RptTracingConfig read_config() { std::shared_lock<std::shared_mutex> l(config_lock); return config; }
};

// Harness for synchronous tracing: sets `running` and spawns a thread that clears it once
// `samples >= n` (src/trace.rs:331-344).
std::shared_ptr<TracingState> setup_trace(uint32_t width, uint32_t height, uint32_t samples);

struct World {  // src/asset.rs:9-16
    std::vector<RptPerVertexData> per_vertex_buffer;
    std::vector<uint32_t> index_buffer;  // 4 per triangle, BVH leaf order
    std::vector<RptBVHNode> nodes;
    std::vector<RptMaterialData> material_data_buffer;
    std::vector<RptLightPickEntry> light_pick_buffer;
    std::vector<uint8_t> atlas;  // RGBA8 or empty
    uint32_t atlas_w = 0, atlas_h = 0;
    static std::optional<World> from_path(const std::string& path);
};

// Returns silently when the scene cannot be loaded, like the reference (`else { return; }`); device
// failures are reported on stderr and end the loop (the reference panics).  Returns the last status.
int trace_gpu(const std::string& scene_path, const char* skybox_path, std::shared_ptr<TracingState> state, int device = 0);

}  // namespace rpt
