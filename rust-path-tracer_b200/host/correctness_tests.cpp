// correctness_tests.cpp — the reference's tests/correctness_tests.rs, against the CUDA backend.
//   usage: correctness_tests <FurnaceTest.rptw>
// furnace_test(use_mis): 128x128, 32 spp, pixel (65,75), every channel ^(1/2.2) within 0.02 of 0.8.
// (The reference also runs the two cases with use_cpu = true; its CPU path is this repo's oracle and is
// tested in tests/test_oracle_furnace.py.)
#include <cmath>
#include <cstdio>

#include "trace.hpp"

static bool furnace_test(const char* scene, bool use_mis) {
    const size_t size = 128, coord_x = 65, coord_y = 75;
    const float albedo = 0.8f, tolerance = 0.02f;
    auto state = rpt::setup_trace((uint32_t)size, (uint32_t)size, 32);
    if (use_mis) {
        std::unique_lock<std::shared_mutex> l(state->config_lock);
        state->config.nee = RPT_NEE_MIS;  // NextEventEstimation::MultipleImportanceSampling.to_u32()
    }
    if (rpt::trace_gpu(scene, nullptr, state) != RPT_OK) return false;
    std::shared_lock<std::shared_mutex> l(state->framebuffer_lock);
    bool ok = state->samples.load() >= 32;
    for (size_t c = 0; c < 3; ++c) {
        const float pixel = std::pow(state->framebuffer[(size * 3) * coord_y + coord_x * 3 + c], 1.0f / 2.2f);
        std::printf("  furnace_test_gpu%s channel %zu: %.4f\n", use_mis ? "_mis" : "", c, pixel);
        ok = ok && std::fabs(pixel - albedo) < tolerance;
    }
    return ok;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s FurnaceTest.rptw\n", argv[0]); return 2; }
    int failed = 0;
    for (bool mis : {false, true}) {
        const bool ok = furnace_test(argv[1], mis);
        std::printf("test furnace_test_gpu%s ... %s\n", mis ? "_mis" : "", ok ? "ok" : "FAILED");
        failed += !ok;
    }
    return failed ? 1 : 0;
}
