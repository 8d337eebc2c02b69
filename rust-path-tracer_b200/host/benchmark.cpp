// benchmark.cpp — the reference's benches/benchmark.rs cases, against the CUDA backend.
//   usage: benchmark <DarkCornell.rptw> [<BreakTime(.proxy).rptw>]
// "Startup time (GPU)": trace_gpu(BreakTime, None, setup_trace(1280, 720, 0))    (reference comment: 3.021 s)
// "160 samples (GPU)":  trace_gpu(DarkCornell, None, setup_trace(1280, 720, 160)) (reference comment: 2.408 s)
// Like criterion's, every timing includes scene load, BVH build and upload.
#include <chrono>
#include <cstdio>

#include "trace.hpp"

static double run(const char* scene, uint32_t samples) {
    const auto t0 = std::chrono::steady_clock::now();
    rpt::trace_gpu(scene, nullptr, rpt::setup_trace(1280, 720, samples));
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s DarkCornell.rptw [BreakTime.rptw]\n", argv[0]); return 2; }
    run(argv[1], 32);  // warm-up: CUDA context creation, module load
    double best = 1e30;
    if (argc > 2) {  // (fresh boxes page the file and the allocator in on the first pass: best of three, like the case below)
        for (int i = 0; i < 3; ++i) best = std::min(best, run(argv[2], 0));
        std::printf("Startup time (GPU): %.3f s   [reference comment: 3.021 s on an unstated GPU]\n", best);
    }
    best = 1e30;
    for (int i = 0; i < 3; ++i) best = std::min(best, run(argv[1], 160));
    std::printf("160 samples (GPU): %.3f s  -> %.1f Mpaths/s incl. load   [reference comment: 2.408 s, >= 61.2 Mpaths/s]\n", best,
                1280.0 * 720.0 * 160.0 / best / 1e6);
    return 0;
}
