// wavefront_miss.cu — the HDR lat-long sky lookup (lib.rs:70-78) for the compacted queue of escaped paths.
//
// Its own translation unit because it is built with IEEE division and sqrt, unlike wavefront_shade.cu (which holds
// the procedural-sky variant of this stage): an HDR sky can put a 5 000-nit sun on a few texels, and that gradient
// multiplies a 2-ulp error in the lat-long coordinate — atan2f divides internally — into visible radiance error.
#include "device_scene.h"

namespace rpt {

__global__ void __launch_bounds__(128) wf_miss_hdr_kernel(FrameParams f, WaveState s) {
    const uint32_t n = s.ctl->n_miss;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = __ldg(s.q_miss + i);
        const f3 rd = xyz(s.ray_d[slot]), throughput = xyz(s.thr[slot]);
        const f3 c = throughput * f.sky.lookup(rd);  // NOT NaN-masked in the reference (lib.rs:77)
        float4 r = s.rad[slot];
        r.x += c.x; r.y += c.y; r.z += c.z;
        s.rad[slot] = r;
    }
}

void launch_wf_miss_procedural(const WaveLaunch& l, const FrameParams& f, const WaveState& s);  // wavefront_shade.cu

void launch_wf_miss(const WaveLaunch& l, const FrameParams& f, const WaveState& s) {
    if (f.has_skybox) wf_miss_hdr_kernel<<<l.grid * 8, 128, 0, l.stream>>>(f, s);
    else launch_wf_miss_procedural(l, f, s);
}

}  // namespace rpt
