// wavefront_trace_nofma.cu — the ray-casting stages of the wavefront pipeline:
//   generate        camera rays for every slot of a wave          (kernels/src/lib.rs:36-51)
//   extend          nearest-hit traversal + hit/miss compaction   (intersection.rs:165-167 intersect_nearest)
//   shadow-connect  any-hit traversal of the NEE shadow rays      (intersection.rs:169-171 intersect_any)
//
// Built with -fmad=false: ray generation and the ray/triangle test evaluate the reference's fp32
// operations in its order (dev/exact.cuh), which is what makes primary-hit ids bit-exact.
//
// All three are persistent kernels: the grid is a fixed multiple of the SM count and warps
// stride over the queue, whose length lives in device memory (WaveCtl) so the host never has to
// read it back.  Each lane keeps its traversal stack in a per-warp shared-memory slab
// (stack[depth][lane]: conflict-free 8-byte accesses), and queue appends are warp-aggregated:
// one atomicAdd per warp, slots handed out by ballot + popc.
#include "device_scene.h"
#include "wide_bvh.h"

namespace rpt {

constexpr int kTraceBlock = 128;  // 4 warps; 24 KB of stack slabs per block
constexpr int kTraceWarps = kTraceBlock / 32;

struct SmemStack {
    uint2* column;  // this lane's column of the warp slab; entries are 32 lanes apart
    int n;
    __device__ __forceinline__ void push(uint2 v) {
        if (n < (int)kWideStackCapacity) column[n * 32] = v;  // depth is validated at upload; never drop silently there
        ++n;
    }
    __device__ __forceinline__ uint2 pop() { --n; return column[n * 32]; }
    __device__ __forceinline__ bool empty() const { return n == 0; }
};

// Append `value` to a queue for every lane with `pred`; one atomic per warp.
__device__ __forceinline__ void warp_append(bool pred, uint32_t* queue, uint32_t* count, uint32_t value) {
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, pred);
    if (mask == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs((int)mask) - 1)) base = atomicAdd(count, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs((int)mask) - 1);
    if (pred) queue[base + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = value;
}

__device__ __forceinline__ uint32_t wave_pixel(const WaveDesc& d, uint32_t j) {
    const uint32_t i = d.pix_base + j;
    return d.pixel_map ? __ldg(d.pixel_map + i) : i;
}

__global__ void __launch_bounds__(256) wf_generate_kernel(FrameParams f, WaveState s, WaveDesc d, const uint2* __restrict__ rng) {
    const uint32_t nslots = d.npix * d.k_samples;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nslots; slot += gridDim.x * blockDim.x) {
        const uint32_t j = slot % d.npix, k = slot / d.npix;
        const uint32_t pixel = wave_pixel(d, j);
        const uint2 seed = __ldg(rng + pixel);
        Rng r{seed.x + k + seed.y, 0u};
        f3 ro, rd;
        camera_ray(f.camera, pixel % f.width, pixel / f.width, r, ro, rd);
        s.ray_o[slot] = mk4(ro, 0.0f);
        s.ray_d[slot] = mk4(rd, __uint_as_float(r.dim));  // dimension cursor = 2, last lobe = diffuse (BSDFSample::default)
        s.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        s.rad[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

__global__ void __launch_bounds__(kTraceBlock) wf_extend_kernel(WideScene bvh, WaveState s, int in_queue, bool identity, uint32_t n_identity) {
    __shared__ uint2 slabs[kTraceWarps][kWideStackCapacity][32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n = identity ? n_identity : s.ctl->n_ext[in_queue];
    const uint32_t* __restrict__ queue = s.q_ext[in_queue];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.counters + 1, (unsigned long long)n);
    const uint32_t stride = gridDim.x * blockDim.x;
    // whole warps iterate together so the ballots below always see 32 lanes
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += stride) {
        const uint32_t i = base + lane;
        const bool active = i < n;
        uint32_t slot = 0;
        WideHit h{1000000.0f, 0u, false, false};
        if (active) {
            slot = identity ? i : __ldg(queue + i);
            const float4 o = s.ray_o[slot], dv = s.ray_d[slot];
            SmemStack st{&slabs[warp][0][lane], 0};
            h = wide_intersect<true>(bvh, xyz(o), xyz(dv), 0.0f, st);
            if (h.hit) s.hit[slot] = make_uint2(__float_as_uint(h.t), h.triangle | (h.backface ? 0x80000000u : 0u));
        }
        __syncwarp();
        warp_append(active && h.hit, s.q_hit, &s.ctl->n_hit, slot);
        warp_append(active && !h.hit, s.q_miss, &s.ctl->n_miss, slot);
    }
}

__global__ void __launch_bounds__(kTraceBlock) wf_shadow_kernel(WideScene bvh, WaveState s) {
    __shared__ uint2 slabs[kTraceWarps][kWideStackCapacity][32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n = s.ctl->n_shadow;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.counters + 2, (unsigned long long)n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o = s.sh_o[i], dv = s.sh_d[i];
        SmemStack st{&slabs[warp][0][lane], 0};
        const WideHit h = wide_intersect<false>(bvh, xyz(o), xyz(dv), o.w, st);
        if (!h.hit) {  // unoccluded: radiance += mask_nan(contribution) (lib.rs:164; masked when queued)
            const uint32_t slot = __float_as_uint(dv.w);
            const float4 c = s.sh_c[i];
            float4 r = s.rad[slot];
            r.x += c.x; r.y += c.y; r.z += c.z;
            s.rad[slot] = r;
        }
    }
}

// Diagnostics: bounce-0 triangle ids of a wave, translated back to the reference's numbering.
__global__ void wf_export_primary_kernel(WideScene bvh, WaveState s, WaveDesc d, uint32_t* __restrict__ ids) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.npix) return;
    ids[wave_pixel(d, j)] = 0xFFFFFFFFu;
}
__global__ void wf_export_primary_hits_kernel(WideScene bvh, WaveState s, WaveDesc d, uint32_t* __restrict__ ids) {
    const uint32_t n = s.ctl->n_hit;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = s.q_hit[i];
        const uint32_t tri = s.hit[slot].y & 0x7FFFFFFFu;
        ids[wave_pixel(d, slot % d.npix)] = __float_as_uint(__ldg(bvh.tri_pos + 3u * (size_t)tri).w);
    }
}

void launch_wf_generate(const WaveLaunch& l, const FrameParams& f, const WaveState& s, const WaveDesc& d, const uint2* rng) {
    wf_generate_kernel<<<l.grid * 2, 256, 0, l.stream>>>(f, s, d, rng);
}
void launch_wf_extend(const WaveLaunch& l, const WideScene& bvh, const WaveState& s, int in_queue, bool identity_queue, uint32_t n_identity) {
    wf_extend_kernel<<<l.grid * 4, kTraceBlock, 0, l.stream>>>(bvh, s, in_queue, identity_queue, n_identity);
}
void launch_wf_shadow(const WaveLaunch& l, const WideScene& bvh, const WaveState& s) {
    wf_shadow_kernel<<<l.grid * 4, kTraceBlock, 0, l.stream>>>(bvh, s);
}
void launch_wf_export_primary(const WaveLaunch& l, const WideScene& bvh, const WaveState& s, const WaveDesc& d, uint32_t* ids) {
    wf_export_primary_kernel<<<(d.npix + 255) / 256, 256, 0, l.stream>>>(bvh, s, d, ids);
    wf_export_primary_hits_kernel<<<l.grid * 2, 256, 0, l.stream>>>(bvh, s, d, ids);
}

}  // namespace rpt
