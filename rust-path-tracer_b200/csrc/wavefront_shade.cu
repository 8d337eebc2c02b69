// wavefront_shade.cu — the shading stages of the wavefront pipeline:
//   shade       everything kernels/src/lib.rs:81-181 does at a surface hit: emission / MIS rules,
//               attribute interpolation, normal map, PBR lobe sampling, the NEE light sample
//               (stored as a shadow ray with its would-be contribution), throughput update,
//               Russian roulette, and the next ray
//   miss        procedural sky for escaped paths (lib.rs:66-69; the HDR lookup is wavefront_miss.cu)
//   accumulate  output[pixel] += (radiance, 1) per sample in sample order; rng.x += samples
//               (lib.rs:225-226 / src/trace.rs:295-296)
//   normalize   packed RGB framebuffer = output.xyz / samples (src/trace.rs:199-204)
//
// Materials and — when they fit kSmemLightBytes — the light bins and light records are staged in shared memory once
// per persistent block (kernels/src/light_pick.rs:8-16 reads one bin and one record per NEE sample, the MIS term
// one record per emitter hit).
#include "device_scene.h"

namespace rpt {

#ifndef RPT_SHADE_BLOCK
#define RPT_SHADE_BLOCK 256
#endif
#ifndef RPT_SHADE_MIN_BLOCKS
#define RPT_SHADE_MIN_BLOCKS 4
#endif
constexpr int kShadeBlock = RPT_SHADE_BLOCK;
constexpr int kShadeMinBlocks = RPT_SHADE_MIN_BLOCKS;  // resident blocks per SM the register allocation aims at
constexpr uint32_t kSmemMaterials = 64;  // 6 KB

// A light-table word: plain (shared-memory) load when the table is staged, read-only global load otherwise.
template <bool SMEM, class T>
__device__ __forceinline__ T ldl(const T* p) { return SMEM ? *p : __ldg(p); }

__device__ __forceinline__ uint32_t wave_pixel(const WaveDesc& d, uint32_t j) {
    const uint32_t i = d.pix_base + j;
    return d.pixel_map ? __ldg(d.pixel_map + i) : i;
}

// One thread per hit, no block-level synchronisation: what a hit produces is recorded in the word
// q_shaded[i] = slot | kShadedNoNext | kShadedShadow (i = the hit's position in q_hit) and
// wf_compact_shaded_kernel turns those words into the next extend queue and the shadow queue, in order.
// The shadow ray itself is stored at index i (sh_o / sh_d / sh_c), the next ray in the path's own slot.
template <bool MATS_IN_SMEM, bool LIGHTS_IN_SMEM, bool TANGENTS>
__global__ void __launch_bounds__(kShadeBlock, kShadeMinBlocks) wf_shade_kernel(FrameParams f, WideWorld w, WaveState s, WaveDesc d, const uint2* __restrict__ rng,
                                                               uint32_t bounce) {
    __shared__ RptMaterialData sm_materials[MATS_IN_SMEM ? kSmemMaterials : 1];
    __shared__ __align__(16) uint32_t sm_lights[LIGHTS_IN_SMEM ? kSmemLightBytes / 4 : 4];  // records first (16-byte aligned), then bins
    if (MATS_IN_SMEM) {
        const uint32_t words = w.nmaterials * (uint32_t)(sizeof(RptMaterialData) / 4);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(sm_materials)[i] = reinterpret_cast<const uint32_t*>(w.materials)[i];
    }
    if (LIGHTS_IN_SMEM) {
        const uint32_t rec_words = w.nlights * (uint32_t)(sizeof(LightRecord) / 4), bin_words = w.nbins * (uint32_t)(sizeof(LightBin) / 4);
        for (uint32_t i = threadIdx.x; i < rec_words; i += blockDim.x) sm_lights[i] = reinterpret_cast<const uint32_t*>(w.lights)[i];
        for (uint32_t i = threadIdx.x; i < bin_words; i += blockDim.x) sm_lights[rec_words + i] = reinterpret_cast<const uint32_t*>(w.light_bins)[i];
    }
    if (MATS_IN_SMEM || LIGHTS_IN_SMEM) __syncthreads();
    const LightRecord* const lights = LIGHTS_IN_SMEM ? reinterpret_cast<const LightRecord*>(sm_lights) : w.lights;
    const LightBin* const light_bins = LIGHTS_IN_SMEM ? reinterpret_cast<const LightBin*>(sm_lights + w.nlights * (uint32_t)(sizeof(LightRecord) / 4)) : w.light_bins;
    const uint32_t n = s.ctl->n_hit;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.counters + 3, (unsigned long long)n);  // surface hits shaded
    const bool nee = f.nee != RPT_NEE_NONE;
    const bool last_bounce = bounce + 1u >= f.max_bounces;
    // The first two levels of the dependent load chain (q_hit[i] -> slot -> hit[slot]) are software-pipelined two
    // iterations ahead, so an iteration starts with its slot and hit record already in registers and goes
    // straight to the path-state and triangle fetches.
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot_next = i < n ? __ldg(s.q_hit + i) : 0u;
    uint32_t slot_after = i + stride < n ? __ldg(s.q_hit + i + stride) : 0u;
    uint2 hit_next = i < n ? s.hit[slot_next] : make_uint2(0u, 0u);
    for (; i < n; i += stride) {
        bool want_shadow = false, want_next = false;
        const uint32_t slot = slot_next;
        const uint2 hr = hit_next;
        slot_next = slot_after;
        if (i + stride < n) hit_next = s.hit[slot_next];
        if (i + 2u * stride < n) slot_after = __ldg(s.q_hit + i + 2u * stride);
        const float t = __uint_as_float(hr.x);
        const uint32_t tri = hr.y & 0x7FFFFFFFu;
        const bool backface = (hr.y >> 31) != 0u;
        const float4 o4 = s.ray_o[slot], d4 = s.ray_d[slot], thr4 = s.thr[slot];
        const f3 ro = xyz(o4), rd = xyz(d4);
        const f3 throughput = xyz(thr4);
        const uint32_t flags = __float_as_uint(d4.w);  // rng dimension | last lobe << 8 | light record of the last NEE sample << 9
        const uint32_t last_lobe = (flags >> 8) & 1u;
        const uint2 seed = __ldg(rng + wave_pixel(d, slot % d.npix));
        Rng rstate{seed.x + slot / d.npix + seed.y, flags & 0xFFu};

        // the hit triangle's shading record: one aligned run of sectors (device_scene.h)
        const float4* rec = w.tri_shade + (size_t)(TANGENTS ? kShadeStrideTangents : kShadeStridePlain) * tri;
        const float4 a4 = __ldg(rec), e14 = __ldg(rec + 1), e24 = __ldg(rec + 2);
        const RptMaterialData& mat = MATS_IN_SMEM ? sm_materials[__float_as_uint(a4.w)] : w.materials[__float_as_uint(a4.w)];
        const f3 emissive = mk3(mat.emissive[0], mat.emissive[1], mat.emissive[2]);
        const f3 hit = ro + rd * t;
        bool alive = true;

        if (!zero3(emissive)) {  // lib.rs:86-109
            if (backface) {
                alive = false;
            } else if (!nee || bounce == 0u || last_lobe != kLobeDiffuse) {
                const f3 c = mask_nan(throughput * emissive);
                float4 r = s.rad[slot];
                r.x += c.x; r.y += c.y; r.z += c.z;
                s.rad[slot] = r;
                alive = false;
            } else if (f.nee == RPT_NEE_MIS) {  // calculate_bsdf_mis_contribution, light_pick.rs:179-199
                // (with the "no lights" sentinel table the reference's default DirectLightSample has area 0, i.e.
                // light pdf 0: the term is 0 and the path ends — there is no light record to read)
                // The reference multiplies the throughput BEFORE the last bounce by that bounce's spectrum / pdf; that
                // product is the current throughput times the Russian-roulette probability it was divided by
                // (o4.w; 1 when no roulette ran), so no per-path copy of either factor is kept.
                f3 c = splat3(0.0f);
                const LightRecord& L = lights[w.nbins > 0u ? flags >> 9 : 0u];
                if (w.nbins > 0u && tri == __float_as_uint(ldl<LIGHTS_IN_SMEM>(&L.e2_tri).w)) {
                    const float4 la = ldl<LIGHTS_IN_SMEM>(&L.a_area);
                    const float lp = light_pdf(la.w, t, xyz(ldl<LIGHTS_IN_SMEM>(&L.normal)), rd);
                    if (lp > 0.0f) {
                        const float wgt = power_heuristic(thr4.w, lp);
                        c = (throughput * o4.w) * ((xyz(ldl<LIGHTS_IN_SMEM>(&L.emission)) * wgt) / ldl<LIGHTS_IN_SMEM>(&L.e1_pdf).w);
                    }
                }
                c = mask_nan(c);
                float4 r = s.rad[slot];
                r.x += c.x; r.y += c.y; r.z += c.z;
                s.rad[slot] = r;
                alive = false;
            }
            // RPT_NEE_DIRECT after a diffuse bounce: fall through and shade the emitter as a surface
        }

        if (alive) {
            // lib.rs:111-129 — barycentrics re-derived from the hit point; normal not renormalised
            const float4 na = __ldg(rec + 3), nb = __ldg(rec + 4), nc = __ldg(rec + 5), r6 = __ldg(rec + 6);
            const f3 bary = barycentric(hit, xyz(a4), xyz(e14), xyz(e24));
            f3 normal = (bary.x * xyz(na) + bary.y * xyz(nb)) + bary.z * xyz(nc);
            f2 uv{(bary.x * e14.w + bary.y * na.w) + bary.z * nc.w, (bary.x * e24.w + bary.y * nb.w) + bary.z * r6.x};
            if (fminf(fmaxf(uv.x, 0.0f), 1.0f) != uv.x || fminf(fmaxf(uv.y, 0.0f), 1.0f) != uv.y) uv = f2{uv.x - floorf(uv.x), uv.y - floorf(uv.y)};
            if (TANGENTS && mat.has_normal_texture) {  // lib.rs:131-141
                const f3 nm = f.atlas.sample(mat.normals, uv) * 2.0f - splat3(1.0f);
                const float4 r7 = __ldg(rec + 7), r8 = __ldg(rec + 8);
                const f3 tangent = (bary.x * mk3(r6.y, r6.z, r6.w) + bary.y * xyz(r7)) + bary.z * mk3(r7.w, r8.x, r8.y);
                const f3 bitangent = cross(tangent, normal);
                normal = normalize((tangent * nm.x + bitangent * nm.y) + normal * nm.z);
            }

            const Pbr bsdf = make_pbr(mat, uv, f.atlas, f.clamp_lo, f.clamp_hi);
            const f3 view = -rd;
            const float r1 = rstate.next(), r2 = rstate.next(), r3 = rstate.next();
            const BsdfSample bs = pbr_sample(bsdf, view, normal, mk3(r1, r2, r3));

            uint32_t light_rec = 0;
            if (nee && bs.lobe == kLobeDiffuse && w.nbins > 0u) {  // sample_direct_lighting, light_pick.rs:100-173
                const float l1 = rstate.next(), l2 = rstate.next();
                uint32_t bin_i = (uint32_t)fminf(l1 * (float)w.nbins, 4294967040.0f);
                bin_i = min(bin_i, w.nbins - 1u);  // l1 == 1.0 would index one past the end (CPU path panics)
                const LightBin* bp = light_bins + bin_i;
                const LightBin bin{ldl<LIGHTS_IN_SMEM>(&bp->light_a), ldl<LIGHTS_IN_SMEM>(&bp->light_b), ldl<LIGHTS_IN_SMEM>(&bp->ratio)};
                light_rec = l2 < bin.ratio ? bin.light_a : bin.light_b;
                const LightRecord& L = lights[light_rec];
                const float4 la = ldl<LIGHTS_IN_SMEM>(&L.a_area), le1 = ldl<LIGHTS_IN_SMEM>(&L.e1_pdf), le2 = ldl<LIGHTS_IN_SMEM>(&L.e2_tri);
                const float q1 = rstate.next(), q2 = rstate.next();
                const float sq = sqrtf(q1);
                // (1-sq) a + sq(1-q2) b + sq q2 c  ==  a + sq(1-q2) e1 + sq q2 e2
                const f3 lp = xyz(la) + xyz(le1) * (sq * (1.0f - q2)) + xyz(le2) * (sq * q2);
                const f3 to_light = lp - hit;
                const float dist = length(to_light);
                const f3 l = to_light / dist;
                const float lpdf = light_pdf(la.w, dist, xyz(ldl<LIGHTS_IN_SMEM>(&L.normal)), l);
                if (lpdf > 0.0f) {
                    f3 fd;
                    float bpdf;
                    pbr_eval_diffuse(bsdf, view, normal, l, fd, bpdf);
                    if (bpdf > 0.0f) {
                        const float wgt = f.nee == RPT_NEE_MIS ? power_heuristic(lpdf, bpdf) : 1.0f;
                        const f3 direct = (fd * xyz(ldl<LIGHTS_IN_SMEM>(&L.emission)) * wgt / lpdf) / le1.w;
                        const f3 c = throughput * direct;
                        // a zero or non-finite contribution adds nothing whether or not the light is visible
                        if (finite3(c) && !zero3(c)) {
                            want_shadow = true;
                            s.sh_o[i] = mk4(hit + l * kEps, dist - kEps * 2.0f);
                            s.sh_d[i] = mk4(l, __uint_as_float(slot));
                            s.sh_c[i] = mk4(c, 0.0f);
                        }
                    }
                }
            }

            if (!last_bounce) {
                f3 next_thr = throughput * (bs.spectrum / bs.pdf);
                float roulette = 1.0f;
                want_next = true;
                if (bounce > f.min_bounces) {  // Russian roulette, lib.rs:175-181
                    roulette = max_element(next_thr);
                    if (rstate.next() > roulette) want_next = false;
                    next_thr = next_thr * (1.0f / roulette);
                } else if (f.retire_dead_paths && zero3(next_thr) && !rstate.draws_a_one()) {
                    // Throughput exactly zero before the roulette bounces (a specular sample under the horizon, a black
                    // texel): everything the rest of this path can add is 0 x (something finite) — the reference walks
                    // it to its first roulette, which ends it.  5-17 % of all path vertices on the scenes here.
                    // "Finite" has one exception that random numbers alone produce, hence the guard: a lobe selector of
                    // exactly 1.0 against a specular weight of exactly 1.0 picks the diffuse lobe with 1 / (1 - 1) = inf,
                    // 0 x inf = NaN, and the unmasked sky term carries the NaN into the accumulator (4 of the reference's
                    // 20 NaN pixels of the proxy at sample 4534).  Such a path is traced on, like the reference does.
                    want_next = false;
                }
                if (want_next) {
                    s.ray_o[slot] = mk4(hit + bs.direction * kEps, roulette);
                    s.ray_d[slot] = mk4(bs.direction, __uint_as_float((rstate.dim & 0xFFu) | (bs.lobe << 8) | (light_rec << 9)));
                    s.thr[slot] = mk4(next_thr, bs.pdf);
                }
            }
        }
        s.q_shaded[i] = slot | (want_next ? 0u : kShadedNoNext) | (want_shadow ? kShadedShadow : 0u);
    }
}

// Procedural sky (skybox.rs:46-94, lib.rs:66-69) for the compacted queue of escaped paths, at full SIMT efficiency.
// (The HDR lat-long variant of this stage lives in wavefront_miss.cu, built with IEEE division.)
__global__ void __launch_bounds__(128) wf_miss_procedural_kernel(FrameParams f, WaveState s) {
    const uint32_t n = s.ctl->n_miss;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = __ldg(s.q_miss + i);
        const f3 ro = xyz(s.ray_o[slot]), rd = xyz(s.ray_d[slot]), throughput = xyz(s.thr[slot]);
        const f3 c = throughput * sky::scatter(f.sun_dir, f.sun_intensity, ro, rd);  // NOT NaN-masked in the reference (lib.rs:69)
        float4 r = s.rad[slot];
        r.x += c.x; r.y += c.y; r.z += c.z;
        s.rad[slot] = r;
    }
}

__global__ void __launch_bounds__(256) wf_accumulate_kernel(WaveState s, WaveDesc d, uint2* __restrict__ rng, float4* __restrict__ output) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.counters, (unsigned long long)d.npix * d.k_samples);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < d.npix; j += gridDim.x * blockDim.x) {
        const uint32_t pixel = wave_pixel(d, j);
        float4 acc = output[pixel];
        for (uint32_t k = 0; k < d.k_samples; ++k) {  // sample order, like consecutive dispatches
            const float4 r = s.rad[(size_t)k * d.npix + j];
            acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += 1.0f;
        }
        output[pixel] = acc;
        uint2 seed = rng[pixel];
        seed.x += d.k_samples;
        rng[pixel] = seed;
    }
}

// Multi-GPU tile combine: the pixels a rank owns, packed in pixel-map order for the wire, and their scatter into the
// combined frame on the root (every pixel belongs to exactly one rank, so the frame is written exactly once).
__global__ void __launch_bounds__(256) pack_pixels_kernel(const float4* __restrict__ frame, const uint32_t* __restrict__ map, float4* __restrict__ packed, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) packed[i] = frame[__ldg(map + i)];
}
__global__ void __launch_bounds__(256) unpack_pixels_kernel(const float4* __restrict__ packed, const uint32_t* __restrict__ map, float4* __restrict__ frame, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) frame[__ldg(map + i)] = packed[i];
}

// Clears the per-bounce queue lengths (whole = true also clears the current extend queue).
__global__ void wf_reset_kernel(WaveState s, int next_queue, bool whole) {
    if (threadIdx.x == 0) {
        s.ctl->n_hit = 0; s.ctl->n_miss = 0; s.ctl->n_shadow = 0;
        s.ctl->fetch_extend = 0; s.ctl->fetch_shadow = 0;
        if (whole || next_queue == 0) s.ctl->n_ext[0] = 0;
        if (whole || next_queue == 1) s.ctl->n_ext[1] = 0;
    }
}

void launch_wf_shade(const WaveLaunch& l, const FrameParams& f, const WideWorld& w, const WaveState& s, const WaveDesc& d, const uint2* rng,
                     uint32_t bounce) {
    const bool mats = w.nmaterials <= kSmemMaterials;
    const bool lights = w.nbins > 0u && (size_t)w.nlights * sizeof(LightRecord) + (size_t)w.nbins * sizeof(LightBin) <= kSmemLightBytes;
    const bool tangents = w.shade_stride == kShadeStrideTangents;
    const dim3 grid(l.grid * kShadeMinBlocks);
#define RPT_SHADE(M, L, T) wf_shade_kernel<M, L, T><<<grid, kShadeBlock, 0, l.stream>>>(f, w, s, d, rng, bounce)
    if (mats && lights && tangents) RPT_SHADE(true, true, true);
    else if (mats && lights) RPT_SHADE(true, true, false);
    else if (mats && tangents) RPT_SHADE(true, false, true);
    else if (mats) RPT_SHADE(true, false, false);
    else if (tangents) RPT_SHADE(false, false, true);   // (more than 64 materials: everything from global memory)
    else RPT_SHADE(false, false, false);
#undef RPT_SHADE
}
void launch_pack_pixels(const float4* frame, const uint32_t* map, float4* packed, uint32_t n, int grid, cudaStream_t stream) {
    if (n) pack_pixels_kernel<<<grid * 8, 256, 0, stream>>>(frame, map, packed, n);
}
void launch_unpack_pixels(const float4* packed, const uint32_t* map, float4* frame, uint32_t n, int grid, cudaStream_t stream) {
    if (n) unpack_pixels_kernel<<<grid * 8, 256, 0, stream>>>(packed, map, frame, n);
}
void launch_wf_reset(const WaveLaunch& l, const WaveState& s, int next_queue, bool whole) { wf_reset_kernel<<<1, 32, 0, l.stream>>>(s, next_queue, whole); }
void launch_wf_miss_procedural(const WaveLaunch& l, const FrameParams& f, const WaveState& s) {
    wf_miss_procedural_kernel<<<l.grid * 8, 128, 0, l.stream>>>(f, s);
}
void launch_wf_accumulate(const WaveLaunch& l, const WaveState& s, const WaveDesc& d, uint2* rng, float4* output) {
    wf_accumulate_kernel<<<l.grid * 8, 256, 0, l.stream>>>(s, d, rng, output);
}

}  // namespace rpt
