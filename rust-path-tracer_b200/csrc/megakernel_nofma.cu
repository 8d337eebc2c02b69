// megakernel_nofma.cu — the comparison arm: the reference kernel's shape on sm_100a.
//
// One thread per pixel runs the whole path of kernels/src/lib.rs:21-186 (camera ray, up to
// max_bounces ordered binary-BVH traversals, shading, NEE shadow ray, sky) against the
// reference's own buffer layouts, exactly like `trace_kernel` (lib.rs:189-227) — except that
// it loops over `n_samples` sample indices per launch and keeps the running sum in registers,
// so the accumulator is read and written once per launch instead of once per sample.
// Compiled with -fmad=false: traversal order, box tests and triangle tests are the
// reference's, so primary-hit ids match the CPU path bit for bit (ties included).
#include "dev/bvh2.cuh"
#include "dev/shading.cuh"
#include "device_scene.h"

namespace rpt {

struct MisState {  // what calculate_bsdf_mis_contribution reads from the previous bounce
    f3 spectrum, throughput, light_normal, emission;
    float pdf, area, pick_pdf;
    uint32_t light_triangle;
};

__device__ f3 trace_path(const MegaParams& p, uint32_t px, uint32_t py, uint32_t key, unsigned long long* n_nearest,
                         unsigned long long* n_any, uint32_t* primary_id) {
    const Bvh2Scene bvh{p.nodes, p.triangles, p.vertices};
    const Atlas atlas{p.atlas, p.atlas_w, p.atlas_h, pow2_mask(p.atlas_w), pow2_mask(p.atlas_h)};
    const uint32_t nee_mode = p.nee;
    const bool nee = nee_mode != RPT_NEE_NONE;

    Rng rng{key, 0u};
    f3 ro, rd;
    camera_ray(p.camera, px, py, rng, ro, rd);

    f3 throughput = splat3(1.0f), radiance = splat3(0.0f);
    uint32_t last_lobe = kLobeDiffuse;  // BSDFSample::default()
    f3 last_dir = splat3(0.0f);
    MisState mis{};
    uint32_t nearest = 0, any = 0;

    for (uint32_t bounce = 0; bounce < p.max_bounces; ++bounce) {
        ++nearest;
        const Hit h = bvh2_intersect<true>(bvh, ro, rd, 0.0f);
        if (bounce == 0 && primary_id) *primary_id = h.hit ? h.triangle : 0xFFFFFFFFu;
        if (!h.hit) {
            if (!p.has_skybox) radiance = radiance + throughput * sky::scatter(p.sun_dir, p.sun_intensity, ro, rd);
            else radiance = radiance + throughput * p.sky.lookup(rd);
            break;
        }
        const f3 hit = ro + rd * h.t;
        const uint4 tri = __ldg(p.triangles + h.triangle);
        const RptMaterialData& mat = p.materials[tri.w];
        const f3 emissive = mk3(mat.emissive[0], mat.emissive[1], mat.emissive[2]);
        if (!zero3(emissive)) {  // lib.rs:86-109
            if (h.backface) break;
            if (!nee || bounce == 0 || last_lobe != kLobeDiffuse) {
                radiance = radiance + mask_nan(throughput * emissive);
                break;
            }
            if (nee_mode == RPT_NEE_MIS) {  // light_pick.rs:179-199
                f3 c = splat3(0.0f);
                if (h.triangle == mis.light_triangle) {
                    const float lp = light_pdf(mis.area, h.t, mis.light_normal, last_dir);
                    if (lp > 0.0f) {
                        const float w = power_heuristic(mis.pdf, lp);
                        c = mis.throughput * ((mis.spectrum * mis.emission * w / mis.pdf) / mis.pick_pdf);
                    }
                }
                radiance = radiance + mask_nan(c);
                break;
            }
        }

        // lib.rs:111-129
        const float4 va = __ldg(p.vertices + 4 * tri.x), vb = __ldg(p.vertices + 4 * tri.y), vc = __ldg(p.vertices + 4 * tri.z);
        const f3 a = xyz(va);
        const f3 bary = barycentric(hit, a, xyz(vb) - a, xyz(vc) - a);
        f3 normal = (bary.x * xyz(__ldg(p.vertices + 4 * tri.x + 1)) + bary.y * xyz(__ldg(p.vertices + 4 * tri.y + 1))) +
                    bary.z * xyz(__ldg(p.vertices + 4 * tri.z + 1));
        const float4 ta = __ldg(p.vertices + 4 * tri.x + 3), tb = __ldg(p.vertices + 4 * tri.y + 3), tc = __ldg(p.vertices + 4 * tri.z + 3);
        f2 uv{(bary.x * ta.x + bary.y * tb.x) + bary.z * tc.x, (bary.x * ta.y + bary.y * tb.y) + bary.z * tc.y};
        if (fminf(fmaxf(uv.x, 0.0f), 1.0f) != uv.x || fminf(fmaxf(uv.y, 0.0f), 1.0f) != uv.y) uv = f2{uv.x - floorf(uv.x), uv.y - floorf(uv.y)};
        if (mat.has_normal_texture) {  // lib.rs:131-141
            const f3 nm = atlas.sample(mat.normals, uv) * 2.0f - splat3(1.0f);
            const f3 tangent = (bary.x * xyz(__ldg(p.vertices + 4 * tri.x + 2)) + bary.y * xyz(__ldg(p.vertices + 4 * tri.y + 2))) +
                               bary.z * xyz(__ldg(p.vertices + 4 * tri.z + 2));
            const f3 bitangent = cross(tangent, normal);
            normal = normalize((tangent * nm.x + bitangent * nm.y) + normal * nm.z);
        }

        const Pbr bsdf = make_pbr(mat, uv, atlas, p.clamp_lo, p.clamp_hi);
        const f3 view = -rd;
        const float r1 = rng.next(), r2 = rng.next(), r3 = rng.next();
        const BsdfSample bs = pbr_sample(bsdf, view, normal, mk3(r1, r2, r3));
        last_lobe = bs.lobe;
        last_dir = bs.direction;

        if (nee && bs.lobe == kLobeDiffuse && !(p.lights[0].ratio < 0.0f)) {  // light_pick.rs:100-173
            const float l1 = rng.next(), l2 = rng.next();
            uint32_t slot = (uint32_t)fminf(l1 * (float)p.nlights, 4294967040.0f);
            slot = min(slot, p.nlights - 1u);  // l1 can be exactly 1.0: clamp where the CPU path would panic
            const RptLightPickEntry e = p.lights[slot];
            const bool first = l2 < e.ratio;
            const uint32_t li = first ? e.triangle_index_a : e.triangle_index_b;
            const float area = first ? e.triangle_area_a : e.triangle_area_b;
            const float pick_pdf = first ? e.triangle_pick_pdf_a : e.triangle_pick_pdf_b;
            const uint4 lt = __ldg(p.triangles + li);
            const f3 la = xyz(__ldg(p.vertices + 4 * lt.x)), lb = xyz(__ldg(p.vertices + 4 * lt.y)), lc = xyz(__ldg(p.vertices + 4 * lt.z));
            const f3 ln = ((xyz(__ldg(p.vertices + 4 * lt.x + 1)) + xyz(__ldg(p.vertices + 4 * lt.y + 1))) + xyz(__ldg(p.vertices + 4 * lt.z + 1))) / 3.0f;
            const RptMaterialData& lm = p.materials[lt.w];
            const f3 le = mk3(lm.emissive[0], lm.emissive[1], lm.emissive[2]);
            const float q1 = rng.next(), q2 = rng.next();
            const float sq = sqrtf(q1);
            const f3 lp = ((1.0f - sq) * la + (sq * (1.0f - q2)) * lb) + (sq * q2) * lc;
            const f3 to_light = lp - hit;
            const float dist = length(to_light);
            const f3 l = to_light / dist;
            f3 direct = splat3(0.0f);
            ++any;
            const Hit sh = bvh2_intersect<false>(bvh, hit + l * kEps, l, dist - kEps * 2.0f);
            if (!sh.hit) {
                const float lpdf = light_pdf(area, dist, ln, l);
                if (lpdf > 0.0f) {
                    f3 f;
                    float bpdf;
                    pbr_eval_diffuse(bsdf, view, normal, l, f, bpdf);
                    if (bpdf > 0.0f) {
                        const float w = nee_mode == RPT_NEE_MIS ? power_heuristic(lpdf, bpdf) : 1.0f;
                        direct = (f * le * w / lpdf) / pick_pdf;
                    }
                }
            }
            mis.area = area; mis.light_normal = ln; mis.pick_pdf = pick_pdf; mis.emission = le;
            mis.light_triangle = li; mis.throughput = throughput;
            radiance = radiance + mask_nan(throughput * direct);
        }
        mis.spectrum = bs.spectrum;
        mis.pdf = bs.pdf;

        throughput = throughput * (bs.spectrum / bs.pdf);
        rd = bs.direction;
        ro = hit + rd * kEps;

        if (bounce > p.min_bounces) {  // lib.rs:175-181
            const float prob = max_element(throughput);
            if (rng.next() > prob) break;
            throughput = throughput * (1.0f / prob);
        }
    }
    if (n_nearest) { atomicAdd(n_nearest, (unsigned long long)nearest); atomicAdd(n_any, (unsigned long long)any); }
    return radiance;
}

__global__ void __launch_bounds__(64) mega_trace_kernel(MegaParams p, uint32_t n_samples) {
    // 8x8 pixel tiles per 64-thread block (the reference's workgroup shape, lib.rs:189)
    const uint32_t tiles_x = (p.width + 7u) / 8u;
    const uint32_t tile = blockIdx.x, lane = threadIdx.x;
    const uint32_t px = (tile % tiles_x) * 8u + (lane & 7u), py = (tile / tiles_x) * 8u + (lane >> 3);
    if (px >= p.width || py >= p.height) return;
    const uint32_t i = py * p.width + px;
    if (p.tile_count > 1u) {
        const uint32_t t32 = (py / 32u) * ((p.width + 31u) / 32u) + (px / 32u);
        if (t32 % p.tile_count != p.tile_rank) return;
    }
    const uint2 seed = p.rng[i];
    float4 acc = p.output[i];
    unsigned long long *cn = nullptr, *ca = nullptr;
    if (p.counters) { cn = p.counters + 1; ca = p.counters + 2; }
    for (uint32_t s = 0; s < n_samples; ++s) {
        const f3 r = trace_path(p, px, py, seed.x + s + seed.y, cn, ca, nullptr);
        acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += 1.0f;
    }
    p.output[i] = acc;
    p.rng[i] = make_uint2(seed.x + n_samples, seed.y);
}

__global__ void __launch_bounds__(64) mega_primary_kernel(MegaParams p, uint32_t* ids) {
    const uint32_t tiles_x = (p.width + 7u) / 8u;
    const uint32_t px = (blockIdx.x % tiles_x) * 8u + (threadIdx.x & 7u), py = (blockIdx.x / tiles_x) * 8u + (threadIdx.x >> 3);
    if (px >= p.width || py >= p.height) return;
    const uint32_t i = py * p.width + px;
    const uint2 seed = p.rng[i];
    MegaParams q = p;
    q.max_bounces = 1;
    uint32_t id = 0xFFFFFFFFu;
    // one bounce, no shading side effects: the miss branch only adds sky radiance we discard
    const Bvh2Scene bvh{p.nodes, p.triangles, p.vertices};
    Rng rng{seed.x + seed.y, 0u};
    f3 ro, rd;
    camera_ray(p.camera, px, py, rng, ro, rd);
    const Hit h = bvh2_intersect<true>(bvh, ro, rd, 0.0f);
    if (h.hit) id = h.triangle;
    ids[i] = id;
}

void launch_mega_trace(const MegaParams& p, uint32_t n_samples, cudaStream_t stream) {
    const uint32_t tiles = ((p.width + 7u) / 8u) * ((p.height + 7u) / 8u);
    mega_trace_kernel<<<tiles, 64, 0, stream>>>(p, n_samples);
}
void launch_mega_primary(const MegaParams& p, uint32_t* ids, cudaStream_t stream) {
    const uint32_t tiles = ((p.width + 7u) / 8u) * ((p.height + 7u) / 8u);
    mega_primary_kernel<<<tiles, 64, 0, stream>>>(p, ids);
}

}  // namespace rpt
