// wavefront_trace_nofma.cu — the ray-casting stages of the wavefront pipeline:
//   generate        camera rays for every slot of a wave          (kernels/src/lib.rs:36-51)
//   extend          nearest-hit traversal + hit/miss compaction   (intersection.rs:165-167 intersect_nearest)
//   shadow-connect  any-hit traversal of the NEE shadow rays      (intersection.rs:169-171 intersect_any)
//
// Built with -fmad=false: ray generation and the ray/triangle test evaluate the reference's fp32
// operations in its order (dev/exact.cuh), which is what makes primary-hit ids bit-exact.
//
// All three are persistent kernels: the grid is a fixed multiple of the SM count and warps pull work
// from the queue, whose length lives in device memory (WaveCtl) so the host never has to read it
// back.  Each lane keeps its traversal stack in a per-warp shared-memory slab (stack[depth][lane]:
// conflict-free 8-byte accesses); rays are handed out with one atomicAdd per warp (ballot + popc),
// and the output queues are built afterwards by the order-preserving compaction kernels below.
#include "device_scene.h"
#include "wide_bvh.h"

namespace rpt {

// Build switches of the trace kernels; every one was measured on a B200 (profiles/r2_variant_sweeps.txt):
//   RPT_TRACE_LATE_WRITE   ON   a finished ray's result is written when the warp next reconverges, by all lanes that ended
//                               since then together, instead of by one lane alone (shadow-connect -5 %, extend -1 %)
//   RPT_TRACE_STEPWISE     off  every lane takes ONE step per round — a pending triangle or the next node — instead of "a
//                               node, then all of its triangles while the other lanes wait" (a wash / -10 % on DarkCornell)
//   RPT_TRI_FIRST_INLINE   off  the first triangle of a visit tested inside the node block (1: both ray kinds, 2: nearest
//                               only): extend +6 %
//   RPT_TRACE_MIN_BLOCKS   9    resident blocks per SM the register allocation aims at (8: -3 %, 10: -6 %)
// (dev/exact.cuh: RPT_TRI_STRAIGHT, ON; dev/wide_bvh.cuh: RPT_NODE_SHORT_PAD, ON, RPT_PLANES_FP32 / RPT_NODE_INDEXED_LOADS, off;
// wide_bvh.h: RPT_STACK_SHARED, 12.)
#ifndef RPT_TRACE_STEPWISE
#define RPT_TRACE_STEPWISE 0
#endif
#ifndef RPT_TRACE_LATE_WRITE
#define RPT_TRACE_LATE_WRITE 1
#endif
#ifndef RPT_TRI_FIRST_INLINE
#define RPT_TRI_FIRST_INLINE 0
#endif
#ifndef RPT_TRACE_MIN_BLOCKS
#define RPT_TRACE_MIN_BLOCKS 9  // resident blocks per SM the register allocation aims at: 9 x 128 threads x 56 registers
#endif
constexpr int kTraceBlock = 128;  // 4 warps; 16 KB of stack slabs + the 2 KB permutation table per block
constexpr int kTraceWarps = kTraceBlock / 32;
size_t trace_stack_overflow_entries(int grid_blocks) { return (size_t)grid_blocks * kTraceBlock * (kWideStackCapacity - kWideStackShared); }

// DEEP = false: the whole stack fits the shared slab (trees of at most kWideStackShared levels: every scene seen so
// far); DEEP = true adds a global-memory overflow area for entries kWideStackShared.. so that deeper trees trace
// correctly too (1 % slower, so the host only picks that variant when the tree needs it).
template <bool DEEP>
struct SmemStack {
    uint2* column;    // this lane's column of the warp slab; entries are 32 lanes apart
    int n;
    const uint8_t* perm;  // the block's octant permutation table, [octant][child set]
    uint2* overflow;  // this thread's column of the overflow area; entries are `overflow_stride` apart
    uint32_t overflow_stride;
    __device__ __forceinline__ uint32_t permute(uint32_t oct, uint32_t m) const { return perm[oct * 256u + m]; }
    __device__ __forceinline__ void push(uint2 v) {  // (the tree's depth is validated at upload)
        if (!DEEP || n < (int)kWideStackShared) column[n * 32] = v;
        else overflow[(size_t)(n - (int)kWideStackShared) * overflow_stride] = v;
        ++n;
    }
    __device__ __forceinline__ uint2 pop() {
        --n;
        return (!DEEP || n < (int)kWideStackShared) ? column[n * 32] : overflow[(size_t)(n - (int)kWideStackShared) * overflow_stride];
    }
    __device__ __forceinline__ bool empty() const { return n == 0; }
    __device__ __forceinline__ void clear() { n = 0; }
};

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

// Persistent trace kernel for both ray kinds.
//   NEAREST: items are path slots (queue q_ext[in_queue], or 0..n_identity-1 for bounce 0); the hit
//            (or miss) record goes to s.hit[slot]; wf_compact_kernel then splits the slots into
//            q_hit / q_miss IN QUEUE ORDER, so the shading stages read path state coalesced
//            although rays finish in any order here.
//   ANY:     items are the shadow rays listed in q_shadow; an unoccluded ray adds its contribution to rad[slot].
// Lanes pull rays one at a time from a device-side cursor: when a lane's ray terminates it waits
// only until the warp's live-lane count drops below `refill_below`, then every idle lane is handed
// a new ray (ray refill keeps the warp full although rays need very different numbers of steps).

constexpr uint32_t kMissRecord = 0xFFFFFFFFu;  // hit[].y of a ray that hit nothing (triangle indices are < 2^31)

// STATS = true is the diagnostic build (rpt_set_trace_statistics): it also counts the node visits and ray/triangle
// tests of the launch — the kernel's work in ITS OWN layout (80 B per node, 48 B per triangle record) — into
// counters[4..7]; the product launches STATS = false.
template <bool NEAREST, bool DEEP, bool STATS>
__global__ void __launch_bounds__(kTraceBlock, RPT_TRACE_MIN_BLOCKS) wf_trace_kernel(WideScene bvh, WaveState s, int in_queue, bool identity, uint32_t n_identity,
                                                               int refill_below, uint2* stack_overflow) {
    __shared__ uint2 slabs[kTraceWarps][kWideStackShared][32];
    __shared__ uint8_t perm_table[8 * 256];
    for (uint32_t i = threadIdx.x; i < 8u * 256u; i += kTraceBlock) perm_table[i] = (uint8_t)octant_permute(i >> 8, i & 0xFFu);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // (ternaries, not s.q_ext[in_queue]: dynamic indexing would force the parameter block into local memory)
    const uint32_t n = NEAREST ? (identity ? n_identity : (in_queue ? s.ctl->n_ext[1] : s.ctl->n_ext[0])) : s.ctl->n_shadow;
    const uint32_t* __restrict__ queue = in_queue ? s.q_ext[1] : s.q_ext[0];
    uint32_t* fetch = NEAREST ? &s.ctl->fetch_extend : &s.ctl->fetch_shadow;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.counters + (NEAREST ? 1 : 2), (unsigned long long)n);

    SmemStack<DEEP> st{&slabs[warp][0][lane], 0, perm_table, DEEP ? stack_overflow + (size_t)blockIdx.x * kTraceBlock + threadIdx.x : nullptr,
                       gridDim.x * kTraceBlock};
    WideCursor<NEAREST> c;
    uint32_t item = 0;       // path slot (NEAREST) / index of the shadow ray (ANY)
    bool busy = false;       // this lane holds an unfinished ray
    bool exhausted = false;  // the cursor ran past the end of the queue (warp-uniform)
    uint32_t stat_visits = 0, stat_tests = 0;
    // RPT_TRACE_LATE_WRITE: a ray's result is written when the warp next reconverges (every lane that ended since then
    // writes together) instead of by each lane alone at the moment its ray ends.
    bool unwritten = false;
    auto write_result = [&] {
        if (NEAREST) {  // every traced slot gets a record; kMissRecord marks "no hit" for wf_compact_kernel
            s.hit[item] = make_uint2(__float_as_uint(c.best_t), c.hit_tri);  // kMissRecord == kNoNode marks "no hit"
        } else if (c.hit_tri == kNoNode) {  // unoccluded: radiance += mask_nan(contribution) (lib.rs:164; masked when queued)
            const uint32_t slot = __float_as_uint(s.sh_d[item].w);
            const float4 add = s.sh_c[item];
            float4 r = s.rad[slot];
            r.x += add.x; r.y += add.y; r.z += add.z;
            s.rad[slot] = r;
        }
    };

    for (;;) {
#if RPT_TRACE_LATE_WRITE
        // ---- converged: the results of the rays that ended since the warp was last here ------------
        __syncwarp();
        if (unwritten) {
            unwritten = false;
            write_result();
        }
#endif
        // ---- converged: refill idle lanes -------------------------------------------------------
        if (!exhausted) {
            const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !busy);
            if (idle) {
                const int leader = __ffs((int)idle) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(fetch, (uint32_t)__popc(idle));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                exhausted = base + (uint32_t)__popc(idle) >= n;
                if (!busy && i < n) {
                    float4 o, dv;
                    float max_t = 0.0f;
                    if (NEAREST) {
                        item = identity ? i : __ldg(queue + i);
                        o = s.ray_o[item];
                        dv = s.ray_d[item];
                    } else {
                        item = __ldg(s.q_shadow + i);
                        o = s.sh_o[item];
                        dv = s.sh_d[item];
                        max_t = o.w;
                    }
                    c.begin(xyz(o), xyz(dv), max_t);
                    st.n = 0;
                    busy = true;
                }
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, busy) == 0u) {
            if (exhausted) break;
            continue;
        }

        // ---- traverse until the ray ends or the warp wants a refill ---------------------------
        while (busy) {
            bool finished = false;
#if RPT_TRACE_STEPWISE
            // one step per lane per round: a pending triangle if there is one, else the next node — the same per-ray
            // sequence of visits and tests, so the same result; no lane waits for another lane's triangles
            if (c.has_triangles()) {
                if (STATS) ++stat_tests;
                if (c.test_triangle(bvh)) finished = true;
            } else if (c.has_nodes()) {
                c.visit_node(bvh, st);
                if (STATS) ++stat_visits;
            }
            if (!c.has_nodes() && !c.has_triangles()) finished = true;
#else
            if (c.has_nodes()) {  // (only a non-finite ray starts without nodes)
                // RPT_TRI_FIRST_INLINE (1: both ray kinds, 2: nearest-hit rays only)
                constexpr bool kInlineFirst = RPT_TRI_FIRST_INLINE == 1 || (RPT_TRI_FIRST_INLINE == 2 && NEAREST);
                if constexpr (kInlineFirst) {
                    // The visit's first triangle is tested as part of the visit, by every lane of the node block (a lane
                    // whose visit found none tests record 0 and discards the result): with ~26 lanes visiting, some lane
                    // has one in practically every round, so the triangle block ran once per round anyway — this way its
                    // loads are issued before the next-node selection instead of after it, and the first test needs no branch.
                    const typename WideCursor<NEAREST>::Visit v = c.test_children(bvh);
                    bool any;
                    const uint32_t ti = c.take_first_triangle(any);
                    const float4* rec = bvh.tri_pos + 3u * (size_t)ti;
                    const float4 ta = __ldg(rec), te1 = __ldg(rec + 1), te2 = __ldg(rec + 2);
                    c.select_next(bvh, st, v);
                    if (STATS) { ++stat_visits; stat_tests += any ? 1u : 0u; }
                    if (c.template test_record<false>(ti, ta, te1, te2, any)) finished = true;
                } else {
                    c.visit_node(bvh, st);
                    if (STATS) ++stat_visits;
                }
            }
            while (!finished && c.has_triangles()) {
                if (STATS) ++stat_tests;
                if (c.test_triangle(bvh)) { finished = true; break; }
            }
            if (!c.has_nodes()) finished = true;
#endif
            if (finished) {
                busy = false;
#if RPT_TRACE_LATE_WRITE
                unwritten = true;
#else
                write_result();
#endif
                break;
            }
            if (!exhausted && __popc(__activemask()) < refill_below) break;
        }
    }
    if (STATS) {
        for (int d = 16; d > 0; d >>= 1) {
            stat_visits += __shfl_xor_sync(0xFFFFFFFFu, stat_visits, d);
            stat_tests += __shfl_xor_sync(0xFFFFFFFFu, stat_tests, d);
        }
        if (lane == 0u) {
            atomicAdd(s.counters + (NEAREST ? 4 : 6), (unsigned long long)stat_visits);
            atomicAdd(s.counters + (NEAREST ? 5 : 7), (unsigned long long)stat_tests);
        }
    }
}

// ---- extend / shadow-connect with DEFERRED triangle tests -------------------------------------------------------
// In the kernel above a lane tests the triangles of a visit right after the visit, so the ray/triangle code runs for
// the two or three lanes of a warp that happen to have found a leaf in that round: it is a third of the kernel's
// instructions at a tenth of its width.  Here a visit only QUEUES its triangles (wide indices, a per-lane stack of
// kPendingCapacity words in shared memory) and the lane goes on traversing; the warp, which stays converged round by
// round, runs triangle rounds when at least `flush_at` lanes hold a queued triangle — or when no lane has a node
// left to visit — and keeps running them while at least `flush_keep` lanes still do; what is left stays queued for
// the next flush.  A ray whose traversal is over keeps its slot until its queue is empty.  Modelled on the CPU first
// (tools/warp_sim.py partial, harness_warp_sim2): triangle rounds 2.1 -> 0.65 per ray at 11 instead of 3 lanes, for
// 2 % more node visits (the culling bound is refreshed later) and node rounds 5 % narrower (waiting lanes).
// Results are the in-order kernel's, except that among hits with bit-equal t the smallest wide triangle index wins
// (test_one<true>): the order in which a ray's candidates are tested depends on the other lanes here.
constexpr int kPendingCapacity = 4;  // triangle groups (one per visit that found any) a lane can hold besides the cursor's own

template <bool NEAREST, bool DEEP>
__global__ void __launch_bounds__(kTraceBlock, 9) wf_trace_deferred_kernel(WideScene bvh, WaveState s, int in_queue, bool identity, uint32_t n_identity,
                                                                        int refill_below, int flush_at, int flush_keep, uint2* stack_overflow) {
    __shared__ uint2 slabs[kTraceWarps][kWideStackShared][32];
    // queued groups, as visit_node leaves them in the cursor: first triangle of the node, hit mask, valid mask
    __shared__ uint32_t pending[kTraceWarps][kPendingCapacity][3][32];
    __shared__ uint8_t perm_table[8 * 256];
    for (uint32_t i = threadIdx.x; i < 8u * 256u; i += kTraceBlock) perm_table[i] = (uint8_t)octant_permute(i >> 8, i & 0xFFu);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n = NEAREST ? (identity ? n_identity : (in_queue ? s.ctl->n_ext[1] : s.ctl->n_ext[0])) : s.ctl->n_shadow;
    const uint32_t* __restrict__ queue = in_queue ? s.q_ext[1] : s.q_ext[0];
    uint32_t* fetch = NEAREST ? &s.ctl->fetch_extend : &s.ctl->fetch_shadow;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(s.counters + (NEAREST ? 1 : 2), (unsigned long long)n);

    SmemStack<DEEP> st{&slabs[warp][0][lane], 0, perm_table, DEEP ? stack_overflow + (size_t)blockIdx.x * kTraceBlock + threadIdx.x : nullptr,
                       gridDim.x * kTraceBlock};
    uint32_t* const pend = &pending[warp][0][0][lane];  // group g, word w at pend[(3 g + w) * 32]
    int qn = 0;                                          // queued groups (the cursor's own group comes on top of these)
    WideCursor<NEAREST> c;
    uint32_t item = 0;
    bool busy = false;
    bool exhausted = false;
    constexpr uint32_t kAll = 0xFFFFFFFFu;

    for (;;) {
        // ---- refill idle lanes (as in wf_trace_kernel)
        if (!exhausted) {
            const uint32_t idle = __ballot_sync(kAll, !busy);
            if (idle) {
                const int leader = __ffs((int)idle) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(fetch, (uint32_t)__popc(idle));
                base = __shfl_sync(kAll, base, leader);
                const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                exhausted = base + (uint32_t)__popc(idle) >= n;
                if (!busy && i < n) {
                    float4 o, dv;
                    float max_t = 0.0f;
                    if (NEAREST) {
                        item = identity ? i : __ldg(queue + i);
                        o = s.ray_o[item];
                        dv = s.ray_d[item];
                    } else {
                        item = __ldg(s.q_shadow + i);
                        o = s.sh_o[item];
                        dv = s.sh_d[item];
                        max_t = o.w;
                    }
                    c.begin(xyz(o), xyz(dv), max_t);
                    st.n = 0;
                    qn = 0;
                    busy = true;
                }
            }
        }
        uint32_t live = __ballot_sync(kAll, busy);
        if (live == 0u) {
            if (exhausted) break;
            continue;
        }
        // ---- converged rounds until the warp wants a refill
        for (;;) {
            // node round: a lane walks on if it has a node and room to keep what the visit may find
            const bool walking = busy && c.has_nodes() && (c.tgroup.y == 0u || qn < kPendingCapacity);
            if (walking) {
                if (c.tgroup.y != 0u) {  // the group of an earlier visit moves from the cursor to the queue
                    pend[(3 * qn + 0) * 32] = c.tgroup.x;
                    pend[(3 * qn + 1) * 32] = c.tgroup.y;
                    pend[(3 * qn + 2) * 32] = c.tvalid;
                    ++qn;
                }
                c.visit_node(bvh, st);
            }
            const bool has_work = busy && (c.tgroup.y != 0u || qn > 0);
            uint32_t holding = __ballot_sync(kAll, has_work);
            // flush when enough lanes hold triangles, or when nobody walked this round (every live lane waits for its triangles)
            if (holding != 0u && (__popc(holding) >= flush_at || __ballot_sync(kAll, walking) == 0u)) {
                do {
                    if (busy) {
                        if (c.tgroup.y == 0u && qn > 0) {
                            --qn;
                            c.tgroup.x = pend[(3 * qn + 0) * 32];
                            c.tgroup.y = pend[(3 * qn + 1) * 32];
                            c.tvalid = pend[(3 * qn + 2) * 32];
                        }
                        if (c.tgroup.y != 0u && c.template test_next<true>(bvh)) {  // true: an any-hit ray is decided
                            c.abandon(st);
                            qn = 0;
                        }
                    }
                    holding = __ballot_sync(kAll, busy && (c.tgroup.y != 0u || qn > 0));
                } while (__popc(holding) >= flush_keep);
            }
            if (busy && !c.has_nodes() && c.tgroup.y == 0u && qn == 0) {
                busy = false;
                if (NEAREST) {
                    s.hit[item] = make_uint2(__float_as_uint(c.best_t), c.hit_tri);
                } else if (c.hit_tri == kNoNode) {
                    const uint32_t slot = __float_as_uint(s.sh_d[item].w);
                    const float4 add = s.sh_c[item];
                    float4 r = s.rad[slot];
                    r.x += add.x; r.y += add.y; r.z += add.z;
                    s.rad[slot] = r;
                }
            }
            live = __ballot_sync(kAll, busy);
            if (live == 0u || (!exhausted && __popc(live) < refill_below)) break;
        }
    }
}

// ---- order-preserving two-way stream compaction ------------------------------------------------
// Both compaction kernels split one input stream into two output queues IN INPUT ORDER (the shading stages
// then read path state coalesced although rays finish in any order).  A block handles kCompactItems
// consecutive items per thread; the per-thread counts of both outputs ride in one word through a warp
// shuffle scan and a warp-total scan in shared memory, and the block reserves its two output ranges with ONE
// atomic pair (a single device counter only sustains a few atomics per nanosecond).
constexpr int kCompactBlock = 256;
constexpr int kCompactItems = 4;
constexpr uint32_t kCompactTile = kCompactBlock * kCompactItems;

// `mine` = (count for queue B) << 16 | (count for queue A) of this thread; returns this thread's first output
// index in both queues, packed the same way relative to the block's reservations base_a / base_b.
__device__ __forceinline__ void compact_reserve(uint32_t mine, uint32_t* count_a, uint32_t* count_b, uint32_t& out_a, uint32_t& out_b) {
    __shared__ uint32_t warp_total[kCompactBlock / 32], block_base[2];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += up;
    }
    if (lane == 31u) warp_total[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t wt = lane < kCompactBlock / 32 ? warp_total[lane] : 0u;
        uint32_t wincl = wt;
#pragma unroll
        for (int d = 1; d < kCompactBlock / 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, wincl, d);
            if ((int)lane >= d) wincl += up;
        }
        if (lane < kCompactBlock / 32) warp_total[lane] = wincl - wt;  // exclusive
        if (lane == kCompactBlock / 32 - 1) {
            const uint32_t ta = wincl & 0xFFFFu, tb = wincl >> 16;
            block_base[0] = ta ? atomicAdd(count_a, ta) : 0u;
            block_base[1] = tb ? atomicAdd(count_b, tb) : 0u;
        }
    }
    __syncthreads();
    const uint32_t excl = warp_total[warp] + incl - mine;
    out_a = block_base[0] + (excl & 0xFFFFu);
    out_b = block_base[1] + (excl >> 16);
    __syncthreads();  // the shared words are rewritten by the next tile
}

// After extend: the traced slots go to q_hit or q_miss; counts the rays actually traced.
__global__ void __launch_bounds__(kCompactBlock) wf_compact_kernel(WaveState s, int in_queue, bool identity, uint32_t n_identity) {
    const uint32_t n = identity ? n_identity : (in_queue ? s.ctl->n_ext[1] : s.ctl->n_ext[0]);
    const uint32_t* __restrict__ queue = in_queue ? s.q_ext[1] : s.q_ext[0];
    for (uint32_t tile = blockIdx.x * kCompactTile; tile < n; tile += gridDim.x * kCompactTile) {  // block-uniform trip count
        const uint32_t first = tile + threadIdx.x * kCompactItems;
        uint32_t slot[kCompactItems];
        uint32_t hit_bits = 0, valid_bits = 0;
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            const uint32_t i = first + k;
            slot[k] = 0;
            if (i < n) {
                slot[k] = identity ? i : __ldg(queue + i);
                valid_bits |= 1u << k;
                if (s.hit[slot[k]].y != kMissRecord) hit_bits |= 1u << k;
            }
        }
        const uint32_t miss_bits = valid_bits & ~hit_bits;
        uint32_t out_hit, out_miss;
        compact_reserve((uint32_t)__popc(hit_bits) | ((uint32_t)__popc(miss_bits) << 16), &s.ctl->n_hit, &s.ctl->n_miss, out_hit, out_miss);
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            if ((hit_bits >> k) & 1u) s.q_hit[out_hit++] = slot[k];
            if ((miss_bits >> k) & 1u) s.q_miss[out_miss++] = slot[k];
        }
    }
}

// After shade: q_shaded[i] = slot | flags for every hit i.  Paths that go on are queued for the next extend
// pass (their slot), hits that sampled a light are queued for shadow-connect (the index i of their shadow ray).
__global__ void __launch_bounds__(kCompactBlock) wf_compact_shaded_kernel(WaveState s, int out_queue) {
    const uint32_t n = s.ctl->n_hit;
    uint32_t* __restrict__ q_next = out_queue ? s.q_ext[1] : s.q_ext[0];
    uint32_t* count_next = out_queue ? &s.ctl->n_ext[1] : &s.ctl->n_ext[0];
    for (uint32_t tile = blockIdx.x * kCompactTile; tile < n; tile += gridDim.x * kCompactTile) {
        const uint32_t first = tile + threadIdx.x * kCompactItems;
        uint32_t word[kCompactItems];
        uint32_t next_bits = 0, shadow_bits = 0;
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            const uint32_t i = first + k;
            word[k] = i < n ? s.q_shaded[i] : kShadedNoNext;
            if (!(word[k] & kShadedNoNext)) next_bits |= 1u << k;
            if (word[k] & kShadedShadow) shadow_bits |= 1u << k;
        }
        uint32_t out_next, out_shadow;
        compact_reserve((uint32_t)__popc(next_bits) | ((uint32_t)__popc(shadow_bits) << 16), count_next, &s.ctl->n_shadow, out_next, out_shadow);
#pragma unroll
        for (int k = 0; k < kCompactItems; ++k) {
            if ((next_bits >> k) & 1u) q_next[out_next++] = word[k] & kShadedSlotMask;
            if ((shadow_bits >> k) & 1u) s.q_shadow[out_shadow++] = first + k;
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
    wf_generate_kernel<<<l.grid * 8, 256, 0, l.stream>>>(f, s, d, rng);
}
void launch_wf_extend(const WaveLaunch& l, const WideScene& bvh, const WaveState& s, int in_queue, bool identity_queue, uint32_t n_identity) {
    if (l.defer_extend && l.stack_overflow)
        wf_trace_deferred_kernel<true, true><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, in_queue, identity_queue, n_identity, l.refill_below, l.flush_at, l.flush_keep, l.stack_overflow);
    else if (l.defer_extend)
        wf_trace_deferred_kernel<true, false><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, in_queue, identity_queue, n_identity, l.refill_below, l.flush_at, l.flush_keep, nullptr);
    else if (l.trace_statistics && l.stack_overflow)
        wf_trace_kernel<true, true, true><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, in_queue, identity_queue, n_identity, l.refill_below, l.stack_overflow);
    else if (l.trace_statistics)
        wf_trace_kernel<true, false, true><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, in_queue, identity_queue, n_identity, l.refill_below, nullptr);
    else if (l.stack_overflow)
        wf_trace_kernel<true, true, false><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, in_queue, identity_queue, n_identity, l.refill_below, l.stack_overflow);
    else
        wf_trace_kernel<true, false, false><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, in_queue, identity_queue, n_identity, l.refill_below, nullptr);
    wf_compact_kernel<<<l.grid * 8, kCompactBlock, 0, l.stream>>>(s, in_queue, identity_queue, n_identity);
}
void launch_wf_compact_shaded(const WaveLaunch& l, const WaveState& s, int out_queue) {
    wf_compact_shaded_kernel<<<l.grid * 8, kCompactBlock, 0, l.stream>>>(s, out_queue);
}
void launch_wf_shadow(const WaveLaunch& l, const WideScene& bvh, const WaveState& s) {
    if (l.defer_shadow && l.stack_overflow)
        wf_trace_deferred_kernel<false, true><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, 0, false, 0u, l.refill_below, l.flush_at, l.flush_keep, l.stack_overflow);
    else if (l.defer_shadow)
        wf_trace_deferred_kernel<false, false><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, 0, false, 0u, l.refill_below, l.flush_at, l.flush_keep, nullptr);
    else if (l.trace_statistics && l.stack_overflow)
        wf_trace_kernel<false, true, true><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, 0, false, 0u, l.refill_below, l.stack_overflow);
    else if (l.trace_statistics)
        wf_trace_kernel<false, false, true><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, 0, false, 0u, l.refill_below, nullptr);
    else if (l.stack_overflow)
        wf_trace_kernel<false, true, false><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, 0, false, 0u, l.refill_below, l.stack_overflow);
    else
        wf_trace_kernel<false, false, false><<<l.grid * l.trace_blocks_per_sm, kTraceBlock, 0, l.stream>>>(bvh, s, 0, false, 0u, l.refill_below, nullptr);
}
void launch_wf_export_primary(const WaveLaunch& l, const WideScene& bvh, const WaveState& s, const WaveDesc& d, uint32_t* ids) {
    wf_export_primary_kernel<<<(d.npix + 255) / 256, 256, 0, l.stream>>>(bvh, s, d, ids);
    wf_export_primary_hits_kernel<<<l.grid * 8, 256, 0, l.stream>>>(bvh, s, d, ids);
}

}  // namespace rpt
