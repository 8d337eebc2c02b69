// device_build.cu — the 8-wide BVH built and refitted ON THE DEVICE (build half).
//
// The reference builds its binary BVH on the host (src/bvh.rs:257-323, binned SAH) and the default path of this
// backend collapses that tree into the wide layout on the host too (wide_bvh_build.cpp), which keeps the reference's
// tree quality.  This file is the path for hosts that do not want to wait for either — scene edits, animated vertices,
// start-up — `rpt_upload_world(nodes = NULL)` and `rpt_refit_world`:
//
//   build   Morton codes of the triangle centroids (63 bits) -> radix sort (cub) -> the wide tree is cut straight out of
//           the sorted codes, level by level: a node owns a run of codes and cuts it into up to eight sub-runs by
//           repeated binary splits (largest sub-run first, at the highest bit in which it differs), so nodes come out
//           full; the axes those splits decide give each sub-run its octant slot, which is what the traversal's
//           `slot ^ octant` order expects (x the most significant bit of a Morton triple).  Sub-runs of one triangle (up to
//           three with RPT_DEVICE_BUILD_MAX_LEAF) become leaf slots, longer ones child nodes (contiguous, in slot order); a node's leaf triangles get one
//           contiguous range of the triangle stream, in slot order, as wide_bvh.h lays down.  Identical codes split in halves.
//   fit     bottom-up over the levels: slot boxes from the triangles' vertices / the child nodes' boxes, node box,
//           power-of-two cell, conservative 8-bit quantisation — the arithmetic of Collapser::encode.
//   emit    triangle position stream and shading records in the new leaf order.
//   refit   emit + fit again on a tree whose topology stays (vertices moved); works for host-collapsed trees as well:
//           the level lists come from a breadth-first walk of the node array.
//
// Nearest-hit results do not depend on the tree (only exact-t ties do), so ids and radiance stay within the same bars
// as with the reference's tree; a Morton tree costs more node visits per ray than the SAH tree (measured: DESIGN.md).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "device_build.h"

namespace rpt {
namespace {

constexpr int kThreads = 256;
inline int blocks_for(size_t n) { return (int)std::min<size_t>((n + kThreads - 1) / kThreads, 65535u * 8u); }

// ---- Morton codes ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {  // 21 bits -> every third bit
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__device__ __forceinline__ float3 vertex_of(const RptPerVertexData* v, uint32_t i) { return make_float3(v[i].vertex[0], v[i].vertex[1], v[i].vertex[2]); }
__device__ __forceinline__ float finite_or(float x, float fallback) { return (x == x && fabsf(x) < 3.0e38f) ? x : fallback; }

// scene bounds of the centroids, as ordered integers so that atomicMin / atomicMax work on floats
__device__ __forceinline__ int ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float unordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void centroid_bounds_kernel(const RptPerVertexData* __restrict__ verts, const uint4* __restrict__ tris, uint32_t ntris, int* __restrict__ bounds) {
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntris; t += gridDim.x * blockDim.x) {
        const uint4 tri = tris[t];
        const float3 a = vertex_of(verts, tri.x), b = vertex_of(verts, tri.y), c = vertex_of(verts, tri.z);
        const float cx = finite_or((a.x + b.x + c.x) * (1.0f / 3.0f), 0.0f), cy = finite_or((a.y + b.y + c.y) * (1.0f / 3.0f), 0.0f),
                    cz = finite_or((a.z + b.z + c.z) * (1.0f / 3.0f), 0.0f);
        lo[0] = fminf(lo[0], cx); lo[1] = fminf(lo[1], cy); lo[2] = fminf(lo[2], cz);
        hi[0] = fmaxf(hi[0], cx); hi[1] = fmaxf(hi[1], cy); hi[2] = fmaxf(hi[2], cz);
    }
    for (int k = 0; k < 3; ++k) {
        for (int d = 16; d > 0; d >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], d));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], d));
        }
        if ((threadIdx.x & 31u) == 0u) {
            atomicMin(bounds + k, ordered(lo[k]));
            atomicMax(bounds + 3 + k, ordered(hi[k]));
        }
    }
}

__global__ void morton_kernel(const RptPerVertexData* __restrict__ verts, const uint4* __restrict__ tris, uint32_t ntris, const int* __restrict__ bounds,
                              unsigned long long* __restrict__ codes, uint32_t* __restrict__ ids) {
    const float lo[3] = {unordered(bounds[0]), unordered(bounds[1]), unordered(bounds[2])};
    const float hi[3] = {unordered(bounds[3]), unordered(bounds[4]), unordered(bounds[5])};
    double scale[3];
    for (int k = 0; k < 3; ++k) scale[k] = hi[k] > lo[k] ? 2097151.0 / ((double)hi[k] - (double)lo[k]) : 0.0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntris; t += gridDim.x * blockDim.x) {
        const uint4 tri = tris[t];
        const float3 a = vertex_of(verts, tri.x), b = vertex_of(verts, tri.y), c = vertex_of(verts, tri.z);
        const float ce[3] = {finite_or((a.x + b.x + c.x) * (1.0f / 3.0f), 0.0f), finite_or((a.y + b.y + c.y) * (1.0f / 3.0f), 0.0f),
                             finite_or((a.z + b.z + c.z) * (1.0f / 3.0f), 0.0f)};
        unsigned long long q[3];
        for (int k = 0; k < 3; ++k) {
            const double g = ((double)ce[k] - (double)lo[k]) * scale[k];
            q[k] = (unsigned long long)fmin(fmax(g, 0.0), 2097151.0);
        }
        codes[t] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);  // x is the most significant bit of a triple
        ids[t] = t;
    }
}

// ---- topology, level by level ------------------------------------------------------------------------------------
struct Task {
    uint32_t lo, hi;  // run of the sorted codes
    uint32_t node;    // node this run becomes
};

struct BuildCounters {
    uint32_t nodes;      // nodes allocated so far
    uint32_t triangles;  // triangle-stream positions allocated so far
    uint32_t next_tasks; // tasks appended to the next level
    uint32_t pad;
};

// first index in [lo, hi) whose code has bit `bit` set (the run is sorted and agrees on every higher bit)
__device__ __forceinline__ uint32_t first_with_bit(const unsigned long long* codes, uint32_t lo, uint32_t hi, int bit) {
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((codes[mid] >> bit) & 1ull) hi = mid; else lo = mid + 1;
    }
    return lo;
}


// One node per thread.  The node's run of sorted codes is cut into up to eight sub-runs by repeated BINARY splits — always
// the largest remaining sub-run, at the highest bit in which it differs — so nodes come out full (a fixed split at one
// bit triple leaves most octree cells empty).  Every split decides one axis of the two halves' octant (bit % 3: x is the
// most significant bit of a Morton triple); a sub-run takes the slot of its octant, its undecided axes used to resolve
// collisions, so `slot ^ ray octant` still visits near children first.  Sub-runs of <= kMaxLeafTriangles become leaf slots.
__global__ void split_level_kernel(const unsigned long long* __restrict__ codes, const uint32_t* __restrict__ sorted_ids, const Task* __restrict__ tasks,
                                   uint32_t ntasks, Task* __restrict__ next, BuildCounters* __restrict__ ctr, uint32_t* __restrict__ node_words,
                                   uint32_t* __restrict__ orig_index, uint32_t kMaxLeafTriangles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntasks) return;
    const Task t = tasks[i];
    uint32_t lo[8], hi[8], known[8], side[8];
    uint32_t nruns = 1;
    lo[0] = t.lo; hi[0] = t.hi; known[0] = 0u; side[0] = 0u;
    while (nruns < 8u) {
        uint32_t best = 8u, best_count = 1u;
        for (uint32_t r = 0; r < nruns; ++r)
            if (hi[r] - lo[r] > best_count) { best = r; best_count = hi[r] - lo[r]; }
        if (best == 8u) break;  // nothing left with more than one triangle
        const unsigned long long x = codes[lo[best]] ^ codes[hi[best] - 1u];
        uint32_t mid, axis = 3u;
        if (x == 0ull) {
            mid = lo[best] + best_count / 2u;  // identical codes (coincident centroids): halves
        } else {
            const int bit = 63 - __clzll((long long)x);
            mid = first_with_bit(codes, lo[best], hi[best], bit);
            axis = (uint32_t)bit % 3u;
        }
        lo[nruns] = mid; hi[nruns] = hi[best]; known[nruns] = known[best]; side[nruns] = side[best];
        hi[best] = mid;
        if (axis < 3u && !((known[best] >> axis) & 1u)) {  // the first split along an axis places the halves on its two sides
            known[best] |= 1u << axis; known[nruns] |= 1u << axis;
            side[nruns] |= 1u << axis;
        }
        ++nruns;
    }
    // octant slots: the decided axes are binding, the undecided ones are free to dodge a collision
    int run_in_slot[8];
    for (int k = 0; k < 8; ++k) run_in_slot[k] = -1;
    for (uint32_t r = 0; r < nruns; ++r) {
        int slot = -1;
        for (uint32_t f = 0; f < 8u && slot < 0; ++f) {  // f: bits tried on the undecided axes
            if (f & known[r]) continue;
            const uint32_t cand = (side[r] & known[r]) | f;
            if (run_in_slot[cand] < 0) slot = (int)cand;
        }
        for (uint32_t cand = 0; cand < 8u && slot < 0; ++cand)
            if (run_in_slot[cand] < 0) slot = (int)cand;
        run_in_slot[slot] = (int)r;
    }
    uint32_t imask = 0, valid = 0, n_inner = 0, n_leaf_tris = 0;
    for (uint32_t s = 0; s < 8u; ++s) {
        if (run_in_slot[s] < 0) continue;
        const uint32_t cnt = hi[run_in_slot[s]] - lo[run_in_slot[s]];
        if (cnt <= kMaxLeafTriangles) { valid |= ((1u << cnt) - 1u) << (3u * s); n_leaf_tris += cnt; }
        else { imask |= 1u << s; ++n_inner; }
    }
    const uint32_t child_base = n_inner ? atomicAdd(&ctr->nodes, n_inner) : 0u;
    const uint32_t task_base = n_inner ? atomicAdd(&ctr->next_tasks, n_inner) : 0u;
    const uint32_t tri_base = n_leaf_tris ? atomicAdd(&ctr->triangles, n_leaf_tris) : 0u;
    uint32_t ci = 0, ti = 0;
    for (uint32_t s = 0; s < 8u; ++s) {
        if (run_in_slot[s] < 0) continue;
        const uint32_t first = lo[run_in_slot[s]], cnt = hi[run_in_slot[s]] - first;
        if (cnt <= kMaxLeafTriangles) {
            for (uint32_t k = 0; k < cnt; ++k) orig_index[tri_base + ti++] = sorted_ids[first + k];
        } else {
            next[task_base + ci] = Task{first, first + cnt, child_base + ci};
            ++ci;
        }
    }
    uint32_t* w = node_words + 20u * (size_t)t.node;
    w[4] = child_base; w[5] = tri_base; w[6] = valid | (imask << 24);
}

// breadth-first level lists of ANY wide node array (refit of host-collapsed trees): children of the nodes in `in`
__global__ void expand_level_kernel(const uint32_t* __restrict__ node_words, const uint32_t* __restrict__ in, uint32_t nin, uint32_t* __restrict__ out,
                                    uint32_t* __restrict__ nout) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nin) return;
    const uint32_t* w = node_words + 20u * (size_t)in[i];
    const uint32_t kids = (uint32_t)__popc(w[6] >> 24);
    if (!kids) return;
    const uint32_t base = atomicAdd(nout, kids);
    for (uint32_t k = 0; k < kids; ++k) out[base + k] = w[4] + k;
}

// ---- triangle streams in leaf order ------------------------------------------------------------------------------
__global__ void emit_triangles_kernel(const RptPerVertexData* __restrict__ verts, const uint4* __restrict__ tris, const uint32_t* __restrict__ orig_index,
                                      uint32_t ntris, float4* __restrict__ tri_pos, float4* __restrict__ tri_shade, uint32_t shade_stride,
                                      uint32_t* __restrict__ wide_index) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < ntris; p += gridDim.x * blockDim.x) {
        const uint32_t t = orig_index[p];
        const uint4 tri = tris[t];
        const RptPerVertexData &a = verts[tri.x], &b = verts[tri.y], &c = verts[tri.z];
        // e1 = b - a, e2 = c - a: single IEEE subtractions, as on the host (wide_bvh_build.cpp emit_triangle)
        const float4 r0 = make_float4(a.vertex[0], a.vertex[1], a.vertex[2], __uint_as_float(t));
        const float4 r1 = make_float4(__fsub_rn(b.vertex[0], a.vertex[0]), __fsub_rn(b.vertex[1], a.vertex[1]), __fsub_rn(b.vertex[2], a.vertex[2]), __uint_as_float(tri.w));
        const float4 r2 = make_float4(__fsub_rn(c.vertex[0], a.vertex[0]), __fsub_rn(c.vertex[1], a.vertex[1]), __fsub_rn(c.vertex[2], a.vertex[2]), 0.0f);
        float4* pos = tri_pos + 3u * (size_t)p;
        pos[0] = r0; pos[1] = r1; pos[2] = r2;
        float4* rec = tri_shade + (size_t)shade_stride * p;  // layout: device_scene.h
        rec[0] = make_float4(r0.x, r0.y, r0.z, r1.w);
        rec[1] = make_float4(r1.x, r1.y, r1.z, a.uv0[0]);
        rec[2] = make_float4(r2.x, r2.y, r2.z, a.uv0[1]);
        rec[3] = make_float4(a.normal[0], a.normal[1], a.normal[2], b.uv0[0]);
        rec[4] = make_float4(b.normal[0], b.normal[1], b.normal[2], b.uv0[1]);
        rec[5] = make_float4(c.normal[0], c.normal[1], c.normal[2], c.uv0[0]);
        if (shade_stride == kShadeStrideTangents) {
            rec[6] = make_float4(c.uv0[1], a.tangent[0], a.tangent[1], a.tangent[2]);
            rec[7] = make_float4(b.tangent[0], b.tangent[1], b.tangent[2], c.tangent[0]);
            rec[8] = make_float4(c.tangent[1], c.tangent[2], 0.0f, 0.0f);
            rec[9] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        } else {
            rec[6] = make_float4(c.uv0[1], 0.0f, 0.0f, 0.0f);
            rec[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        if (wide_index) wide_index[t] = p;
    }
}

// ---- boxes, bottom-up ----------------------------------------------------------------------------------------------
struct Box { float lo[3], hi[3]; };
__device__ __forceinline__ void grow(Box& b, float3 p) {  // fminf / fmaxf ignore NaN, like the host's min / max
    b.lo[0] = fminf(b.lo[0], p.x); b.lo[1] = fminf(b.lo[1], p.y); b.lo[2] = fminf(b.lo[2], p.z);
    b.hi[0] = fmaxf(b.hi[0], p.x); b.hi[1] = fmaxf(b.hi[1], p.y); b.hi[2] = fmaxf(b.hi[2], p.z);
}

__global__ void fit_level_kernel(const RptPerVertexData* __restrict__ verts, const uint4* __restrict__ tris, const uint32_t* __restrict__ orig_index,
                                 const uint32_t* __restrict__ level_nodes, uint32_t nlevel, uint32_t* __restrict__ node_words, Box* __restrict__ node_box) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlevel) return;
    const uint32_t id = level_nodes[i];
    uint32_t* w = node_words + 20u * (size_t)id;
    const uint32_t child_base = w[4], tri_base = w[5], valid = w[6] & 0x00FFFFFFu, imask = w[6] >> 24;
    Box slot[8];
    Box nb{{3e38f, 3e38f, 3e38f}, {-3e38f, -3e38f, -3e38f}};
    uint32_t used = 0;
    for (uint32_t s = 0; s < 8u; ++s) {
        Box b{{3e38f, 3e38f, 3e38f}, {-3e38f, -3e38f, -3e38f}};
        const uint32_t cnt = (uint32_t)__popc((valid >> (3u * s)) & 7u);
        if ((imask >> s) & 1u) {
            b = node_box[child_base + (uint32_t)__popc(imask & ((1u << s) - 1u))];
            used |= 1u << s;
        } else if (cnt) {
            const uint32_t first = tri_base + (uint32_t)__popc(valid & ((1u << (3u * s)) - 1u));
            for (uint32_t k = 0; k < cnt; ++k) {
                const uint4 tri = tris[orig_index[first + k]];
                grow(b, vertex_of(verts, tri.x)); grow(b, vertex_of(verts, tri.y)); grow(b, vertex_of(verts, tri.z));
            }
            used |= 1u << s;
        }
        slot[s] = b;
        if ((used >> s) & 1u)
            for (int k = 0; k < 3; ++k) { nb.lo[k] = fminf(nb.lo[k], b.lo[k]); nb.hi[k] = fmaxf(nb.hi[k], b.hi[k]); }
    }
    if (nb.lo[0] > nb.hi[0] || nb.lo[1] > nb.hi[1] || nb.lo[2] > nb.hi[2]) {  // nothing but NaN vertices below: an empty box at the origin
        for (int k = 0; k < 3; ++k) nb.lo[k] = nb.hi[k] = 0.0f;
    }
    node_box[id] = nb;
    // power-of-two cells and conservative 8-bit planes: Collapser::encode (wide_bvh_build.cpp), in the same double arithmetic
    uint32_t e[3];
    double cell[3];
    for (int k = 0; k < 3; ++k) {
        const double extent = (double)nb.hi[k] - (double)nb.lo[k];
        int ex = -126;
        if (extent > 0.0 && extent < 1e37) {
            ex = (int)ceil(log2(extent / 255.0));
            while (ceil(extent / ldexp(1.0, ex)) > 255.0) ++ex;
        } else if (extent > 0.0) {
            ex = 127;
        }
        ex = min(max(ex, -126), 127);
        e[k] = (uint32_t)(ex + 127);
        cell[k] = ldexp(1.0, ex);
    }
    uint32_t q[6][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (uint32_t s = 0; s < 8u; ++s) {
        if (!((used >> s) & 1u)) continue;
        for (int k = 0; k < 3; ++k) {
            // an empty slot box (all-NaN triangles) quantises to lo = 255 > hi = 0: never entered
            const double lo = floor(((double)slot[s].lo[k] - (double)nb.lo[k]) / cell[k]);
            const double hi = ceil(((double)slot[s].hi[k] - (double)nb.lo[k]) / cell[k]);
            const uint32_t qlo = (uint32_t)fmin(fmax(lo, 0.0), 255.0), qhi = (uint32_t)fmin(fmax(hi, 0.0), 255.0);
            q[k][s >> 2] |= qlo << (8u * (s & 3u));
            q[3 + k][s >> 2] |= qhi << (8u * (s & 3u));
        }
    }
    w[0] = __float_as_uint(nb.lo[0]); w[1] = __float_as_uint(nb.lo[1]); w[2] = __float_as_uint(nb.lo[2]);
    w[3] = e[0] << 23;
    w[7] = (e[1] << 7) | (e[2] << 23);
    w[8] = q[0][0]; w[9] = q[0][1]; w[10] = q[1][0]; w[11] = q[1][1];
    w[12] = q[2][0]; w[13] = q[2][1]; w[14] = q[3][0]; w[15] = q[3][1];
    w[16] = q[4][0]; w[17] = q[4][1]; w[18] = q[5][0]; w[19] = q[5][1];
}

template <class T>
cudaError_t dev_alloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T)); }

#define RPT_CK(expr)                         \
    do {                                     \
        const cudaError_t rpt_e_ = (expr);   \
        if (rpt_e_ != cudaSuccess) { status = rpt_e_; goto done; } \
    } while (0)

}  // namespace

// Level lists of a node array by breadth-first expansion: level_nodes holds the node ids level after level,
// level_offsets[L] .. level_offsets[L + 1] is level L.  Returns cudaSuccess or the failing call's error.
static cudaError_t list_levels(const uint32_t* node_words, uint32_t nnodes, uint32_t* level_nodes, uint32_t* d_counter, std::vector<uint32_t>& level_offsets,
                               cudaStream_t stream) {
    level_offsets.assign(1, 0u);
    const uint32_t zero = 0;
    cudaError_t e = cudaMemcpyAsync(level_nodes, &zero, 4, cudaMemcpyHostToDevice, stream);  // level 0 = the root
    if (e != cudaSuccess) return e;
    uint32_t begin = 0, count = 1;
    while (count) {
        level_offsets.push_back(begin + count);
        if (begin + count >= nnodes) break;
        if ((e = cudaMemsetAsync(d_counter, 0, 4, stream)) != cudaSuccess) return e;
        expand_level_kernel<<<blocks_for(count), kThreads, 0, stream>>>(node_words, level_nodes + begin, count, level_nodes + begin + count, d_counter);
        uint32_t next = 0;
        if ((e = cudaMemcpyAsync(&next, d_counter, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        if (begin + count + next > nnodes) return cudaErrorInvalidValue;  // not a tree
        begin += count;
        count = next;
    }
    return cudaGetLastError();
}

static cudaError_t fit_levels(const RptPerVertexData* verts, const uint4* tris, const uint32_t* orig_index, const uint32_t* level_nodes,
                              const std::vector<uint32_t>& level_offsets, uint32_t* node_words, void* node_box, cudaStream_t stream) {
    for (size_t L = level_offsets.size() - 1; L-- > 0;) {
        const uint32_t begin = level_offsets[L], count = level_offsets[L + 1] - begin;
        if (count) fit_level_kernel<<<blocks_for(count), kThreads, 0, stream>>>(verts, tris, orig_index, level_nodes + begin, count, node_words, static_cast<Box*>(node_box));
    }
    return cudaGetLastError();
}

cudaError_t device_build_wide_bvh(const RptPerVertexData* d_verts, const uint4* d_tris, uint32_t ntris, uint32_t shade_stride, DeviceBuildResult& out,
                                  cudaStream_t stream) {
    cudaError_t status = cudaSuccess;
    out = DeviceBuildResult{};
    int* d_bounds = nullptr;
    unsigned long long *d_codes = nullptr, *d_codes_sorted = nullptr;
    uint32_t *d_ids = nullptr, *d_ids_sorted = nullptr;
    void* d_sort_temp = nullptr;
    Task *d_tasks_a = nullptr, *d_tasks_b = nullptr;
    BuildCounters* d_ctr = nullptr;
    size_t sort_bytes = 0;
    const uint32_t max_nodes = std::max(ntris, 1u);  // every inner node has at least two children: fewer nodes than leaves
    std::vector<uint32_t> level_offsets;
    uint32_t levels = 0;
    // triangles per leaf slot (1..3, wide_bvh.h): fewer = tighter leaf boxes and fewer ray/triangle tests, more nodes
    uint32_t max_leaf = 1;  // measured (proxy / DarkCornell, Mpaths/s): 3 -> 596 / 939, 2 -> 624 / 940, 1 -> 642 / 962
    if (const char* v = std::getenv("RPT_DEVICE_BUILD_MAX_LEAF")) max_leaf = (uint32_t)std::min(3, std::max(1, std::atoi(v)));

    RPT_CK(dev_alloc(&d_bounds, 6));
    RPT_CK(dev_alloc(&d_codes, ntris)); RPT_CK(dev_alloc(&d_codes_sorted, ntris));
    RPT_CK(dev_alloc(&d_ids, ntris)); RPT_CK(dev_alloc(&d_ids_sorted, ntris));
    RPT_CK(dev_alloc(&d_tasks_a, max_nodes)); RPT_CK(dev_alloc(&d_tasks_b, max_nodes));
    RPT_CK(dev_alloc(&d_ctr, 1));
    RPT_CK(dev_alloc(&out.nodes, (size_t)max_nodes * 5));
    RPT_CK(dev_alloc(&out.tri_pos, (size_t)ntris * 3));
    RPT_CK(dev_alloc(&out.tri_shade, (size_t)ntris * shade_stride));
    RPT_CK(dev_alloc(&out.orig_index, ntris));
    RPT_CK(dev_alloc(&out.wide_index, ntris));
    RPT_CK(dev_alloc(&out.level_nodes, max_nodes));
    RPT_CK(cudaMalloc(&out.node_box, (size_t)max_nodes * sizeof(Box)));
    {
        auto ordered_host = [](float f) { int i; std::memcpy(&i, &f, 4); return i >= 0 ? i : i ^ 0x7FFFFFFF; };
        const int init[6] = {ordered_host(3e38f), ordered_host(3e38f), ordered_host(3e38f), ordered_host(-3e38f), ordered_host(-3e38f), ordered_host(-3e38f)};
        RPT_CK(cudaMemcpyAsync(d_bounds, init, sizeof init, cudaMemcpyHostToDevice, stream));
        RPT_CK(cudaStreamSynchronize(stream));  // (stack array)
    }
    centroid_bounds_kernel<<<std::min(blocks_for(ntris), 1184), kThreads, 0, stream>>>(d_verts, d_tris, ntris, d_bounds);
    morton_kernel<<<blocks_for(ntris), kThreads, 0, stream>>>(d_verts, d_tris, ntris, d_bounds, d_codes, d_ids);
    RPT_CK(cudaGetLastError());
    RPT_CK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_codes, d_codes_sorted, d_ids, d_ids_sorted, (int)ntris, 0, 63, stream));
    RPT_CK(cudaMalloc(&d_sort_temp, std::max<size_t>(sort_bytes, 1)));
    RPT_CK(cub::DeviceRadixSort::SortPairs(d_sort_temp, sort_bytes, d_codes, d_codes_sorted, d_ids, d_ids_sorted, (int)ntris, 0, 63, stream));

    // ---- topology: one launch per level, the next level's size read back in between
    RPT_CK(cudaMemsetAsync(out.nodes, 0, (size_t)max_nodes * 80, stream));
    {
        const BuildCounters start{1u, 0u, 0u, 0u};
        const Task root{0u, ntris, 0u};
        RPT_CK(cudaMemcpyAsync(d_ctr, &start, sizeof start, cudaMemcpyHostToDevice, stream));
        RPT_CK(cudaMemcpyAsync(d_tasks_a, &root, sizeof root, cudaMemcpyHostToDevice, stream));
        RPT_CK(cudaStreamSynchronize(stream));
    }
    {
        uint32_t ntasks = 1;
        Task *cur = d_tasks_a, *nxt = d_tasks_b;
        while (ntasks) {
            ++levels;
            split_level_kernel<<<blocks_for(ntasks), kThreads, 0, stream>>>(d_codes_sorted, d_ids_sorted, cur, ntasks, nxt, d_ctr, reinterpret_cast<uint32_t*>(out.nodes), out.orig_index, max_leaf);
            BuildCounters c{};
            RPT_CK(cudaMemcpyAsync(&c, d_ctr, sizeof c, cudaMemcpyDeviceToHost, stream));
            RPT_CK(cudaStreamSynchronize(stream));
            ntasks = c.next_tasks;
            out.nnodes = c.nodes;
            if (ntasks) {
                const uint32_t zero = 0;
                RPT_CK(cudaMemcpyAsync(&d_ctr->next_tasks, &zero, 4, cudaMemcpyHostToDevice, stream));
                RPT_CK(cudaStreamSynchronize(stream));
            }
            std::swap(cur, nxt);
            if (levels > 4096) { status = cudaErrorInvalidValue; goto done; }
        }
    }
    out.max_depth = levels - 1;
    // ---- triangle streams, level lists, boxes
    emit_triangles_kernel<<<blocks_for(ntris), kThreads, 0, stream>>>(d_verts, d_tris, out.orig_index, ntris, out.tri_pos, out.tri_shade, shade_stride, out.wide_index);
    RPT_CK(cudaGetLastError());
    RPT_CK(list_levels(reinterpret_cast<const uint32_t*>(out.nodes), out.nnodes, out.level_nodes, &d_ctr->pad, level_offsets, stream));
    if (level_offsets.back() != out.nnodes) { status = cudaErrorInvalidValue; goto done; }
    RPT_CK(fit_levels(d_verts, d_tris, out.orig_index, out.level_nodes, level_offsets, reinterpret_cast<uint32_t*>(out.nodes), out.node_box, stream));
    RPT_CK(cudaStreamSynchronize(stream));
    out.level_offsets = level_offsets;
done:
    for (void* p : {(void*)d_bounds, (void*)d_codes, (void*)d_codes_sorted, (void*)d_ids, (void*)d_ids_sorted, d_sort_temp, (void*)d_tasks_a, (void*)d_tasks_b, (void*)d_ctr})
        if (p) cudaFree(p);
    if (status != cudaSuccess) out.release();
    return status;
}

cudaError_t device_refit_wide_bvh(const RptPerVertexData* d_verts, const uint4* d_tris, uint32_t ntris, uint32_t shade_stride, DeviceBuildResult& tree,
                                  cudaStream_t stream) {
    cudaError_t status = cudaSuccess;
    uint32_t* d_counter = nullptr;
    if (tree.level_offsets.empty()) {  // a host-collapsed tree: list its levels once
        RPT_CK(dev_alloc(&d_counter, 1));
        if (!tree.level_nodes) RPT_CK(dev_alloc(&tree.level_nodes, tree.nnodes));
        if (!tree.node_box) RPT_CK(cudaMalloc(&tree.node_box, (size_t)tree.nnodes * sizeof(Box)));
        RPT_CK(list_levels(reinterpret_cast<const uint32_t*>(tree.nodes), tree.nnodes, tree.level_nodes, d_counter, tree.level_offsets, stream));
        if (tree.level_offsets.back() != tree.nnodes) { tree.level_offsets.clear(); status = cudaErrorInvalidValue; goto done; }
    }
    emit_triangles_kernel<<<blocks_for(ntris), kThreads, 0, stream>>>(d_verts, d_tris, tree.orig_index, ntris, tree.tri_pos, tree.tri_shade, shade_stride, nullptr);
    RPT_CK(cudaGetLastError());
    RPT_CK(fit_levels(d_verts, d_tris, tree.orig_index, tree.level_nodes, tree.level_offsets, reinterpret_cast<uint32_t*>(tree.nodes), tree.node_box, stream));
done:
    if (d_counter) cudaFree(d_counter);
    return status;
}

void DeviceBuildResult::release() {
    for (void* p : {(void*)nodes, tri_pos_in_nodes_block ? nullptr : (void*)tri_pos, (void*)tri_shade, (void*)orig_index, (void*)wide_index, (void*)level_nodes, node_box})
        if (p) cudaFree(p);
    *this = DeviceBuildResult{};
}

}  // namespace rpt
