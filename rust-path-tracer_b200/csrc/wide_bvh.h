// wide_bvh.h — the backend's private scene layout, built at rpt_upload_world from the
// reference's buffers (BVHNode[], UVec4 index buffer, PerVertexData[]).
//
// * 8-wide compressed BVH: each node is 80 bytes = five 16-byte vector loads.  Child boxes are
//   quantised to 8 bits per plane on a per-node power-of-two grid anchored at the node's min
//   corner (conservative: lo floored, hi ceiled); children sit in octant-ordered slots so a
//   ray's visiting order is `slot ^ octant` with no distance sort.
//       word 0: origin.x, origin.y, origin.z, cell.x (f32; cells are powers of two)
//       word 1: first_child_node, first_triangle, [valid_triangles (24 bits) | inner_mask<<24],
//               [upper half of cell.y's bits | upper half of cell.z's bits << 16]
//       word 2: qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7]
//       word 3: qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7]
//       word 4: qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7]
//   A slot is an inner child (its inner_mask bit; child nodes are contiguous in slot order), a leaf
//   of 1..3 triangles (valid_triangles bits 3*slot .. 3*slot+2 set in unary; the node's triangles
//   are contiguous in bit order, so bit k is triangle first_triangle + popc(valid below k)), or empty.
// * triangle position stream, in wide-leaf order, 3 x float4 per triangle:
//       (a.xyz, bits(reference triangle index)), (e1 = b-a, bits(material)), (e2 = c-a, 0)
//   e1/e2 are single IEEE subtractions done on the host, so the device-side ray/triangle test
//   computes exactly what the reference computes from a, b, c.
// * triangle shading stream, 4 x float4: (n_a, uv_a.x) (n_b, uv_a.y) (n_c, uv_b.x)
//   (uv_b.y, uv_c.x, uv_c.y, 0); optional tangent stream 3 x float4.
#pragma once

#include <cstdint>
#include <memory>
#include <new>
#include <utility>
#include <vector>

#include "../../include/rpt_shared_structs.h"

namespace rpt {

struct WideNode {
    uint32_t w[20];
};
static_assert(sizeof(WideNode) == 80, "wide node is five 16-byte words");

// std::vector whose resize() leaves new elements uninitialised: the big per-triangle arrays are written exactly once,
// by the build threads, and a serial zero fill (plus its page faults) of ~60 bytes per triangle is start-up time.
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
    template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using UninitVector = std::vector<T, DefaultInitAllocator<T>>;

struct WideBvh {
    std::vector<WideNode> nodes;
    UninitVector<float> tri_pos;        // 12 floats per triangle
    UninitVector<uint32_t> orig_index;  // wide order -> reference triangle index
    std::vector<uint32_t> wide_index;  // reference triangle index -> wide order
    uint32_t max_depth = 0;            // levels below the root (stack entries needed <= max_depth + 1)
    uint32_t inner_children = 0, leaf_children = 0;
};

// Collapse the reference's binary BVH into the wide layout.  Returns false (with a message) if
// the input is malformed.
bool build_wide_bvh(const RptBVHNode* nodes, uint32_t nnodes, const uint32_t* triangles, uint32_t ntriangles,
                    const RptPerVertexData* vertices, uint32_t nvertices, WideBvh& out, const char** error);

// Traversal stack: the first kWideStackShared entries of a lane live in shared memory (8-wide: a 1M-triangle scene
// is 10 levels deep), deeper ones in a global-memory overflow area; trees deeper than kWideStackCapacity are
// rejected at upload.
#ifndef RPT_STACK_SHARED
#define RPT_STACK_SHARED 12
#endif
constexpr uint32_t kWideStackShared = RPT_STACK_SHARED;
constexpr uint32_t kWideStackCapacity = 64;

}  // namespace rpt
