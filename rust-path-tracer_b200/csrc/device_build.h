// device_build.h — build / refit of the 8-wide BVH on the device (device_build.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../../include/rpt_shared_structs.h"
#include "device_scene.h"

namespace rpt {

// Device buffers of a wide tree plus what a refit needs.  Owned by whoever holds the struct (release()).
struct DeviceBuildResult {
    uint4* nodes = nullptr;          // 5 per node (wide_bvh.h)
    float4* tri_pos = nullptr;       // 3 per triangle, leaf order
    float4* tri_shade = nullptr;     // shade_stride per triangle, leaf order (device_scene.h)
    uint32_t* orig_index = nullptr;  // leaf order -> caller's triangle index
    uint32_t* wide_index = nullptr;  // caller's triangle index -> leaf order
    uint32_t* level_nodes = nullptr; // node ids, level after level (root first)
    void* node_box = nullptr;        // one float[6] box per node (scratch of the bottom-up fit)
    std::vector<uint32_t> level_offsets;  // level L = level_nodes[level_offsets[L] .. level_offsets[L + 1])
    uint32_t nnodes = 0, max_depth = 0;
    bool tri_pos_in_nodes_block = false;  // tri_pos points into the allocation of `nodes` (one range for the L2 window)
    void release();
};

// d_verts / d_tris: the caller's vertex records and (i0, i1, i2, material) triangles, on the device.
cudaError_t device_build_wide_bvh(const RptPerVertexData* d_verts, const uint4* d_tris, uint32_t ntris, uint32_t shade_stride, DeviceBuildResult& out,
                                  cudaStream_t stream);
// Vertices moved, topology unchanged: triangle streams and boxes again.  `tree` may describe a host-collapsed tree
// (nodes, tri_pos, tri_shade, orig_index and nnodes set, level lists empty: they are derived here, once).
cudaError_t device_refit_wide_bvh(const RptPerVertexData* d_verts, const uint4* d_tris, uint32_t ntris, uint32_t shade_stride, DeviceBuildResult& tree,
                                  cudaStream_t stream);

}  // namespace rpt
