"""`World` — the scene buffers the tracing loop reads (mirror of src/asset.rs:9-16, 55-224).

`World.from_path` plays the role of the reference's `World::from_path`: import the scene, build
the binned-SAH BVH (which permutes the index buffer), build the light-pick table from the
permuted triangles, pack `PerVertexData`.  The heavy lifting is native (csrc/world_build.cpp
through the C ABI of include/rpt_host.h); this class only owns the numpy buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass

import numpy as np

from . import capi
from .glb import MATERIAL_DTYPE, VERTEX_DTYPE, BakedScene, load_glb


@dataclass
class World:
    per_vertex_buffer: np.ndarray  # (V,) VERTEX_DTYPE
    index_buffer: np.ndarray  # (T,4) u32, BVH leaf order
    nodes: np.ndarray  # (N,) BVH_NODE_DTYPE
    material_data_buffer: np.ndarray  # (M,) MATERIAL_DTYPE
    light_pick_buffer: np.ndarray  # (L,) LIGHT_DTYPE, L >= 1
    atlas: np.ndarray | None = None  # (H,W,4) u8 or None when no material is textured
    build_seconds: dict | None = None

    @staticmethod
    def from_baked(scene: BakedScene, sah_samples: int = 128, atlas: np.ndarray | None = None) -> "World":
        lib = capi.lib()
        verts = np.ascontiguousarray(scene.vertices, np.float32)
        tris = np.ascontiguousarray(scene.indices, np.uint32).copy()
        mats = np.ascontiguousarray(scene.materials)
        nv, nt = len(verts), len(tris)
        nodes = np.zeros(max(2 * nt - 1, 1), capi.BVH_NODE_DTYPE)
        nnodes = C.c_uint32(0)
        t0 = time.perf_counter()
        capi.check(lib.rpt_build_bvh(capi.ptr(verts), C.c_uint32(nv), capi.ptr(tris), C.c_uint32(nt), C.c_uint32(sah_samples),
                                     capi.ptr(nodes), C.byref(nnodes)), "rpt_build_bvh")
        t1 = time.perf_counter()
        lights = np.zeros(max(nt, 1), capi.LIGHT_DTYPE)
        nlights = C.c_uint32(0)
        capi.check(lib.rpt_build_light_pick_table(capi.ptr(verts), C.c_uint32(nv), capi.ptr(tris), C.c_uint32(nt), capi.ptr(mats),
                                                  C.c_uint32(len(mats)), capi.ptr(lights), C.byref(nlights)), "rpt_build_light_pick_table")
        t2 = time.perf_counter()
        packed = np.zeros(nv, VERTEX_DTYPE)
        capi.check(lib.rpt_pack_per_vertex(capi.ptr(verts), capi.ptr(np.ascontiguousarray(scene.normals, np.float32)),
                                           capi.ptr(np.ascontiguousarray(scene.tangents, np.float32)),
                                           capi.ptr(np.ascontiguousarray(scene.uvs, np.float32)), C.c_uint32(nv), capi.ptr(packed)),
                   "rpt_pack_per_vertex")
        return World(packed, tris, nodes[: nnodes.value].copy(), mats.copy(), lights[: nlights.value].copy(), atlas,
                     {"bvh": t1 - t0, "light_table": t2 - t1})

    @staticmethod
    def from_path(path: str) -> "World | None":
        """.glb (imported here) or a baked .npz fixture; returns None on failure like the reference."""
        try:
            if path.endswith(".npz"):
                return World.from_baked(BakedScene.load(path))
            scene = load_glb(path)
            atlas = None
            if any(scene.textures):
                from .atlas import pack_scene_textures

                atlas = pack_scene_textures(scene)
            return World.from_baked(scene, atlas=atlas)
        except (OSError, ValueError, KeyError, NotImplementedError):
            return None

    @property
    def ntriangles(self) -> int:
        return len(self.index_buffer)


def blue_noise_r8() -> np.ndarray:
    """R channel of the reference's blue-noise tile after `into_rgba8()` (src/trace.rs:2,5):
    256x256 u8, committed as a fixture (tools/make_fixtures.py converts the 16-bit PNG with
    image-0.24's `(c + 128) / 257`)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources", "bluenoise_r8.npy")
    return np.load(path)


def make_rng_seeds(width: int, height: int, use_blue_noise: bool = True, uniform_seed: int = 0) -> np.ndarray:
    """Per-pixel (x, y) seeds of src/trace.rs:149-160: blue-noise mode x = 0, y = scaled tile value."""
    seeds = np.zeros((height * width, 2), np.uint32)
    blue = np.ascontiguousarray(blue_noise_r8()) if use_blue_noise else None
    bh, bw = (blue.shape if blue is not None else (0, 0))
    capi.check(capi.lib().rpt_make_rng_seeds(capi.ptr(blue), C.c_uint32(bw), C.c_uint32(bh), C.c_uint32(width), C.c_uint32(height),
                                             C.c_uint64(uniform_seed), capi.ptr(seeds)), "rpt_make_rng_seeds")
    return seeds
