"""Texture atlas packing — host-side input producer: a thin binding of the C++ restatement
of src/atlas.rs and the texture part of src/asset.rs:135-192 (csrc/atlas_build.cpp, include/rpt_host.h).

* quadtree split of the atlas until there are more leaves than textures, leaves sorted by
  descending width (stable) and truncated (src/atlas.rs:26-69);
* every texture resized to its leaf (Lanczos3), flipped vertically, copied in (src/atlas.rs:71-87);
* the rect handed to the kernel is (x/W, y/W, w/W, h/H) — the y offset really is divided by the
  atlas WIDTH (src/atlas.rs:16-23; harmless for the square atlas the reference uses);
* albedo textures are gamma-2.2 decoded in 8 bits before packing (src/asset.rs:140-147);
* per material the order is albedo, metallic, roughness, normals (src/asset.rs:138-163,179-192).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def packing_rects(ntextures: int, atlas_w: int, atlas_h: int):
    rects = np.zeros((ntextures, 4), np.uint32)
    capi.check(capi.lib().rpt_atlas_rects(C.c_uint32(ntextures), C.c_uint32(atlas_w), C.c_uint32(atlas_h), capi.ptr(rects)), "rpt_atlas_rects")
    return [tuple(int(v) for v in r) for r in rects]


def pack_textures(textures, atlas_w: int = 4096, atlas_h: int = 4096):
    """textures: list of (H, W, 4) uint8.  Returns (atlas (atlas_h, atlas_w, 4) uint8, list of rects f32[4])."""
    texs = [np.ascontiguousarray(t, np.uint8) for t in textures]
    n = len(texs)
    atlas = np.empty((atlas_h, atlas_w, 4), np.uint8)
    sts = np.zeros((n, 4), np.float32)
    pointers = (C.c_void_p * max(n, 1))(*[t.ctypes.data for t in texs])
    widths = np.array([t.shape[1] for t in texs], np.uint32)
    heights = np.array([t.shape[0] for t in texs], np.uint32)
    capi.check(capi.lib().rpt_atlas_pack(pointers, capi.ptr(widths), capi.ptr(heights), C.c_uint32(n), C.c_uint32(atlas_w), C.c_uint32(atlas_h),
                                         capi.ptr(atlas), capi.ptr(sts)), "rpt_atlas_pack")
    return atlas, [sts[i].copy() for i in range(n)]


def decode_albedo_gamma(tex: np.ndarray) -> np.ndarray:
    """`((p / 255).powf(2.2) * 255) as u8` on RGB; the result is an RGB image (alpha dropped -> 255)."""
    tex = np.ascontiguousarray(tex, np.uint8)
    out = np.empty_like(tex)
    capi.check(capi.lib().rpt_decode_albedo_gamma(capi.ptr(tex), C.c_size_t(tex.shape[0] * tex.shape[1]), capi.ptr(out)), "rpt_decode_albedo_gamma")
    return out


def pack_scene_textures(scene, atlas_w: int = 4096, atlas_h: int = 4096) -> np.ndarray:
    """Pack `scene.textures` (one dict per material) and rewrite the materials' rect fields."""
    ordered, slots = [], []
    for mi, tex in enumerate(scene.textures):
        for key, field, flag in (("albedo", "albedo", "has_albedo_texture"), ("metallic", "metallic", "has_metallic_texture"),
                                 ("roughness", "roughness", "has_roughness_texture"), ("normals", "normals", "has_normal_texture")):
            if key in tex:
                img = decode_albedo_gamma(tex[key]) if key == "albedo" else tex[key]
                ordered.append(img)
                slots.append((mi, field, flag))
    atlas, sts = pack_textures(ordered, atlas_w, atlas_h)
    for (mi, field, flag), st in zip(slots, sts):
        scene.materials[mi][field] = st
        scene.materials[mi][flag] = 1
    return atlas
