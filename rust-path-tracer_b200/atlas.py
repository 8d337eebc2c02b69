"""Texture atlas packing — host-side input producer, restating src/atlas.rs and
the texture part of the material loop in src/asset.rs:135-192.

* quadtree split of the atlas until there are more leaves than textures, leaves sorted by
  descending width (stable) and truncated (src/atlas.rs:26-69);
* every texture resized to its leaf, flipped vertically, copied in (src/atlas.rs:71-87);
* the rect handed to the kernel is (x/W, y/W, w/W, h/H) — the y offset really is divided by the
  atlas WIDTH (src/atlas.rs:16-23; harmless for the square atlas the reference uses);
* albedo textures are gamma-2.2 decoded in 8 bits before packing (src/asset.rs:140-147);
* per material the order is albedo, metallic, roughness, normals (src/asset.rs:138-163,179-192).

The reference resizes with fast_image_resize's Lanczos3; here PIL's LANCZOS is used when a texture
is not already leaf-sized (not bit-identical; the synthetic scenes of this repo generate textures
at leaf size, so no resampling happens on the paths that are parity-tested).
"""
from __future__ import annotations

from collections import deque

import numpy as np


def packing_rects(ntextures: int, atlas_w: int, atlas_h: int):
    queue = deque([(0, 0, atlas_w, atlas_h)])
    while len(queue) <= ntextures:
        x, y, w, h = queue.popleft()
        hw, hh = w // 2, h // 2
        queue.extend([(x, y, hw, hh), (x + hw, y, hw, hh), (x, y + hh, hw, hh), (x + hw, y + hh, hw, hh)])
    leaves = sorted(queue, key=lambda r: -r[2])  # stable, like slice::sort_by
    return leaves[:ntextures]


def pack_textures(textures, atlas_w: int = 4096, atlas_h: int = 4096):
    """textures: list of (H, W, 4) uint8.  Returns (atlas (atlas_h, atlas_w, 4) uint8, list of rects f32[4])."""
    atlas = np.zeros((atlas_h, atlas_w, 4), np.uint8)
    rects = packing_rects(len(textures), atlas_w, atlas_h) if textures else []
    sts = []
    for tex, (x, y, w, h) in zip(textures, rects):
        tex = np.ascontiguousarray(tex, np.uint8)
        if tex.shape[0] != h or tex.shape[1] != w:
            from PIL import Image

            tex = np.asarray(Image.fromarray(tex, "RGBA").resize((w, h), Image.LANCZOS), np.uint8)
        atlas[y:y + h, x:x + w] = tex[::-1]  # flipv
        f = np.float32
        sts.append(np.array([f(x) / f(atlas_w), f(y) / f(atlas_w), f(w) / f(atlas_w), f(h) / f(atlas_h)], np.float32))
    return atlas, sts


def decode_albedo_gamma(tex: np.ndarray) -> np.ndarray:
    """`((p / 255).powf(2.2) * 255) as u8` on RGB; the result is an RGB image (alpha dropped -> 255)."""
    rgb = tex[..., :3].astype(np.float32) / np.float32(255.0)
    lin = np.power(rgb, np.float32(2.2), dtype=np.float32) * np.float32(255.0)
    out = np.empty(tex.shape[:2] + (4,), np.uint8)
    out[..., :3] = np.clip(np.floor(lin), 0, 255).astype(np.uint8)
    out[..., 3] = 255
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
