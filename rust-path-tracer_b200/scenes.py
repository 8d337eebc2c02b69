"""Deterministic synthetic inputs for the configs the shipped assets cannot cover.

* `textured_pbr_variant` — BASELINE.json config 3 names a "textured metallic/roughness atlas", but
  none of the four shipped .glb files contains an image: this assigns procedurally generated
  albedo / metallic / roughness / normal textures to PBRTest's sphere materials.
* `synthetic_hdr_sky` — a lat-long float4 sky with a sun disk, standing in for an .hdr file.
* `breaktime_proxy` — `scenes/BreakTime.glb` is absent from the reference checkout
  (.MISSING_LARGE_BLOBS).  This builds a LABELLED PROXY of comparable weight: a room with window
  openings, a grid of tessellated objects with textured materials, ceiling emitters — about one
  million triangles by default.  Every report that uses it says "BreakTime proxy".

Everything is a pure function of its arguments (fixed numpy seeds).
"""
from __future__ import annotations

import numpy as np

from .glb import MATERIAL_DTYPE, BakedScene


def _value_noise(rs, size: int, octaves: int = 4) -> np.ndarray:
    out = np.zeros((size, size), np.float32)
    amp, total = 1.0, 0.0
    for o in range(octaves):
        n = 4 << o
        grid = rs.random((n + 1, n + 1), dtype=np.float32)
        grid[-1], grid[:, -1] = grid[0], grid[:, 0]  # tileable
        t = np.linspace(0, n, size, endpoint=False, dtype=np.float32)
        i = t.astype(np.int32)
        f = t - i
        f = f * f * (3 - 2 * f)
        a = grid[i][:, i] * (1 - f)[None, :] + grid[i][:, i + 1] * f[None, :]
        b = grid[i + 1][:, i] * (1 - f)[None, :] + grid[i + 1][:, i + 1] * f[None, :]
        out += amp * (a * (1 - f)[:, None] + b * f[:, None])
        total += amp
        amp *= 0.5
    return out / total


def procedural_textures(seed: int, size: int) -> dict:
    """One material's albedo / metallic / roughness / normal textures, (size, size, 4) uint8 each."""
    rs = np.random.default_rng(seed)
    n1, n2, n3 = _value_noise(rs, size), _value_noise(rs, size), _value_noise(rs, size)
    base = rs.random(3).astype(np.float32) * 0.7 + 0.25
    albedo = np.clip(base[None, None, :] * (0.55 + 0.9 * n1[..., None]), 0, 1)
    checker = ((np.indices((size, size)).sum(axis=0) // max(size // 8, 1)) % 2).astype(np.float32)
    metallic = np.clip(0.15 + 0.8 * checker * (n2 > 0.45), 0, 1)
    roughness = np.clip(0.08 + 0.85 * n3, 0, 1)
    gy, gx = np.gradient(n2 * 6.0)
    nrm = np.stack([-gx, -gy, np.ones_like(gx)], axis=-1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)

    def rgba(a):
        a = np.asarray(a, np.float32)
        if a.ndim == 2:
            a = np.repeat(a[..., None], 3, axis=2)
        out = np.empty(a.shape[:2] + (4,), np.uint8)
        out[..., :3] = np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8)
        out[..., 3] = 255
        return out

    return {"albedo": rgba(albedo), "metallic": rgba(metallic), "roughness": rgba(roughness), "normals": rgba(nrm * 0.5 + 0.5)}


def textured_pbr_variant(pbr_scene: BakedScene, texture_size: int = 512, atlas_size: int = 4096, seed: int = 0):
    """(scene with textured materials, atlas).  texture_size must make 4 x nmaterials leaves of that size."""
    from .atlas import pack_scene_textures

    scene = BakedScene(pbr_scene.vertices, pbr_scene.normals, pbr_scene.tangents, pbr_scene.uvs, pbr_scene.indices,
                       pbr_scene.materials.copy(), [dict() for _ in pbr_scene.materials])
    for mi in range(len(scene.materials) - 1):  # the appended default material (ground plane) stays untextured
        scene.textures[mi] = procedural_textures(seed * 1000 + mi, texture_size)
    atlas = pack_scene_textures(scene, atlas_size, atlas_size)
    return scene, atlas


def synthetic_hdr_sky(width: int = 2048, height: int = 1024, sun_nits: float = 5000.0, seed: int = 1) -> np.ndarray:
    rs = np.random.default_rng(seed)
    v = np.linspace(0, 1, height, dtype=np.float32)[:, None]
    u = np.linspace(0, 1, width, endpoint=False, dtype=np.float32)[None, :]
    img = np.zeros((height, width, 4), np.float32)
    horizon = np.exp(-((v - 0.5) * 6.0) ** 2)
    img[..., 0] = 0.25 + 0.9 * horizon + 0.1 * v
    img[..., 1] = 0.45 + 0.7 * horizon
    img[..., 2] = 1.1 - 0.5 * v + 0.3 * horizon
    img[height // 2:, :, :3] *= np.float32(0.35)  # ground half
    clouds = _value_noise(rs, 256)
    reps = (-(-height // 256), -(-width // 256))
    img[..., :3] *= (0.8 + 0.4 * np.tile(clouds, reps)[:height, :width, None])
    d2 = ((u - 0.3) * 2.0) ** 2 + (v - 0.2) ** 2
    img[..., :3] += np.float32(sun_nits) * np.exp(-d2 / np.float32(2e-5))[..., None]
    img[..., 3] = 1.0
    return img.astype(np.float32)


def _icosphere(subdiv: int):
    t = (1 + 5 ** 0.5) / 2
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        mid = {}
        verts = list(v)
        nf = []

        def m(a, b):
            key = (min(a, b), max(a, b))
            if key not in mid:
                p = verts[a] + verts[b]
                verts.append(p / np.linalg.norm(p))
                mid[key] = len(verts) - 1
            return mid[key]

        for a, b, c in f:
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.array(verts), np.array(nf, np.int64)
    return v, f


def breaktime_proxy(target_triangles: int = 1_000_000, n_materials: int = 16, texture_size: int = 512, atlas_size: int = 4096, seed: int = 0):
    """Returns (BakedScene, atlas).  ~target_triangles triangles; deterministic in its arguments."""
    from .atlas import pack_scene_textures

    rs = np.random.default_rng(seed)
    verts, nrms, uvs, tris = [], [], [], []
    count = 0

    def add(p, n, uv, f, mat):
        nonlocal count
        verts.append(p); nrms.append(n); uvs.append(uv)
        tris.append(np.concatenate([f + count, np.full((len(f), 1), mat, np.int64)], axis=1))
        count += len(p)

    def quad(origin, eu, ev, mat, res=1, uv_scale=1.0):
        s = np.linspace(0, 1, res + 1)
        gu, gv = np.meshgrid(s, s, indexing="xy")
        p = origin[None, :] + gu.reshape(-1, 1) * eu[None, :] + gv.reshape(-1, 1) * ev[None, :]
        n = np.cross(eu, ev); n = n / np.linalg.norm(n)
        idx = np.arange((res + 1) ** 2).reshape(res + 1, res + 1)
        a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
        f = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
        add(p, np.tile(n, (len(p), 1)), np.stack([gu.ravel(), gv.ravel()], 1) * uv_scale, f, mat)

    # scene frame = the tracer's (x right, y up, z forward); the default camera sits at (0,1,-5) looking +z
    X, Y0, Y1, Z0, Z1 = 8.0, 0.0, 6.0, -7.0, 11.0
    m_floor, m_wall, m_ceil, m_light = 0, 1, 2, n_materials  # emitter = first untextured material
    quad(np.array([-X, Y0, Z0]), np.array([0, 0, Z1 - Z0]), np.array([2 * X, 0, 0]), m_floor, 64, 6.0)      # floor (normal +y)
    quad(np.array([-X, Y1, Z0]), np.array([2 * X, 0, 0]), np.array([0, 0, Z1 - Z0]), m_ceil, 32, 4.0)       # ceiling (normal -y)
    quad(np.array([-X, Y0, Z1]), np.array([0, Y1, 0]), np.array([2 * X, 0, 0]), m_wall, 32, 3.0)            # back wall
    quad(np.array([-X, Y0, Z0]), np.array([2 * X, 0, 0]), np.array([0, Y1, 0]), m_wall, 32, 3.0)            # wall behind the camera
    for side in (-1.0, 1.0):  # side walls with three window openings each (strips between the windows)
        for k in range(7):
            z0 = Z0 + (Z1 - Z0) * k / 7
            z1 = Z0 + (Z1 - Z0) * (k + 1) / 7
            if k % 2 == 1:  # window: sill and lintel only
                parts = [(Y0, 1.2), (4.2, Y1)]
            else:
                parts = [(Y0, Y1)]
            for ya, yb in parts:
                o = np.array([side * X, ya, z0 if side < 0 else z1])
                quad(o, np.array([0, yb - ya, 0]), np.array([0, 0, (z1 - z0) * (1 if side < 0 else -1)]), m_wall, 8, 2.0)  # normal faces the room
    for k in range(4):  # ceiling emitters, facing down
        cx, cz = (-3.5 if k % 2 == 0 else 3.5), (-1.0 if k < 2 else 6.0)
        quad(np.array([cx - 0.9, Y1 - 0.02, cz - 0.9]), np.array([1.8, 0, 0]), np.array([0, 0, 1.8]), m_light, 2)

    room_tris = sum(len(t) for t in tris)
    sv, sf = _icosphere(5)  # 20480 triangles per object
    n_objects = max(1, (target_triangles - room_tris) // len(sf))
    cols = int(np.ceil(np.sqrt(n_objects * 1.6)))
    su = 0.5 + np.arctan2(sv[:, 2], sv[:, 0]) / (2 * np.pi)
    sw = 0.5 - np.arcsin(np.clip(sv[:, 1], -1, 1)) / np.pi
    for k in range(n_objects):
        gx, gz = k % cols, k // cols
        radius = 0.22 + 0.3 * rs.random()
        squash = 0.6 + 0.8 * rs.random(3)
        centre = np.array([-X + 1.0 + (2 * X - 2.0) * (gx + 0.5) / cols + 0.2 * rs.standard_normal(),
                           radius * squash[1] + (0.0 if rs.random() < 0.7 else 1.0 + 2.5 * rs.random()),
                           Z0 + 3.5 + (Z1 - Z0 - 4.5) * (gz + 0.5) / max(1, -(-n_objects // cols)) + 0.2 * rs.standard_normal()])
        bump = 1.0 + 0.08 * np.sin(sv @ (rs.standard_normal(3) * 7.0))
        p = centre[None, :] + sv * bump[:, None] * (radius * squash)[None, :]
        n = sv / squash[None, :]
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        add(p, n, np.stack([su, sw], 1) * 2.0, sf, 3 + k % (n_materials - 3))

    materials = np.zeros(n_materials + 1, MATERIAL_DTYPE)
    textures = [dict() for _ in range(n_materials + 1)]
    for mi in range(n_materials):
        materials[mi]["albedo"] = (0.8, 0.8, 0.8, 1.0)
        materials[mi]["roughness"] = 0.5
        materials[mi]["metallic"] = 0.0
        textures[mi] = procedural_textures(seed * 1000 + 500 + mi, texture_size)
    materials[m_light]["albedo"] = (0.8, 0.8, 0.8, 1.0)
    materials[m_light]["emissive"] = (15.0 * 1.2, 15.0 * 1.1, 15.0 * 0.9, 15.0)  # emissiveFactor x 15, src/asset.rs:165-168
    materials[m_light]["roughness"] = 1.0

    pos = np.concatenate(verts).astype(np.float32)
    nrm = np.concatenate(nrms).astype(np.float32)
    uv = np.concatenate(uvs).astype(np.float32)
    idx = np.concatenate(tris).astype(np.uint32)
    from .glb import _tangents

    tan = _tangents(pos, uv, nrm, idx[:, :3].astype(np.int64))
    one = np.ones((len(pos), 1), np.float32)
    zero = np.zeros((len(pos), 1), np.float32)
    scene = BakedScene(np.concatenate([pos, one], 1), np.concatenate([nrm, zero], 1), np.concatenate([tan, zero], 1), uv, idx, materials, textures)
    atlas = pack_scene_textures(scene, atlas_size, atlas_size)
    return scene, atlas
