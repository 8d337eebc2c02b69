"""Known answers for the tracing path that do NOT come from the oracle.

The reference holds one golden value for this path (tests/correctness_tests.rs:14-33: FurnaceTest, pixel
(65, 75), 0.8 +- 0.02 in gamma space), and that pixel only sees the diffuse lobe of an untextured grey sphere.
Everything here widens the pins with answers derived from the reference's own formulas by closed form or by
numpy quadrature, or from laws the reference's estimators obey by construction (an emitter seen from inside is a
constant environment; a constant texture is a constant; NEE / MIS / Russian roulette are unbiased when the
throughput stays below 1).  Each case is a (world, config, sky) triple plus a prediction, rendered by whoever
calls it: the CPU oracle (tests/test_known_answers.py, CPU suite) and the CUDA path (same file, -m gpu).

Scene: FurnaceTest.glb's fixture — a grey unit icosphere (material 0: albedo 0.18, roughness 1, metallic 0) at
the origin inside an emissive shell of radius 6.85 (material 1: emission 3.0), camera (0, 1, -5) looking +z.
"""
from __future__ import annotations

import functools
import os

import numpy as np

import helpers

SPHERE_CENTRE = np.zeros(3)
SPHERE_RADIUS = 1.0
SHELL_EMISSION = 3.0


def _furnace_baked():
    from rust_path_tracer_b200.glb import BakedScene

    return BakedScene.load(os.path.join(helpers.SCENE_DIR, "FurnaceTest.npz"))


def _clone(scene, indices=None, materials=None, textures=None):
    from rust_path_tracer_b200.glb import BakedScene

    return BakedScene(scene.vertices, scene.normals, scene.tangents, scene.uvs, scene.indices.copy() if indices is None else indices,
                      scene.materials.copy() if materials is None else materials, textures if textures is not None else [dict() for _ in scene.materials])


@functools.lru_cache(maxsize=None)
def sphere_only_world(albedo=(0.18, 0.18, 0.18), roughness=1.0, metallic=0.0):
    """The grey sphere without the shell (for constant-environment cases)."""
    from rust_path_tracer_b200.world import World

    base = _furnace_baked()
    mats = base.materials.copy()
    mats[0]["albedo"] = (*albedo, 1.0)
    mats[0]["roughness"] = roughness
    mats[0]["metallic"] = metallic
    return World.from_baked(_clone(base, indices=base.indices[base.indices[:, 3] == 0].copy(), materials=mats))


@functools.lru_cache(maxsize=None)
def furnace_world(shell_albedo=None):
    """The full furnace; shell_albedo = 0 makes the emitter reflect nothing in the diffuse lobe (NEE mode 2 shades
    emitters hit after a diffuse bounce as ordinary surfaces, lib.rs:97-109)."""
    from rust_path_tracer_b200.world import World

    base = _furnace_baked()
    mats = base.materials.copy()
    if shell_albedo is not None:
        mats[1]["albedo"] = (shell_albedo, shell_albedo, shell_albedo, 1.0)
    return World.from_baked(_clone(base, materials=mats))


@functools.lru_cache(maxsize=None)
def constant_texture_furnace_world(byte=118):
    """The furnace with the sphere's albedo / roughness / metallic coming from CONSTANT textures through the atlas,
    packed by the product's atlas packer (albedo texels are gamma-decoded in 8 bits on the way, src/asset.rs:140-147:
    118 -> 46, i.e. 0.1804).  Returns (world, equivalent untextured world)."""
    from rust_path_tracer_b200.atlas import decode_albedo_gamma, pack_scene_textures
    from rust_path_tracer_b200.world import World

    base = _furnace_baked()
    tex = [dict() for _ in base.materials]

    def const(v):
        img = np.empty((16, 16, 4), np.uint8)
        img[..., :3] = v
        img[..., 3] = 255
        return img

    tex[0] = {"albedo": const(byte), "roughness": const(255), "metallic": const(0)}
    textured = _clone(base, textures=tex)
    # keep the lookups strictly inside each rect: the CPU polyfill reads the texel at ceil(coord) too, which at uv = 1
    # belongs to the neighbouring rect (image_polyfill.rs:38-46 — faithful bleeding, but not a constant any more)
    textured.uvs = (0.25 + 0.5 * (base.uvs - np.floor(base.uvs))).astype(np.float32)
    atlas = pack_scene_textures(textured, 256, 256)
    w_tex = World.from_baked(textured, atlas=atlas)
    mats = base.materials.copy()
    a = np.float32(decode_albedo_gamma(const(byte))[0, 0, 0]) / np.float32(255.0)
    mats[0]["albedo"] = (a, a, a, 1.0)
    w_ref = World.from_baked(_clone(base, materials=mats))
    return w_tex, w_ref


@functools.lru_cache(maxsize=None)
def flat_normal_map_sphere_world():
    """The grey sphere with a FLAT normal map (texel (128, 128, 255): tangent-space normal (0.004, 0.004, 1)) through the
    atlas: lib.rs:131-141 must hand the BSDF (almost exactly) the geometric normal, normalised."""
    from rust_path_tracer_b200.atlas import pack_scene_textures
    from rust_path_tracer_b200.world import World

    base = _furnace_baked()
    flat = np.empty((16, 16, 4), np.uint8)
    flat[...] = (128, 128, 255, 255)
    tex = [dict() for _ in base.materials]
    tex[0] = {"normals": flat}
    scene = _clone(base, indices=base.indices[base.indices[:, 3] == 0].copy(), textures=tex)
    scene.uvs = (0.25 + 0.5 * (base.uvs - np.floor(base.uvs))).astype(np.float32)
    atlas = pack_scene_textures(scene, 256, 256)
    return World.from_baked(scene, atlas=atlas)


@functools.lru_cache(maxsize=None)
def gradient_albedo_sphere_world():
    """The sphere with a smooth, non-constant albedo texture through the atlas (red rises with x, green with y of the
    texture) — what a constant texture cannot show: uv interpolation (H1), the rect mapping and the bilinear polyfill on
    RGBA8 texels (H3).  Returns the world (its atlas is what the prediction reads)."""
    from rust_path_tracer_b200.atlas import pack_scene_textures
    from rust_path_tracer_b200.world import World

    base = _furnace_baked()
    n = 64
    ramp = np.linspace(40, 250, n).astype(np.uint8)
    img = np.empty((n, n, 4), np.uint8)
    img[..., 0] = ramp[None, :]
    img[..., 1] = ramp[:, None]
    img[..., 2] = 128
    img[..., 3] = 255
    tex = [dict() for _ in base.materials]
    tex[0] = {"albedo": img}
    scene = _clone(base, indices=base.indices[base.indices[:, 3] == 0].copy(), textures=tex)
    scene.uvs = (0.25 + 0.5 * (base.uvs - np.floor(base.uvs))).astype(np.float32)
    atlas = pack_scene_textures(scene, 256, 256)
    return World.from_baked(scene, atlas=atlas)


def albedo_texture_prediction(world, width, height, cam=(0.0, 1.0, -5.0)):
    """Albedo the reference's lookup yields at every pixel-centre primary hit, float64 numpy straight from the scene
    buffers: brute-force nearest triangle, barycentrics of the hit point (util.rs:238-251), uv interpolation
    (lib.rs:119-129, out-of-range uvs wrapped by fract), the material's rect (bsdf.rs:356) and the CPU polyfill's bilinear
    with floor / ceil taps and modulo wrap on texels / 255 (image_polyfill.rs:32-55).  NaN where the ray misses."""
    ys, xs = np.mgrid[0:height, 0:width]
    u = ((xs + 0.5) / width) * 2 - 1
    v = ((1 - (ys + 0.5) / height) * 2 - 1) * (height / width)
    d = np.stack([u, v, np.ones_like(u)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.asarray(cam, np.float64)
    pos = world.per_vertex_buffer["vertex"][:, :3].astype(np.float64)
    uvs = np.stack([world.per_vertex_buffer["uv0"][:, 0], world.per_vertex_buffer["uv0"][:, 1]], 1).astype(np.float64) if "uv0" in world.per_vertex_buffer.dtype.names else None
    tri = world.index_buffer[:, :3]
    best_t = np.full(len(d), np.inf)
    best = np.full(len(d), -1)
    for i in range(len(tri)):
        a, e1, e2 = pos[tri[i, 0]], pos[tri[i, 1]] - pos[tri[i, 0]], pos[tri[i, 2]] - pos[tri[i, 0]]
        pv = np.cross(d, e2)
        det = pv @ e1
        ok = np.abs(det) >= 1e-6
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = o - a
        uu = (pv @ tv) * inv
        qv = np.cross(tv, e1)
        vv = (d @ qv) * inv
        t = (qv @ e2) * inv
        hit = ok & (uu >= 0) & (uu <= 1) & (vv >= 0) & (uu + vv <= 1) & (t > 0.001) & (t < best_t)
        best_t = np.where(hit, t, best_t)
        best[hit] = i
    hitmask = best >= 0
    p = o + d * np.where(hitmask, best_t, 0.0)[:, None]
    ia, ib, ic = (tri[np.maximum(best, 0), k] for k in range(3))
    v0, v1, v2 = pos[ib] - pos[ia], pos[ic] - pos[ia], p - pos[ia]
    d00, d01, d11 = (v0 * v0).sum(1), (v0 * v1).sum(1), (v1 * v1).sum(1)
    d20, d21 = (v2 * v0).sum(1), (v2 * v1).sum(1)
    denom = d00 * d11 - d01 * d01
    bv = (d11 * d20 - d01 * d21) / denom
    bw = (d00 * d21 - d01 * d20) / denom
    bu = 1 - bv - bw
    uv = bu[:, None] * uvs[ia] + bv[:, None] * uvs[ib] + bw[:, None] * uvs[ic]
    outside = ((uv < 0) | (uv > 1)).any(1)
    uv = np.where(outside[:, None], uv - np.floor(uv), uv)
    rect = world.material_data_buffer["albedo"][0].astype(np.float64)  # (u0, v0, su, sv)
    au, av = rect[0] + uv[:, 0] * rect[2], rect[1] + uv[:, 1] * rect[3]
    atlas = world.atlas.astype(np.float64)[..., :3] / 255.0
    h, w = atlas.shape[:2]
    px, py = au * w, av * h
    fx, fy = px - np.floor(px), py - np.floor(py)
    x0, x1 = np.floor(px).astype(int) % w, np.ceil(px).astype(int) % w
    y0, y1 = np.floor(py).astype(int) % h, np.ceil(py).astype(int) % h
    top = atlas[y0, x0] + (atlas[y0, x1] - atlas[y0, x0]) * fx[:, None]
    bot = atlas[y1, x0] + (atlas[y1, x1] - atlas[y1, x0]) * fx[:, None]
    out = top + (bot - top) * fy[:, None]
    out[~hitmask] = np.nan
    return out.reshape(height, width, 3)


def constant_sky(value=SHELL_EMISSION, width=8, height=4):
    img = np.empty((height, width, 4), np.float32)
    img[..., :3] = value
    img[..., 3] = 1.0
    return img


# ---- analytic side ---------------------------------------------------------------------------------------------
def sphere_pixel_cosines(width, height, cam=(0.0, 1.0, -5.0), margin=0.9):
    """For pixel centres of the default camera (lib.rs:38-51, rotation 0): cos(angle between the view direction and
    the sphere normal) where the primary ray hits the unit sphere inside `margin` of its silhouette radius, else NaN."""
    ys, xs = np.mgrid[0:height, 0:width]
    u = ((xs + 0.5) / width) * 2 - 1
    v = (1 - (ys + 0.5) / height) * 2 - 1
    v = v * (height / width)
    d = np.stack([u, v, np.ones_like(u)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.asarray(cam, np.float64) - SPHERE_CENTRE
    b = d @ o
    c = o @ o - SPHERE_RADIUS ** 2
    disc = b * b - c
    hit = disc > 0
    t = -b - np.sqrt(np.where(hit, disc, 0))
    p = o + d * t[..., None]
    n = p / SPHERE_RADIUS
    cosv = -(n * d).sum(-1)
    # impact parameter of the ray relative to the sphere radius: stay away from the silhouette, where the
    # tessellated sphere and its interpolated normals differ from the ideal one
    impact = np.sqrt(np.maximum(o @ o - b * b, 0)) / SPHERE_RADIUS
    return np.where(hit & (impact < margin), cosv, np.nan)


def schlick(cos_theta, f0):
    return f0 + (1 - f0) * (1 - cos_theta) ** 5


@functools.lru_cache(maxsize=None)
def _diffuse_table(f0=0.04, n_table=256, n_theta=400, n_phi=400):
    """E over cosine-distributed directions d of 1 - Schlick(h . v), h = normalize(v + d), as a function of
    cos(theta_v): the expected diffuse-lobe weight of `PBR::sample` per unit albedo when the lobe is always chosen
    (bsdf.rs:202-211, 301-315: kd = (1 - ks)(1 - metallic); spectrum / pdf = kd * albedo / (1 - specular_weight))."""
    cv = np.linspace(0.0, 1.0, n_table)
    # midpoint rule in (r1, phi): cosine-weighted sampling has cos^2(theta) = r1 uniformly distributed
    r1 = (np.arange(n_theta) + 0.5) / n_theta
    phi = (np.arange(n_phi) + 0.5) / n_phi * 2 * np.pi
    ct = np.sqrt(r1)[:, None]
    st = np.sqrt(1 - r1)[:, None]
    dx, dy, dz = st * np.cos(phi)[None, :], st * np.sin(phi)[None, :], ct * np.ones_like(phi)[None, :]
    out = np.empty(n_table)
    for i, c in enumerate(cv):
        vx, vz = np.sqrt(max(0.0, 1 - c * c)), c
        hx, hy, hz = dx + vx, dy, dz + vz
        hv = (hx * vx + hz * vz) / np.sqrt(hx * hx + hy * hy + hz * hz)
        out[i] = (1 - schlick(np.maximum(hv, 0.0), f0)).mean()
    return cv, out


def diffuse_only_prediction(width, height, albedo, environment):
    """Image of the sphere under a constant environment when only the diffuse lobe exists (specular_weight_clamp =
    (0, 0)); NaN outside the sphere's interior."""
    cosv = sphere_pixel_cosines(width, height)
    cv, table = _diffuse_table()
    e = np.interp(np.nan_to_num(cosv, nan=1.0), cv, table)
    img = environment * np.asarray(albedo, np.float64)[None, None, :] * e[..., None]
    img[np.isnan(cosv)] = np.nan
    return img


def mirror_prediction(width, height, albedo, metallic, environment):
    """Roughness -> 0 and specular_weight_clamp = (1, 1): `sample_ggx` returns the mirror direction, D cancels
    between spectrum and pdf, the Schlick-GGX geometry term is 1, so spectrum / pdf = Schlick(n . v, f0) with
    f0 = lerp(0.04, albedo, metallic) (bsdf.rs:212-233, 316-334; util.rs:58-85, 211-236)."""
    cosv = sphere_pixel_cosines(width, height)
    f0 = 0.04 + (np.asarray(albedo, np.float64) - 0.04) * metallic
    img = environment * schlick(np.nan_to_num(cosv, nan=1.0)[..., None], f0[None, None, :])
    img[np.isnan(cosv)] = np.nan
    return img


@functools.lru_cache(maxsize=None)
def _rough_specular_table(roughness, f0, n_table=128, n_phi=256, n_r2=1024):
    """E over (r1, r2) uniform in [0, 1)^2 of the specular lobe's spectrum / pdf as `PBR::sample` computes it with
    specular_weight = 1, per channel of f0, as a function of cos(theta_v) — float64 numpy straight from the reference's
    formulas, none of the repo's code:
      sample_ggx (util.rs:67-85)      l = GGX-distributed direction AROUND THE MIRROR DIRECTION R (a = roughness^2,
                                      cos = sqrt((1 - r2) / (r2 (a^2 - 1) + 1)), phi = 2 pi r1) — the sampled direction
                                      itself, not a half vector
      ggx_distribution (util.rs:58-64) D = r^2 / max(pi ((n.h)^2 (r^2 - 1) + 1)^2, EPS), h = normalize(v + l)
      geometry (util.rs:211-227)      G = g(n.v) g(n.l), g(x) = max(x, 0) / (max(x, 0) (1 - k) + k), k = r^2 / 8
      spectrum (bsdf.rs:212-227)      D G F / max(4 max(n.v, 0) c, EPS) * c, c = max(n.l, EPS), F = Schlick(max(h.v, 0), f0)
      pdf (bsdf.rs:233-241)           D (n.h) / (4 v.h)
    A sample below the horizon has G = 0 and contributes nothing; every other one leaves the convex sphere."""
    eps = 1e-3
    cv = np.linspace(0.05, 1.0, n_table)
    r1 = (np.arange(n_phi) + 0.5) / n_phi
    r2 = ((np.arange(n_r2) + 0.5) / n_r2)[:, None]
    a = roughness * roughness
    cos_t = np.sqrt((1 - r2) / (r2 * (a * a - 1) + 1))
    sin_t = np.sqrt(np.maximum(1 - cos_t * cos_t, 0))
    phi = 2 * np.pi * r1[None, :]
    k = roughness * roughness / 8
    g = lambda x: np.maximum(x, 0) / (np.maximum(x, 0) * (1 - k) + k)
    f0 = np.asarray(f0, np.float64)
    out = np.empty((n_table, 3))
    for i, c in enumerate(cv):
        s_v = np.sqrt(max(0.0, 1 - c * c))
        v = np.array([s_v, 0.0, c])           # normal = +z
        refl = np.array([-s_v, 0.0, c])       # reflect(-v, n)
        t = np.cross([0.0, 1.0, 0.0], refl)   # any frame around R: the lobe is isotropic in phi
        t /= np.linalg.norm(t)
        b = np.cross(refl, t)
        l = (t[None, None, :] * (np.cos(phi) * sin_t)[..., None] + b[None, None, :] * (np.sin(phi) * sin_t)[..., None]
             + refl[None, None, :] * (cos_t * np.ones_like(phi))[..., None])
        l /= np.linalg.norm(l, axis=-1, keepdims=True)
        nl = l[..., 2]
        h = l + v
        h /= np.linalg.norm(h, axis=-1, keepdims=True)
        nh, vh = h[..., 2], h @ v
        d = roughness ** 2 / np.maximum(np.pi * (np.maximum(nh, 0) ** 2 * (roughness ** 2 - 1) + 1) ** 2, eps)
        cth = np.maximum(nl, eps)
        fres = schlick(np.maximum(vh, 0)[..., None], f0[None, None, :])
        spectrum = (d * g(c) * g(nl))[..., None] * fres / np.maximum(4 * max(c, 0.0) * cth, eps)[..., None] * cth[..., None]
        pdf = d * nh / (4 * vh)
        w = np.where((nl > 0)[..., None], spectrum / pdf[..., None], 0.0)
        out[i] = w.mean((0, 1))
    return cv, out


def rough_metal_prediction(width, height, albedo, roughness, metallic, environment):
    """Image of the sphere under a constant environment when only the specular lobe exists (specular_weight_clamp =
    (1, 1)) at a finite roughness: environment * E[spectrum / pdf] by quadrature; NaN outside the sphere's interior
    and where the view grazes (cos < 0.05)."""
    cosv = sphere_pixel_cosines(width, height)
    f0 = tuple(float(x) for x in 0.04 + (np.asarray(albedo, np.float64) - 0.04) * metallic)
    cv, table = _rough_specular_table(float(roughness), f0)
    c = np.nan_to_num(cosv, nan=1.0)
    img = np.stack([environment * np.interp(c, cv, table[:, k]) for k in range(3)], -1)
    img[np.isnan(cosv) | (c < 0.05)] = np.nan
    return img


def relative_error_of_mean(image, prediction):
    """|mean(image) / mean(prediction) - 1| per channel over the pixels the prediction covers."""
    mask = np.isfinite(prediction).all(-1)
    return np.abs(image[mask].mean(0) / prediction[mask].mean(0) - 1.0), int(mask.sum())


def gradient_sky(width=256, height=128):
    """A smooth lat-long image whose bilinear interpolation is (almost) exact: texel = linear in u and v."""
    v, u = np.mgrid[0:height, 0:width]
    img = np.empty((height, width, 4), np.float32)
    img[..., 0] = u / width
    img[..., 1] = v / height
    img[..., 2] = 0.25
    img[..., 3] = 1.0
    return img


def sky_lookup_prediction(width, height, sky, sun_direction, cam_rotation=(0.0, 0.0)):
    """numpy (float64) restatement of lib.rs:70-78 + image_polyfill.rs:32-55 for pixel-centre primary rays that
    miss everything: yaw by atan2(sun.z, sun.x), lat-long (u, v), bilinear with floor / ceil and wrap, x sun.w / 15."""
    ys, xs = np.mgrid[0:height, 0:width]
    u = ((xs + 0.5) / width) * 2 - 1
    v = ((1 - (ys + 0.5) / height) * 2 - 1) * (height / width)
    d = np.stack([u, v, np.ones_like(u)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rx, ry = cam_rotation
    cx, sx, cy, sy = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry)
    rot_x = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    rot_y = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    d = d @ (rot_y @ rot_x).T
    yaw = np.arctan2(sun_direction[2], sun_direction[0])
    c, s = np.cos(yaw), np.sin(yaw)
    r = d @ np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]).T
    su = 0.5 + np.arctan2(r[..., 2], r[..., 0]) / (2 * np.pi)
    sv = 1 - (0.5 + np.arcsin(np.clip(r[..., 1], -1, 1)) / np.pi)
    h, w = sky.shape[:2]
    px, py = su * w, sv * h
    fx, fy = px - np.floor(px), py - np.floor(py)
    x0, x1 = np.floor(px).astype(int) % w, np.ceil(px).astype(int) % w
    y0, y1 = np.floor(py).astype(int) % h, np.ceil(py).astype(int) % h
    s64 = sky.astype(np.float64)[..., :3]
    top = s64[y0, x0] + (s64[y0, x1] - s64[y0, x0]) * fx[..., None]
    bot = s64[y1, x0] + (s64[y1, x1] - s64[y1, x0]) * fx[..., None]
    return (top + (bot - top) * fy[..., None]) * (sun_direction[3] / 15.0)


def procedural_sky_prediction(width, height, sun_direction, cam_position=(0.0, 1.0, -5.0), cam_rotation=(0.0, 0.0)):
    """numpy (float64) restatement of kernels/src/skybox.rs:18-94 for pixel-centre primary rays that miss everything:
    12-step single scattering (Rayleigh + Mie) from the camera position, two-sample optical depth towards the sun,
    phase functions, sqrt then ^2.2."""
    ys, xs = np.mgrid[0:height, 0:width]
    u = ((xs + 0.5) / width) * 2 - 1
    v = ((1 - (ys + 0.5) / height) * 2 - 1) * (height / width)
    d = np.stack([u, v, np.ones_like(u)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rx, ry = cam_rotation
    cx, sx, cy, sy = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry)
    d = d @ (np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])).T
    o = np.broadcast_to(np.asarray(cam_position, np.float64), d.shape)
    sun = np.asarray(sun_direction[:3], np.float64)
    ray_c = np.array([58e-7, 135e-7, 331e-7])
    mie_s = np.full(3, 2e-5)
    mie_e = mie_s * 1.1
    earth, atmo, h_ray, h_mie = 6360e3, 6380e3, 8e3, 12e2
    centre = np.array([0.0, -earth, 0.0])

    def escape(p, dirs, r):
        vv = p - centre
        b = (vv * dirs).sum(-1)
        det = b * b - (vv * vv).sum(-1) + r * r
        root = np.sqrt(np.maximum(det, 0))
        t = np.where(-b - root >= 0, -b - root, -b + root)
        return np.where(det < 0, -1.0, t)

    def densities(p):
        h = np.maximum(np.linalg.norm(p - centre, axis=-1) - earth, 0.0)
        return np.exp(-h / h_ray), np.exp(-h / h_mie)

    steps = 12
    depth = escape(o, d, atmo) / steps
    i_r = np.zeros(d.shape)
    i_m = np.zeros(d.shape)
    tot_r = np.zeros(d.shape[:2])
    tot_m = np.zeros(d.shape[:2])
    sun_dirs = np.broadcast_to(sun, d.shape)
    for i in range(steps):
        p = o + d * (depth * i)[..., None]
        dr, dm = densities(p)
        dr, dm = dr * depth, dm * depth
        tot_r += dr
        tot_m += dm
        l = escape(p, sun_dirs, atmo)
        r0, m0 = densities(p)
        r1, m1 = densities(p + sun_dirs * l[..., None])
        sum_r = tot_r + r0 * (l / 2) + r1 * (l / 2)
        sum_m = tot_m + m0 * (l / 2) + m1 * (l / 2)
        a = np.exp(-ray_c[None, None, :] * sum_r[..., None] - mie_e[None, None, :] * sum_m[..., None])
        i_r += a * dr[..., None]
        i_m += a * dm[..., None]
    mu = (d * sun).sum(-1)[..., None]
    res = sun_direction[3] * (1 + mu * mu) * (i_r * ray_c * 0.0597 + i_m * mie_s * 0.0196 / (1.58 - 1.52 * mu) ** 1.5)
    out = np.sqrt(res)
    out = np.where(np.isfinite(out).all(-1, keepdims=True), out, 0.0)
    return out ** 2.2
