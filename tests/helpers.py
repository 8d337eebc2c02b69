"""Shared helpers for the parity tests: fixture scenes, configs, comparison metrics."""
from __future__ import annotations

import functools
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE_DIR = os.path.join(REPO, "tests", "golden", "scenes")
GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
SCENES = ["FurnaceTest", "DarkCornell", "PBRTest", "VeachMIS"]


@functools.lru_cache(maxsize=None)
def world(name: str):
    from rust_path_tracer_b200.world import World

    w = World.from_path(os.path.join(SCENE_DIR, name + ".npz"))
    assert w is not None, name
    return w


def config(width: int, height: int, nee: int = 0, **kw):
    from rust_path_tracer_b200.capi import TracingConfig

    cfg = TracingConfig.default(width, height)
    cfg.nee = nee
    for k, v in kw.items():
        if isinstance(v, (list, tuple)):
            getattr(cfg, k)[:] = v
        else:
            setattr(cfg, k, v)
    return cfg


def seeds(width: int, height: int):
    from rust_path_tracer_b200.world import make_rng_seeds

    return make_rng_seeds(width, height, use_blue_noise=True)


def mae(a: np.ndarray, b: np.ndarray):
    """Mean absolute error over pixels and channels of two normalised linear RGB images; NaN pixels
    are counted separately and excluded."""
    a = np.asarray(a, np.float64).reshape(-1, 3)
    b = np.asarray(b, np.float64).reshape(-1, 3)
    bad = ~(np.isfinite(a).all(axis=1) & np.isfinite(b).all(axis=1))
    return float(np.abs(a[~bad] - b[~bad]).mean()), int(bad.sum())


def record_parity(case: str, **measured):
    """Append one measured parity line (id mismatch fraction, MAE, ...) to gpurun_out/parity_report.jsonl, so the
    numbers quoted in DESIGN.md / profiles/ come from the test run itself.  Best effort: never fails a test."""
    import json

    try:
        out_dir = os.path.join(REPO, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"case": case, **measured}) + "\n")
    except OSError:
        pass


def synthetic_sky(width: int = 64, height: int = 32, seed: int = 7) -> np.ndarray:
    """Small deterministic lat-long HDR image (float4) with a bright disk."""
    rs = np.random.default_rng(seed)
    img = np.zeros((height, width, 4), np.float32)
    v = np.linspace(0, 1, height, dtype=np.float32)[:, None]
    img[..., 0] = 0.3 + 0.5 * v
    img[..., 1] = 0.4 + 0.4 * v
    img[..., 2] = 0.9 - 0.3 * v
    img[..., :3] += rs.random((height, width, 3), dtype=np.float32) * 0.05
    img[height // 4 : height // 4 + 3, width // 3 : width // 3 + 3, :3] = 40.0
    img[..., 3] = 1.0
    return img


@functools.lru_cache(maxsize=None)
def textured_world(texture_size: int = 64, atlas_size: int = 1024):
    """PBRTest with procedural albedo / metallic / roughness / normal textures on its 25 sphere materials
    (the shipped scenes contain no image)."""
    from rust_path_tracer_b200.glb import BakedScene
    from rust_path_tracer_b200.scenes import textured_pbr_variant
    from rust_path_tracer_b200.world import World

    scene, atlas = textured_pbr_variant(BakedScene.load(os.path.join(SCENE_DIR, "PBRTest.npz")), texture_size, atlas_size)
    return World.from_baked(scene, atlas=atlas)


@functools.lru_cache(maxsize=None)
def proxy_world(triangles: int = 60000):
    """Small instance of the labelled BreakTime proxy (textured, emitters, windows)."""
    from rust_path_tracer_b200.scenes import breaktime_proxy
    from rust_path_tracer_b200.world import World

    scene, atlas = breaktime_proxy(triangles, texture_size=64, atlas_size=512)
    return World.from_baked(scene, atlas=atlas)
