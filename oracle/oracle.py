"""ctypes wrapper of the CPU oracle (oracle.cpp).  TEST INFRASTRUCTURE ONLY — see oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import
this module.  It takes the same `World` buffers and `TracingConfig` the CUDA backend takes, so
both sides see identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "_build", "liboracle.so")
_lib = None


class _World(C.Structure):
    _fields_ = [
        ("verts", C.c_void_p), ("nverts", C.c_uint32), ("tris", C.c_void_p), ("ntris", C.c_uint32),
        ("nodes", C.c_void_p), ("nnodes", C.c_uint32), ("mats", C.c_void_p), ("nmats", C.c_uint32),
        ("lights", C.c_void_p), ("nlights", C.c_uint32), ("atlas", C.c_void_p), ("atlas_w", C.c_uint32),
        ("atlas_h", C.c_uint32), ("sky", C.c_void_p), ("sky_w", C.c_uint32), ("sky_h", C.c_uint32),
        ("atlas_f32", C.c_void_p),
    ]


class _Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "nearest_rays", "any_rays", "nodes_popped", "boxes_tested", "tris_tested",
                                           "stack_overflows", "light_index_clamped", "rng_exhausted", "boxes_tested_any", "tris_tested_any")]


def build() -> str:
    src = os.path.join(_DIR, "oracle.cpp")
    if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        p = subprocess.run(["make", "-C", _DIR], capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + p.stdout + p.stderr)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p).value


class OracleScene:
    """Pins the numpy buffers of a `World` (+ optional sky texels) for oracle calls."""

    def __init__(self, world, skybox: np.ndarray | None = None):
        self.keep = [
            np.ascontiguousarray(world.per_vertex_buffer), np.ascontiguousarray(world.index_buffer, np.uint32),
            np.ascontiguousarray(world.nodes), np.ascontiguousarray(world.material_data_buffer),
            np.ascontiguousarray(world.light_pick_buffer),
            None if world.atlas is None else np.ascontiguousarray(world.atlas, np.uint8),
            None if skybox is None else np.ascontiguousarray(skybox, np.float32),
        ]
        v, t, n, m, l, a, s = self.keep
        # the CPU path converts the atlas to float texels once, before its sample loop (src/trace.rs:268-271)
        self.atlas_f32 = None
        if a is not None:
            self.atlas_f32 = np.empty(a.shape[:2] + (4,), np.float32)
            if lib().oracle_convert_atlas(C.c_void_p(_ptr(a)), C.c_uint32(a.shape[1]), C.c_uint32(a.shape[0]), C.c_void_p(_ptr(self.atlas_f32))) != 0:
                raise RuntimeError("oracle_convert_atlas failed")
        self.c = _World(_ptr(v), len(v), _ptr(t), len(t), _ptr(n), len(n), _ptr(m), len(m), _ptr(l), len(l),
                        _ptr(a), 0 if a is None else a.shape[1], 0 if a is None else a.shape[0],
                        _ptr(s), 0 if s is None else s.shape[1], 0 if s is None else s.shape[0], _ptr(self.atlas_f32))


def trace(config, scene: OracleScene, seeds: np.ndarray, n_samples: int, output: np.ndarray | None = None, threads: int = 0,
          want_primary_ids: bool = False):
    """n passes of the trace_cpu loop body.  Returns (output[N,4] running sum, seeds advanced, counters dict, primary ids or None)."""
    npix = config.width * config.height
    seeds = np.ascontiguousarray(seeds, np.uint32).copy().reshape(npix, 2)
    out = np.zeros((npix, 4), np.float32) if output is None else np.ascontiguousarray(output, np.float32).copy()
    ids = np.zeros(npix, np.uint32) if want_primary_ids else None
    ctr = _Counters()
    rc = lib().oracle_trace(C.byref(config), C.byref(scene.c), C.c_void_p(_ptr(seeds)), C.c_void_p(_ptr(out)), C.c_uint32(n_samples),
                            C.c_int(threads), C.c_void_p(_ptr(ids)), C.byref(ctr))
    if rc != 0:
        raise RuntimeError(f"oracle_trace failed: {rc}")
    counters = {n: getattr(ctr, n) for n, _ in _Counters._fields_}
    if counters["stack_overflows"] or counters["rng_exhausted"]:
        raise RuntimeError(f"oracle hit a condition where the reference panics: {counters}")
    return out, seeds, counters, ids


def set_retire_dead_paths(on: bool):
    """Test switch (off = the reference): stop paths whose throughput is exactly zero, like the CUDA backend does."""
    lib().oracle_set_retire_dead_paths(C.c_int(int(on)))


def intersect(scene: OracleScene, rays: np.ndarray, any_hit: bool = False, max_t: np.ndarray | None = None):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = len(rays)
    max_t = np.zeros(n, np.float32) if max_t is None else np.ascontiguousarray(max_t, np.float32)
    hit, tri, t, back = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
    rc = lib().oracle_intersect(C.byref(scene.c), C.c_void_p(_ptr(rays)), C.c_uint32(n), C.c_int(int(any_hit)), C.c_void_p(_ptr(max_t)),
                                C.c_void_p(_ptr(hit)), C.c_void_p(_ptr(tri)), C.c_void_p(_ptr(t)), C.c_void_p(_ptr(back)))
    if rc != 0:
        raise RuntimeError(f"oracle_intersect failed: {rc}")
    return hit, tri, t, back


def camera_rays(config, seeds: np.ndarray) -> np.ndarray:
    npix = config.width * config.height
    seeds = np.ascontiguousarray(seeds, np.uint32)
    rays = np.zeros((npix, 6), np.float32)
    lib().oracle_camera_rays(C.byref(config), C.c_void_p(_ptr(seeds)), C.c_void_p(_ptr(rays)))
    return rays
