"""ctypes binding of the C ABI declared in include/rpt_b200.h, include/rpt_host.h.

The library is the product: if it cannot be loaded there is nothing to fall back to, so the
loader raises.  (The shared object is built in-tree by `build.py`; `lib()` builds it on first
use if it is missing and nvcc exists.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

OK = 0
ERR_INVALID_ARGUMENT, ERR_NO_DEVICE, ERR_CUDA, ERR_NOT_READY = -1, -2, -3, -4
ERR_RNG_DIMENSIONS, ERR_SIZE_MISMATCH, ERR_NCCL, ERR_UNSUPPORTED = -5, -6, -7, -8
PIPELINE_WAVEFRONT, PIPELINE_MEGAKERNEL = 0, 1
# `Tonemapping` (src/app.rs:20-28)
TONEMAPS = ["none", "reinhard", "aces_narkowicz", "aces_narkowicz_overexposed", "aces_hill", "neutral", "uncharted"]

# every symbol the headers declare (tests check the library exports each of them)
HOST_SYMBOLS = ["rpt_build_bvh", "rpt_build_light_pick_table", "rpt_pack_per_vertex", "rpt_make_rng_seeds", "rpt_camera_matrix",
                "rpt_tile_partition_pixels", "rpt_atlas_rects", "rpt_atlas_pack", "rpt_decode_albedo_gamma", "rpt_decode_hdr", "rpt_sky_texels"]
DEVICE_SYMBOLS = [
    "rpt_create", "rpt_destroy", "rpt_last_error", "rpt_set_pipeline", "rpt_set_wave_slots", "rpt_upload_world", "rpt_refit_world",
    "rpt_set_config", "rpt_write_rng", "rpt_read_rng", "rpt_write_output", "rpt_set_tile_partition", "rpt_enqueue",
    "rpt_sync", "rpt_enqueue_interruptible", "rpt_read_output", "rpt_read_framebuffer", "rpt_read_framebuffer_async", "rpt_readback_wait", "rpt_set_frame_hook", "rpt_read_display", "rpt_read_display_rgba8", "rpt_read_primary_ids", "rpt_get_counters",
    "rpt_reset_counters", "rpt_get_device_ms", "rpt_set_stage_timing", "rpt_get_stage_timing", "rpt_set_trace_statistics", "rpt_get_trace_statistics", "rpt_get_sm_count", "rpt_timer_start", "rpt_timer_stop", "rpt_comm_unique_id", "rpt_comm_init", "rpt_comm_reduce_output",
    "rpt_comm_destroy", "rpt_host_alloc", "rpt_host_free",
]


class TracingConfig(C.Structure):
    """shared_structs/src/lib.rs:12-42 (layout + Default)."""

    _fields_ = [
        ("cam_position", C.c_float * 4),
        ("cam_rotation", C.c_float * 4),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("min_bounces", C.c_uint32),
        ("max_bounces", C.c_uint32),
        ("sun_direction", C.c_float * 4),
        ("nee", C.c_uint32),
        ("has_skybox", C.c_uint32),
        ("specular_weight_clamp", C.c_float * 2),
    ]

    @staticmethod
    def default(width: int = 1280, height: int = 720) -> "TracingConfig":
        f = np.float32
        sun = np.array([0.5, 1.3, 1.0], f)
        inv = f(1.0) / np.sqrt(f(f(sun[0] * sun[0] + sun[1] * sun[1]) + sun[2] * sun[2]))  # glam normalize
        sun = sun * inv
        cfg = TracingConfig()
        cfg.cam_position[:] = [0.0, 1.0, -5.0, 0.0]
        cfg.cam_rotation[:] = [0.0, 0.0, 0.0, 0.0]
        cfg.width, cfg.height, cfg.min_bounces, cfg.max_bounces = width, height, 3, 4
        cfg.sun_direction[:] = [float(sun[0]), float(sun[1]), float(sun[2]), 15.0]
        cfg.nee, cfg.has_skybox = 0, 0
        cfg.specular_weight_clamp[:] = [0.1, 0.9]
        return cfg

    def copy(self) -> "TracingConfig":
        out = TracingConfig()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(TracingConfig))
        return out


assert C.sizeof(TracingConfig) == 80


class Counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("nearest_rays", C.c_uint64), ("any_rays", C.c_uint64), ("kernel_launches", C.c_uint64)]


class TraceStatistics(C.Structure):
    _fields_ = [("nearest_rays", C.c_uint64), ("nearest_node_visits", C.c_uint64), ("nearest_triangle_tests", C.c_uint64),
                ("any_rays", C.c_uint64), ("any_node_visits", C.c_uint64), ("any_triangle_tests", C.c_uint64), ("shaded_hits", C.c_uint64),
                ("node_bytes", C.c_uint32), ("triangle_bytes", C.c_uint32)]


STAGES = ["generate", "extend", "miss", "shade", "shadow", "accumulate", "megakernel", "other"]


class StageTiming(C.Structure):
    _fields_ = [("ms", C.c_float * 8), ("launches", C.c_uint64 * 8)]


FRAME_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p)  # rpt_frame_hook

BVH_NODE_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("triangle_count", "<u4"), ("aabb_max", "<f4", 3), ("left_or_first", "<u4")])
LIGHT_DTYPE = np.dtype(
    [("triangle_index_a", "<u4"), ("triangle_area_a", "<f4"), ("triangle_pick_pdf_a", "<f4"), ("triangle_index_b", "<u4"),
     ("triangle_area_b", "<f4"), ("triangle_pick_pdf_b", "<f4"), ("ratio", "<f4")]
)
assert BVH_NODE_DTYPE.itemsize == 32 and LIGHT_DTYPE.itemsize == 28

_lib = None


class RptError(RuntimeError):
    def __init__(self, code: int, where: str, message: str = ""):
        self.code = code
        super().__init__(f"{where} failed with status {code}" + (f": {message}" if message else ""))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("RPT_B200_LIBRARY") or _build.LIB_PATH  # (override: A/B runs of differently built libraries)
        if not os.path.exists(path):
            _build.build_library()
        _lib = C.CDLL(path)
        _lib.rpt_last_error.restype = C.c_char_p
        _lib.rpt_last_error.argtypes = [C.c_void_p]
        for name in HOST_SYMBOLS + DEVICE_SYMBOLS:
            fn = getattr(_lib, name)
            if name != "rpt_last_error":
                fn.restype = C.c_int
    return _lib


class _PinnedBlock:
    """Owner of one rpt_host_alloc block.  It exposes the block through `__array_interface__`, so the numpy array made
    from it — and, through numpy's base chain, every view, slice or reshape of that array — holds a reference to this
    object; the page-locked memory is released when the last of them is collected."""

    def __init__(self, nbytes: int):
        self.address = C.c_void_p()
        check(lib().rpt_host_alloc(C.c_size_t(nbytes), C.byref(self.address)), "rpt_host_alloc")
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.address.value, False), "version": 3}

    def __del__(self):
        try:
            if self.address:
                lib().rpt_host_free(self.address)
                self.address = C.c_void_p()
        except Exception:
            pass


def _array_over(owner, shape, dtype) -> np.ndarray:
    """ndarray of `dtype` / `shape` over the bytes `owner` exposes; `owner` stays alive while any view of it does."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    return np.asarray(owner)[: count * dtype.itemsize].view(dtype).reshape(shape)


def pinned_empty(shape, dtype) -> np.ndarray:
    """An uninitialised numpy array in page-locked host memory (rpt_host_alloc); needs a CUDA device."""
    dtype = np.dtype(dtype)
    return _array_over(_PinnedBlock(max(1, int(np.prod(shape)) * dtype.itemsize)), shape, dtype)


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(code: int, where: str, ctx=None):
    if code != OK:
        msg = lib().rpt_last_error(ctx)
        raise RptError(code, where, msg.decode() if msg else "")
