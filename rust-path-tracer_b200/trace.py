"""Host-side mirror of the reference's trace driver (src/trace.rs) over the CUDA backend.

Same names and meaning as the Rust API so tests and benches read like the reference's:

    state = setup_trace(128, 128, 32)                    # src/trace.rs:331-344
    state.config.nee = 1                                 # NextEventEstimation::MultipleImportanceSampling
    trace_gpu("tests/golden/scenes/FurnaceTest.npz", None, state)   # src/trace.rs:136-224
    state.framebuffer                                    # packed RGB f32, output.xyz / samples

`trace_gpu` keeps the reference loop's structure — batches of `sync_rate` samples (ONE sample while
`interacting | dirty`, like the reference's early exit from its dispatch loop), `samples` advanced
per batch, readback + normalisation per batch, flush on `interacting | dirty` — but each batch is
ONE `rpt_enqueue(n)` (the device loops over the samples) instead of n x (dispatch + poll).  There is no `trace_cpu` here: the CPU path of the reference is
the oracle (oracle/), which is test infrastructure, not product.

`Renderer` is the thin object over the C ABI (include/rpt_b200.h) that `trace_gpu`, the tests
and bench.py share.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import capi
from .capi import TracingConfig
from .world import World, make_rng_seeds


def _tonemap_index(tonemap) -> int:
    return capi.TONEMAPS.index(tonemap) if isinstance(tonemap, str) else int(tonemap)


class Renderer:
    """One CUDA tracing context (one GPU).  Raises capi.RptError on any failure."""

    def __init__(self, device: int = 0, pipeline: int = capi.PIPELINE_WAVEFRONT):
        self._lib = capi.lib()
        self._ctx = C.c_void_p()
        code = self._lib.rpt_create(C.c_int(device), C.byref(self._ctx))
        capi.check(code, "rpt_create")
        self.device = device
        self._call("rpt_set_pipeline", C.c_int(pipeline))
        self.npixels = 0

    def _call(self, name, *args):
        capi.check(getattr(self._lib, name)(self._ctx, *args), name, self._ctx)

    def close(self):
        if self._ctx:
            self._lib.rpt_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene / state -------------------------------------------------------------------
    def refit_world(self, per_vertex_buffer: np.ndarray, light_pick_buffer: np.ndarray | None = None):
        """Vertices moved, topology unchanged: refit the tree on the device (rpt_refit_world)."""
        v = np.ascontiguousarray(per_vertex_buffer)
        l = None if light_pick_buffer is None else np.ascontiguousarray(light_pick_buffer)
        self._call("rpt_refit_world", capi.ptr(v), C.c_uint32(len(v)), capi.ptr(l), C.c_uint32(0 if l is None else len(l)))

    def upload_world(self, world: World, skybox: np.ndarray | None = None, build_on_device: bool = False):
        """skybox: (H, W, 4) float32 lat-long texels or None (2x2 magenta fallback).  build_on_device: do not hand over
        the reference BVH; the backend builds its tree on the GPU (world.nodes may then be None)."""
        atlas = None if world.atlas is None else np.ascontiguousarray(world.atlas, np.uint8)
        sky = None if skybox is None else np.ascontiguousarray(skybox, np.float32)
        u32 = C.c_uint32
        self._call(
            "rpt_upload_world",
            capi.ptr(world.per_vertex_buffer), u32(len(world.per_vertex_buffer)),
            capi.ptr(world.index_buffer), u32(len(world.index_buffer)),
            None if build_on_device else capi.ptr(world.nodes), u32(0 if build_on_device else len(world.nodes)),
            capi.ptr(world.material_data_buffer), u32(len(world.material_data_buffer)),
            capi.ptr(world.light_pick_buffer), u32(len(world.light_pick_buffer)),
            capi.ptr(atlas), u32(0 if atlas is None else atlas.shape[1]), u32(0 if atlas is None else atlas.shape[0]),
            capi.ptr(sky), u32(0 if sky is None else sky.shape[1]), u32(0 if sky is None else sky.shape[0]),
        )

    def set_config(self, config: TracingConfig):
        self._call("rpt_set_config", C.byref(config))
        self.npixels = config.width * config.height

    def set_pipeline(self, pipeline: int):
        self._call("rpt_set_pipeline", C.c_int(pipeline))

    def set_wave_slots(self, slots: int):
        self._call("rpt_set_wave_slots", C.c_uint32(slots))

    def set_tile_partition(self, rank: int, count: int):
        self._call("rpt_set_tile_partition", C.c_uint32(rank), C.c_uint32(count))

    def write_rng(self, seeds: np.ndarray):
        seeds = np.ascontiguousarray(seeds, np.uint32)
        self._call("rpt_write_rng", capi.ptr(seeds), C.c_size_t(seeds.size // 2))

    def read_rng(self) -> np.ndarray:
        out = np.empty((self.npixels, 2), np.uint32)
        self._call("rpt_read_rng", capi.ptr(out), C.c_size_t(self.npixels))
        return out

    def write_output(self, rgba: np.ndarray | None):
        if rgba is None:
            self._call("rpt_write_output", None, C.c_size_t(self.npixels))
        else:
            rgba = np.ascontiguousarray(rgba, np.float32)
            self._call("rpt_write_output", capi.ptr(rgba), C.c_size_t(rgba.size // 4))

    # ---- run / read ----------------------------------------------------------------------
    def enqueue(self, n_samples: int):
        self._call("rpt_enqueue", C.c_uint32(n_samples))

    def sync(self):
        self._call("rpt_sync")

    def enqueue_interruptible(self, n_samples: int, stop_flag: np.ndarray | None, poll_samples: int = 1) -> int:
        """rpt_enqueue that reads stop_flag[0] (a uint32 array another thread may write) after every `poll_samples`
        samples, like the reference's dispatch loop (src/trace.rs:182-193); returns the samples rendered."""
        finished = C.c_uint32(0)
        self._call("rpt_enqueue_interruptible", C.c_uint32(n_samples), capi.ptr(stop_flag), C.c_uint32(poll_samples), C.byref(finished))
        return finished.value

    def read_framebuffer_async(self, samples: float, out: np.ndarray):
        """`out`: page-locked (capi.pinned_empty); valid after readback_wait()."""
        self._call("rpt_read_framebuffer_async", capi.ptr(out), C.c_size_t(self.npixels), C.c_float(samples))

    def readback_wait(self):
        self._call("rpt_readback_wait")

    def set_frame_hook(self, hook):
        """hook(rgb_device_pointer, width, height, cuda_stream) or None — the reference's denoiser slot."""
        self._hook = capi.FRAME_HOOK(lambda p, w, h, s, _u: hook(p, w, h, s)) if hook else C.cast(None, capi.FRAME_HOOK)
        self._call("rpt_set_frame_hook", self._hook, None)

    def read_output(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.npixels, 4), np.float32)
        self._call("rpt_read_output", capi.ptr(out), C.c_size_t(self.npixels))
        return out

    def read_framebuffer(self, samples: float, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.npixels * 3, np.float32)
        self._call("rpt_read_framebuffer", capi.ptr(out), C.c_size_t(self.npixels), C.c_float(samples))
        return out

    def read_display(self, samples: float, tonemap: int | str = 0, out: np.ndarray | None = None) -> np.ndarray:
        """render.wgsl's fragment stage on the device: packed RGB f32 of tonemap(output.xyz / samples)."""
        if out is None:
            out = np.empty(self.npixels * 3, np.float32)
        self._call("rpt_read_display", capi.ptr(out), C.c_size_t(self.npixels), C.c_float(samples), C.c_uint32(_tonemap_index(tonemap)))
        return out

    def read_display_rgba8(self, samples: float, tonemap: int | str = 0, srgb: bool = True, out: np.ndarray | None = None) -> np.ndarray:
        """What save_render writes (src/app.rs:759-840): (npixels, 4) bytes R,G,B,255 of the tonemapped frame."""
        if out is None:
            out = np.empty((self.npixels, 4), np.uint8)
        self._call("rpt_read_display_rgba8", capi.ptr(out), C.c_size_t(self.npixels), C.c_float(samples), C.c_uint32(_tonemap_index(tonemap)),
                   C.c_uint32(1 if srgb else 0))
        return out

    def read_primary_ids(self) -> np.ndarray:
        out = np.empty(self.npixels, np.uint32)
        self._call("rpt_read_primary_ids", capi.ptr(out), C.c_size_t(self.npixels))
        return out

    def counters(self) -> dict:
        c = capi.Counters()
        self._call("rpt_get_counters", C.byref(c))
        return {"paths": c.paths, "nearest_rays": c.nearest_rays, "any_rays": c.any_rays, "kernel_launches": c.kernel_launches}

    def reset_counters(self):
        self._call("rpt_reset_counters")

    def device_ms(self) -> float:
        ms = C.c_float(0)
        self._call("rpt_get_device_ms", C.byref(ms))
        return ms.value

    def timer_start(self):
        self._call("rpt_timer_start")

    def timer_stop(self) -> float:
        """Device milliseconds since timer_start (enqueues and the combine); waits for that work."""
        ms = C.c_float(0)
        self._call("rpt_timer_stop", C.byref(ms))
        return ms.value

    def sm_count(self) -> int:
        n = C.c_int(0)
        self._call("rpt_get_sm_count", C.byref(n))
        return n.value

    def set_trace_statistics(self, enable: bool):
        self._call("rpt_set_trace_statistics", C.c_int(int(enable)))

    def trace_statistics(self) -> dict:
        """Node visits / triangle tests counted by the diagnostic build of the trace kernels since reset_counters."""
        t = capi.TraceStatistics()
        self._call("rpt_get_trace_statistics", C.byref(t))
        return {name: getattr(t, name) for name, _ in capi.TraceStatistics._fields_}

    def set_stage_timing(self, enable: bool):
        self._call("rpt_set_stage_timing", C.c_int(int(enable)))

    def stage_timing(self) -> dict:
        """{stage: (summed launch ms, launches)} since the last call (syncs)."""
        t = capi.StageTiming()
        self._call("rpt_get_stage_timing", C.byref(t))
        return {name: (t.ms[i], t.launches[i]) for i, name in enumerate(capi.STAGES)}

    # ---- multi-GPU -----------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        capi.check(capi.lib().rpt_comm_unique_id(buf), "rpt_comm_unique_id")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._call("rpt_comm_init", buf, C.c_int(rank), C.c_int(nranks))

    def comm_reduce_output(self, root: int = 0):
        self._call("rpt_comm_reduce_output", C.c_int(root))

    def comm_destroy(self):
        self._call("rpt_comm_destroy")


class TracingState:
    """src/trace.rs:40-92.  Atomics become attributes; `running`, `interacting` and `dirty` also keep one word of
    shared memory up to date (`stop_flag`: interacting | dirty | !running) that `rpt_enqueue_interruptible` polls from
    inside a batch, the way the reference's dispatch loop reads its atomics after every sample."""

    def __init__(self, width: int, height: int):
        self.stop_flag = np.zeros(1, np.uint32)
        self._running = self._interacting = self._dirty = False
        self.config = TracingConfig.default(width, height)
        self.framebuffer = np.zeros(width * height * 3, np.float32)
        self.running = False
        self.samples = 0
        self.denoise = False
        self.sync_rate = 32
        self.use_blue_noise = True
        self.interacting = False
        self.dirty = False
        self.poll_samples = 1  # samples between two looks at the control flags inside a batch (the reference: 1)
        self._stop_at = None  # set by setup_trace: the watcher thread's threshold
        self.lock = threading.Lock()

    def _refresh(self):
        self.stop_flag[0] = 1 if (self._interacting or self._dirty or not self._running) else 0

    running = property(lambda self: self._running, lambda self, v: (setattr(self, "_running", bool(v)), self._refresh())[1])
    interacting = property(lambda self: self._interacting, lambda self, v: (setattr(self, "_interacting", bool(v)), self._refresh())[1])
    dirty = property(lambda self: self._dirty, lambda self, v: (setattr(self, "_dirty", bool(v)), self._refresh())[1])

    def _watch(self):
        # the reference spawns a thread that clears `running` once `samples >= N`
        # (src/trace.rs:335-341); evaluated inline here, i.e. an infinitely fast watcher
        if self._stop_at is not None and self.samples >= self._stop_at:
            self.running = False


def setup_trace(width: int, height: int, samples: int) -> TracingState:
    """Harness for synchronous tracing, src/trace.rs:331-344."""
    state = TracingState(width, height)
    state.running = True
    state._stop_at = samples
    state._watch()  # samples == 0 requested: the watcher stops the loop at once ("startup" benches)
    return state


def decode_hdr(data: bytes) -> np.ndarray | None:
    """Radiance .hdr bytes -> (H, W, 3) float32 (csrc/image_io.cpp, `HdrDecoder` semantics); None if malformed."""
    lib = capi.lib()
    buf = np.frombuffer(data, np.uint8)
    w, h = C.c_uint32(0), C.c_uint32(0)
    if lib.rpt_decode_hdr(capi.ptr(buf), C.c_size_t(len(buf)), None, C.byref(w), C.byref(h)) != capi.OK:
        return None
    out = np.empty((h.value, w.value, 3), np.float32)
    if lib.rpt_decode_hdr(capi.ptr(buf), C.c_size_t(len(buf)), capi.ptr(out), C.byref(w), C.byref(h)) != capi.OK:
        return None
    return out


def load_skybox(path: str | None, cpu_path_rgb8: bool = False) -> np.ndarray | None:
    """`load_dynamic_image` + the texel conversion of the sky image (src/asset.rs:238-273): a Radiance .hdr file
    (decoded natively), a .npy array (H,W,3|4 float32) or any PIL-readable 8-bit image.  Returns (H,W,4) float32 texels —
    as the GPU path uploads them (Rgba32Float, `dynamic_image_to_gpu_image`), or with `cpu_path_rgb8` as the CPU path
    reads them (`dynamic_image_to_cpu_buffer`: clamped to [0,1] and quantised to 8 bits) — or None on failure."""
    if path is None:
        return None
    try:
        if path.endswith(".hdr"):
            with open(path, "rb") as f:
                img = decode_hdr(f.read())
            if img is None:
                return None
        elif path.endswith(".npy"):
            img = np.load(path).astype(np.float32)
        else:
            from PIL import Image

            img = np.asarray(Image.open(path).convert("RGB"), np.float32) / np.float32(255.0)
    except (OSError, ValueError):
        return None
    if img.ndim != 3 or img.shape[2] not in (3, 4):
        return None
    rgb = np.ascontiguousarray(img[..., :3], np.float32)
    out = np.empty(rgb.shape[:2] + (4,), np.float32)
    capi.check(capi.lib().rpt_sky_texels(capi.ptr(rgb), C.c_uint32(rgb.shape[1]), C.c_uint32(rgb.shape[0]), C.c_int(int(cpu_path_rgb8)), capi.ptr(out)),
               "rpt_sky_texels")
    return out


def trace_gpu(scene_path: str, skybox_path: str | None, state: TracingState, device: int = 0,
              pipeline: int = capi.PIPELINE_WAVEFRONT, world: World | None = None) -> None:
    """src/trace.rs:136-224 with the wgpu dispatch replaced by the CUDA backend."""
    if world is None:
        world = World.from_path(scene_path)
        if world is None:
            return  # `let Some(world) = ... else { return; }`
    skybox = load_skybox(skybox_path)
    width, height = state.config.width, state.config.height
    seeds = make_rng_seeds(width, height, use_blue_noise=state.use_blue_noise)

    with Renderer(device, pipeline) as r:
        # the buffers that cross the boundary every batch live in page-locked memory (rpt_host_alloc)
        pinned_seeds = capi.pinned_empty(seeds.shape, np.uint32)
        pinned_seeds[...] = seeds
        seeds = pinned_seeds
        framebuffer = capi.pinned_empty(state.framebuffer.shape, np.float32)
        framebuffer[...] = state.framebuffer
        state.framebuffer = framebuffer
        r.upload_world(world, skybox)
        r.set_config(state.config)
        r.write_rng(seeds)
        # restore previous state ("continue previous", src/trace.rs:163-164)
        if state.samples > 0:
            init = np.concatenate([state.framebuffer.reshape(-1, 3), np.ones((width * height, 1), np.float32)], axis=1)
            r.write_output(init * np.float32(state.samples))

        while state.running:
            # The reference leaves its dispatch loop as soon as a sample ends with `interacting | dirty` set, and
            # abandons the batch when `running` is cleared (src/trace.rs:182-193).  A synchronous harness (setup_trace)
            # has nobody to set the flags mid-batch, so it takes the whole batch in one uninterrupted enqueue.
            if state._stop_at is not None and not (state.interacting or state.dirty):
                r.enqueue(state.sync_rate)
                r.sync()
                finished = state.sync_rate
            else:
                finished = r.enqueue_interruptible(state.sync_rate, state.stop_flag, state.poll_samples)
                if not state.running:
                    return
            flush = state.interacting or state.dirty
            state.samples += finished
            state._watch()
            with state.lock:
                r.read_framebuffer(float(state.samples), state.framebuffer)
            if flush:  # src/trace.rs:216-222
                state.dirty = False
                state.samples = 0
                if (state.config.width, state.config.height) != (width, height):
                    # The reference sizes every buffer once, before the loop, and would index out of bounds after a
                    # resize; here a resized config gets fresh seeds and a fresh framebuffer.
                    width, height = state.config.width, state.config.height
                    fresh = make_rng_seeds(width, height, use_blue_noise=state.use_blue_noise)
                    seeds = capi.pinned_empty(fresh.shape, np.uint32)
                    seeds[...] = fresh
                    with state.lock:
                        state.framebuffer = capi.pinned_empty(width * height * 3, np.float32)
                        state.framebuffer[...] = 0.0
                r.set_config(state.config)
                r.write_output(None)
                r.write_rng(seeds)
