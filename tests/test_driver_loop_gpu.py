"""The driver-loop side of the boundary: batches that can be cut short like the reference's
dispatch loop, readback that does not stall the next batch, the denoiser slot, and an .hdr sky end to end."""
import ctypes as C
import threading
import time

import numpy as np
import pytest

import helpers
from rust_path_tracer_b200 import capi
from rust_path_tracer_b200.trace import Renderer, TracingState, load_skybox, setup_trace, trace_gpu

pytestmark = pytest.mark.gpu


def _renderer(w=96, h=64, nee=1, scene="DarkCornell"):
    r = Renderer(0)
    r.upload_world(helpers.world(scene))
    r.set_config(helpers.config(w, h, nee))
    r.write_rng(helpers.seeds(w, h))
    return r


def test_interruptible_enqueue_matches_plain_enqueue_and_stops_on_the_flag():
    flag = np.zeros(1, np.uint32)
    with _renderer() as r:
        r.enqueue(12)
        whole = r.read_output()
        r.write_rng(helpers.seeds(96, 64)); r.write_output(None)
        assert r.enqueue_interruptible(12, flag, 5) == 12  # groups of 5, 5, 2
        np.testing.assert_array_equal(r.read_output(), whole)
        # a set flag ends the batch after the first group, like `flush |= interacting || dirty; if flush { break }`
        r.write_rng(helpers.seeds(96, 64)); r.write_output(None)
        flag[0] = 1
        assert r.enqueue_interruptible(12, flag, 1) == 1
        assert (r.read_output()[:, 3] == 1.0).all()
        assert r.enqueue_interruptible(12, None, 4) == 12  # no flag: never stops


def test_a_moving_camera_gets_one_sample_per_frame():
    """trace_gpu with `interacting` set: every pass renders ONE sample, reads it back and restarts (src/trace.rs:182-222)."""
    state = TracingState(64, 48)
    state.config.nee = 1
    state.running = True
    state.interacting = True
    seen = []

    def ui():  # the UI thread: let a few frames go by, then stop interacting, then stop the render
        deadline = time.time() + 20
        while state.samples == 0 and not seen and time.time() < deadline:
            time.sleep(0.001)
        time.sleep(0.05)
        seen.append(state.samples)
        state.interacting = False
        while state.samples < 64 and time.time() < deadline:
            time.sleep(0.001)
        state.running = False

    t = threading.Thread(target=ui)
    t.start()
    trace_gpu("", None, state, world=helpers.world("DarkCornell"))
    t.join()
    assert seen and seen[0] <= 1  # while interacting the count is reset after every single-sample frame
    assert state.samples >= 32 and np.isfinite(state.framebuffer).all()


def test_async_readback_equals_blocking_readback():
    with _renderer(128, 96) as r:
        a = capi.pinned_empty(128 * 96 * 3, np.float32)
        b = capi.pinned_empty(128 * 96 * 3, np.float32)
        c = capi.pinned_empty(128 * 96 * 3, np.float32)
        r.enqueue(4)
        r.read_framebuffer_async(4.0, a)
        r.enqueue(4)  # overlaps the copy of `a`
        r.read_framebuffer_async(8.0, b)
        r.enqueue(4)
        r.read_framebuffer_async(12.0, c)  # third frame: the first buffer is reused once its copy has landed
        r.readback_wait()
        after12 = r.read_framebuffer(12.0)
        np.testing.assert_array_equal(c, after12)
        r.write_rng(helpers.seeds(128, 96)); r.write_output(None)
        r.enqueue(4)
        np.testing.assert_array_equal(a, r.read_framebuffer(4.0))
        r.enqueue(4)
        np.testing.assert_array_equal(b, r.read_framebuffer(8.0))


def test_frame_hook_sees_the_normalised_frame_on_the_device():
    """The denoiser slot (src/trace.rs:207-210): the hook gets the device pointer of the normalised frame and the stream;
    here it overwrites the first row on that stream, and the readback must show it."""
    cudart = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            cudart = C.CDLL(name)
            break
        except OSError:
            continue
    if cudart is None:
        import glob
        import os

        import torch

        hits = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        if not hits:
            pytest.skip("no libcudart to call from the hook")
        cudart = C.CDLL(hits[0])
    cudart.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
    calls = []

    def hook(ptr, w, h, stream):
        calls.append((w, h))
        assert cudart.cudaMemsetAsync(ptr, 0, w * 3 * 4, stream) == 0  # first row -> 0.0

    w, h = 96, 64
    with _renderer(w, h) as r:
        r.enqueue(8)
        plain = r.read_framebuffer(8.0).reshape(h, w, 3)
        r.set_frame_hook(hook)
        hooked = r.read_framebuffer(8.0).reshape(h, w, 3)
        pinned = capi.pinned_empty(w * h * 3, np.float32)
        r.read_framebuffer_async(8.0, pinned)
        r.readback_wait()
        r.set_frame_hook(None)
        again = r.read_framebuffer(8.0).reshape(h, w, 3)
    assert calls == [(w, h), (w, h)]
    assert (hooked[0] == 0).all() and (plain[0] != 0).any()
    np.testing.assert_array_equal(hooked[1:], plain[1:])
    np.testing.assert_array_equal(pinned.reshape(h, w, 3), hooked)
    np.testing.assert_array_equal(again, plain)


def test_hdr_file_as_skybox_end_to_end(tmp_path):
    """An .hdr file written here, loaded by load_skybox in both texel conventions, rendered on the GPU and by the oracle."""
    import oracle as om
    from test_image_io import header, rle_scanline, to_rgbe

    sky = helpers.synthetic_sky(64, 32)[..., :3].astype(np.float64)
    rgbe = to_rgbe(sky)
    path = tmp_path / "sky.hdr"
    path.write_bytes(header(64, 32) + b"".join(rle_scanline(rgbe[y]) for y in range(32)))
    world = helpers.world("PBRTest")
    cfg = helpers.config(128, 72, 0, has_skybox=1)
    seeds = helpers.seeds(128, 72)
    for cpu_path in (False, True):
        texels = load_skybox(str(path), cpu_path_rgb8=cpu_path)
        assert texels is not None and (texels[..., :3].max() > 1.0) != cpu_path  # the CPU path's 8-bit sky has no range above 1
        ref, _, _, _ = om.trace(cfg, om.OracleScene(world, texels), seeds, 8)
        with Renderer(0) as r:
            r.upload_world(world, texels); r.set_config(cfg); r.write_rng(seeds)
            r.enqueue(8)
            out = r.read_output()
        assert helpers.mae(out[:, :3] / 8, ref[:, :3] / 8)[0] <= 1e-3
