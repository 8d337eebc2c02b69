"""Scratch: a small render of every pipeline feature (textured proxy, MIS, HDR sky, tile partition) for compute-sanitizer."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import helpers
from rust_path_tracer_b200.trace import Renderer

w, h = 48, 32
with Renderer(0) as r:
    r.upload_world(helpers.proxy_world(), helpers.synthetic_sky(64, 32))
    r.set_config(helpers.config(w, h, 1, has_skybox=1)); r.write_rng(helpers.seeds(w, h))
    r.read_primary_ids(); r.enqueue(2); r.enqueue(2)
    r.set_tile_partition(1, 2); r.write_rng(helpers.seeds(w, h)); r.enqueue(1); r.set_tile_partition(0, 1)
    a = r.read_framebuffer(5.0)
    r.upload_world(helpers.world("PBRTest"))
    r.set_config(helpers.config(w, h, 0)); r.write_rng(helpers.seeds(w, h)); r.enqueue(2)
    b = r.read_output()
# round 2: tree built and refitted on the device, asynchronous readback, interruptible batches, traversal statistics
from rust_path_tracer_b200 import capi
world = helpers.world("DarkCornell")
with Renderer(0) as r:
    r.upload_world(world, build_on_device=True)
    r.set_config(helpers.config(w, h, 1)); r.write_rng(helpers.seeds(w, h))
    r.enqueue(2)
    moved = world.per_vertex_buffer.copy(); moved["vertex"][:, :3] *= np.float32(1.25)
    r.refit_world(moved)
    r.enqueue(2)
    pinned = capi.pinned_empty(w * h * 3, np.float32)
    r.read_framebuffer_async(4.0, pinned); r.enqueue(1); r.readback_wait()
    flag = np.zeros(1, np.uint32)
    r.enqueue_interruptible(3, flag, 1)
    r.set_trace_statistics(True); r.enqueue(1); st = r.trace_statistics(); r.set_trace_statistics(False)
    c = r.read_output()
with Renderer(0) as r:  # host-collapsed tree, then refit
    r.upload_world(world); r.set_config(helpers.config(w, h, 1)); r.write_rng(helpers.seeds(w, h))
    r.enqueue(1); r.refit_world(moved); r.enqueue(1)
    d = r.read_output()
print("ok", np.isfinite(a).all(), np.isfinite(b[:, 3]).all(), np.isfinite(pinned).all(), st["nearest_node_visits"] > 0, np.isfinite(c[:, 3]).all(), np.isfinite(d[:, 3]).all())
