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
print("ok", np.isfinite(a).all(), np.isfinite(b[:, 3]).all())
