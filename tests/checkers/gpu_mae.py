"""Scratch: MAE of the wavefront pipeline against the oracle on the HDR stress cases (run on the GPU box)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np
import helpers
import oracle as oracle_mod
from rust_path_tracer_b200 import capi
from rust_path_tracer_b200.trace import Renderer

def run(name, world, cfg, seeds, spp, sky=None):
    scene = oracle_mod.OracleScene(world, sky)
    ref, *_ = oracle_mod.trace(cfg, scene, seeds, spp)
    with Renderer(0) as r:
        r.upload_world(world, sky); r.set_config(cfg); r.write_rng(seeds); r.enqueue(spp); out = r.read_output()
    err, bad = helpers.mae(out[:, :3] / spp, ref[:, :3] / spp)
    d = np.abs(out[:, :3] - ref[:, :3]).max(axis=1) / spp
    print(f"{name}: MAE {err:.3e}  pixels differing by > 1e-3: {(d > 1e-3).mean():.2e}  > 1e-5: {(d > 1e-5).mean():.2e}", flush=True)

w, h = 320, 180
run("proxy 60k MIS HDR sky 16spp", helpers.proxy_world(), helpers.config(w, h, 1, has_skybox=1), helpers.seeds(w, h), 16, helpers.synthetic_sky(128, 64))
run("PBRTest HDR sky rotated 16spp", helpers.world("PBRTest"), helpers.config(w, h, 0, has_skybox=1, cam_rotation=[0.1, 0.4, 0.0, 0.0], cam_position=[1.0, 2.0, -6.0, 0.0]), helpers.seeds(w, h), 16, helpers.synthetic_sky())
run("VeachMIS MIS 32spp", helpers.world("VeachMIS"), helpers.config(w, h, 1), helpers.seeds(w, h), 32)
run("PBRTest textured 16spp", helpers.textured_world(), helpers.config(w, h, 0), helpers.seeds(w, h), 16)
