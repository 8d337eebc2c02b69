"""Scratch: the reference's `interacting` loop (one sample, readback, flush with a new camera; src/trace.rs:182-222)
timed through the C ABI.  usage: python tools/gpu_interactive.py [workload] [frames]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import bench
from rust_path_tracer_b200 import capi
from rust_path_tracer_b200.trace import Renderer

workload = sys.argv[1] if len(sys.argv) > 1 else "cornell"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 200
world, cfg, seeds, _, label, scene, sky = bench.load_workload(workload)
cfg.width, cfg.height = 1280, 720
from rust_path_tracer_b200.world import make_rng_seeds
seeds_host = make_rng_seeds(1280, 720)
seeds = capi.pinned_empty(seeds_host.shape, np.uint32); seeds[...] = seeds_host
fb = capi.pinned_empty(1280 * 720 * 3, np.float32)
with Renderer(0) as r:
    r.upload_world(world, sky); r.set_config(cfg); r.write_rng(seeds)
    for moving in (True, False):
        for _ in range(20):
            r.enqueue(1)
        r.sync()
        t0 = time.perf_counter()
        for k in range(frames):
            if moving:  # flush: new camera, zeroed accumulator, pristine seeds
                cfg.cam_rotation[1] = 0.001 * k
                r.set_config(cfg); r.write_output(None); r.write_rng(seeds)
            r.enqueue(1); r.sync()
            r.read_framebuffer(1.0 if moving else float(k + 1), fb)
        dt = time.perf_counter() - t0
        print(f"{workload} 1280x720 {'camera moving (flush every frame)' if moving else 'camera still (accumulating)'}: {1e3 * dt / frames:.3f} ms/frame, {frames / dt:.0f} fps, {1280 * 720 * frames / dt / 1e6:.0f} Mpaths/s", flush=True)
