"""Scratch: where the time of one `trace_gpu(DarkCornell, 1280x720, 160 samples)` goes, call by call (run on the GPU box)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import helpers
from rust_path_tracer_b200.trace import Renderer

world = helpers.world("DarkCornell")
w, h = 1280, 720
cfg = helpers.config(w, h, 0)
seeds = helpers.seeds(w, h)
for rep in range(3):
    t = [time.perf_counter()]
    def lap(): t.append(time.perf_counter())
    r = Renderer(0); lap()
    r.upload_world(world); lap()
    r.set_config(cfg); r.write_rng(seeds); lap()
    r.enqueue(128); r.sync(); lap()
    fb = r.read_framebuffer(128.0); lap()
    r.enqueue(128); r.sync(); lap()
    fb = r.read_framebuffer(256.0); lap()
    r.enqueue(128); r.sync(); lap()
    r.close() if hasattr(r, "close") else r.__exit__(None, None, None); lap()
    names = ["create", "upload", "config+rng", "enqueue#1", "read#1", "enqueue#2", "read#2", "enqueue#3", "destroy"]
    print(rep, " ".join(f"{n}={1e3 * (b - a):.1f}" for n, a, b in zip(names, t, t[1:])), flush=True)
