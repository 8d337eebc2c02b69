"""Scratch: the BreakTime proxy grown beyond L2 (default 8 M triangles: ~215 MB of nodes + 384 MB of triangle
positions + 512 MB of shading records) — where does the extend kernel go when the scene no longer fits on chip?
usage: python tools/gpu_bigscene.py [triangles] [spp]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from rust_path_tracer_b200.capi import TracingConfig
from rust_path_tracer_b200.scenes import breaktime_proxy, synthetic_hdr_sky
from rust_path_tracer_b200.trace import Renderer
from rust_path_tracer_b200.world import World, make_rng_seeds

ntri = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t0 = time.perf_counter()
baked, atlas = breaktime_proxy(ntri)
t1 = time.perf_counter()
world = World.from_baked(baked, atlas=atlas)
t2 = time.perf_counter()
print(f"{len(world.index_buffer)} triangles, {len(world.nodes)} binary nodes: generate {t1 - t0:.1f} s, SAH build + tables {t2 - t1:.1f} s", flush=True)
cfg = TracingConfig.default(1920, 1080); cfg.nee = 1; cfg.has_skybox = 1
seeds = make_rng_seeds(1920, 1080)
with Renderer(0) as r:
    t3 = time.perf_counter()
    r.upload_world(world, synthetic_hdr_sky())
    print(f"rpt_upload_world (wide re-layout + copies) {time.perf_counter() - t3:.1f} s", flush=True)
    r.set_config(cfg); r.write_rng(seeds)
    for _ in range(3):  # direct run, graph capture, first replay
        r.enqueue(spp)
    r.sync()
    r.reset_counters(); r.enqueue(spp); ms = r.device_ms(); c = r.counters()
    r.set_stage_timing(True); r.enqueue(spp); st = r.stage_timing()
rays = c["nearest_rays"] + c["any_rays"]
print(f"{c['paths'] / ms / 1e3:.1f} Mpaths/s {rays / ms / 1e3:.1f} Mrays/s | " + " ".join(f"{k}={v[0]:.1f}" for k, v in st.items() if v[1]), flush=True)
