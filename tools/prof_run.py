"""Scratch: one short render for ncu captures (run under `ncu ... python tools/prof_run.py [workload] [spp]` on the GPU box;
set RPT_GRAPHS=0 so every launch is visible)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from rust_path_tracer_b200.trace import Renderer

workload = sys.argv[1] if len(sys.argv) > 1 else "breaktime"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
world, cfg, seeds, _, label, scene, sky = bench.load_workload(workload)
with Renderer(0) as r:
    r.upload_world(world, sky); r.set_config(cfg); r.write_rng(seeds)
    r.enqueue(spp)
    r.sync()
    print(workload, spp, r.counters())
