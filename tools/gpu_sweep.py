"""Scratch: time the wavefront stages for tunable settings (run on the GPU box).
usage: python tools/gpu_sweep.py workload[,workload] spp "ENV=VAL,ENV=VAL" ...      (RPT_B200_LIBRARY picks a variant build)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import bench
from rust_path_tracer_b200.trace import Renderer

workloads, spp = sys.argv[1].split(","), int(sys.argv[2])
tag = os.path.basename(os.environ.get("RPT_B200_LIBRARY", "default"))
for workload, setting in ((w, s) for w in workloads for s in (sys.argv[3:] or [""])):
    world, cfg, seeds, _, label, scene, sky = bench.load_workload(workload) if workload != globals().get("_loaded") else _cache
    _loaded, _cache = workload, (world, cfg, seeds, _, label, scene, sky)
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    with Renderer(0) as r:
        r.upload_world(world, sky, build_on_device=os.environ.get('SWEEP_DEVICE_BUILD') == '1'); r.set_config(cfg); r.write_rng(seeds)
        for _ in range(3):  # direct run, graph capture, first replay
            r.enqueue(spp)
        r.sync()
        r.reset_counters(); r.enqueue(spp); ms = r.device_ms(); c = r.counters()
        r.set_stage_timing(True); r.enqueue(spp); st = r.stage_timing(); r.set_stage_timing(False)
        r.set_trace_statistics(True); r.reset_counters(); r.enqueue(spp); ts = r.trace_statistics(); r.set_trace_statistics(False)
    rays = c["nearest_rays"] + c["any_rays"]
    print(f"{tag} {workload} [{setting}] {c['paths']/ms/1e3:8.1f} Mpaths/s {rays/ms/1e3:8.1f} Mrays/s | " +
          " ".join(f"{k}={v[0]:.1f}" for k, v in st.items() if v[1]) +
          f" | per nearest ray: {ts['nearest_node_visits'] / max(ts['nearest_rays'], 1):.2f} visits, {ts['nearest_triangle_tests'] / max(ts['nearest_rays'], 1):.2f} tests"
          f"; per any ray: {ts['any_node_visits'] / max(ts['any_rays'], 1):.2f}, {ts['any_triangle_tests'] / max(ts['any_rays'], 1):.2f}", flush=True)
    for kv in filter(None, setting.split(",")):
        os.environ.pop(kv.split("=")[0], None)
