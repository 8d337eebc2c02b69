"""Throughput sweeps of SURVEY.md §8(d): camera yaw x position at a fixed spp, and spp at the default camera, for the
named scenes at their BASELINE frame sizes (run on the GPU box).  usage: python tools/camera_sweep.py > profiles/...txt"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np

import bench
from rust_path_tracer_b200.trace import Renderer

YAWS = [0.0, 0.5, -0.5, float(np.pi)]
POSITIONS = [(0.0, 1.0, -5.0), (2.5, 1.5, -3.0), (0.0, 3.0, 0.5)]
SPPS = [1, 16, 256, 1024]


def measure(r, cfg, seeds, spp):
    r.set_config(cfg)
    r.write_rng(seeds)
    r.write_output(None)
    npix = cfg.width * cfg.height
    r.enqueue(spp if npix * spp <= (1 << 26) else max(1, (1 << 26) // npix))  # warm-up with the timed run's wave shapes (allocation, graph capture)
    r.sync()
    r.write_rng(seeds)
    r.write_output(None)
    r.reset_counters()
    r.enqueue(spp)
    ms = r.device_ms()
    c = r.counters()
    frame = r.read_output()
    nan = int((~np.isfinite(frame[:, :3]).all(axis=1)).sum())
    return c["paths"] / ms / 1e3, (c["nearest_rays"] + c["any_rays"]) / ms / 1e3, (c["nearest_rays"] + c["any_rays"]) / c["paths"], nan


def main():
    workloads = sys.argv[1:] or ["cornell", "pbr", "veach", "furnace", "breaktime"]
    for name in workloads:
        world, cfg0, seeds, spp0, label, scene, sky = bench.load_workload(name)
        print(f"## {label}  [{cfg0.width}x{cfg0.height}, nee={cfg0.nee}]", flush=True)
        with Renderer(0) as r:
            r.upload_world(world, sky)
            print("camera sweep at 16 spp: position, yaw -> Mpaths/s, Mrays/s, rays/path, NaN pixels")
            for pos in POSITIONS:
                for yaw in YAWS:
                    cfg = cfg0.copy()
                    cfg.cam_position[:] = [pos[0], pos[1], pos[2], 0.0]
                    cfg.cam_rotation[:] = [0.0, yaw, 0.0, 0.0]
                    mp, mr, rpp, nan = measure(r, cfg, seeds, 16)
                    print(f"  pos {pos} yaw {yaw:+.2f}: {mp:8.1f} {mr:8.1f} {rpp:5.2f} {nan}", flush=True)
            print("spp sweep at the default camera: spp -> Mpaths/s, Mrays/s")
            for spp in SPPS:
                if cfg0.width * cfg0.height * spp > 2.5e9:
                    continue
                mp, mr, rpp, nan = measure(r, cfg0.copy(), seeds, spp)
                print(f"  spp {spp:5d}: {mp:8.1f} {mr:8.1f}", flush=True)


if __name__ == "__main__":
    main()
