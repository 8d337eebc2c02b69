"""Which samples of a long render are NaN?  (run on the GPU box)  The reference adds the sky term without a NaN mask
(kernels/src/lib.rs:69), so a pixel whose sample k is NaN stays NaN for good; at thousands of samples per pixel a
handful of pixels of the BreakTime proxy are (bench.py's reduce_check counts 16 at 5120 spp).  This finds, for every such
pixel, the FIRST sample index that is NaN (chunks of `chunk` samples, then the chunk again one sample at a time), and
writes the list to gpurun_out/nan_samples.json — tests/checkers/nan_samples_vs_oracle.py then asks the CPU oracle for exactly
those samples.
usage: python tools/gpu_nan_hunt.py [workload] [total_spp] [chunk]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import bench
from rust_path_tracer_b200.trace import Renderer

workload = sys.argv[1] if len(sys.argv) > 1 else "breaktime"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 5120
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 64
world, cfg, seeds, _, label, scene, sky = bench.load_workload(workload)
found = {}  # pixel -> first NaN sample index
with Renderer(0) as r, Renderer(0) as probe:
    for x in (r, probe):
        x.upload_world(world, sky); x.set_config(cfg)
    r.write_rng(seeds)
    out = None
    bad_before = np.zeros(cfg.width * cfg.height, bool)
    for first in range(0, total, chunk):
        r.enqueue(chunk)
        out = r.read_output(out)
        bad = ~np.isfinite(out[:, :3]).all(axis=1)
        new = np.flatnonzero(bad & ~bad_before)
        bad_before = bad
        if len(new) == 0:
            continue
        s = seeds.copy(); s[:, 0] += np.uint32(first)  # the accumulate kernel advances seeds.x by the samples done
        probe.write_rng(s)
        left = set(int(p) for p in new)
        one = None
        for k in range(chunk):
            probe.write_output(None)
            probe.enqueue(1)
            one = probe.read_output(one)
            for p in [p for p in left if not np.isfinite(one[p, :3]).all()]:
                found[p] = first + k
                left.discard(p)
            if not left:
                break
        print(f"samples {first}..{first + chunk - 1}: {len(new)} new NaN pixel(s) -> {[(p, found.get(p)) for p in new]}", flush=True)
res = {"workload": workload, "width": cfg.width, "height": cfg.height, "samples": total, "nan_pixels": len(found),
       "first_nan_sample": {str(p): k for p, k in sorted(found.items())}}
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(REPO, "gpurun_out", "nan_samples.json"), "w"), indent=1)
print(json.dumps(res))
