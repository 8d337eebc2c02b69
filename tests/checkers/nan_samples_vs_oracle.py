"""CPU side of tools/gpu_nan_hunt.py: render exactly the sample indices the GPU found NaN with the oracle and compare,
pixel by pixel, which are NaN there.  usage: python tests/checkers/nan_samples_vs_oracle.py [gpurun_out/nan_samples.json]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np
import bench
import oracle as om

res = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "nan_samples.json")))
world, cfg, seeds, _, label, scene, sky = bench.load_workload(res["workload"])
osc = om.OracleScene(world, sky)
by_sample = {}
for p, k in res["first_nan_sample"].items():
    by_sample.setdefault(k, []).append(int(p))
print(f"{res['workload']} {res['width']}x{res['height']}, {res['samples']} samples per pixel: {res['nan_pixels']} NaN pixels on the GPU, first NaN at sample(s) {sorted(by_sample)}")
for k, pixels in sorted(by_sample.items()):
    s = seeds.copy(); s[:, 0] += np.uint32(k)
    out, _, _, _ = om.trace(cfg, osc, s, 1)
    bad = np.flatnonzero(~np.isfinite(out[:, :3]).all(axis=1))
    print(f"sample {k}: GPU NaN pixels {sorted(pixels)}")
    print(f"           oracle NaN pixels {bad.tolist()}")
    print(f"           identical sets: {sorted(pixels) == bad.tolist()};  blue-noise offsets seeds.y of those pixels: {sorted(set(int(v) for v in seeds[bad, 1]))}")
