"""Scratch diagnostics on the GPU box: primary ids of both pipelines vs the oracle."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, REPO + "/oracle", REPO + "/tests"):
    sys.path.insert(0, p)
import numpy as np
import helpers, oracle as om
from rust_path_tracer_b200 import capi
from rust_path_tracer_b200.trace import Renderer

scene = sys.argv[1] if len(sys.argv) > 1 else "DarkCornell"
W, H, nee, spp = 64, 48, 1, 4
world = helpers.world(scene); cfg = helpers.config(W, H, nee); seeds = helpers.seeds(W, H)
osc = om.OracleScene(world)
_, _, _, oids = om.trace(cfg, osc, seeds, 1, want_primary_ids=True)
oout, _, octr, _ = om.trace(cfg, osc, seeds, spp)
print("oracle ids", oids[:8], "ctr", octr["nearest_rays"], octr["any_rays"])
for name, pipe in (("mega", capi.PIPELINE_MEGAKERNEL), ("wave", capi.PIPELINE_WAVEFRONT)):
    with Renderer(0, pipe) as r:
        r.upload_world(world); r.set_config(cfg); r.write_rng(seeds)
        ids = r.read_primary_ids()
        r.enqueue(spp)
        out = r.read_output(); ctr = r.counters()
    mism = (ids != oids)
    print(name, "ids", ids[:8], "mismatch", mism.mean(), "miss count", (ids == 0xFFFFFFFF).sum(), "ctr", ctr)
    print(name, "MAE", np.abs(out[:, :3] / spp - oout[:, :3] / spp).mean(), "mean", out[:, :3].mean() / spp, oout[:, :3].mean() / spp, "w", out[:4, 3])
    if mism.any():
        k = np.nonzero(mism)[0][:6]
        print("  first mismatches at", k, ids[k], oids[k])
