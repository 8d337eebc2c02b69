"""Generate the committed golden vectors under tests/golden/ by running the CPU oracle.

    python tests/checkers/make_golden.py

Each .npz holds, for one small case: the accumulator after `spp` samples (float32 sum rgb + count),
the bounce-0 triangle ids of sample 0, and the oracle's ray counters.  tests/test_golden.py checks
the oracle still reproduces them bit for bit (CPU), tests/test_parity_gpu.py checks the CUDA path
against them on the GPU box (where /root/reference and this generator's inputs are the same
committed fixtures).
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
import oracle as om  # noqa: E402

CASES = [
    # name, scene, w, h, spp, nee, extra config
    ("furnace_nee0", "FurnaceTest", 64, 64, 8, 0, {}),
    ("furnace_mis", "FurnaceTest", 64, 64, 8, 1, {}),
    ("cornell_mis", "DarkCornell", 64, 48, 16, 1, {}),
    ("cornell_direct", "DarkCornell", 64, 48, 16, 2, {}),
    ("pbr_sky", "PBRTest", 80, 44, 8, 0, {}),
    ("veach_mis", "VeachMIS", 80, 44, 8, 1, {}),
    ("pbr_rotated", "PBRTest", 64, 36, 4, 0, {"cam_rotation": [0.15, -0.6, 0.0, 0.0], "cam_position": [-2.0, 1.5, -4.0, 0.0]}),
]


def main():
    for name, scene, w, h, spp, nee, extra in CASES:
        world = helpers.world(scene)
        cfg = helpers.config(w, h, nee, **extra)
        seeds = helpers.seeds(w, h)
        osc = om.OracleScene(world)
        _, _, _, ids = om.trace(cfg, osc, seeds, 1, want_primary_ids=True)
        out, rng, ctr, _ = om.trace(cfg, osc, seeds, spp)
        path = os.path.join(helpers.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, scene=scene, width=w, height=h, spp=spp, nee=nee, config=np.frombuffer(bytes(cfg), np.uint8),
                            output=out, primary_ids=ids, nearest_rays=ctr["nearest_rays"], any_rays=ctr["any_rays"])
        print(name, os.path.getsize(path), "B", "mean", out[:, :3].mean() / spp)


if __name__ == "__main__":
    main()
