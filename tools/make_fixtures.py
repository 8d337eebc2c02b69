"""Regenerate the committed input fixtures from the reference checkout's ASSETS (data, not source).

Run in the build container (needs /root/reference):  python tools/make_fixtures.py

  tests/golden/scenes/<Scene>.npz                    baked, pre-BVH scene arrays of the four shipped .glb scenes
                                                     (rust-path-tracer_b200/glb.py restates the assimp bake)
  rust-path-tracer_b200/resources/bluenoise_r8.npy   R8 of src/resources/bluenoise.png after `into_rgba8()`:
                                                     16-bit gray -> u8 with image-0.24's (c + 128) / 257

The GPU box has no /root/reference; tests, smoke() and bench.py read only these fixtures.
"""
import os
import sys

import numpy as np
from PIL import Image

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rust_path_tracer_b200.glb import load_glb  # noqa: E402

REF = "/root/reference"
SCENES = ["FurnaceTest", "DarkCornell", "PBRTest", "VeachMIS"]


def main():
    out_dir = os.path.join(REPO, "tests", "golden", "scenes")
    os.makedirs(out_dir, exist_ok=True)
    for name in SCENES:
        scene = load_glb(os.path.join(REF, "scenes", name + ".glb"))
        path = os.path.join(out_dir, name + ".npz")
        scene.save(path)
        print(f"{name}: {len(scene.indices)} tris, {len(scene.vertices)} verts, {len(scene.materials)} materials -> {os.path.getsize(path)} B")
    blue16 = np.asarray(Image.open(os.path.join(REF, "src", "resources", "bluenoise.png"))).astype(np.uint32)
    assert blue16.shape == (256, 256)
    r8 = ((blue16 + 128) // 257).astype(np.uint8)
    res = os.path.join(REPO, "rust-path-tracer_b200", "resources")
    os.makedirs(res, exist_ok=True)
    np.save(os.path.join(res, "bluenoise_r8.npy"), r8)
    print("bluenoise_r8.npy", r8.shape, r8.dtype)


if __name__ == "__main__":
    main()
