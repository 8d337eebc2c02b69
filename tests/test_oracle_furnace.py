"""Pins the CPU oracle to the ONLY known answer the reference holds for this path:
tests/correctness_tests.rs:14-33 — FurnaceTest, 128x128, 32 spp, default config, blue-noise seeds,
pixel (65, 75): every channel ^(1/2.2) within 0.02 of 0.8, with nee = 0 and nee = MIS."""
import numpy as np
import pytest

import helpers
import oracle as om


@pytest.mark.parametrize("nee", [0, 1], ids=["nee_off", "mis"])
def test_furnace_known_answer(nee):
    size, coord, albedo, tolerance, spp = 128, (65, 75), 0.8, 0.02, 32
    world = helpers.world("FurnaceTest")
    cfg = helpers.config(size, size, nee)
    out, rng, ctr, _ = om.trace(cfg, om.OracleScene(world), helpers.seeds(size, size), spp)
    frame = (out[:, :3] / np.float32(spp)).reshape(-1)  # framebuffer = output.xyz / samples, src/trace.rs:303-308
    for c in range(3):
        v = frame[(size * 3) * coord[1] + coord[0] * 3 + c] ** (1.0 / 2.2)
        assert abs(v - albedo) < tolerance, (c, v)
    # loop bookkeeping of src/trace.rs:295-296 / kernels/src/rng.rs:47-49
    assert (out[:, 3] == spp).all() and (rng[:, 0] == spp).all()
    assert ctr["light_index_clamped"] == 0


def test_furnace_is_energy_conserving_over_the_sphere():
    """Same scene at 64x64: the 7x7 block of pixels at the centre of the grey sphere averages to the
    furnace value (0.8 in gamma space) with and without NEE, and nothing is NaN."""
    world = helpers.world("FurnaceTest")
    for nee in (0, 1):
        cfg = helpers.config(64, 64, nee)
        out, *_ = om.trace(cfg, om.OracleScene(world), helpers.seeds(64, 64), 32)
        img = (out[:, :3] / 32).reshape(64, 64, 3)
        assert np.isfinite(img).all()
        assert abs(img[34:41, 29:36].mean() ** (1 / 2.2) - 0.8) < 0.02
