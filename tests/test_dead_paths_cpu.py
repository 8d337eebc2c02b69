"""The CUDA backend retires paths whose throughput became exactly (0, 0, 0) before the roulette bounces instead of
tracing them to their first roulette bounce like kernels/src/lib.rs:145-181 does (wavefront_shade.cu).
Checked here on the CPU with the restatement itself: with the same rule switched on in the oracle (a test switch,
off by default), the accumulator of every scene / NEE mode / sky kind is the same, bit for bit, while fewer rays are
traced."""
import numpy as np
import pytest

import helpers
import oracle as om

CASES = [(scene, nee, sky) for scene in helpers.SCENES for nee in (0, 1, 2) for sky in (False,)] + [("PBRTest", 0, True), ("VeachMIS", 1, True)]


@pytest.fixture
def retire_switch():
    yield om.set_retire_dead_paths
    om.set_retire_dead_paths(False)


@pytest.mark.parametrize("scene,nee,sky", CASES)
def test_retired_paths_add_nothing(scene, nee, sky, retire_switch):
    world = helpers.world(scene)
    w, h, spp = 96, 54, 8
    cfg = helpers.config(w, h, nee, has_skybox=1 if sky else 0)
    sc = om.OracleScene(world, helpers.synthetic_sky() if sky else None)
    seeds = helpers.seeds(w, h)
    retire_switch(False)
    out_ref, rng_ref, ctr_ref, _ = om.trace(cfg, sc, seeds, spp)
    retire_switch(True)
    out, rng, ctr, _ = om.trace(cfg, sc, seeds, spp)
    np.testing.assert_array_equal(out.view(np.uint32), out_ref.view(np.uint32))
    np.testing.assert_array_equal(rng, rng_ref)
    assert ctr["nearest_rays"] <= ctr_ref["nearest_rays"] and ctr["any_rays"] <= ctr_ref["any_rays"]


def test_retirement_on_the_textured_proxy(retire_switch):
    """Textures with black texels, normal maps, MIS and an HDR sky: the case the bench runs."""
    world, sky = helpers.proxy_world(), helpers.synthetic_sky()
    w, h, spp = 160, 90, 4
    cfg = helpers.config(w, h, 1, has_skybox=1)
    sc = om.OracleScene(world, sky)
    seeds = helpers.seeds(w, h)
    retire_switch(False)
    out_ref, _, ctr_ref, _ = om.trace(cfg, sc, seeds, spp)
    retire_switch(True)
    out, _, ctr, _ = om.trace(cfg, sc, seeds, spp)
    np.testing.assert_array_equal(out.view(np.uint32), out_ref.view(np.uint32))
    assert ctr["nearest_rays"] < ctr_ref["nearest_rays"]


NAN_SAMPLE = 4534  # found on a B200 by tools/gpu_nan_hunt.py: the first (and only) NaN sample of the proxy's first 5120


def test_a_dead_path_that_will_draw_a_one_is_traced_on(retire_switch):
    """The rule's one guard, on the case that showed it is needed: at sample 4534 of the BreakTime proxy's 1080p frame
    every pixel with the blue-noise offset 640034368 draws exactly 1.0 in dimension 11 (0xFFFFFF90 rounds up to 2^32); as a
    lobe selector against a specular weight of exactly 1.0 it picks the diffuse lobe with 1 / (1 - 1) = inf, and the
    reference has 20 NaN pixels.  Four of those paths were dead (throughput exactly 0) when they got there: 0 x inf = NaN
    — retiring them would leave 16.  With the guard the retiring restatement equals the reference, NaN for NaN."""
    import bench

    world, cfg, seeds, _spp, _label, _scene, sky = bench.load_workload("breaktime")
    sc = om.OracleScene(world, sky)
    at = seeds.copy()
    at[:, 0] += np.uint32(NAN_SAMPLE)
    retire_switch(False)
    out_ref, _, ctr_ref, _ = om.trace(cfg, sc, at.copy(), 1)
    retire_switch(True)
    out, _, ctr, _ = om.trace(cfg, sc, at.copy(), 1)
    nan_ref = np.flatnonzero(~np.isfinite(out_ref[:, :3]).all(axis=1))
    assert len(nan_ref) == 20 and len(set(seeds[nan_ref, 1].tolist())) == 1
    np.testing.assert_array_equal(out.view(np.uint32), out_ref.view(np.uint32))  # (NaN payloads included)
    assert ctr["nearest_rays"] < ctr_ref["nearest_rays"]
