"""The oracle reproduces the committed golden vectors bit for bit (tests/checkers/make_golden.py made them);
this guards the checker itself against drift between rounds."""
import glob
import os

import numpy as np
import pytest

import helpers
import oracle as om
from rust_path_tracer_b200.capi import TracingConfig

GOLDENS = sorted(glob.glob(os.path.join(helpers.GOLDEN_DIR, "*.npz")))


def load_case(path):
    z = np.load(path)
    cfg = TracingConfig.from_buffer_copy(z["config"].tobytes())
    return z, cfg


@pytest.mark.parametrize("path", GOLDENS, ids=[os.path.basename(p)[:-4] for p in GOLDENS])
def test_oracle_reproduces_golden(path):
    z, cfg = load_case(path)
    world = helpers.world(str(z["scene"]))
    seeds = helpers.seeds(cfg.width, cfg.height)
    osc = om.OracleScene(world)
    _, _, _, ids = om.trace(cfg, osc, seeds, 1, want_primary_ids=True)
    out, _, ctr, _ = om.trace(cfg, osc, seeds, int(z["spp"]))
    np.testing.assert_array_equal(ids, z["primary_ids"])
    np.testing.assert_array_equal(out.view(np.uint32), z["output"].view(np.uint32))
    assert ctr["nearest_rays"] == int(z["nearest_rays"]) and ctr["any_rays"] == int(z["any_rays"])


def test_oracle_is_thread_count_independent():
    world = helpers.world("DarkCornell")
    cfg = helpers.config(48, 32, 1)
    seeds = helpers.seeds(48, 32)
    a, *_ = om.trace(cfg, om.OracleScene(world), seeds, 4, threads=1)
    b, *_ = om.trace(cfg, om.OracleScene(world), seeds, 4, threads=4)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
