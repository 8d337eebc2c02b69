"""Two contexts on two B200s combined with ncclReduce (skipped on a single-GPU box)."""
import threading

import numpy as np
import pytest
import torch

import helpers
from rust_path_tracer_b200 import dist as rdist
from rust_path_tracer_b200.trace import Renderer

pytestmark = pytest.mark.gpu


def _two_gpus():
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.skipif(not _two_gpus(), reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_nccl_reduce_matches_single_gpu(mode):
    world = helpers.world("DarkCornell")
    w, h, spp, n = 96, 64, 8, 2
    cfg = helpers.config(w, h, 1)
    seeds = helpers.seeds(w, h)
    with Renderer(0) as r:
        r.upload_world(world); r.set_config(cfg); r.write_rng(seeds)
        r.enqueue(spp)
        single = r.read_output()

    uid = Renderer.comm_unique_id()
    results, errors = {}, []

    def rank_main(rank):
        try:
            with Renderer(rank) as r:
                r.upload_world(world); r.set_config(cfg)
                if mode == "samples":
                    s0, s1 = rdist.sample_range(spp, rank, n)
                    r.write_rng(rdist.offset_seeds(seeds, s0))
                    count = s1 - s0
                else:
                    r.set_tile_partition(rank, n)
                    r.write_rng(seeds)
                    count = spp
                r.comm_init(uid, rank, n)
                r.enqueue(count)
                own = None
                if rank != 0:
                    own = r.read_output()
                r.comm_reduce_output(0)
                r.comm_reduce_output(0)  # the combine modifies no accumulator: combining twice gives the same frame
                results[rank] = r.read_output()
                if rank != 0:  # only the root holds a combined frame; the others still read their own accumulator
                    np.testing.assert_array_equal(results[rank], own)
                r.comm_destroy()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=rank_main, args=(k,)) for k in range(n)]
    [t.start() for t in threads]
    [t.join(timeout=120) for t in threads]
    assert not errors, errors
    combined = results[0]
    np.testing.assert_array_equal(combined[:, 3], single[:, 3])
    if mode == "tiles":
        np.testing.assert_array_equal(combined, single)
    else:
        np.testing.assert_allclose(combined[:, :3], single[:, :3], rtol=1e-6, atol=1e-6)
