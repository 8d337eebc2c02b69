"""Host-side logic of the one-process-per-GPU split, exercised with world_size = 2 over gloo on the
CPU: the sample-range and tile partitions each rank derives, and the combine (a sum-reduce of
full-frame accumulators, which the GPU path does with ncclReduce).  The per-rank "render" is the
oracle, so this checks the PARTITION, not the CUDA kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from rust_path_tracer_b200 import dist as rdist

W, H, SPP, NRANKS = 48, 40, 6, 2


def _worker(rank, port, mode, out_dir):
    for p in (helpers.REPO, os.path.join(helpers.REPO, "oracle"), os.path.join(helpers.REPO, "tests")):
        sys.path.insert(0, p)
    import oracle as om

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=NRANKS)
    world = helpers.world("DarkCornell")
    cfg = helpers.config(W, H, 1)
    seeds = helpers.seeds(W, H)
    scene = om.OracleScene(world)
    if mode == "samples":
        s0, s1 = rdist.sample_range(SPP, rank, NRANKS)
        acc, _, _, _ = om.trace(cfg, scene, rdist.offset_seeds(seeds, s0), s1 - s0)
    else:  # tiles: this rank owns a subset of pixels, everything else stays zero
        full, _, _, _ = om.trace(cfg, scene, seeds, SPP)
        acc = np.zeros_like(full)
        mine = rdist.tile_pixels(W, H, rank, NRANKS)
        acc[mine] = full[mine]
    t = torch.from_numpy(acc)
    dist.reduce(t, 0, op=dist.ReduceOp.SUM)  # ncclReduce(sum, root 0) on the GPU path
    if rank == 0:
        np.save(os.path.join(out_dir, f"{mode}.npy"), t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_rank_partition_matches_single_render(tmp_path, mode):
    import oracle as om

    port = 29500 + (os.getpid() % 2000) + (0 if mode == "samples" else 1)
    mp.spawn(_worker, args=(port, mode, str(tmp_path)), nprocs=NRANKS, join=True)
    combined = np.load(tmp_path / f"{mode}.npy")
    world = helpers.world("DarkCornell")
    single, _, _, _ = om.trace(helpers.config(W, H, 1), om.OracleScene(world), helpers.seeds(W, H), SPP)
    np.testing.assert_array_equal(combined[:, 3], single[:, 3])  # every pixel got exactly SPP samples
    if mode == "tiles":
        np.testing.assert_array_equal(combined, single)  # disjoint pixels: bit-identical
    else:
        np.testing.assert_allclose(combined[:, :3], single[:, :3], rtol=1e-6, atol=1e-6)  # fp32 summation order only


def test_sample_ranges_tile_the_interval():
    for total in (1, 7, 64, 1000):
        for n in (1, 2, 3, 8):
            r = [rdist.sample_range(total, k, n) for k in range(n)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(n - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


@pytest.mark.parametrize("w,h,n", [(64, 64, 2), (100, 70, 3), (1920, 1080, 8), (31, 5, 4)])
def test_tile_partition_is_a_partition(w, h, n):
    parts = [rdist.tile_pixels(w, h, r, n) for r in range(n)]
    allpix = np.concatenate(parts)
    assert len(allpix) == w * h and len(np.unique(allpix)) == w * h
    # tiles are 32x32 and round-robin: pixel (x, y) belongs to rank ((y//32) * ceil(w/32) + x//32) % n
    for r, p in enumerate(parts):
        x, y = p % w, p // w
        assert (((y // 32) * ((w + 31) // 32) + x // 32) % n == r).all()


def test_tracing_state_keeps_the_stop_word_in_step_with_its_flags():
    """`rpt_enqueue_interruptible` polls ONE word; TracingState derives it from running / interacting / dirty the way the
    reference's dispatch loop combines its atomics (src/trace.rs:187-193): stop = interacting | dirty | !running."""
    from rust_path_tracer_b200.trace import TracingState, setup_trace

    s = TracingState(8, 8)
    assert s.stop_flag[0] == 1  # not running yet
    s.running = True
    assert s.stop_flag[0] == 0
    s.interacting = True
    assert s.stop_flag[0] == 1 and s.interacting
    s.interacting = False
    s.dirty = True
    assert s.stop_flag[0] == 1
    s.dirty = False
    assert s.stop_flag[0] == 0
    s.running = False
    assert s.stop_flag[0] == 1
    t = setup_trace(8, 8, 4)
    assert t.running and t.stop_flag[0] == 0
    t.samples = 4
    t._watch()  # the watcher clears `running` once enough samples are in
    assert not t.running and t.stop_flag[0] == 1
    assert not setup_trace(8, 8, 0).running  # 0 samples requested: stopped at once ("startup" benches)
