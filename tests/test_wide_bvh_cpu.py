"""The backend's 8-wide BVH collapse + traversal, compiled for the host by tests/cpu_harness, must
return the same triangle, the same t bits and the same back-face flag as the reference traversal
(oracle) — nearest-hit and any-hit — on camera rays and on random / axis-aligned / surface rays."""
import ctypes as C
import os
import subprocess

import warnings

import numpy as np
import pytest

import helpers
import oracle as om

HARNESS_DIR = os.path.join(helpers.REPO, "tests", "cpu_harness")
HARNESS_SO = os.path.join(HARNESS_DIR, "_build", "libwide_harness.so")


@pytest.fixture(scope="module")
def harness():
    srcs = [os.path.join(HARNESS_DIR, "wide_harness.cpp"), os.path.join(helpers.REPO, "rust-path-tracer_b200", "csrc", "wide_bvh_build.cpp")]
    hdrs = [os.path.join(helpers.REPO, "rust-path-tracer_b200", "csrc", "dev", h) for h in ("wide_bvh.cuh", "exact.cuh", "vec.cuh")]
    os.makedirs(os.path.dirname(HARNESS_SO), exist_ok=True)
    if not os.path.exists(HARNESS_SO) or any(os.path.getmtime(s) > os.path.getmtime(HARNESS_SO) for s in srcs + hdrs):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", HARNESS_SO, *srcs], check=True,
                       capture_output=True)
    return C.CDLL(HARNESS_SO)


def wide_intersect(lib, world, rays, any_hit=False, max_t=None):
    n = len(rays)
    rays = np.ascontiguousarray(rays, np.float32)
    max_t = np.zeros(n, np.float32) if max_t is None else np.ascontiguousarray(max_t, np.float32)
    hit, tri, t, back = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
    stats = np.zeros(5, np.uint32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.harness_wide_intersect(P(world.per_vertex_buffer), C.c_uint32(len(world.per_vertex_buffer)), P(world.index_buffer),
                                    C.c_uint32(len(world.index_buffer)), P(world.nodes), C.c_uint32(len(world.nodes)), P(rays), C.c_uint32(n),
                                    C.c_int(int(any_hit)), P(max_t), P(hit), P(tri), P(t), P(back), P(stats))
    assert rc == 0
    return hit, tri, t, back, stats


def make_rays(world, n, rs):
    pos = world.per_vertex_buffer["vertex"][:, :3]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        lo, hi = np.nanmin(pos, axis=0), np.nanmax(pos, axis=0)  # (some test scenes carry NaN vertices)
    c, ext = (lo + hi) / 2, (hi - lo).max()
    o = (c + (rs.random((n, 3)) - 0.5) * ext * 1.2).astype(np.float32)
    d = rs.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    k = n // 20
    d[:k, 0] = 0; d[k:2 * k, 1] = 0; d[2 * k:3 * k] = [0, 0, 1]; d[3 * k:4 * k] = [0, -1, 0]
    tris = world.index_buffer
    tp = pos[tris[rs.integers(0, len(tris), n // 2), :3]]
    bw = rs.dirichlet([1, 1, 1], n // 2).astype(np.float32)
    o[n // 2:] = (tp * bw[:, :, None]).sum(1) + d[n // 2:] * np.float32(0.001)  # bounce-like rays leaving a surface
    return np.concatenate([o, d], axis=1).astype(np.float32), (rs.random(n) * ext).astype(np.float32)


@pytest.mark.parametrize("scene", helpers.SCENES)
def test_wide_traversal_matches_reference_traversal(harness, scene):
    world = helpers.world(scene)
    osc = om.OracleScene(world)
    rays, max_t = make_rays(world, 40000, np.random.default_rng(3))
    cfg = helpers.config(160, 90)
    rays = np.concatenate([rays, om.camera_rays(cfg, helpers.seeds(160, 90))])
    max_t = np.concatenate([max_t, np.full(160 * 90, 3.0, np.float32)])

    oh, ot, ott, ob = om.intersect(osc, rays)
    wh, wt, wtt, wb, stats = wide_intersect(harness, world, rays)
    np.testing.assert_array_equal(wh, oh)
    both = oh == 1
    tie = both & (wt != ot)  # a different triangle is only acceptable as an exact-t tie
    assert (wtt[tie].view(np.uint32) == ott[tie].view(np.uint32)).all()
    assert tie.mean() <= 1e-4
    np.testing.assert_array_equal(wtt[both & ~tie].view(np.uint32), ott[both & ~tie].view(np.uint32))
    np.testing.assert_array_equal(wb[both & ~tie], ob[both & ~tie])
    assert stats[4] <= 24 and stats[1] + 1 <= 24  # stack high-water / tree depth within the smem stack

    oh, *_ = om.intersect(osc, rays, any_hit=True, max_t=max_t)
    wh, *_ = wide_intersect(harness, world, rays, any_hit=True, max_t=max_t)
    np.testing.assert_array_equal(wh, oh)


def test_nan_and_degenerate_rays_miss(harness):
    world = helpers.world("DarkCornell")
    rays = np.array([[0, 1, -5, np.nan, 0, 1], [np.nan, 1, -5, 0, 0, 1], [0, 1, -5, 0, 0, 0], [0, 1, -5, np.inf, 0, 1]], np.float32)
    oh, *_ = om.intersect(om.OracleScene(world), rays)
    wh, *_ = wide_intersect(harness, world, rays)
    np.testing.assert_array_equal(wh, oh)
    assert (wh == 0).all()


def test_unorm8_is_the_ieee_division(harness):
    """The shading kernels decode texels with one Newton step instead of a division; it must give exactly
    `x as f32 / 255.0` (src/asset.rs:266-273) for every byte value."""
    harness.harness_unorm8.restype = C.c_float
    got = np.array([harness.harness_unorm8(C.c_uint32(x)) for x in range(256)], np.float32)
    want = np.arange(256, dtype=np.float32) / np.float32(255.0)
    np.testing.assert_array_equal(got, want)


def wide_digest(lib, world, nodes=None):
    nodes = world.nodes if nodes is None else nodes
    out = np.zeros(6, np.uint32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.harness_wide_digest(P(world.per_vertex_buffer), C.c_uint32(len(world.per_vertex_buffer)), P(world.index_buffer),
                                 C.c_uint32(len(world.index_buffer)), P(nodes), C.c_uint32(len(nodes)), P(out))
    return rc, out.tolist()


@pytest.mark.parametrize("mode", ["dp", "greedy"])
def test_collapse_is_independent_of_thread_count(harness, mode, monkeypatch):
    """The collapse builds every level of the wide tree with all host threads; nodes, triangle records and both
    index maps must come out bit for bit the same for any thread count."""
    world = helpers.world("PBRTest")  # 47 k binary nodes: large enough to be shared out
    monkeypatch.setenv("RPT_COLLAPSE", mode)
    digests = []
    for threads in ("1", "2", "5", "16"):
        monkeypatch.setenv("RPT_BUILD_THREADS", threads)
        rc, d = wide_digest(harness, world)
        assert rc == 0
        digests.append(d)
    assert all(d == digests[0] for d in digests), digests


def test_collapse_rejects_malformed_trees(harness, monkeypatch):
    world = helpers.world("PBRTest")
    inner = np.flatnonzero(world.nodes["triangle_count"] == 0)
    for threads in ("1", "8"):
        monkeypatch.setenv("RPT_BUILD_THREADS", threads)
        bad = world.nodes.copy()
        bad["left_or_first"][inner[-1]] = len(bad) - 1          # right child out of range
        assert wide_digest(harness, world, bad)[0] == -1
        bad = world.nodes.copy()
        bad["left_or_first"][inner[-1]] = bad["left_or_first"][inner[10]]  # a subtree referenced twice (and its own orphaned)
        assert wide_digest(harness, world, bad)[0] == -1
        bad = world.nodes.copy()
        bad["left_or_first"][inner[-1]] = 0                     # a cycle through the root
        assert wide_digest(harness, world, bad)[0] == -1


def _soup_world(verts_xyz, name):
    """A World built by the product's own host builders (rpt_build_bvh etc.) from a bare triangle soup."""
    from rust_path_tracer_b200.glb import BakedScene
    from rust_path_tracer_b200.world import World

    base = BakedScene.load(helpers.SCENE_DIR + "/DarkCornell.npz")
    nt = len(verts_xyz) // 3
    scene = BakedScene.__new__(BakedScene)
    scene.__dict__.update(base.__dict__)
    scene.vertices = np.concatenate([np.asarray(verts_xyz, np.float32), np.ones((nt * 3, 1), np.float32)], axis=1)
    scene.normals = np.tile(np.array([[0, 1, 0, 0]], np.float32), (nt * 3, 1))[:, : base.normals.shape[1]]
    scene.tangents = np.tile(np.array([[1, 0, 0, 0]], np.float32), (nt * 3, 1))[:, : base.tangents.shape[1]]
    scene.uvs = np.zeros((nt * 3, base.uvs.shape[1]), np.float32)
    idx = np.zeros((nt, 4), np.uint32)
    idx[:, :3] = np.arange(nt * 3, dtype=np.uint32).reshape(nt, 3)
    scene.indices = idx
    scene.textures = [{} for _ in base.materials]
    return World.from_baked(scene)


DEGENERATE = ["coincident", "flat", "duplicates", "slivers", "one", "nan_vertices", "all_nan"]


@pytest.mark.parametrize("kind", DEGENERATE)
def test_degenerate_geometry_traces_like_the_reference(harness, kind):
    """Geometry that stresses the builders' corner cases — all centroids equal (no axis can split: one over-full
    leaf), everything in one plane, exact duplicates (exact-t ties), needle triangles, a single triangle — must still
    trace exactly like the reference traversal of the same binary tree."""
    rs = np.random.default_rng(23)
    if kind == "coincident":  # 300 different triangles around one common centroid
        a = rs.normal(size=(300, 3)); b = rs.normal(size=(300, 3))
        v = np.stack([a, b, -(a + b)], axis=1).reshape(-1, 3) * 0.5 + [0.0, 1.0, 0.0]
    elif kind == "flat":  # 2000 triangles in the plane y = 1
        c = rs.random((2000, 1, 3)) * [4, 0, 4]
        v = (c + rs.normal(0, 0.05, (2000, 3, 3)) * [1, 0, 1] + [-2.0, 1.0, -2.0]).reshape(-1, 3)
    elif kind == "duplicates":  # every triangle four times
        c = rs.random((250, 1, 3)) * 3
        t = (c + rs.normal(0, 0.2, (250, 3, 3))) + [-1.5, 0.0, -1.5]
        v = np.repeat(t, 4, axis=0).reshape(-1, 3)
    elif kind == "slivers":  # needles spanning the whole scene
        p = rs.random((400, 3)) * 4 - 2
        q = rs.random((400, 3)) * 4 - 2
        v = np.stack([p, q, p + rs.normal(0, 1e-4, (400, 3))], axis=1).reshape(-1, 3) + [0.0, 1.0, 0.0]
    elif kind in ("nan_vertices", "all_nan"):  # NaN coordinates: ignored by every min / max, such triangles are never hit
        v = (rs.random((1500, 1, 3)) * 4 - 2 + rs.normal(0, 0.15, (1500, 3, 3))).reshape(-1, 3)
        v[:: (1 if kind == "all_nan" else 1201)] = np.nan  # (all of them, or 4 scattered vertices: more would make the binary tree deeper than the reference's own stack)
    else:
        v = np.array([[-1, 0, 0], [1, 0, 0], [0, 2, 0]], np.float64)
    world = _soup_world(v.astype(np.float32), kind)
    osc = om.OracleScene(world)
    rays, max_t = make_rays(world, 20000, rs)
    oh, ot, ott, ob = om.intersect(osc, rays)
    wh, wt, wtt, wb, _ = wide_intersect(harness, world, rays)
    np.testing.assert_array_equal(wh, oh)
    both = oh == 1
    tie = both & (wt != ot)  # (duplicates: any copy may win, at the same t)
    if kind in ("flat", "duplicates"):
        # Overlapping coplanar triangles: their t differ by rounding only, and the reference's own box test (tmin computed
        # by a division, compared with a t from Moeller-Trumbore) can cull the node of the triangle that is an ulp
        # nearer.  The wide tree's boxes are conservative, so it returns the true minimum: never farther, at most 2 ulp nearer.
        ulps = ott[tie].view(np.int32).astype(np.int64) - wtt[tie].view(np.int32).astype(np.int64)
        assert (ulps >= 0).all() and (ulps <= 2).all()
    else:
        assert (wtt[tie].view(np.uint32) == ott[tie].view(np.uint32)).all()
        assert tie.mean() <= 1e-3
    np.testing.assert_array_equal(wtt[both & ~tie].view(np.uint32), ott[both & ~tie].view(np.uint32))
    oh, *_ = om.intersect(osc, rays, any_hit=True, max_t=max_t)
    wh, *_ = wide_intersect(harness, world, rays, any_hit=True, max_t=max_t)
    np.testing.assert_array_equal(wh, oh)
    assert both.sum() > 20 or kind in ("one", "all_nan")
    if kind == "all_nan":
        assert both.sum() == 0


def test_infinite_or_huge_coordinates_are_refused(harness):
    """The quantised-box arithmetic would overflow into NaN and hide whole subtrees: refused at once (and not after
    2^31 iterations of a rounding guard, which is what an infinite extent used to cost)."""
    rs = np.random.default_rng(4)
    base = (rs.random((500, 1, 3)) * 4 + rs.normal(0, 0.1, (500, 3, 3))).reshape(-1, 3).astype(np.float32)
    for bad in (np.inf, -np.inf, 3e15, 1e30):
        v = base.copy()
        v[11, 1] = bad
        world = _soup_world(v, "bad")
        assert wide_digest(harness, world)[0] == -1
    v = base.copy()
    v[11, 1] = 9e14  # large but fine
    assert wide_digest(harness, _soup_world(v, "large"))[0] == 0
