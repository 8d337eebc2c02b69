"""The traversal structure built and refitted ON THE DEVICE (build half): rpt_upload_world without a
reference BVH, rpt_refit_world.  Nearest hits do not depend on the tree, so the bars are the usual ones: primary ids vs
the oracle within the 1e-4 budget, radiance MAE <= 1e-3."""
import time

import numpy as np
import pytest

import helpers
import oracle as om
from rust_path_tracer_b200 import capi
from rust_path_tracer_b200.glb import BakedScene
from rust_path_tracer_b200.trace import Renderer
from rust_path_tracer_b200.world import World

pytestmark = pytest.mark.gpu
ID_BUDGET, MAE_TOLERANCE = 1e-4, 1e-3


def render(world, cfg, seeds, spp, build_on_device, sky=None, refit=None):
    with Renderer(0) as r:
        r.upload_world(world, sky, build_on_device=build_on_device)
        if refit is not None:
            r.refit_world(*refit)
        r.set_config(cfg); r.write_rng(seeds)
        ids = r.read_primary_ids()
        r.enqueue(spp)
        return r.read_output(), ids, r.counters()


@pytest.mark.parametrize("scene,nee", [("FurnaceTest", 1), ("DarkCornell", 1), ("PBRTest", 0), ("VeachMIS", 1)])
def test_device_built_tree_matches_the_oracle(scene, nee):
    world = helpers.world(scene)
    w, h, spp = 160, 96, 8
    cfg, seeds = helpers.config(w, h, nee), helpers.seeds(w, h)
    osc = om.OracleScene(world)
    _, _, _, o_ids = om.trace(cfg, osc, seeds, 1, want_primary_ids=True)
    o_out, _, o_ctr, _ = om.trace(cfg, osc, seeds, spp)
    d_out, d_ids, d_ctr = render(world, cfg, seeds, spp, build_on_device=True)
    h_out, h_ids, h_ctr = render(world, cfg, seeds, spp, build_on_device=False)
    mismatch = float((d_ids != o_ids).mean())
    err = helpers.mae(d_out[:, :3] / spp, o_out[:, :3] / spp)[0]
    helpers.record_parity(f"{scene} {w}x{h} {spp}spp, tree built on the device", id_mismatch=mismatch, mae=err,
                          rays_vs_host_tree=d_ctr["nearest_rays"] / h_ctr["nearest_rays"])
    assert mismatch <= ID_BUDGET and err <= MAE_TOLERANCE
    assert float((d_ids != h_ids).mean()) <= ID_BUDGET
    assert (d_out[:, 3] == spp).all()


def test_megakernel_arm_needs_the_reference_bvh():
    world = helpers.world("DarkCornell")
    with Renderer(0, capi.PIPELINE_MEGAKERNEL) as r:
        r.upload_world(world, build_on_device=True)
        r.set_config(helpers.config(32, 32, 0)); r.write_rng(helpers.seeds(32, 32))
        with pytest.raises(capi.RptError) as e:
            r.enqueue(1)
        assert e.value.code == capi.ERR_NOT_READY


def _soup(n, seed, coincident=False):
    """n random small triangles in front of the camera (optionally many with the SAME centroid: Morton ties)."""
    from rust_path_tracer_b200.glb import MATERIAL_DTYPE

    rs = np.random.default_rng(seed)
    centre = rs.uniform([-2, -0.5, 1], [2, 2.5, 6], (n, 3))
    if coincident:
        centre[: n // 2] = centre[0]
    ofs = rs.normal(size=(n, 3, 3)) * 0.15
    ofs -= ofs.mean(axis=1, keepdims=True)  # the centroid stays where it is
    v = (centre[:, None, :] + ofs).reshape(-1, 3).astype(np.float32)
    verts = np.concatenate([v, np.ones((len(v), 1), np.float32)], 1)
    nrm = np.tile(np.array([[0, 0, -1, 0]], np.float32), (len(v), 1))
    tris = np.concatenate([np.arange(3 * n, dtype=np.uint32).reshape(n, 3), np.zeros((n, 1), np.uint32)], 1)
    mats = np.zeros(1, MATERIAL_DTYPE)
    mats[0]["albedo"] = (0.6, 0.6, 0.6, 1)
    mats[0]["roughness"] = 1.0
    return BakedScene(verts, nrm, np.zeros_like(nrm), np.zeros((len(v), 2), np.float32), tris, mats)


@pytest.mark.parametrize("n,coincident", [(1, False), (3, False), (4, False), (40, True), (500, False), (3000, True)])
def test_small_and_degenerate_scenes(n, coincident):
    world = World.from_baked(_soup(n, seed=n, coincident=coincident))
    cfg, seeds = helpers.config(96, 64, 0), helpers.seeds(96, 64)
    _, _, _, o_ids = om.trace(cfg, om.OracleScene(world), seeds, 1, want_primary_ids=True)
    out, ids, _ = render(world, cfg, seeds, 2, build_on_device=True)
    assert float((ids != o_ids).mean()) <= 2e-3  # (overlapping random triangles: a few exact-t / edge ties at this frame size)
    assert np.isfinite(out).all()


def _scaled(scene: BakedScene, k: float) -> BakedScene:
    v = scene.vertices.copy()
    v[:, :3] *= np.float32(k)
    return BakedScene(v, scene.normals, scene.tangents, scene.uvs, scene.indices.copy(), scene.materials.copy(), [dict() for _ in scene.materials])


@pytest.mark.parametrize("build_on_device", [False, True], ids=["host_collapsed_tree", "device_built_tree"])
def test_refit_after_an_exact_scaling_matches_the_oracle(build_on_device):
    """Scaling by 2 is exact in binary floating point, so the reference's SAH build of the scaled scene makes the same
    choices and permutes the index buffer the same way: triangle ids of the refitted world and of the oracle's scaled
    world are comparable."""
    import os

    base = BakedScene.load(os.path.join(helpers.SCENE_DIR, "DarkCornell.npz"))
    world = World.from_baked(base)
    big = World.from_baked(_scaled(base, 2.0))
    np.testing.assert_array_equal(world.index_buffer, big.index_buffer)
    w, h, spp = 128, 96, 8
    cfg = helpers.config(w, h, 1, cam_position=[0.0, 2.0, -10.0, 0.0])
    seeds = helpers.seeds(w, h)
    osc = om.OracleScene(big)
    _, _, _, o_ids = om.trace(cfg, osc, seeds, 1, want_primary_ids=True)
    o_out, _, _, _ = om.trace(cfg, osc, seeds, spp)
    out, ids, _ = render(world, cfg, seeds, spp, build_on_device, refit=(big.per_vertex_buffer, big.light_pick_buffer))
    assert float((ids != o_ids).mean()) <= ID_BUDGET
    assert helpers.mae(out[:, :3] / spp, o_out[:, :3] / spp)[0] <= MAE_TOLERANCE
    assert (o_ids != 0xFFFFFFFF).mean() > 0.5  # the camera does look at the scaled scene


@pytest.mark.parametrize("build_on_device", [False, True], ids=["host_collapsed_tree", "device_built_tree"])
def test_refit_after_a_deformation_matches_a_fresh_build(build_on_device):
    """Vertices pushed around (not a similarity): the refitted tree against a tree built from scratch for the deformed
    vertices over the SAME index buffer."""
    world = helpers.world("PBRTest")
    moved = world.per_vertex_buffer.copy()
    p = moved["vertex"][:, :3]
    p += (0.25 * np.sin(3.0 * p[:, [1, 2, 0]] + 1.0)).astype(np.float32)
    deformed = World(moved, world.index_buffer, None, world.material_data_buffer, world.light_pick_buffer, world.atlas)
    w, h, spp = 160, 96, 4
    cfg, seeds = helpers.config(w, h, 0), helpers.seeds(w, h)
    fresh_out, fresh_ids, _ = render(deformed, cfg, seeds, spp, build_on_device=True)
    out, ids, _ = render(world, cfg, seeds, spp, build_on_device, refit=(moved, None))
    assert float((ids != fresh_ids).mean()) <= ID_BUDGET
    assert helpers.mae(out[:, :3] / spp, fresh_out[:, :3] / spp)[0] <= MAE_TOLERANCE
    assert (ids != helpers_ids(world, cfg, seeds)).mean() > 0.05  # and the deformation did change the picture


def helpers_ids(world, cfg, seeds):
    with Renderer(0) as r:
        r.upload_world(world); r.set_config(cfg); r.write_rng(seeds)
        return r.read_primary_ids()


def test_start_up_with_a_device_build_of_the_proxy():
    """BASELINE's start-up bench shape: time from buffers on the host to a context ready to trace, 1M-triangle proxy."""
    import bench

    world, cfg, seeds, _spp, _label, _scene, sky = bench.load_workload("breaktime")
    times = {}
    for mode in (True, False, True):
        with Renderer(0) as r:
            t0 = time.perf_counter()
            r.upload_world(world, sky, build_on_device=mode)
            r.sync()
            times[mode] = time.perf_counter() - t0
            r.set_config(helpers.config(320, 180, 1, has_skybox=1)); r.write_rng(helpers.seeds(320, 180))
            ids = r.read_primary_ids()
            r.enqueue(4)
            out = r.read_output()
            ctr = r.counters()
        if mode:
            d_ids, d_out = ids, out
    helpers.record_parity("BreakTime proxy start-up (upload + tree), seconds", device_build_s=times[True], host_collapse_s=times[False],
                          triangles=int(world.ntriangles))
    assert float((d_ids != ids).mean()) <= ID_BUDGET
    assert helpers.mae(d_out[:, :3] / 4, out[:, :3] / 4)[0] <= MAE_TOLERANCE
    assert times[True] < 0.5, times


@pytest.mark.parametrize("kind", ["some_nan", "all_nan", "huge_range"])
def test_non_finite_and_extreme_vertices(kind):
    """NaN coordinates are ignored by every min / max and such triangles are never hit (as in the reference); a scene
    spanning many orders of magnitude still quantises conservatively; infinite coordinates are refused."""
    scene = _soup(600, seed=5)
    v = scene.vertices.copy()
    if kind == "some_nan":
        v[::97, 0] = np.nan
    elif kind == "all_nan":
        v[:, :3] = np.nan
    else:
        v[:300, :3] *= np.float32(1e-3)   # a cloud of tiny triangles around the origin ...
        v[300:330, :3] *= np.float32(3e3)  # ... and a few enormous ones
    scene = BakedScene(v, scene.normals, scene.tangents, scene.uvs, scene.indices, scene.materials)
    world = World.from_baked(scene)
    cfg, seeds = helpers.config(96, 64, 0), helpers.seeds(96, 64)
    _, _, _, o_ids = om.trace(cfg, om.OracleScene(world), seeds, 1, want_primary_ids=True)
    out, ids, _ = render(world, cfg, seeds, 2, build_on_device=True)
    host_out, host_ids, _ = render(world, cfg, seeds, 2, build_on_device=False)
    np.testing.assert_array_equal(ids, host_ids)  # both of the product's trees agree exactly ...
    assert float((ids != o_ids).mean()) <= 2e-3   # ... and with the reference up to ties between overlapping random triangles
    if kind == "all_nan":
        assert (ids == 0xFFFFFFFF).all()
    inf = v.copy()
    inf[5, 1] = np.inf
    bad = World(world.per_vertex_buffer.copy(), world.index_buffer, None, world.material_data_buffer, world.light_pick_buffer)
    bad.per_vertex_buffer["vertex"][5, 1] = np.inf
    with Renderer(0) as r:
        with pytest.raises(capi.RptError) as e:
            r.upload_world(bad, build_on_device=True)
        assert e.value.code == capi.ERR_INVALID_ARGUMENT
