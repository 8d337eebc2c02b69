"""The .glb import (rust-path-tracer_b200/glb.py, row f1 / f3: what `World::from_path` + assimp produce, src/asset.rs:55-175)
on a file the test writes itself: a textured quad under a scaled, translated node plus a primitive without a material.
This is the path a real BreakTime.glb would take (the shipped scenes have no textures; the fixtures are baked .npz)."""
import io
import json
import struct

import numpy as np

from rust_path_tracer_b200.glb import load_glb
from rust_path_tracer_b200.world import World


def _png(rgba):
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(rgba, "RGBA").save(buf, format="PNG")
    return buf.getvalue()


def _write_glb(path):
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    rs = np.random.default_rng(1)
    albedo = rs.integers(0, 256, (4, 4, 4), dtype=np.uint8); albedo[..., 3] = 255
    mr = rs.integers(0, 256, (2, 2, 4), dtype=np.uint8); mr[..., 3] = 255
    nm = np.full((2, 2, 4), (128, 128, 255, 255), np.uint8)
    blobs = [pos.tobytes(), nrm.tobytes(), uv.tobytes(), idx.tobytes(), _png(albedo), _png(mr), _png(nm)]
    views, binary = [], b""
    for b in blobs:
        binary += b"\0" * (-len(binary) % 4)
        views.append({"buffer": 0, "byteOffset": len(binary), "byteLength": len(b)})
        binary += b
    binary += b"\0" * (-len(binary) % 4)
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"mesh": 0, "translation": [1.0, 2.0, 3.0], "scale": [2.0, 2.0, 2.0]}],
        "meshes": [{"primitives": [
            {"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 2}, "indices": 3, "material": 0},
            {"attributes": {"POSITION": 0, "NORMAL": 1}, "indices": 3}]}],
        "accessors": [
            {"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3", "min": [0, 0, 0], "max": [1, 1, 0]},
            {"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC3"},
            {"bufferView": 2, "componentType": 5126, "count": 4, "type": "VEC2"},
            {"bufferView": 3, "componentType": 5123, "count": 6, "type": "SCALAR"}],
        "bufferViews": views, "buffers": [{"byteLength": len(binary)}],
        "images": [{"bufferView": 4, "mimeType": "image/png"}, {"bufferView": 5, "mimeType": "image/png"}, {"bufferView": 6, "mimeType": "image/png"}],
        "textures": [{"source": 0}, {"source": 1}, {"source": 2}],
        "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.25, 1.0, 1.0], "baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 1}, "roughnessFactor": 0.5},
                       "normalTexture": {"index": 2}, "emissiveFactor": [0.0, 1.0, 0.5]}],
    }
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    with open(path, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(binary)))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(binary), 0x004E4942) + binary)
    return albedo, mr, nm


def test_textured_glb_bakes_like_the_reference_import(tmp_path):
    path = str(tmp_path / "quad.glb")
    albedo, mr, nm = _write_glb(path)
    scene = load_glb(path)
    # two primitives x 4 vertices, world space = 2 * p + (1, 2, 3), stored (x, z, y, 1) (src/asset.rs:101-104)
    assert scene.vertices.shape == (8, 4)
    np.testing.assert_array_equal(scene.vertices[:4], np.array([[1, 3, 2, 1], [3, 3, 2, 1], [3, 3, 4, 1], [1, 3, 4, 1]], np.float32))
    # winding (f0, f2, f1), material in .w; the primitive without a material uses the appended default one (assimp)
    np.testing.assert_array_equal(scene.indices, np.array([[0, 2, 1, 0], [0, 3, 2, 0], [4, 6, 5, 1], [4, 7, 6, 1]], np.uint32))
    # normals: node rotation only (scale divided out), swizzled like positions
    np.testing.assert_allclose(scene.normals[:, :3], np.tile([[0, 1, 0]], (8, 1)), atol=1e-6)
    # V flipped (assimp's glTF2 importer); the primitive without TEXCOORD_0 gets zeros
    np.testing.assert_array_equal(scene.uvs[:4], np.array([[0, 1], [1, 1], [1, 0], [0, 0]], np.float32))
    np.testing.assert_array_equal(scene.uvs[4:], np.zeros((4, 2), np.float32))
    m = scene.materials
    assert len(m) == 2
    np.testing.assert_array_equal(m[0]["albedo"], np.array([0.5, 0.25, 1.0, 1.0], np.float32))
    np.testing.assert_array_equal(m[0]["emissive"][:3], np.array([0.0, 15.0, 7.5], np.float32))  # x 15, src/asset.rs:165-168
    assert m[0]["metallic"][0] == 1.0 and m[0]["roughness"][0] == 0.5  # metallicFactor defaults to 1
    np.testing.assert_array_equal(m[1]["albedo"], np.ones(4, np.float32))
    assert m[1]["metallic"][0] == 1.0 and m[1]["roughness"][0] == 1.0 and not m[1]["emissive"][:3].any()
    assert set(scene.textures[0]) == {"albedo", "metallic", "roughness", "normals"} and scene.textures[1] == {}
    np.testing.assert_array_equal(scene.textures[0]["albedo"], albedo)
    np.testing.assert_array_equal(scene.textures[0]["roughness"], mr)
    np.testing.assert_array_equal(scene.textures[0]["normals"], nm)


def test_textured_glb_becomes_a_world_with_an_atlas(tmp_path):
    path = str(tmp_path / "quad.glb")
    _write_glb(path)
    world = World.from_path(path)
    assert world is not None and world.ntriangles == 4
    mat = world.material_data_buffer
    assert mat["has_albedo_texture"][0] == 1 and mat["has_metallic_texture"][0] == 1 and mat["has_roughness_texture"][0] == 1 and mat["has_normal_texture"][0] == 1
    assert not mat["has_albedo_texture"][1]
    assert world.atlas is not None and world.atlas.dtype == np.uint8 and world.atlas.shape[2] == 4 and world.atlas[..., :3].any()
    # the four rects (u0, v0, su, sv) the materials carry lie inside the atlas and do not coincide
    rects = [tuple(np.round(mat[k][0], 6)) for k in ("albedo", "metallic", "roughness", "normals")]
    assert len(set(rects)) == 4
    for u0, v0, su, sv in rects:
        assert 0 <= u0 < 1 and 0 <= v0 < 1 and 0 < su <= 1 and 0 < sv <= 1 and u0 + su <= 1 + 1e-6 and v0 + sv <= 1 + 1e-6
    # emissive triangles made it into the light table; the BVH covers every triangle exactly once
    assert len(world.light_pick_buffer) >= 1 and world.light_pick_buffer["ratio"][0] >= 0
    leaves = world.nodes[world.nodes["triangle_count"] > 0]
    assert leaves["triangle_count"].sum() == 4
    assert World.from_path(str(tmp_path / "missing.glb")) is None


def test_oracle_renders_the_textured_glb(tmp_path):
    import helpers
    import oracle as om

    path = str(tmp_path / "quad.glb")
    _write_glb(path)
    world = World.from_path(path)
    cfg = helpers.config(32, 18, 1)
    cfg.cam_position[:] = [2.0, 6.0, 3.0, 0.0]   # above the quad (which lies in the plane y = 3 after the swizzle) ...
    cfg.cam_rotation[:] = [1.5707964, 0.0, 0.0, 0.0]  # ... looking straight down
    out, _, ctr, ids = om.trace(cfg, om.OracleScene(world), helpers.seeds(32, 18), 4, want_primary_ids=True)
    assert np.isfinite(out).all() and (out[:, 3] == 4).all()
    assert (ids != 0xFFFFFFFF).any() and out[:, :3].max() > 0  # the emissive, textured quad is in view
