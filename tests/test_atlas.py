"""Atlas packing rules of src/atlas.rs and the oracle's texture fetch (image_polyfill.rs)."""
import numpy as np

import helpers
import oracle as om
from rust_path_tracer_b200 import atlas as atlas_mod


def test_quadtree_rects_follow_the_reference_rule():
    # 1 texture: the root is split once (queue.len() <= textures.len() holds for 1), first quadrant used
    assert atlas_mod.packing_rects(1, 4096, 4096) == [(0, 0, 2048, 2048)]
    r = atlas_mod.packing_rects(4, 4096, 4096)  # 4 textures: split until 7 leaves, four largest first
    assert len(r) == 4 and all(w == 2048 for _, _, w, _ in r[:3])
    r = atlas_mod.packing_rects(64, 4096, 4096)
    assert len(r) == 64 and len({(x, y) for x, y, _, _ in r}) == 64 and all(w >= 256 for _, _, w, _ in r)


def test_pack_flips_vertically_and_reports_uv_rect():
    tex = np.zeros((2048, 2048, 4), np.uint8)
    tex[0, :, 0] = 200  # top row marked
    atlas, sts = atlas_mod.pack_textures([tex], 4096, 4096)
    assert atlas[2047, 0, 0] == 200 and atlas[0, 0, 0] == 0  # flipv (src/atlas.rs:85)
    np.testing.assert_array_equal(sts[0], np.array([0, 0, 0.5, 0.5], np.float32))


def test_albedo_gamma_decode_is_8bit_truncating():
    tex = np.full((1, 3, 4), 128, np.uint8)
    out = atlas_mod.decode_albedo_gamma(tex)
    assert out[0, 0, 0] == int(np.float32(np.float32(128 / 255) ** np.float32(2.2)) * np.float32(255.0)) and out[0, 0, 3] == 255


def test_textured_scene_traces_and_uses_the_atlas():
    w = helpers.textured_world()
    assert w.atlas is not None and w.material_data_buffer["has_albedo_texture"][:25].all()
    plain = helpers.world("PBRTest")
    cfg = helpers.config(64, 36, 0)
    a, *_ = om.trace(cfg, om.OracleScene(w), helpers.seeds(64, 36), 4)
    b, *_ = om.trace(cfg, om.OracleScene(plain), helpers.seeds(64, 36), 4)
    assert np.isfinite(a).all() and np.abs(a[:, :3] - b[:, :3]).mean() > 1e-3  # textures change the image


def test_lanczos_resize_matches_the_published_algorithm():
    """A texture that is not leaf-sized goes through the Lanczos3 convolution (src/atlas.rs:71-83, fast_image_resize);
    the restatement follows Pillow's published resampling, so Pillow is the checker here (test-only dependency)."""
    from PIL import Image

    rs = np.random.default_rng(3)
    for shape in ((96, 160), (300, 200), (256, 256), (700, 513)):
        tex = rs.integers(0, 256, shape + (4,), dtype=np.uint8)
        tex[: shape[0] // 2] //= 4  # some structure, not only noise
        atlas, sts = atlas_mod.pack_textures([tex], 512, 512)
        # channel by channel: Pillow premultiplies alpha when it resizes an RGBA image, the reference's resizer does not
        want = np.stack([np.asarray(Image.fromarray(tex[..., c], "L").resize((256, 256), Image.LANCZOS), np.uint8) for c in range(4)], axis=2)[::-1]
        got = atlas[:256, :256]
        diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
        assert diff.max() <= 1, (shape, int(diff.max()))  # same taps; rounding of the 8-bit intermediate may differ by 1
        assert (atlas[256:] == 0).all() and (atlas[:, 256:] == 0).all()
        np.testing.assert_array_equal(sts[0], np.array([0, 0, 0.5, 0.5], np.float32))


def test_atlas_errors_are_status_codes():
    import ctypes as C

    from rust_path_tracer_b200 import capi

    assert capi.lib().rpt_atlas_rects(C.c_uint32(1), C.c_uint32(0), C.c_uint32(16), None) == capi.ERR_INVALID_ARGUMENT
    rects = np.zeros((100, 4), np.uint32)  # 100 textures cannot fit a 4x4 atlas: leaves would be empty
    assert capi.lib().rpt_atlas_rects(C.c_uint32(100), C.c_uint32(4), C.c_uint32(4), capi.ptr(rects)) == capi.ERR_UNSUPPORTED
