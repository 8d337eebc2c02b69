"""Radiance .hdr decoding and the sky image's texel conventions (src/asset.rs:238-273) — csrc/image_io.cpp.
The test writes its own .hdr files (flat, new-style RLE, old-style run markers); expected values follow the RGBE
definition the `image` crate's HdrDecoder implements: channel * 2^(e - 136), e == 0 -> 0."""
import numpy as np
import pytest

from rust_path_tracer_b200.trace import decode_hdr, load_skybox


def to_rgbe(rgb):
    """float (H,W,3) -> uint8 (H,W,4), the usual frexp encoding."""
    rgb = np.asarray(rgb, np.float64)
    m = rgb.max(axis=-1)
    out = np.zeros(rgb.shape[:2] + (4,), np.uint8)
    mant, exp = np.frexp(m)
    ok = m > 1e-32
    scale = np.where(ok, mant * 256.0 / np.where(ok, m, 1.0), 0.0)
    out[..., :3] = np.clip(np.floor(rgb * scale[..., None]), 0, 255).astype(np.uint8)
    out[..., 3] = np.where(ok, exp + 128, 0).astype(np.uint8)
    return out


def from_rgbe(rgbe):
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e == 0, 0.0, np.ldexp(1.0, e - 136)).astype(np.float32)
    return rgbe[..., :3].astype(np.float32) * scale[..., None]


def header(w, h, signature=b"#?RADIANCE"):
    return signature + b"\n# written by the test\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=2.0\n\n" + f"-Y {h} +X {w}\n".encode()


def rle_scanline(row):
    """New-style RLE of one (W,4) scanline: 2 2 hi lo, then each component as runs / literals."""
    w = len(row)
    out = bytearray([2, 2, w >> 8, w & 255])
    for c in range(4):
        comp = row[:, c]
        x = 0
        while x < w:
            run = 1
            while x + run < w and run < 127 and comp[x + run] == comp[x]:
                run += 1
            if run >= 3:
                out += bytes([128 + run, int(comp[x])])
                x += run
            else:
                lit = 1
                while x + lit < w and lit < 128 and not (x + lit + 2 < w and comp[x + lit] == comp[x + lit + 1] == comp[x + lit + 2]):
                    lit += 1
                out += bytes([lit]) + bytes(int(v) for v in comp[x:x + lit])
                x += lit
    return bytes(out)


def image(w, h, seed=0):
    rs = np.random.default_rng(seed)
    img = rs.random((h, w, 3)) * np.exp(rs.normal(size=(h, w, 1)) * 3.0)  # several stops of range
    img[0, : w // 2] = img[0, 0]  # runs for the RLE
    img[-1, 0] = 0.0              # e == 0
    return img


def test_new_style_rle_and_flat_files_decode_to_the_rgbe_values():
    for w, h in ((37, 9), (8, 3), (5, 4)):  # width < 8 cannot be RLE-encoded: flat pixels
        rgbe = to_rgbe(image(w, h, seed=w))
        flat = header(w, h) + rgbe.tobytes()
        np.testing.assert_array_equal(decode_hdr(flat), from_rgbe(rgbe))
        if w >= 8:
            rle = header(w, h) + b"".join(rle_scanline(rgbe[y]) for y in range(h))
            assert len(rle) < len(flat) + 4 * h + 64
            np.testing.assert_array_equal(decode_hdr(rle), from_rgbe(rgbe))


def test_old_style_run_markers():
    w, h = 12, 2
    px = np.array([10, 20, 30, 130], np.uint8)
    other = np.array([200, 100, 50, 120], np.uint8)
    # row 0: px, then "repeat 7 times", then 4 x other; row 1: 12 literal pixels
    row0 = bytes(px) + bytes([1, 1, 1, 7]) + bytes(other) * 4
    row1 = bytes(other) * 12
    want = np.zeros((h, w, 4), np.uint8)
    want[0, :8] = px
    want[0, 8:] = other
    want[1] = other
    np.testing.assert_array_equal(decode_hdr(header(w, h) + row0 + row1), from_rgbe(want))


@pytest.mark.parametrize("data", [b"", b"#?RADIANCE\n", b"P6\n1 1\n255\n", header(4, 4)[:-2], header(4, 4) + b"\x00" * 10,
                                  header(4, 4).replace(b"-Y 4 +X 4", b"+X 4 -Y 4") + b"\x00" * 64,
                                  header(4, 4).replace(b"32-bit_rle_rgbe", b"32-bit_rle_xyze") + b"\x00" * 64])
def test_malformed_files_are_refused(data):
    assert decode_hdr(data) is None


def test_sky_texel_conventions(tmp_path):
    img = image(16, 8, seed=3)
    img[1, 1] = (0.5, 1.7, 300.0)
    rgbe = to_rgbe(img)
    path = tmp_path / "sky.hdr"
    path.write_bytes(header(16, 8, b"#?RGBE") + b"".join(rle_scanline(rgbe[y]) for y in range(8)))
    decoded = from_rgbe(rgbe)
    gpu = load_skybox(str(path))
    assert gpu.shape == (8, 16, 4) and gpu.dtype == np.float32
    np.testing.assert_array_equal(gpu[..., :3], decoded)  # Rgba32Float: the decoded values, alpha 1
    assert (gpu[..., 3] == 1.0).all()
    cpu = load_skybox(str(path), cpu_path_rgb8=True)  # into_rgb8: clamp, x255, round half away from zero, / 255
    q = np.floor(np.clip(decoded.astype(np.float64), 0, 1).astype(np.float32) * np.float32(255.0) + np.float32(0.5))
    np.testing.assert_array_equal(cpu[..., :3], (q / 255.0).astype(np.float32))
    assert cpu[..., :3].max() == 1.0 and (cpu[..., 3] == 1.0).all()  # the HDR range is gone on the CPU path
    assert load_skybox(str(tmp_path / "missing.hdr")) is None
