"""CPU restatement of the reference's display stage.  TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

Follows src/resources/render.wgsl (the fragment stage fs_main, :150-185, and the operators above it) with every
operation rounded to float32 in source order — numpy float32 arrays do exactly that and never contract a
multiply-add — and the colour-attachment store `save_render` (src/app.rs:759-840) reads back.

Parity status: WGSL leaves the precision of `/` and of the transfer function open and the reference has no test or
golden image for this stage, so the restatement is pinned only by the operators' published known answers
(tests/test_display.py: Reinhard(1) = 1/2, Uncharted(11.2 / 2) = 1, curves through 0, Narkowicz saturation, the Hill
fit's grey point) — "parity unpinned" beyond those.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
TONEMAPS = ["none", "reinhard", "aces_narkowicz", "aces_narkowicz_overexposed", "aces_hill", "neutral", "uncharted"]  # src/app.rs:20-28


def _clamp01(x):
    return np.fmin(np.fmax(x, f32(0.0)), f32(1.0))  # fmax/fmin drop NaN: NaN -> 0


def aces_narkowicz(x):
    """render.wgsl:36-43"""
    a, b, c, d, e = f32(2.51), f32(0.03), f32(2.43), f32(0.59), f32(0.14)
    return _clamp01((x * (a * x + b)) / (x * (c * x + d) + e))


def aces_hill(rgb):
    """render.wgsl:46-71; the WGSL builds both matrices as transpose(mat3x3(rows))."""
    m_in = np.array([[0.59719, 0.35458, 0.04823], [0.07600, 0.90834, 0.01566], [0.02840, 0.13383, 0.83777]], f32)
    m_out = np.array([[1.60475, -0.53108, -0.07367], [-0.10208, 1.10813, -0.00605], [-0.00327, -0.07276, 1.07602]], f32)

    def mul(m, v):  # row . v, summed left to right
        return np.stack([(m[r, 0] * v[..., 0] + m[r, 1] * v[..., 1]) + m[r, 2] * v[..., 2] for r in range(3)], axis=-1)

    c = mul(m_in, rgb)
    a = c * (c + f32(0.0245786)) - f32(0.000090537)
    b = c * (f32(0.983729) * c + f32(0.4329510)) + f32(0.238081)
    return _clamp01(mul(m_out, a / b))


def _filmic_curve(x, a, b, c, d, e, f):
    """render.wgsl:77-79 (neutralCurve) and :104-112 (unchartedPartial): one curve, two constant sets."""
    a, b, c, d, e, f = (f32(v) for v in (a, b, c, d, e, f))
    return ((x * (a * x + c * b) + d * e) / (x * (a * x + b) + d * f)) - e / f


def neutral(x):
    """render.wgsl:81-102"""
    k = (0.2, 0.29, 0.24, 0.272, 0.02, 0.3)
    white_scale = f32(1.0) / _filmic_curve(f32(5.3), *k)
    return (_filmic_curve(x * white_scale, *k) * white_scale) / f32(1.0)


def uncharted(x):
    """render.wgsl:114-121"""
    k = (0.15, 0.50, 0.10, 0.20, 0.02, 0.30)
    return _filmic_curve(x * f32(2.0), *k) * (f32(1.0) / _filmic_curve(f32(11.2), *k))


def tonemap(rgb: np.ndarray, op: int) -> np.ndarray:
    """The switch of fs_main (render.wgsl:162-184) on an (..., 3) float32 array."""
    rgb = np.asarray(rgb, f32)
    with np.errstate(all="ignore"):
        if op == 1:
            return rgb / (rgb + f32(1.0))
        if op == 2:
            return aces_narkowicz(rgb * f32(0.6))
        if op == 3:
            return aces_narkowicz(rgb)
        if op == 4:
            return aces_hill(rgb)
        if op == 5:
            return neutral(rgb)
        if op == 6:
            return uncharted(rgb)
    return rgb


def display(output: np.ndarray, samples: float, op: int) -> np.ndarray:
    """(npixels, 4) accumulator -> (npixels, 3) displayed colour (src/trace.rs:199-204, then fs_main)."""
    with np.errstate(all="ignore"):
        return tonemap(np.asarray(output, f32)[:, :3] / f32(samples), op)


def to_rgba8(rgb: np.ndarray, srgb: bool) -> np.ndarray:
    """Store to a (s)RGB 8-bit unorm attachment, alpha 1, R,G,B,A byte order (src/app.rs:829 swizzle applied)."""
    x = _clamp01(np.asarray(rgb, f32))
    if srgb:
        x = np.where(x <= f32(0.0031308), f32(12.92) * x, f32(1.055) * np.power(x, f32(1.0) / f32(2.4)) - f32(0.055)).astype(f32)
    out = np.full(x.shape[:-1] + (4,), 255, np.uint8)
    out[..., :3] = np.rint(x * f32(255.0)).astype(np.uint8)
    return out
