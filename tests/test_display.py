"""Display stage (display half): the tonemap operators of src/resources/render.wgsl and the 8-bit
store of save_render (src/app.rs:759-840).  CPU tests pin the restatement (oracle/display_oracle.py) to the operators'
known answers; GPU tests compare the device resolve (rpt_read_display, rpt_read_display_rgba8) with it."""
import ctypes as C

import numpy as np
import pytest

import display_oracle as disp
import helpers

f32 = np.float32


def test_operator_known_answers():
    one = np.ones((1, 3), f32)
    zero = np.zeros((1, 3), f32)
    assert np.array_equal(disp.tonemap(one * f32(0.37), 0), one * f32(0.37))               # None: untouched
    assert np.array_equal(disp.tonemap(one * f32(0.37), 9), one * f32(0.37))               # default arm of the switch
    assert np.array_equal(disp.tonemap(one, 1), one * f32(0.5))                            # Reinhard(1) = 1/2
    np.testing.assert_allclose(disp.tonemap(one * f32(5.6), 6), one, rtol=3e-7)            # Uncharted(W / exposure bias) = 1
    for op in (1, 2, 3, 5, 6):
        np.testing.assert_allclose(disp.tonemap(zero, op), zero, atol=2e-8)                # curves through the origin
    assert np.array_equal(disp.tonemap(one * f32(1e4), 3), one)                            # Narkowicz saturates (clamped)
    np.testing.assert_allclose(disp.tonemap(one, 3), one * f32((2.51 + 0.03) / (2.43 + 0.59 + 0.14)), rtol=3e-7)
    np.testing.assert_allclose(disp.tonemap(one / f32(0.6), 2), disp.tonemap(one, 3), rtol=1e-6)  # op 2 = op 3 on 0.6 x
    # Hill: grey stays (nearly) grey — the input and output matrices' rows each sum to ~1 — and it saturates at 1
    grey = disp.tonemap(one * f32(0.18), 4)
    assert np.ptp(grey) < 2e-3 and 0.05 < grey[0, 0] < 0.25
    assert np.array_equal(disp.tonemap(one * f32(1e4), 4), one)
    # Neutral: the curve is scaled so that its white level ends at 1: neutral(5.3 / ws) = curve(5.3) * ws = 1
    ws = f32(1.0) / disp._filmic_curve(f32(5.3), 0.2, 0.29, 0.24, 0.272, 0.02, 0.3)
    np.testing.assert_allclose(disp.tonemap(one * (f32(5.3) / ws), 5), one, rtol=1e-6)


def test_operators_are_monotonic_and_float32():
    x = np.linspace(0, 16, 4097, dtype=f32)
    rgb = np.stack([x, x, x], axis=-1)
    for op in range(7):
        y = disp.tonemap(rgb, op)
        assert y.dtype == f32
        assert np.all(np.diff(y[:, 0].astype(np.float64)) >= -1e-6), disp.TONEMAPS[op]


def test_rgba8_store():
    rgb = np.array([[0.0, 0.5, 1.0], [-1.0, 2.0, np.nan], [0.0031308, 0.2, 0.73536]], f32)
    lin = disp.to_rgba8(rgb, srgb=False)
    assert lin.tolist() == [[0, 128, 255, 255], [0, 255, 0, 255], [1, 51, 188, 255]]  # 127.5 rounds to even = 128
    enc = disp.to_rgba8(rgb, srgb=True)
    assert enc[0].tolist() == [0, 188, 255, 255]   # sRGB(0.5) = 0.7354 -> 188
    assert enc[1].tolist() == [0, 255, 0, 255]
    assert enc[2, 0] == 10                         # 12.92 * 0.0031308 * 255 = 10.3 (the linear toe)


# ---- device resolve ---------------------------------------------------------------------------------------------

def _rendered(scene="DarkCornell", w=160, h=96, spp=8):
    from rust_path_tracer_b200.trace import Renderer
    r = Renderer(0)
    r.upload_world(helpers.world(scene)); r.set_config(helpers.config(w, h, 1)); r.write_rng(helpers.seeds(w, h))
    r.enqueue(spp)
    return r


@pytest.mark.gpu
def test_framebuffer_is_the_ieee_quotient():
    """src/trace.rs:199-204 divides on the host: the device normalisation must be that exact float32 quotient."""
    with _rendered() as r:
        out = r.read_output()
        for samples in (8.0, 3.0, 7.0):
            want = (out[:, :3] / f32(samples)).reshape(-1)
            np.testing.assert_array_equal(r.read_framebuffer(samples), want)


@pytest.mark.gpu
@pytest.mark.parametrize("op", range(8))
def test_display_matches_restatement(op):
    with _rendered() as r:
        # spread the accumulator over the operators' whole range (dark corners to 40x over-exposure, NaN, negatives)
        out = r.read_output().copy()
        out[::7, :3] *= f32(40.0)
        out[5, :3] = np.nan
        out[6, :3] = -1.0
        r.write_output(out)
        got = r.read_display(8.0, op).reshape(-1, 3)
        want = disp.display(out, 8.0, op)
        np.testing.assert_array_equal(got, want)  # float32, source order, IEEE division on both sides: bit for bit
        helpers.record_parity(f"display op {op} ({disp.TONEMAPS[op] if op < 7 else 'default arm'})", f32_mismatches=0)
        for srgb in (False, True):
            got8 = r.read_display_rgba8(8.0, op, srgb)
            want8 = disp.to_rgba8(want, srgb)
            diff = np.abs(got8.astype(np.int32) - want8.astype(np.int32))
            assert diff.max() <= (1 if srgb else 0)       # powf may differ by an ulp at a rounding boundary
            assert (diff != 0).mean() < 1e-3
            assert np.all(got8[:, 3] == 255)


@pytest.mark.gpu
def test_display_by_name_and_errors():
    from rust_path_tracer_b200 import capi
    with _rendered(w=64, h=64, spp=2) as r:
        np.testing.assert_array_equal(r.read_display(2.0, "reinhard"), r.read_display(2.0, 1))
        np.testing.assert_array_equal(r.read_display(2.0, "none"), r.read_framebuffer(2.0))
        with pytest.raises(capi.RptError) as e:
            r._call("rpt_read_display", capi.ptr(np.empty(30, np.float32)), C.c_size_t(10), C.c_float(2.0), C.c_uint32(1))
        assert e.value.code == capi.ERR_SIZE_MISMATCH
        with pytest.raises(capi.RptError) as e:
            r._call("rpt_read_display_rgba8", capi.ptr(np.empty(40, np.uint8)), C.c_size_t(10), C.c_float(2.0), C.c_uint32(1), C.c_uint32(1))
        assert e.value.code == capi.ERR_SIZE_MISMATCH
