"""The C-ABI library loads, exports every symbol the headers declare, and keeps the reference's
record layouts.  No compute calls (runs without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rust_path_tracer_b200 import capi
from rust_path_tracer_b200.glb import MATERIAL_DTYPE, VERTEX_DTYPE

INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")


def declared_symbols():
    names = set()
    for fn in os.listdir(INCLUDE):
        text = open(os.path.join(INCLUDE, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(rpt_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_headers_and_binding_agree():
    assert declared_symbols() == sorted(capi.HOST_SYMBOLS + capi.DEVICE_SYMBOLS)


@pytest.mark.parametrize("name", declared_symbols())
def test_symbol_exported(name):
    assert hasattr(capi.lib(), name)


def test_record_layouts():
    # shared_structs/src/lib.rs: TracingConfig 80, MaterialData 96, PerVertexData 64, LightPickEntry 28, BVHNode 32
    assert C.sizeof(capi.TracingConfig) == 80
    assert capi.TracingConfig.width.offset == 32 and capi.TracingConfig.sun_direction.offset == 48
    assert capi.TracingConfig.nee.offset == 64 and capi.TracingConfig.specular_weight_clamp.offset == 72
    assert MATERIAL_DTYPE.itemsize == 96 and VERTEX_DTYPE.itemsize == 64
    assert capi.LIGHT_DTYPE.itemsize == 28 and capi.BVH_NODE_DTYPE.itemsize == 32


def test_default_config_matches_reference():
    cfg = capi.TracingConfig.default()
    assert (cfg.width, cfg.height, cfg.min_bounces, cfg.max_bounces, cfg.nee, cfg.has_skybox) == (1280, 720, 3, 4, 0, 0)
    assert list(cfg.cam_position) == [0.0, 1.0, -5.0, 0.0]
    sun = np.array(cfg.sun_direction[:3], np.float64)
    assert abs(np.linalg.norm(sun) - 1.0) < 1e-6 and cfg.sun_direction[3] == 15.0
    np.testing.assert_allclose(sun, np.array([0.5, 1.3, 1.0]) / np.linalg.norm([0.5, 1.3, 1.0]), rtol=1e-6)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the backend must refuse to create a context (no silent CPU path)."""
    import torch

    ctx = C.c_void_p()
    code = capi.lib().rpt_create(C.c_int(0), C.byref(ctx))
    if torch.cuda.is_available():
        assert code == capi.OK
        capi.lib().rpt_destroy(ctx)
    else:
        assert code == capi.ERR_NO_DEVICE
        assert b"no CPU fallback" in capi.lib().rpt_last_error(None)


def test_null_arguments_are_status_codes():
    lib = capi.lib()
    assert lib.rpt_create(C.c_int(0), None) == capi.ERR_INVALID_ARGUMENT
    assert lib.rpt_destroy(None) == capi.ERR_INVALID_ARGUMENT
    assert lib.rpt_enqueue(None, C.c_uint32(1)) == capi.ERR_INVALID_ARGUMENT
    assert lib.rpt_build_bvh(None, 0, None, 0, 128, None, None) == capi.ERR_INVALID_ARGUMENT


def test_staging_block_outlives_the_array_it_was_made_for():
    """`pinned_empty` hands out views of a block whose owner must live as long as ANY view does (a reshaped
    framebuffer outliving the original array was freed under it once).  Checked on the lifetime logic itself with a
    malloc-free stand-in for the page-locked block, so it runs without a GPU."""
    import gc

    freed = []

    class Block:
        def __init__(self, nbytes):
            self.store = np.zeros(nbytes, np.uint8)
            self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.store.ctypes.data, False), "version": 3}

        def __del__(self):
            freed.append(True)

    arr = capi._array_over(Block(4 * 6), (2, 3), np.float32)
    view = arr.reshape(3, 2)[1:]
    del arr
    gc.collect()
    assert not freed, "block released while a view of it is alive"
    view[...] = 1.0
    del view
    gc.collect()
    assert freed


def test_integration_doc_binds_every_declared_symbol():
    """INTEGRATION.md shows the Rust `extern "C"` block a maintainer of the reference would add: it has to name every
    function include/*.h declares (and nothing the library does not export)."""
    import helpers

    declared = set()
    for header in ("rpt_b200.h", "rpt_host.h"):
        text = open(os.path.join(helpers.REPO, "include", header)).read()
        declared |= set(re.findall(r"^(?:int|const char\*) (rpt_[a-z0-9_]+)\(", text, flags=re.M))
    bound = set(re.findall(r"pub fn (rpt_[a-z0-9_]+)\(", open(os.path.join(helpers.REPO, "INTEGRATION.md")).read()))
    assert declared and declared <= bound, sorted(declared - bound)
    assert bound <= set(capi.HOST_SYMBOLS + capi.DEVICE_SYMBOLS), sorted(bound - set(capi.HOST_SYMBOLS + capi.DEVICE_SYMBOLS))
