"""The C++ host mirror of src/trace.rs (rust-path-tracer_b200/host/): TracingState / setup_trace /
trace_gpu over the C ABI, with the reference's tests/correctness_tests.rs re-stated in C++."""
import os
import subprocess

import pytest

import helpers
from rust_path_tracer_b200 import build
from rust_path_tracer_b200.glb import BakedScene


@pytest.fixture(scope="module")
def tools():
    return build.build_host_tools()


@pytest.fixture(scope="module")
def furnace_rptw(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("rptw") / "FurnaceTest.rptw")
    BakedScene.load(os.path.join(helpers.SCENE_DIR, "FurnaceTest.npz")).save_rptw(path)
    return path


def test_host_tools_build(tools):
    assert all(os.access(p, os.X_OK) for p in tools.values())


def test_no_cpu_fallback_in_the_cpp_host(tools, furnace_rptw):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_cpp_furnace_tests")
    p = subprocess.run([tools["correctness_tests"], furnace_rptw], capture_output=True, text=True, timeout=300)
    assert p.returncode == 1 and "rpt_create failed (-2)" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_cpp_furnace_tests(tools, furnace_rptw):
    """furnace_test_gpu and furnace_test_gpu_mis of tests/correctness_tests.rs:40-53."""
    p = subprocess.run([tools["correctness_tests"], furnace_rptw], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "test furnace_test_gpu ... ok" in p.stdout and "test furnace_test_gpu_mis ... ok" in p.stdout
