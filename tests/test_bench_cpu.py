"""bench.py's host-side logic that can run without a GPU: the roofline blocks are pure functions of live measurements and
of the committed ncu capture (profiles/kernel_profiles.json) — a schema drift between tools/ncu_profile.py and bench.py
must fail HERE, not on a multi-GPU box."""
import json
import os

import pytest

import bench
import helpers

PROFILES = os.path.join(helpers.REPO, "profiles", "kernel_profiles.json")


def test_workload_config_is_identical_for_both_arms():
    from rust_path_tracer_b200.capi import TracingConfig

    cfg = TracingConfig.default(1920, 1080)
    cfg.nee = 1
    a = bench.workload_config(cfg, "label", "scene", 64, "wavefront", 1, False, 0)
    b = bench.workload_config(cfg, "label", "scene", 64, "wavefront", 1, False, 0)
    assert a == b and set(a) == {"workload", "scene", "width", "height", "nee", "min_bounces", "max_bounces", "spp_per_step", "pipeline", "partition", "l2"}
    assert "tiles" in bench.workload_config(cfg, "l", "s", 16, "wavefront", 8, True, 0)["partition"]


@pytest.mark.skipif(not os.path.exists(PROFILES), reason="no committed ncu capture")
def test_roofline_blocks_from_the_committed_capture():
    table = json.load(open(PROFILES))
    assert table.get("git") and table.get("captured")
    workloads = [w for w, v in table.items() if isinstance(v, dict)]
    assert "breaktime" in workloads
    for w in workloads:
        prof, src = bench.kernel_profiles(w)
        assert src and "git" in src
        ext, sh = prof["extend"], prof["shade"]
        rays, ext_ms = 4.6e8, 100.0  # a 64-spp step of the default workload: 4.6 Grays/s
        roof = bench.extend_roofline(ext, src, rays, ext_ms, 32, 148, 1965.0, {"extend": ext_ms}, "wf_trace_kernel<true> (extend)")
        assert roof["bound"] == "issue" and roof["unit"] == "Gwarp-inst/s"
        assert roof["peak"] == pytest.approx(148 * 4 * 1.965, rel=1e-9)
        assert roof["achieved"] == pytest.approx(ext["warp_instructions_per_ray"] * 4.6, rel=1e-9)
        assert 0 < roof["frac_simt_adjusted"] < roof["frac"]
        stats = {"nearest_rays": 100, "nearest_node_visits": 1200, "nearest_triangle_tests": 270, "node_bytes": 80, "triangle_bytes": 48}
        hbm = bench.extend_hbm_block(ext, stats, 4.6e9, 2256.0, 6532.2, "measured")
        assert hbm["own_layout_bytes_per_ray"] == pytest.approx(12 * 80 + 2.7 * 48)
        assert 0 < hbm["frac"] < 0.2  # measured DRAM traffic of a kernel whose scene is on chip
        shade = bench.shade_roofline(sh, src, 4.4e8, 37.0, 32, 6532.2, "measured", textured=True)
        assert shade["bound"] == "hbm" and shade["algorithmic_bytes_per_hit"] == 272.0
        assert 0 < shade["frac"] < 1.2 and 0 < shade["traffic_frac"] < 1.2
    # the capture of the benchmarked kernel is self-consistent: per-wave sums equal the per-bounce entries
    e = table["breaktime"]["extend"]
    assert e["rays"] == sum(b["rays"] for b in e["per_bounce"])
    assert e["duration_ms"] == pytest.approx(sum(b["duration_ms"] for b in e["per_bounce"]))


def test_roofline_blocks_without_a_capture():
    roof = bench.extend_roofline(None, None, 1e6, 1.0, 1, 148, 1965.0, {}, "mega_trace_kernel")
    assert roof["frac"] is None and roof["achieved"] is None and roof["bound"] == "issue"
    assert "frac" not in bench.extend_hbm_block(None, None, 1e9, 2000.0, 6532.2, "measured")
    assert bench.shade_roofline(None, None, 1e6, 1.0, 1, 6532.2, "measured", textured=False)["algorithmic_bytes_per_hit"] == 240.0


def test_reference_arm_line_on_a_small_workload(capsys):
    """`--impl reference` end to end on the CPU (it never touches the GPU): one JSON line with the contract's keys."""
    import argparse

    args = argparse.Namespace(gpus=1, steps=1, warmup=0, workload="furnace", spp=0, pipeline="wavefront", partition="samples", wave_slots=0)
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpaths/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["spp_per_step"] == 16 and "4 of the step's 16" in line["cpu_baseline"]["sample"]


def test_bench_takes_the_real_asset_when_it_is_supplied(tmp_path, monkeypatch):
    """scenes/BreakTime.glb is absent from the reference checkout; bench.py says so in every line and falls back to the
    labelled proxy — but a file supplied through RPT_BREAKTIME_GLB (or placed at scenes/BreakTime.glb) is what gets
    benchmarked, and the label stops saying PROXY."""
    import bench
    import test_glb_loader

    path = str(tmp_path / "BreakTime.glb")
    test_glb_loader._write_glb(path)
    monkeypatch.setenv("RPT_BREAKTIME_GLB", path)
    world, cfg, seeds, spp, label, scene, sky = bench.load_workload("breaktime")
    assert world.ntriangles == 4 and "PROXY" not in label and "real asset" in scene
    assert cfg.has_skybox == 1 and sky is not None and len(seeds) == cfg.width * cfg.height
