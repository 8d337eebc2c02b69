#!/usr/bin/env python
"""Headline benchmark: Mpaths/s (and Mrays/s) of the tracing hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A *step* is one pass of the hot path over one batch: `spp_per_step` sample indices for every
pixel of the frame (one `rpt_enqueue`).  Workloads (BASELINE.json `configs`; scenes come from the
committed fixtures under tests/golden/scenes, sample indices from the blue-noise seed table):

    breaktime (default) the scene BASELINE.json's metric is quoted on, at 1920x1080 with NEE (MIS), HDR sky and a
              textured atlas.  scenes/BreakTime.glb is ABSENT from the reference checkout (.MISSING_LARGE_BLOBS), so
              this is a LABELLED SYNTHETIC PROXY (~1M triangles, 16 textured materials, emitters, windows to an HDR
              sky; rust-path-tracer_b200/scenes.py).  The line also carries `also.cornell`: configs[1] on the
              shipped DarkCornell asset, measured in the same run.
    cornell   configs[1]: DarkCornell 1024x1024, NEE with MIS; 16 steps x 64 spp = its 1024 spp
    furnace   configs[0]: FurnaceTest 256x256, 64 spp (4 steps x 16)
    pbr       configs[2]: PBRTest 1920x1080, procedural sky;  pbr-textured: + synthetic 4096^2 atlas
    veach     configs[3]: VeachMIS 1920x1080, MIS

N > 1 (torchrun, one rank per GPU): every rank renders the full frame over its own sample-index
range (rank r starts at sample r * K * spp) — per-GPU work is fixed, so scaling is "weak" — and the
per-GPU accumulators are combined once, inside the timed region, by ncclReduce over NVLink.
`--partition tiles` (configs[4]: 4K frame tiled over the GPUs) gives every rank the 32x32 tiles t with
t % N == rank at the same sample indices instead — total work is fixed, scaling is "strong", and the reduce
adds disjoint tiles (bit-identical to one GPU).

Keys beyond the base contract: `mrays_per_s`, `roofline` (extend kernel), `cpu_baseline`
(oracle on the host cores), `e2e` (same metric through the C ABI with host buffers in the timed
region), `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (scene, width, height, nee, spp_per_step, config name)
    "cornell": ("DarkCornell", 1024, 1024, 1, 64, "configs[1] DarkCornell.glb 1024x1024, NEE (MIS)"),
    "furnace": ("FurnaceTest", 256, 256, 0, 16, "configs[0] FurnaceTest.glb 256x256"),
    "pbr": ("PBRTest", 1920, 1080, 0, 16, "configs[2] PBRTest.glb 1920x1080, procedural sky"),
    "veach": ("VeachMIS", 1920, 1080, 1, 32, "configs[3] VeachMIS.glb 1920x1080, MIS"),
    # synthetic inputs (rust-path-tracer_b200/scenes.py): the shipped assets have no textures, and
    # scenes/BreakTime.glb is absent from the reference checkout (.MISSING_LARGE_BLOBS)
    "pbr-textured": ("PBRTest+procedural textures", 1920, 1080, 0, 16, "configs[2] PBRTest.glb 1920x1080 with a synthetic 4096^2 metallic/roughness/albedo/normal atlas"),
    "breaktime": ("BreakTime PROXY (synthetic ~1M-triangle textured interior, HDR sky)", 1920, 1080, 1, 64,
                  "north-star scene BreakTime.glb 1920x1080 — asset absent, LABELLED SYNTHETIC PROXY"),
    "breaktime-4k": ("BreakTime PROXY (synthetic ~1M-triangle textured interior, HDR sky)", 3840, 2160, 1, 4,
                     "configs[4] BreakTime.glb 3840x2160, HDR sky — asset absent, LABELLED SYNTHETIC PROXY (use --partition tiles)"),
}


def extend_traffic(workload):
    """DRAM bytes per nearest ray of the extend kernel, from the committed `ncu --set full` capture of this
    workload (profiles/extend_traffic.json, written by tools/ncu_traffic.py); None if there is no capture."""
    path = os.path.join(REPO, "profiles", "extend_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(workload)


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def load_workload(name):
    from rust_path_tracer_b200.capi import TracingConfig
    from rust_path_tracer_b200.world import World, make_rng_seeds

    scene, w, h, nee, spp, label = WORKLOADS[name]
    cfg = TracingConfig.default(w, h)
    cfg.nee = nee
    sky = None
    if name == "pbr-textured":
        from rust_path_tracer_b200.glb import BakedScene
        from rust_path_tracer_b200.scenes import textured_pbr_variant

        baked, atlas = textured_pbr_variant(BakedScene.load(os.path.join(REPO, "tests", "golden", "scenes", "PBRTest.npz")))
        world = World.from_baked(baked, atlas=atlas)
    elif name in ("breaktime", "breaktime-4k"):
        from rust_path_tracer_b200.scenes import breaktime_proxy, synthetic_hdr_sky

        # The real asset is used when someone supplies it (it is absent from the reference checkout): scenes/BreakTime.glb
        # in this repo, or RPT_BREAKTIME_GLB=<path>; RPT_BREAKTIME_SKY=<.npy float lat-long image> likewise.
        real = os.environ.get("RPT_BREAKTIME_GLB") or os.path.join(REPO, "scenes", "BreakTime.glb")
        world = World.from_path(real) if os.path.exists(real) else None
        if world is not None:
            label = label.replace(" — asset absent, LABELLED SYNTHETIC PROXY", "").replace(" (use --partition tiles)", "")
            scene = "BreakTime.glb (the real asset, supplied at run time)"
        else:
            baked, atlas = breaktime_proxy()
            world = World.from_baked(baked, atlas=atlas)
        sky_path = os.environ.get("RPT_BREAKTIME_SKY")
        if sky_path and os.path.exists(sky_path):
            from rust_path_tracer_b200.trace import load_skybox

            sky = load_skybox(sky_path)
        if sky is None:
            sky = synthetic_hdr_sky()
        cfg.has_skybox = 1
    else:
        world = World.from_path(os.path.join(REPO, "tests", "golden", "scenes", scene + ".npz"))
        scene += " (fixture of the shipped .glb)"
    if world is None:
        raise SystemExit(f"cannot load scene fixture for {scene}")
    return world, cfg, make_rng_seeds(w, h), spp, label, scene, sky


def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class OracleRunner:
    """The CPU oracle on one workload.  The scene object is built once — the CPU path converts the atlas to float
    texels once, before its sample loop (src/trace.rs:268-271), and so does this."""

    def __init__(self, world, sky=None, threads=0):
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import oracle as oracle_mod

        self.mod = oracle_mod
        self.scene = oracle_mod.OracleScene(world, sky)
        self.threads = threads or host_threads()

    def sample(self, cfg, seeds, spp, want_primary_ids=False):
        """Returns (seconds, counters, output running sum, primary ids or None)."""
        t0 = time.perf_counter()
        out, _, ctr, ids = self.mod.trace(cfg, self.scene, seeds, spp, threads=self.threads, want_primary_ids=want_primary_ids)
        return time.perf_counter() - t0, ctr, out, ids


def workload_config(cfg, label, scene, spp, pipeline, world_size, tiles, wave_slots):
    """The `config` object of the JSON line — the same for the GPU arm and the reference arm."""
    npix = cfg.width * cfg.height
    return {"workload": label, "scene": scene, "width": cfg.width, "height": cfg.height,
            "nee": cfg.nee, "min_bounces": cfg.min_bounces, "max_bounces": cfg.max_bounces, "spp_per_step": spp,
            "pipeline": pipeline,
            "partition": (f"32x32 tiles round-robin x{world_size}, owned tiles gathered on rank 0" if tiles else f"sample-index range x{world_size} + ncclReduce") if world_size > 1 else "single GPU",
            "l2": "working set per step (path state %.0f MB) exceeds the 126 MB L2" % (PATH_STATE_BYTES_PER_SLOT * min(npix * spp, wave_slots or (1 << 24)) / 1e6)}


PATH_STATE_BYTES_PER_SLOT = 144.0


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be
    built in this image (no Rust toolchain), so this arm times the oracle port — the C++ restatement
    of trace_cpu + kernels::trace_pixel — with all host threads, on the GPU arm's config.  Each step is a
    BOUNDED SAMPLE of that config's step: `REFERENCE_SPP` sample indices of every pixel of the frame instead of
    `spp_per_step` (Mpaths/s does not depend on the sample count; the scene and its float atlas are set up once,
    outside the timed region, as in trace_cpu)."""
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    world, cfg, seeds, spp, label, scene, sky = load_workload(args.workload)
    if args.spp:
        spp = args.spp
    ref_spp = min(REFERENCE_SPP, spp)
    runner = OracleRunner(world, sky)
    for _ in range(args.warmup):
        runner.sample(cfg, seeds, 1)
    times, rays = [], 0
    for k in range(args.steps):
        step_seeds = seeds.copy()
        step_seeds[:, 0] += np.uint32(k * ref_spp)
        dt, ctr, _, _ = runner.sample(cfg, step_seeds, ref_spp)
        times.append(dt)
        rays += ctr["nearest_rays"] + ctr["any_rays"]
    total = sum(times)
    paths = cfg.width * cfg.height * ref_spp * args.steps
    value = paths / total / 1e6
    sample = f"{ref_spp} of the step's {spp} sample indices for every pixel of the {cfg.width}x{cfg.height} frame, per step; OpenMP rows (warm-up steps: 1 sample index)"
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong" if args.partition == "tiles" and world_size > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, label, scene, spp, args.pipeline, world_size, args.partition == "tiles" and world_size > 1, args.wave_slots),
        "mrays_per_s": rays / total / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": runner.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


REFERENCE_SPP = 4  # sample indices per reference-arm step (about 7 s on 16 cores for the default workload)


def run_b200(args):
    import torch

    from rust_path_tracer_b200 import capi
    from rust_path_tracer_b200.trace import Renderer

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the tracing backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    world, cfg, seeds0, spp, label, scene, sky = load_workload(args.workload)
    if args.spp:
        spp = args.spp
    npix = cfg.width * cfg.height
    pipeline = capi.PIPELINE_MEGAKERNEL if args.pipeline == "megakernel" else capi.PIPELINE_WAVEFRONT
    r = Renderer(local_rank, pipeline)
    r.upload_world(world, sky)
    r.set_config(cfg)
    if args.wave_slots:
        r.set_wave_slots(args.wave_slots)
    # sample-index-range split: rank r owns samples [r*K*spp, (r+1)*K*spp) (+ warm-up samples first)
    seeds = seeds0.copy()
    tiles = args.partition == "tiles" and world_size > 1
    if tiles:
        r.set_tile_partition(rank, world_size)
    else:
        seeds[:, 0] += np.uint32(rank * (args.steps + args.warmup) * spp)
    r.write_rng(seeds)
    if dist is not None:
        from rust_path_tracer_b200.dist import init_comm

        init_comm(r, dist, rank, world_size)

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps --------------
    for _ in range(args.warmup):
        r.enqueue(spp)
    if dist is not None:
        r.comm_reduce_output(0)  # the first collective builds NCCL's channels: keep that out of the timed region
    r.sync()
    r.write_output(None)
    r.reset_counters()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.enqueue(spp)
    if dist is not None:
        r.comm_reduce_output(0)
    r.sync()
    barrier()
    wall_s = time.perf_counter() - t0
    device_ms = r.device_ms()
    clocks = sampler.stop() if sampler else None
    ctr = r.counters()
    # the job's time is the slowest rank's; device time (CUDA events on the launching stream) where available
    times = torch.tensor([wall_s, device_ms / 1e3], dtype=torch.float64, device="cuda")
    counts = torch.tensor([ctr["paths"], ctr["nearest_rays"] + ctr["any_rays"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    wall_s, dev_s = float(times[0]), float(times[1])
    job_s = wall_s if dist is not None else max(dev_s, 1e-9)  # multi-GPU: include the reduce (outside the per-enqueue events)
    total_paths, total_rays = float(counts[0]), float(counts[1])
    value = total_paths / job_s / 1e6

    line = {
        "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * job_s / args.steps, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(cfg, label, scene, spp, args.pipeline, world_size, tiles, args.wave_slots),
        "mrays_per_s": total_rays / job_s / 1e6,
        "wall_ms_per_step": 1e3 * wall_s / args.steps,
        "gpu_launches": ctr["kernel_launches"],
        "clocks": clocks,
    }

    if rank == 0 and not args.quick:
        # ---- roofline of the dominant kernel (extend): per-launch durations from CUDA events --
        r.set_stage_timing(True)
        r.reset_counters()
        r.enqueue(spp)
        stages = r.stage_timing()
        c1 = r.counters()
        r.set_stage_timing(False)
        ext_ms, ext_launches = stages["extend"] if pipeline == capi.PIPELINE_WAVEFRONT else stages["megakernel"]
        # algorithmic bytes per nearest ray, in the reference's layout: 32 B per box slab-tested + 64 B per
        # triangle tested (16 B index + 3 x 16 B positions) — SURVEY.md §8(d); counted by the oracle on 1 spp
        runner = OracleRunner(world, sky)
        threads = runner.threads
        dt1, octr, o_out, o_ids = runner.sample(cfg, seeds0, 1, want_primary_ids=True)
        boxes_n = octr["boxes_tested"] - octr["boxes_tested_any"]
        tris_n = octr["tris_tested"] - octr["tris_tested_any"]
        bytes_per_ray = (32.0 * boxes_n + 64.0 * tris_n) / max(octr["nearest_rays"], 1)
        peak, peak_src = measured_peaks()
        rays_per_launch = c1["nearest_rays"] / max(ext_launches, 1)
        ms_per_launch = ext_ms / max(ext_launches, 1)
        achieved = rays_per_launch * bytes_per_ray / (ms_per_launch * 1e-3) / 1e9 if ms_per_launch > 0 else 0.0
        if pipeline != capi.PIPELINE_WAVEFRONT:  # the megakernel traces shadow rays in the same launch
            bytes_any = (32.0 * octr["boxes_tested_any"] + 64.0 * octr["tris_tested_any"]) / max(octr["any_rays"], 1)
            achieved = (c1["nearest_rays"] * bytes_per_ray + c1["any_rays"] * bytes_any) / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        line["roofline"] = {
            "bound": "hbm", "kernel": "wf_trace_kernel<true> (extend)" if pipeline == capi.PIPELINE_WAVEFRONT else "mega_trace_kernel",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "traffic_source": None,
            "peak_source": peak_src, "algorithmic_bytes_per_ray": bytes_per_ray, "rays_per_launch": rays_per_launch,
            "ms_per_launch": ms_per_launch, "launches_timed": ext_launches,
            "stage_ms": {k: v[0] for k, v in stages.items() if v[1]},
            "note": "scene is L2-resident: algorithmic bytes are served by L1/L2, so frac can exceed DRAM traffic; see DESIGN.md",
        }
    if rank == 0 and not args.quick and not tiles:
        # ---- parity of THIS frame: sample index 0 of every pixel, GPU against the oracle run just made ----------
        r.write_rng(seeds0)
        r.write_output(None)
        g_ids = r.read_primary_ids()
        r.enqueue(1)
        g_out = r.read_output()
        ok = np.isfinite(g_out[:, :3]).all(axis=1) & np.isfinite(o_out[:, :3]).all(axis=1)
        line["parity"] = {
            "against": "CPU oracle (oracle/oracle.cpp), same scene, config, seeds; sample index 0 of every pixel",
            "primary_id_mismatch_fraction": float((g_ids != o_ids).mean()), "primary_id_budget": 1e-4,
            "mae": float(np.abs(g_out[ok, :3].astype(np.float64) - o_out[ok, :3]).mean()), "mae_tolerance": 1e-3,
            "nan_pixels_gpu": int((~np.isfinite(g_out[:, :3]).all(axis=1)).sum()), "nan_pixels_oracle": int((~np.isfinite(o_out[:, :3]).all(axis=1)).sum()),
            "pixels": int(npix), "spp": 1,
        }
        r.write_rng(seeds)
    if rank == 0 and not args.quick and pipeline == capi.PIPELINE_WAVEFRONT:
        tr = extend_traffic(args.workload)
        if tr:  # dram__bytes_read.sum + dram__bytes_write.sum of the profiled launch, scaled to this launch's ray count
            line["roofline"]["traffic"] = tr["dram_bytes_per_ray"] * line["roofline"]["rays_per_launch"]
            line["roofline"]["traffic_source"] = tr["source"]
            # what actually bounds the kernel (the scene is on chip): issue slots at the SIMT efficiency of incoherent rays
            line["roofline"]["issue"] = {k: tr[k] for k in ("issue_active_pct_of_peak", "lanes_per_instruction", "warp_instructions_per_ray") if k in tr}
    if rank == 0 and not args.quick and dist is None:
        # ---- CPU baseline (N = 1 only): the oracle port on the host cores, bounded sample -----
        cpu_spp = max(1, min(8, int(15.0 / max(dt1, 1e-3))))
        dtc, cctr, _, _ = runner.sample(cfg, seeds0, cpu_spp)
        line["cpu_baseline"] = {"value": npix * cpu_spp / dtc / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port",
                                "mrays_per_s": (cctr["nearest_rays"] + cctr["any_rays"]) / dtc / 1e6,
                                "sample": f"{cpu_spp} spp of the full {cfg.width}x{cfg.height} frame ({dtc:.1f} s), OpenMP rows"}

    if rank == 0 and not args.quick and args.workload != "cornell" and dist is None:
        # ---- configs[1] on the shipped asset, same run: DarkCornell 1024^2 MIS, 4 x 64 spp ---------
        w2, cfg2, seeds2, spp2, label2, scene2, sky2 = load_workload("cornell")
        with Renderer(local_rank, pipeline) as r2:
            r2.upload_world(w2, sky2)
            r2.set_config(cfg2)
            r2.write_rng(seeds2)
            for _ in range(3):
                r2.enqueue(spp2)
            r2.sync()
            r2.reset_counters()
            for _ in range(4):
                r2.enqueue(spp2)
            ms2 = r2.device_ms()
            c2 = r2.counters()
        line["also"] = {"cornell": {"workload": label2, "scene": scene2, "steps": 4, "spp_per_step": spp2, "value": c2["paths"] / ms2 / 1e3,
                                    "unit": "Mpaths/s", "mrays_per_s": (c2["nearest_rays"] + c2["any_rays"]) / ms2 / 1e3}}

    # ---- end to end through the C ABI with host buffers: H2D seeds + config, enqueue, D2H frame --
    # (each step's seed table is its input: prepared before the clock starts, in page-locked memory — rpt_host_alloc)
    fb = capi.pinned_empty(npix * 3, np.float32)
    step_seeds = []
    for k in range(args.steps):
        buf = capi.pinned_empty(seeds.shape, np.uint32)
        buf[...] = seeds
        buf[:, 0] += np.uint32(k * spp)
        step_seeds.append(buf)
    r.write_output(None)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        r.set_config(cfg)
        r.write_rng(step_seeds[k])
        r.enqueue(spp)
        r.read_framebuffer(float((k + 1) * spp), fb)
    barrier()
    e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    line["e2e"] = {"value": npix * spp * args.steps * (1 if tiles else world_size) / float(e2e[0]) / 1e6, "unit": "Mpaths/s",
                   "h2d_bytes_per_step": 80 + 8 * npix, "d2h_bytes_per_step": 12 * npix,
                   "what": "per step: rpt_set_config + rpt_write_rng (pinned host seeds) + rpt_enqueue + rpt_read_framebuffer (pinned host RGB)"}
    if not np.isfinite(fb).all():
        line["e2e"]["nan_pixels"] = int((~np.isfinite(fb.reshape(-1, 3)).all(axis=1)).sum())

    if dist is not None:
        r.comm_destroy()
    r.close()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="breaktime")
    ap.add_argument("--pipeline", choices=["wavefront", "megakernel"], default="wavefront")
    ap.add_argument("--partition", choices=["samples", "tiles"], default="samples", help="how N > 1 GPUs share the frame")
    ap.add_argument("--spp", type=int, default=0, help="samples per step (default: the workload's)")
    ap.add_argument("--wave-slots", type=int, default=0)
    ap.add_argument("--quick", action="store_true", help="skip the roofline and cpu_baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 16:
            args.steps = 4
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
