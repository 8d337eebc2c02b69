"""rust-path-tracer_b200 — B200-native tracing hot path behind the reference's trace API.

Layout:
  csrc/      CUDA kernels (sm_100a), the C ABI (include/rpt_b200.h) and host producers
  capi.py    ctypes binding of the C ABI
  world.py   `World` (scene buffers: BVH, light table, per-vertex data) — src/asset.rs
  trace.py   `TracingState`, `setup_trace`, `trace_gpu` — the mirror of src/trace.rs
  glb.py     minimal .glb import (stands in for assimp)
  dist.py    one-process-per-GPU partitioning + NCCL plumbing
"""
