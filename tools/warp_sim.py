"""Scratch: CPU simulation of one warp of the extend kernel (tests/cpu_harness: harness_warp_sim) — warp-level node and
triangle rounds per ray for the current policy and for parked triangle groups.
usage: python tools/warp_sim.py [rays]"""
import ctypes as C, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import bench, helpers
import test_wide_bvh_cpu as tw

so = os.path.join(tw.HARNESS_DIR, "_build", "libwide_harness_sim.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
srcs = [os.path.join(tw.HARNESS_DIR, "wide_harness.cpp"), os.path.join(REPO, "rust-path-tracer_b200", "csrc", "wide_bvh_build.cpp")]
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", so, *srcs], check=True, capture_output=True)
lib = C.CDLL(so)
P = lambda a: a.ctypes.data_as(C.c_void_p)
NODE_INSTR, TRI_INSTR = 225 + 30, 90  # instructions per warp-level node visit (+ next-node selection blocks) / triangle test (profiles/r1k)


def surface_rays(world, n, seed=3):
    """Rays leaving random surface points in random directions (what extend sees after the first bounce)."""
    rs = np.random.default_rng(seed)
    pos = world.per_vertex_buffer["vertex"][:, :3]
    tri = world.index_buffer[rs.integers(0, len(world.index_buffer), n)][:, :3]
    b = rs.random((n, 3)).astype(np.float32); b /= b.sum(1, keepdims=True)
    o = (pos[tri] * b[:, :, None]).sum(1)
    d = rs.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([o + d * 1e-3, d], 1), np.float32)


def sim(world, rays, refill_below, defer, slots=1):
    out = np.zeros(8, np.uint64)
    rc = lib.harness_warp_sim(P(world.per_vertex_buffer), C.c_uint32(len(world.per_vertex_buffer)), P(world.index_buffer), C.c_uint32(len(world.index_buffer)),
                              P(world.nodes), C.c_uint32(len(world.nodes)), P(rays), C.c_uint32(len(rays)), C.c_int(refill_below), C.c_int(defer), C.c_int(slots), P(out))
    assert rc == 0 and out[6] == 0, (rc, out)
    nr, lv, tr, lt, n = (float(x) for x in out[:5])
    return {"node_rounds/ray": nr / n, "lanes/node_round": lv / nr, "visits/ray": lv / n, "tri_rounds/ray": tr / n, "lanes/tri_round": lt / max(tr, 1),
            "tests/ray": lt / n, "instr/ray": (nr * NODE_INSTR + tr * TRI_INSTR) / n}


def sim2(world, rays, refill_below, t_hi, t_lo, capacity, blocked_at=99):
    out = np.zeros(8, np.uint64)
    rc = lib.harness_warp_sim2(P(world.per_vertex_buffer), C.c_uint32(len(world.per_vertex_buffer)), P(world.index_buffer), C.c_uint32(len(world.index_buffer)),
                               P(world.nodes), C.c_uint32(len(world.nodes)), P(rays), C.c_uint32(len(rays)), C.c_int(refill_below), C.c_int(t_hi), C.c_int(t_lo),
                               C.c_int(capacity), C.c_int(blocked_at), P(out))
    assert rc == 0, rc
    nr, lv, tr, lt, n = (float(x) for x in out[:5])
    return {"node_rounds/ray": nr / n, "lanes/node_round": lv / nr, "visits/ray": lv / n, "tri_rounds/ray": tr / n, "lanes/tri_round": lt / max(tr, 1),
            "tests/ray": lt / n, "instr/ray": (nr * (NODE_INSTR + 12) + tr * (TRI_INSTR + 8)) / n, "wrong": float(out[6]), "forced": float(out[7]) / n}


def sim3(world, rays, refill_below, tri_min):
    out = np.zeros(8, np.uint64)
    rc = lib.harness_warp_sim3(P(world.per_vertex_buffer), C.c_uint32(len(world.per_vertex_buffer)), P(world.index_buffer), C.c_uint32(len(world.index_buffer)),
                               P(world.nodes), C.c_uint32(len(world.nodes)), P(rays), C.c_uint32(len(rays)), C.c_int(refill_below), C.c_int(tri_min), P(out))
    assert rc == 0 and out[6] == 0, (rc, out)
    nr, lv, tr, lt, n = (float(x) for x in out[:5])
    return {"node_rounds/ray": nr / n, "lanes/node_round": lv / nr, "visits/ray": lv / n, "tri_rounds/ray": tr / n, "lanes/tri_round": lt / max(tr, 1),
            "tests/ray": lt / n, "instr/ray": (nr * NODE_INSTR + tr * TRI_INSTR) / n}


n = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
if len(sys.argv) > 2 and sys.argv[2] == "step":
    for name in ("breaktime", "cornell"):
        world = bench.load_workload(name)[0]
        rays = surface_rays(world, n)
        r = sim(world, rays, 20, 0, 0)
        print(f"{name:10s} as built:              " + "  ".join(f"{k} {v:7.3f}" for k, v in r.items()), flush=True)
        for refill, tri_min in ((20, 1), (24, 1), (28, 1), (20, 2), (20, 4), (24, 4), (20, 8)):
            r = sim3(world, rays, refill, tri_min)
            print(f"{name:10s} one step per round, refill<{refill} tri_min {tri_min}: " + "  ".join(f"{k} {v:7.3f}" for k, v in r.items()), flush=True)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[2] == "partial":
    for name in ("breaktime", "cornell"):
        world = bench.load_workload(name)[0]
        rays = surface_rays(world, n)
        r = sim(world, rays, 20, 0, 0)
        print(f"{name:10s} as built:                        " + "  ".join(f"{k} {v:7.3f}" for k, v in r.items()), flush=True)
        for t_hi, t_lo, cap, blk in ((12, 6, -16, 99), (12, 6, -16, 6), (12, 6, -16, 4), (12, 6, -16, 3), (12, 4, -16, 3), (16, 6, -16, 4), (12, 2, -16, 4), (12, 1, -16, 99)):
            r = sim2(world, rays, 20, t_hi, t_lo, cap, blk)
            print(f"{name:10s} start {t_hi:2d} keep {t_lo:2d} capacity {cap:2d} blocked {blk:2d}: " + "  ".join(f"{k} {v:7.3f}" for k, v in r.items()), flush=True)
    sys.exit(0)
for name in ("breaktime", "cornell"):
    world = bench.load_workload(name)[0]
    rays = surface_rays(world, n)
    for defer, slots in ((0, 0), (8, 0), (8, 1), (16, 1), (16, 2), (24, 2), (24, 4), (32, 8)):
        r = sim(world, rays, 20, defer, slots)
        print(f"{name:10s} flush at {defer:2d} lanes, {slots} parked slots + the cursor's: " + "  ".join(f"{k} {v:7.3f}" for k, v in r.items()), flush=True)
