"""Scratch (CPU only): would a better binary tree help?  Insertion-based optimisation (Bittner, Hapala, Havran 2013;
tools/tree_reinsert_experiment.cpp) of the reference's binned-SAH tree before the wide collapse — leaves and the index
permutation kept, inner topology re-arranged — and the node visits / triangle tests per ray of the resulting 8-wide tree
on the CPU harness (rays leaving surfaces in random directions).  Result: profiles/r2_tree_reinsertion_cpu.txt.
usage: python tools/tree_reinsert_experiment.py workload [passes] [fraction of the nodes per pass]"""
import ctypes as C, os, subprocess, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import bench

build = os.path.join(REPO, "tests", "cpu_harness", "_build")
os.makedirs(build, exist_ok=True)
harness_so, reinsert_so = os.path.join(build, "libwide_harness_stats.so"), os.path.join(build, "libtree_reinsert.so")
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", harness_so, os.path.join(REPO, "tests", "cpu_harness", "wide_harness.cpp"),
                os.path.join(REPO, "rust-path-tracer_b200", "csrc", "wide_bvh_build.cpp")], check=True)
subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", reinsert_so, os.path.join(REPO, "tools", "tree_reinsert_experiment.cpp")], check=True)
lib, re = C.CDLL(harness_so), C.CDLL(reinsert_so)
re.bvh_reinsert.restype = C.c_int
P = lambda a: a.ctypes.data_as(C.c_void_p)


def surface_rays(world, n, seed=3):
    rs = np.random.default_rng(seed)
    pos = world.per_vertex_buffer["vertex"][:, :3]
    tri = world.index_buffer[rs.integers(0, len(world.index_buffer), n)][:, :3]
    b = rs.random((n, 3)).astype(np.float32); b /= b.sum(1, keepdims=True)
    o = (pos[tri] * b[:, :, None]).sum(1)
    d = rs.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([o + d * 1e-3, d], 1), np.float32)


def stats(world, nodes, rays, tag):
    out = np.zeros(8, np.uint64)
    rc = lib.harness_wide_stats(P(world.per_vertex_buffer), C.c_uint32(len(world.per_vertex_buffer)), P(world.index_buffer), C.c_uint32(len(world.index_buffer)),
                                P(nodes), C.c_uint32(len(nodes)), P(rays), C.c_uint32(len(rays)), P(out))
    assert rc == 0
    print(f"{tag}: {out[0] / out[3]:.3f} node visits per ray ({out[1] / out[3]:.3f} of them hit nothing), {out[2] / out[3]:.3f} triangle tests per ray", flush=True)


name = sys.argv[1]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fraction = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
world = bench.load_workload(name)[0]
rays = surface_rays(world, 200000)
stats(world, world.nodes, rays, f"{name}, reference tree")
out, st = np.zeros(len(world.nodes), world.nodes.dtype), np.zeros(16)
t0 = time.time()
used = re.bvh_reinsert(P(world.nodes), C.c_uint32(len(world.nodes)), C.c_int(passes), C.c_double(fraction), P(out), P(st))
print(f"{name}: {passes} reinsertion passes, {time.time() - t0:.1f} s on one thread; SAH cost (sum of inner areas / root area) {' -> '.join(f'{v:.3f}' for v in st[:passes + 1])}")
stats(world, out[:used], rays, f"{name}, reinserted tree")
