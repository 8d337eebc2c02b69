"""What ARE the primary-hit id mismatches at full size?  Runs on the CPU: the oracle's traversal of the reference BVH and
the product's wide-BVH traversal (host build of the same code, tests/cpu_harness) on the primary rays of sample 0 of
the benchmarked frame, then looks at every pixel where the two name different triangles.
usage: python tests/checkers/id_mismatch_report.py [workload]      (writes profiles/r2_id_mismatches.txt)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, oracle as om
import test_wide_bvh_cpu as tw

workload = sys.argv[1] if len(sys.argv) > 1 else "breaktime"
world, cfg, seeds, *_ = bench.load_workload(workload)
scene = om.OracleScene(world)
rays = om.camera_rays(cfg, seeds)
o_hit, o_tri, o_t, o_back = om.intersect(scene, rays)
import ctypes as C
lib = C.CDLL(tw.HARNESS_SO)  # built by tests/test_wide_bvh_cpu.py's fixture (run that test once first)
w_hit, w_tri, w_t, w_back = tw.wide_intersect(lib, world, rays)[:4]
both = (o_hit == 1) & (w_hit == 1)
diff = np.nonzero((o_hit != w_hit) | (both & (o_tri != w_tri)))[0]
lines = [f"# {workload}: {cfg.width}x{cfg.height}, sample 0, {len(rays)} primary rays; reference traversal (oracle) vs wide-BVH traversal (host build of the product's code)",
         f"# pixels naming different triangles: {len(diff)} = {len(diff) / len(rays):.3e} of the frame (budget 1e-4)"]
pos = world.per_vertex_buffer["vertex"][:, :3]
same_t = shared_edge = ulp_apart = 0
for i in diff:
    if not (o_hit[i] and w_hit[i]):
        lines.append(f"pixel {i}: hit / miss disagreement (oracle {o_hit[i]}, wide {w_hit[i]})")
        continue
    a, b = world.index_buffer[o_tri[i]][:3], world.index_buffer[w_tri[i]][:3]
    pa, pb = pos[a], pos[b]
    shared = sum(any(np.array_equal(v, u) for u in pb) for v in pa)  # vertices in common (by position)
    dt = abs(int(np.float32(o_t[i]).view(np.int32)) - int(np.float32(w_t[i]).view(np.int32)))
    same_t += dt == 0
    ulp_apart += 0 < dt <= 4
    shared_edge += shared >= 2
    lines.append(f"pixel {i}: oracle tri {o_tri[i]} t={o_t[i]!r}  wide tri {w_tri[i]} t={w_t[i]!r}  t differs by {dt} ulp, triangles share {shared} vertices")
lines.insert(2, f"# of these: bit-equal t {same_t}, t within 4 ulp {ulp_apart}, triangles sharing an edge {shared_edge}")
out = os.path.join(REPO, "profiles", f"r2_id_mismatches_{workload}.txt")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
