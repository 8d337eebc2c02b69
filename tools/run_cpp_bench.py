"""Run the C++ mirror of benches/benchmark.rs (rust-path-tracer_b200/host/benchmark.cpp) on the GPU box."""
import os, subprocess, sys, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rust_path_tracer_b200 import build
from rust_path_tracer_b200.glb import BakedScene
from rust_path_tracer_b200.scenes import breaktime_proxy

tools = build.build_host_tools()
tmp = tempfile.mkdtemp()
cornell = os.path.join(tmp, "DarkCornell.rptw")
BakedScene.load(os.path.join(REPO, "tests", "golden", "scenes", "DarkCornell.npz")).save_rptw(cornell)
proxy = os.path.join(tmp, "BreakTimeProxy.rptw")
baked, atlas = breaktime_proxy()
baked.save_rptw(proxy)  # (geometry + materials; the startup case renders 0 samples)
print(subprocess.run([tools["benchmark"], cornell, proxy], capture_output=True, text=True, timeout=900).stdout)
