"""Scratch: how much would two half-waves on two streams gain?  Two contexts of ONE process share the device's primary
CUDA context, so their streams can co-run: thread A and thread B each render half of the samples of the same frame.
Compared with one context rendering all of them.  usage: python tools/gpu_two_streams.py [workload] [spp]"""
import os, sys, threading, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from rust_path_tracer_b200.trace import Renderer

workload = sys.argv[1] if len(sys.argv) > 1 else "breaktime"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
world, cfg, seeds, *_rest = bench.load_workload(workload)
sky = _rest[-1]


def make(slots=0):
    r = Renderer(0)
    r.upload_world(world, sky); r.set_config(cfg); r.write_rng(seeds)
    if slots:
        r.set_wave_slots(slots)
    for _ in range(3):
        r.enqueue(8)
    r.sync()
    return r


def timed(renderers, each):
    def run(r):
        for _ in range(3):
            r.enqueue(each)
        r.sync()
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(r,)) for r in renderers]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    return cfg.width * cfg.height * each * 3 * len(renderers) / dt / 1e6


one = make()
print(f"{workload}: one context, {spp} spp x 3: {timed([one], spp):8.1f} Mpaths/s", flush=True)
one.close()
for slots in (0, 1 << 23):
    pair = [make(slots), make(slots)]
    print(f"{workload}: two contexts ({'16' if not slots else '8'} Mi-slot waves), {spp // 2} spp x 3 each: {timed(pair, spp // 2):8.1f} Mpaths/s", flush=True)
    [r.close() for r in pair]
