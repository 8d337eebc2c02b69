"""In-tree build of the native code: `librpt_b200.so` (CUDA backend + host producers).

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels
with the tree to the GPU box.  `python -m rust_path_tracer_b200.build` rebuilds what is stale.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OUT_DIR = os.path.join(PKG_DIR, "_build")
LIB_PATH = os.path.join(OUT_DIR, "librpt_b200.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
# -ffp-contract=off: host-side fp32 must evaluate in the reference's order (BVH build, light table,
# wide-BVH re-layout, precomputed triangle edges).  Device code keeps FMA contraction except where
# it uses explicit round-to-nearest intrinsics (ray generation, traversal, ray/triangle test).
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-ccbin", HOST_CXX,
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA backend cannot be built (there is no CPU fallback)")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers():
    hdrs = []
    for root in (CSRC, os.path.join(REPO_DIR, "include")):
        for dirpath, _dirs, files in os.walk(root):
            hdrs += [os.path.join(dirpath, f) for f in files if f.endswith((".h", ".cuh", ".hpp"))]
    return hdrs


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = _headers() + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            # *_nofma.cu: the id-critical arithmetic (exact.cuh) — no FMA contraction, IEEE division and sqrt.
            # wavefront_shade.cu: shading is held to a radiance tolerance, not to bit identity.  It is built without
            # FMA contraction too (the kernel is HBM-bound, the extra FMUL/FADD are free, and on chaotic scenes every
            # ulp of agreement with the CPU path counts: MAE 3.3e-4 -> 5.7e-5 on the BreakTime proxy), but with the
            # 2-ulp approximate division and sqrt (IEEE ones cost +75 % shade time: a third of the instructions sat in
            # their special-case subroutine); the specular lobe and the HDR sky lookup ask for IEEE explicitly.
            extra = ["-fmad=false"] if src.endswith("_nofma.cu") else []
            if os.path.basename(src) == "wavefront_shade.cu":
                extra += ["-fmad=false", "-prec-div=false", "-prec-sqrt=false"]
            cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, *extra, "-I", os.path.join(REPO_DIR, "include"), "-I", CSRC, "-x", "cu", "-c", src, "-o", obj]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        p = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OUT_DIR, os.path.basename(src) + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{p.stdout}\n{p.stderr}")
        if verbose:
            sys.stderr.write(p.stderr)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB_PATH, objs):
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC", "-o", LIB_PATH, *objs, "-ldl", "-lpthread", "-cudart", "static"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB_PATH


def build_host_tools() -> dict:
    """The C++ host mirror of src/trace.rs (host/) with its test and bench executables, linked against the library."""
    hdir = os.path.join(PKG_DIR, "host")
    build_library()
    out = {}
    common = [os.path.join(hdir, "trace.cpp")]
    for name in ("correctness_tests", "benchmark"):
        exe = os.path.join(OUT_DIR, name)
        srcs = common + [os.path.join(hdir, name + ".cpp")]
        if _stale(exe, srcs + [os.path.join(hdir, "trace.hpp"), LIB_PATH]):
            cmd = [HOST_CXX, "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", *srcs, "-o", exe, "-L", OUT_DIR, "-lrpt_b200",
                   "-Wl,-rpath," + OUT_DIR]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"host tool build failed:\n{p.stdout}\n{p.stderr}")
        out[name] = exe
    return out


def build_oracle() -> str:
    """Build the CPU oracle (test infrastructure) through its own Makefile."""
    odir = os.path.join(REPO_DIR, "oracle")
    p = subprocess.run(["make", "-C", odir], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{p.stdout}\n{p.stderr}")
    return os.path.join(odir, "_build", "liboracle.so")


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
