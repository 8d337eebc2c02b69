"""Scratch: build librpt variants with different flags for one translation unit (A/B runs via RPT_B200_LIBRARY).
usage: python tools/build_variant.py NAME [--tu wavefront_trace_nofma.cu] FLAG [FLAG ...]
       ->  rust-path-tracer_b200/_build/variants/librpt_NAME.so      (default unit: wavefront_shade.cu)"""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rust_path_tracer_b200 import build as b

name, flags = sys.argv[1], sys.argv[2:]
tu = "wavefront_shade.cu"
if flags and flags[0] == "--tu":
    tu, flags = flags[1], flags[2:]
b.build_library()
nvcc = b._nvcc()
out = os.path.join(b.OUT_DIR, "variants"); os.makedirs(out, exist_ok=True)
objs = [os.path.join(b.OUT_DIR, os.path.basename(s) + ".o") for s in b._sources()]
tus = tu.split(",")  # (several units, comma-separated, when a switch reaches into more than one)
variant_objs = []
for tu in tus:
    src = os.path.join(b.CSRC, tu)
    obj = os.path.join(out, f"{os.path.splitext(tu)[0]}_{name}.o")
    extra = ["-fmad=false"] if tu.endswith("_nofma.cu") else []
    if tu == "wavefront_shade.cu":
        extra += ["-fmad=false", "-prec-div=false", "-prec-sqrt=false"]
    p = subprocess.run([nvcc, *b.ARCH_FLAGS, *b.NVCC_FLAGS, *extra, *flags, "-I", os.path.join(b.REPO_DIR, "include"), "-I", b.CSRC, "-x", "cu", "-c", src, "-o", obj], capture_output=True, text=True)
    if p.returncode:
        sys.exit(p.stderr)
    open(obj + ".log", "w").write(p.stderr)
    variant_objs.append(obj)
lib = os.path.join(out, f"librpt_{name}.so")
others = [o for o in objs if not any(o.endswith(t + ".o") for t in tus)]
subprocess.run([nvcc, *b.ARCH_FLAGS, "-shared", "-ccbin", b.HOST_CXX, "-Xcompiler", "-fPIC", "-o", lib, *variant_objs, *others, "-ldl", "-lpthread", "-cudart", "static"], check=True, capture_output=True)
print(lib)
