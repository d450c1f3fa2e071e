"""Build a kernel variant for A/B measurements: python tools/build_variant.py NAME -DFOO=1 -DBAR=2
-> rtg_slam_b200/variants/NAME.so (git-ignored; travels with gpurun). Select it with RTG_SPLAT_LIB=<path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtg_slam_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.HERE, "variants")
obj_dir = os.path.join(B.HERE, "build", "variant_" + name)
os.makedirs(out_dir, exist_ok=True)
os.makedirs(obj_dir, exist_ok=True)
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
procs, objs = [], []
for src in B.SOURCES:
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    objs.append(obj)
    procs.append(subprocess.Popen([nvcc, *B.NVCC_FLAGS, *defs, "-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", obj],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
for p in procs:
    out, _ = p.communicate()
    if p.returncode:
        print(out)
        raise SystemExit(1)
    for l in out.splitlines():
        if "spill" in l and " 0 bytes spill stores" not in l:
            print(l.strip())
lib = os.path.join(out_dir, name + ".so")
subprocess.check_call([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
print(lib)
