#!/usr/bin/env python
"""Build an experimental variant of libmvf_b200.so with extra nvcc flags (kernel A/B measurements):

    python tools/build_variant.py NAME -DMVFB_SWEEP_MID_UNROLL=2 ...   ->  variants/NAME/libmvf_b200.so

`tools/mvf_microbench.py --lib variants/NAME/libmvf_b200.so` then times that library instead of the in-tree one.
variants/ is git-ignored (it travels with the gpurun snapshot); the product only ever loads mvfnet_b200/libmvf_b200.so.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvfnet_b200 import build as B  # noqa: E402


def main():
    name, extra = sys.argv[1], sys.argv[2:]
    out = os.path.join(ROOT, "variants", name)
    os.makedirs(out, exist_ok=True)
    nvcc = B._nvcc()
    procs, objs = [], []
    for src in B.sources():
        obj = os.path.join(out, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([nvcc] + [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + extra + ["-c", src, "-o", obj]))
    if any(p.wait() for p in procs):
        raise SystemExit("nvcc failed")
    lib = os.path.join(out, "libmvf_b200.so")
    subprocess.run([nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], check=True)
    for o in objs:
        os.remove(o)
    print(lib)


if __name__ == "__main__":
    main()
