#!/usr/bin/env python
"""MVF kernel micro-benchmark (SURVEY.md 8d inputs): achieved algorithmic HBM GB/s of mvf_fwd / mvf_bwd per
R50/R101 slab shape and clip count, CUDA-event timed on the launching stream, L2 flushed between iterations.

    python tools/mvf_microbench.py [--iters 20] [--out gpurun_out/mvf_micro.jsonl]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if "--lib" in sys.argv:                                       # an experimental build (tools/build_variant.py)
    from mvfnet_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])
from mvfnet_b200 import MVF  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "mvf_micro.jsonl"))
    ap.add_argument("--clips", default="12,32,64,128")
    ap.add_argument("--T", default="8")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--only-C", type=int, default=0, help="restrict to the slab shape with this C")
    ap.add_argument("--lib", default=None, help="time this libmvf_b200.so instead of the in-tree one")
    ap.add_argument("--fwd-only", action="store_true")
    args = ap.parse_args()
    peak = 6453.4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    rows = []
    for T in [int(v) for v in args.T.split(",")]:
        for (C, H, Cs) in [(512, 28, 64), (1024, 14, 128), (2048, 7, 256)]:
            if args.only_C and C != args.only_C:
                continue
            for B in [int(v) for v in args.clips.split(",")]:
                for training in (True, False):
                    m = MVF(torch.nn.Identity(), T, C, alpha=0.125).cuda().train(training)
                    x = torch.randn(B * T, C, H, H, device="cuda").to(torch.bfloat16).contiguous(
                        memory_format=torch.channels_last).requires_grad_(True)
                    g = torch.randn(B * T, C, H, H, device="cuda").to(torch.bfloat16).contiguous(
                        memory_format=torch.channels_last)
                    cfg = m._cfg()
                    wt, wh, ww = (w.detach().float().contiguous() for w in m._taps())
                    from mvfnet_b200.mvf import mvf_slab_forward
                    E = B * T * Cs * H * H

                    def fwd():
                        return mvf_slab_forward(x.detach(), cfg, wt, wh, ww, m.bn.weight.detach(), m.bn.bias.detach(),
                                                m.bn.running_mean, m.bn.running_var, out="slab")

                    def timeit(fn):
                        ts = []
                        for i in range(args.iters + 3):
                            if not args.no_flush:
                                flush.zero_()
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            fn()
                            e1.record()
                            torch.cuda.synchronize()
                            if i >= 3:
                                ts.append(e0.elapsed_time(e1) * 1e3)
                        ts.sort()
                        return ts[len(ts) // 2], ts[0]

                    med, best = timeit(fwd)
                    row = dict(kernel="mvf_fwd", T=T, C=C, H=H, Cs=Cs, clips=B, training=training, us_median=med,
                               us_best=best, bytes=2 * E * 2, gbs=2 * E * 2 / med / 1e3, frac=2 * E * 2 / med / 1e3 / peak)
                    rows.append(row)
                    print(json.dumps(row), flush=True)
                    if args.fwd_only:
                        del x, g, m
                        continue
                    y = m.fuse(x)

                    def bwd():
                        x.grad = None
                        y.backward(g, retain_graph=True)

                    # the autograd wrapper clones g (full tensor) before the kernel: time the library call alone
                    from mvfnet_b200 import mvf as mm
                    mm.timing_begin()
                    for i in range(args.iters + 3):
                        if not args.no_flush:
                            flush.zero_()
                        bwd()
                    torch.cuda.synchronize()
                    rec = [s.elapsed_time(e) * 1e3 for k, b, s, e, _ in mm.timing_end() if k == "mvf_bwd"][3:]
                    rec.sort()
                    med = rec[len(rec) // 2]
                    row = dict(kernel="mvf_bwd", T=T, C=C, H=H, Cs=Cs, clips=B, training=training, us_median=med,
                               us_best=rec[0], bytes=3 * E * 2, gbs=3 * E * 2 / med / 1e3, frac=3 * E * 2 / med / 1e3 / peak)
                    rows.append(row)
                    print(json.dumps(row), flush=True)
                    del x, g, y, m
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
