#!/usr/bin/env python
"""Per-layer probe of the tcgen05 GEMM / implicit-GEMM kernels at the R50 8x8 shapes: microseconds, achieved HBM GB/s
(operands + result once) and TF/s per launch, with / without the BatchNorm-statistics epilogue, next to cuBLAS
(torch.matmul) / cuDNN on the same shapes.  L2 is flushed between iterations.

    python tools/gemm_probe.py [--clips 160] [--iters 5] [--out gpurun_out/gemm_probe.jsonl]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvfnet_b200 import ops  # noqa: E402

# (name, HW, K, N) 1x1 layers of R50 (forward shapes; the input-gradient is the same GEMM with K and N swapped)
L1X1 = [("stem 192->64 @112", 112 * 112, 192, 64), ("l1 64->64 @56", 3136, 64, 64), ("l1 64->256 @56", 3136, 64, 256),
        ("l1 256->64 @56", 3136, 256, 64), ("l2 256->128 @56", 3136, 256, 128), ("l2 128->512 @28", 784, 128, 512),
        ("l2 512->128 @28", 784, 512, 128), ("l3 512->256 @28", 784, 512, 256), ("l3 256->1024 @14", 196, 256, 1024),
        ("l3 1024->256 @14", 196, 1024, 256), ("l4 1024->512 @14", 196, 1024, 512), ("l4 512->2048 @7", 49, 512, 2048),
        ("l4 2048->512 @7", 49, 2048, 512)]
L3X3 = [("l1 3x3 64 @56", 56, 64, 1), ("l2 3x3 128 @56s2", 56, 128, 2), ("l2 3x3 128 @28", 28, 128, 1),
        ("l3 3x3 256 @14", 14, 256, 1), ("l4 3x3 512 @7", 7, 512, 1)]


def timeit(fn, iters, flush):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=160)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gemm_probe.jsonl"))
    args = ap.parse_args()
    F = args.clips * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for name, hw, k, n in L1X1:
        m = F * hw
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        nbytes, flops = 2 * (m * k + m * n + n * k), 2 * m * n * k
        for stats in (False, True):
            us = timeit(lambda: ops.gemm_tn(a, b, stats=stats, out=out), args.iters, flush)
            rows.append(dict(kernel="gemm1x1", layer=name, M=m, K=k, N=n, stats=stats, us=us, gbs=nbytes / us / 1e3, tfs=flops / us / 1e6))
            print(json.dumps(rows[-1]), flush=True)
        us = timeit(lambda: torch.matmul(a, b.t(), out=out), args.iters, flush)
        rows.append(dict(kernel="cublas", layer=name, M=m, K=k, N=n, us=us, gbs=nbytes / us / 1e3, tfs=flops / us / 1e6))
        print(json.dumps(rows[-1]), flush=True)
        del a, b, out
    for name, h, c, st in L3X3:
        x = torch.randn(F, h, h, c, device="cuda").to(torch.bfloat16).permute(0, 3, 1, 2)
        w = torch.randn(c, c, 3, 3, device="cuda").to(torch.bfloat16)
        wk = w.permute(0, 2, 3, 1).contiguous()
        ho = (h - 1) // st + 1
        mo = F * ho * ho
        nbytes, flops = 2 * (F * h * h * c + mo * c + 9 * c * c), 2 * mo * c * 9 * c
        for stats in (False, True):
            us = timeit(lambda: ops.conv3x3_raw(x, wk, st, stats), args.iters, flush)
            rows.append(dict(kernel="conv3x3", layer=name, stats=stats, us=us, gbs=nbytes / us / 1e3, tfs=flops / us / 1e6))
            print(json.dumps(rows[-1]), flush=True)
        wcl = w.contiguous(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
        us = timeit(lambda: torch.nn.functional.conv2d(x, wcl, stride=st, padding=1), args.iters, flush)
        rows.append(dict(kernel="cudnn", layer=name, us=us, gbs=nbytes / us / 1e3, tfs=flops / us / 1e6))
        print(json.dumps(rows[-1]), flush=True)
        del x, w, wk, wcl
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
