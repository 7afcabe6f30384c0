#!/usr/bin/env python
"""In-step kernel times grouped by (family, shape): runs the bench's training step (R50 8x8, B clips, bf16, FlatSGD) with
the library's event timers on and prints, per distinct (family, bytes, flops), launches per step, average microseconds,
achieved GB/s and TFLOP/s.   python tools/step_by_shape.py [--batch 160] [--steps 3] [--out gpurun_out/by_shape.json]"""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mvfnet_b200 import build_recognizer  # noqa: E402
from mvfnet_b200 import mvf as mm  # noqa: E402
from mvfnet_b200.tail import FlatSGD, preprocess_frames  # noqa: E402
from mvfnet_b200.utils import to_channels_last  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=160)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "by_shape.json"))
    ap.add_argument("--families", default="gemm1x1,conv3x3,wgrad1x1,wgrad3x3,bn_fwd,bn_bwd")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = to_channels_last(build_recognizer(bench.model_cfg(), None, None).to(dev)).train()
    opt = FlatSGD(model.parameters(), lr=0.015, momentum=0.9, weight_decay=1e-4, nesterov=True, max_norm=40)
    img = torch.randint(0, 256, (args.batch, 8, 224, 224, 3), dtype=torch.uint8).to(dev)
    lbl = torch.randint(0, 400, (args.batch, 1)).to(dev)

    def step():
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(preprocess_frames(img), lbl)["loss_cls"]
        loss.backward()
        opt.step(1)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    mm.timing_begin()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    rec = mm.timing_end()
    groups = collections.defaultdict(list)
    for kind, nbytes, e0, e1, flops in rec:
        groups[(kind, nbytes, flops)].append(e0.elapsed_time(e1) * 1e3)
    fams = set(args.families.split(","))
    rows = []
    for (kind, nbytes, flops), us in groups.items():
        if kind not in fams:
            continue
        avg = sum(us) / len(us)
        rows.append(dict(kind=kind, launches_per_step=len(us) / args.steps, us=avg, ms_per_step=sum(us) / args.steps / 1e3,
                         mb=nbytes / 1e6, gflop=(flops or 0) / 1e9, gbs=nbytes / avg / 1e3, tfs=(flops or 0) / avg / 1e6))
    rows.sort(key=lambda r: -r["ms_per_step"])
    for r in rows:
        print("%-9s x%4.1f  %8.1f us  %7.3f ms/step  %8.1f MB %8.1f GFLOP  %6.0f GB/s %6.0f TF/s" % (
            r["kind"], r["launches_per_step"], r["us"], r["ms_per_step"], r["mb"], r["gflop"], r["gbs"], r["tfs"]))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
