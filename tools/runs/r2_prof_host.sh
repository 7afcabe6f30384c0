#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tail -60
import cProfile, pstats, io, torch, sys, os
sys.path.insert(0, os.getcwd())
import bench
from mvfnet_b200 import build_recognizer
from mvfnet_b200.tail import FlatSGD, preprocess_frames
from mvfnet_b200.utils import to_channels_last
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = to_channels_last(build_recognizer(bench.model_cfg(), None, None).to(dev)).train()
opt = FlatSGD(model.parameters(), lr=0.015, momentum=0.9, weight_decay=1e-4, nesterov=True, max_norm=40)
img = torch.randint(0, 256, (12, 8, 224, 224, 3), dtype=torch.uint8).to(dev)
lbl = torch.randint(0, 400, (12, 1)).to(dev)
def step():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model(preprocess_frames(img), lbl)["loss_cls"]
    loss.backward()
    opt.step(1)
for _ in range(5): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("ms per step (wall):", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(10): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue()[:6000])
PY
