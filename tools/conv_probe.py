import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvfnet_b200 import ops
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (F, Cin, Cout, H, st) in [(512, 64, 64, 56, 1), (512, 128, 128, 28, 1), (512, 256, 256, 14, 1), (512, 512, 512, 7, 1), (512, 128, 128, 56, 2)]:
    x = torch.randn(F, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)).bfloat16()
    wk = w.permute(0, 2, 3, 1).contiguous()
    wcl = w.contiguous(memory_format=torch.channels_last)
    t0 = time.time()
    ours = t(lambda: ops.conv3x3_raw(x, wk, st, stats=True))
    t1 = time.time()
    torch.backends.cudnn.benchmark = True
    lib = t(lambda: torch.nn.functional.conv2d(x, wcl, None, st, 1))
    Ho = (H - 1) // st + 1
    fl = 2 * F * Ho * Ho * Cout * 9 * Cin
    print("F=%d Cin=%d Cout=%d H=%d s=%d: ours %.1f us (%.0f TF/s)  cudnn %.1f us (%.0f TF/s)  [wall ours %.1fs]" % (F, Cin, Cout, H, st, ours, fl / ours / 1e6, lib, fl / lib / 1e6, t1 - t0), flush=True)
