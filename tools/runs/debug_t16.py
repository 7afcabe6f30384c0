import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import mvf_oracle as O
from test_mvf_production_gpu import make_module
from mvfnet_b200 import _lib

def run(C, H, Cs, T, N, training, seed_extra=0):
    m, rm0, rv0 = make_module(C, Cs, T, training, seed=C + H + T + N)
    gen = torch.Generator(device="cuda").manual_seed(17 + N + seed_extra)
    F = N * T
    x = torch.randn((F, H, H, C), generator=gen, device="cuda").to(torch.bfloat16).permute(0, 3, 1, 2)
    y = m(x)
    torch.cuda.synchronize()
    bf = lambda p: p.detach().to(torch.bfloat16).float().cpu().numpy().reshape(Cs, 3)
    kw = dict(wt=bf(m.shift_conv.weight), wh=bf(m.h_conv.weight), ww=bf(m.w_conv.weight),
              gamma=m.bn.weight.detach().cpu().numpy(), beta=m.bn.bias.detach().cpu().numpy(),
              running_mean=rm0, running_var=rv0, mode="THW", use_hs=True, training=training)
    xs = x[:, :Cs].float().cpu().numpy()
    rf = O.mvf_forward(xs, T, Cs, **kw)
    got = y.detach()[:, :Cs].float().cpu().numpy()
    err = np.abs(got - rf["out"]) / np.abs(rf["out"]).max()
    e5 = err.reshape(N, T, Cs, H, H)
    print("case", (C, H, Cs, T, N, training), "kernel", _lib.last_kernel(), "max rel err %.4f" % err.max(), "bad elems", int((err > 0.01).sum()))
    if err.max() > 0.01:
        bad = e5 > 0.01
        print("  per t     :", bad.sum(axis=(0, 2, 3, 4)).tolist())
        print("  per clip  :", [int(v) for v in bad.sum(axis=(1, 2, 3, 4))])
        print("  per chgrp :", bad.reshape(N, T, Cs // 32, 32, H, H).sum(axis=(0, 1, 3, 4, 5)).tolist())
        print("  per row   :", bad.sum(axis=(0, 1, 2, 4)).tolist())
        print("  per col   :", bad.sum(axis=(0, 1, 2, 3)).tolist())
        if training:
            print("  mean err", np.abs(m.bn.running_mean.cpu().numpy() - rf["new_running_mean"]).max(), "var err", np.abs(m.bn.running_var.cpu().numpy() - rf["new_running_var"]).max())

for args in [(1024, 14, 128, 16, 64, True), (1024, 14, 128, 16, 64, False), (1024, 14, 128, 16, 18, True), (1024, 14, 128, 16, 36, True),
             (1024, 14, 128, 16, 37, True), (1024, 14, 128, 8, 64, True), (1024, 14, 128, 4, 64, True), (1024, 14, 128, 4, 300, True), (2048, 7, 256, 16, 64, True)]:
    run(*args)
    run(*args, seed_extra=1)
