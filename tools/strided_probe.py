#!/usr/bin/env python
"""What HBM sustains on the MVF slab access pattern: copy channels [0, Cs) of every pixel of an NHWC bf16 tensor
(runs of 2*Cs bytes every 2*C bytes) into a compact tensor, vs a dense copy of the same number of bytes."""
import ctypes as C, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvfnet_b200 import _lib
L = _lib.lib()
L.copy_cols.restype = C.c_int
L.copy_cols.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=15, clean=True):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        if clean:
            flush.sum()                      # leave CLEAN lines in L2: no dirty write-backs charged to fn
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]

st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (Cc, H, Cs) in [(512, 28, 64), (1024, 14, 128), (2048, 7, 256)]:
    for B in (32, 64, 128):
        M = B * 8 * H * H
        x = torch.randn(M, Cc, device="cuda").bfloat16()
        y = torch.empty(M, Cs, device="cuda", dtype=torch.bfloat16)
        d = torch.randn(M, Cs, device="cuda").bfloat16()
        for clean in (True, False):
            t_s = timeit(lambda: L.copy_cols(x.data_ptr(), Cc, y.data_ptr(), Cs, M, Cs, st), clean=clean)
            t_d = timeit(lambda: L.copy_cols(d.data_ptr(), Cs, y.data_ptr(), Cs, M, Cs, st), clean=clean)
            by = 2 * M * Cs * 2
            print(json.dumps(dict(C=Cc, H=H, Cs=Cs, clips=B, l2="clean" if clean else "dirty", strided_us=t_s, strided_gbs=by / t_s / 1e3,
                                  dense_us=t_d, dense_gbs=by / t_d / 1e3)), flush=True)
