import os, sys, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvfnet_b200 import ops, _lib
from mvfnet_b200._lib import ptr
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
F = 512
print("1x1 wgrad (M, Cout, Cin): ours vs torch.matmul")
for (hw, cout, cin) in [(56, 64, 256), (56, 256, 64), (28, 128, 512), (28, 512, 128), (14, 256, 1024), (14, 1024, 256), (7, 512, 2048), (7, 2048, 512)]:
    M = F * hw * hw
    g = torch.randn(M, cout, device="cuda").bfloat16(); x = torch.randn(M, cin, device="cuda").bfloat16()
    a = t(lambda: ops.gemm_wgrad(g, x)); b = t(lambda: torch.matmul(g.t(), x))
    print("  M=%d %d x %d: ours %.0f us, lib %.0f us, min-traffic %.0f us" % (M, cout, cin, a, b, (M * (cout + cin) * 2) / 6.45e6))
print("3x3 wgrad: ours vs aten.convolution_backward")
torch.backends.cudnn.benchmark = True
for (hw, c, st) in [(56, 64, 1), (28, 128, 1), (14, 256, 1), (7, 512, 1), (56, 128, 2)]:
    x = torch.randn(F, c, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    ho = (hw - 1) // st + 1
    g = torch.randn(F, c, ho, ho, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(c, c, 3, 3, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    d = ops.ConvDesc(); d.F, d.H, d.W, d.Cin, d.Cout, d.stride, d.ksize = F, hw, hw, c, c, st, 3
    dw = torch.empty(c, 3, 3, c, device="cuda")
    L = ops._L()
    _lib.check(L.conv3x3_wgrad(C.byref(d), ptr(g), ptr(x), ptr(dw), ops._stream()), 'conv3x3_wgrad')
    a = t(lambda: L.conv3x3_wgrad(C.byref(d), ptr(g), ptr(x), ptr(dw), ops._stream()))
    b = t(lambda: torch.ops.aten.convolution_backward(g, x, w, None, [st, st], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]))
    print("  hw=%d C=%d s=%d: ours %.0f us, cudnn %.0f us" % (hw, c, st, a, b))
