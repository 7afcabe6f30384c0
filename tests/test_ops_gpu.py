"""GPU parity of the bottleneck operators behind the C ABI: the tcgen05 1x1-convolution GEMM (plain, K-split
MVF variant, fused BatchNorm statistics), the fused BatchNorm(+residual)(+ReLU) kernels, and the whole fused
Bottleneck against the same block executed by torch ops in fp32.  bf16 storage: tolerance 1e-2 relative."""
import numpy as np
import pytest
import torch

from oracle import mvf_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _lib_launches():
    from mvfnet_b200 import _lib
    return _lib.launch_count()


def rel_l2(a, b):
    """||a - b|| / ||b||: for gradients that passed through ReLU masks.  A bf16 and an fp32 pipeline legitimately
    disagree on the mask of the few activations whose pre-activation is within rounding distance of 0, and each flip
    changes that element's gradient by O(1); the max-norm is meaningless there, the energy norm is not."""
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 64), (4704, 512, 2048), (50176, 256, 1024), (12544, 2048, 512),
                                   (1000, 64, 256), (129, 192, 128), (3136 * 4, 256, 64)])
def test_gemm_tn_matches_fp32_matmul(M, N, K):
    from mvfnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    out, s1, s2 = ops.gemm_tn(a, b, stats=True)
    ref = a.float() @ b.float().t()
    assert rel(out, ref) < 1e-2
    # the statistics are those of the ROUNDED output
    assert rel(s1, out.float().sum(0)) < 1e-3 or float((s1 - out.float().sum(0)).abs().max()) < 1e-2 * float(out.float().abs().sum(0).max())
    assert rel(s2, (out.float() ** 2).sum(0)) < 1e-3


@pytest.mark.parametrize("M,N,K,K0", [(1568, 256, 512, 64), (3136, 512, 1024, 128), (392, 512, 2048, 256)])
def test_gemm_k_split_operand(M, N, K, K0):
    """A = [a0[:, :K0] | a1[:, K0:]]: the MVF concatenation without the copy (MVF.py:135)."""
    from mvfnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(K0)
    a1 = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    a0 = torch.randn(M, K0, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    out, _, _ = ops.gemm_tn(a1, b, a0=a0, k0=K0)
    cat = torch.cat([a0, a1[:, K0:]], dim=1).float()
    assert rel(out, cat @ b.float().t()) < 1e-2


def test_gemm_rejects_bad_shapes():
    from mvfnet_b200 import ops, _lib
    a = torch.zeros(128, 48, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(64, 48, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.MvfB200Error):
        ops.gemm_tn(a, b)                                       # K not a multiple of 64


@pytest.mark.parametrize("C,HW,F", [(64, 56, 8), (256, 28, 8), (1024, 14, 16), (2048, 7, 24), (24, 5, 4)])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("training", [True, False])
def test_bn_act_vs_oracle(C, HW, F, relu, res, training):
    """oracle/mvf_oracle.py batchnorm2d(+relu+residual) forward and backward on the same bf16-rounded inputs."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(C + HW)
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C, generator=g))
        bn.bias.copy_(0.3 * torch.randn(C, generator=g))
        bn.running_mean.copy_(0.2 * torch.randn(C, generator=g))
        bn.running_var.copy_(0.5 + torch.rand(C, generator=g))
    rm0, rv0 = bn.running_mean.double().numpy().copy(), bn.running_var.double().numpy().copy()
    bn = bn.cuda().train(training)
    x = (torch.randn(F, C, HW, HW, generator=g) * 1.5 + 0.3).bfloat16()
    r = torch.randn(F, C, HW, HW, generator=g).bfloat16() if res else None
    gy = torch.randn(F, C, HW, HW, generator=g).bfloat16()
    xd = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    rd = r.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True) if res else None
    assert ops.bn_eligible(xd, bn)
    y = ops.bn_act(xd, bn, relu=relu, residual=rd)
    y.backward(gy.cuda().contiguous(memory_format=torch.channels_last))
    gam, bet = bn.weight.detach().double().cpu().numpy(), bn.bias.detach().double().cpu().numpy()
    xn = x.double().numpy()
    a, st = O.batchnorm2d(xn, gam, bet, rm0, rv0, training)
    pre = a + (r.double().numpy() if res else 0.0)
    ref = O.relu(pre) if relu else pre
    gref = gy.double().numpy() * ((pre > 0) if relu else 1.0)
    dx, dg, db = O.batchnorm2d_backward(gref, xn, gam, st["mean"], st["rstd"], training)
    t = lambda v: torch.as_tensor(np.asarray(v))
    assert rel(y.detach().cpu(), t(ref)) < 1e-2
    assert rel(xd.grad.cpu(), t(dx)) < 2e-2
    assert rel(bn.weight.grad.cpu(), t(dg)) < 1e-2
    assert rel(bn.bias.grad.cpu(), t(db)) < 1e-2
    if res:
        assert rel(rd.grad.cpu(), t(gref)) < 1e-2
    if training:
        assert rel(bn.running_mean.cpu(), t(st["new_running_mean"])) < 1e-3
        assert rel(bn.running_var.cpu(), t(st["new_running_var"])) < 1e-3
        assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize("inpl,planes,hw,stride,mvf", [(256, 64, 28, 1, False), (512, 256, 14, 1, True), (1024, 512, 14, 2, True),
                                                        (64, 64, 28, 1, False)])
def test_fused_bottleneck_vs_torch_fp32(inpl, planes, hw, stride, mvf, monkeypatch):
    """The fused block (MVF kernel + tcgen05 GEMMs + fused BN/ReLU/residual) against the SAME module run by torch
    ops in fp32 (the path test_model_gpu pins to the reference's golden vectors).  Three train-mode BatchNorms and
    ReLU masks amplify bf16 rounding, so the yard-stick for the gradients is the error of torch's own bf16 autocast
    execution of the block against the same fp32 run: the fused path must not be worse than 1.5x that."""
    import copy
    import torch.nn as nn
    from mvfnet_b200 import Bottleneck, MVF, _lib
    torch.manual_seed(inpl + planes)
    T, F = 4, 8
    ds = None
    if stride != 1 or inpl != planes * 4:
        ds = nn.Sequential(nn.Conv2d(inpl, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
    blk = Bottleneck(inpl, planes, stride, 1, ds)
    if mvf:
        blk.conv1 = MVF(blk.conv1, T, inpl, 0.125)
    for m in blk.modules():
        if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            with torch.no_grad():
                m.weight.normal_(1, 0.2)
                m.bias.normal_(0, 0.2)
    blk = blk.cuda().train()
    ref, lib = copy.deepcopy(blk), copy.deepcopy(blk)
    x = torch.randn(F, inpl, hw, hw, device="cuda")
    gy = torch.randn(F, planes * 4, hw // stride, hw // stride, device="cuda")

    def run_bf16(module):
        xb = x.bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = module(xb)
        y.backward(gy.bfloat16().contiguous(memory_format=torch.channels_last))
        return xb, y

    before = _lib.launch_count()
    xb, y = run_bf16(blk)
    assert _lib.launch_count() - before >= 10, "fused kernels were not used"
    monkeypatch.setenv("MVFB_CONV1X1", "0")
    monkeypatch.setenv("MVFB_BN", "0")
    xl, yl = run_bf16(lib)                                   # torch / cuDNN bf16 execution of the same block
    monkeypatch.undo()
    xr = x.bfloat16().float().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy.bfloat16().float())
    assert rel(y.detach().float(), yr.detach()) < 3e-2

    def check(ours, library, exact, what):
        e_ours, e_lib = rel_l2(ours.float(), exact), rel_l2(library.float(), exact)
        assert e_ours < max(1.5 * e_lib, 2e-2), (what, e_ours, e_lib)

    check(xb.grad, xl.grad, xr.grad, "dx")
    for (k, p), (_, q), (_, r) in zip(blk.named_parameters(), lib.named_parameters(), ref.named_parameters()):
        check(p.grad, q.grad, r.grad, k)
    for (k, b), (_, c) in zip(blk.named_buffers(), ref.named_buffers()):
        if "running" in k:
            assert rel(b.float(), c.float()) < 2e-2, k


@pytest.mark.parametrize("F,Cin,Cout,H,stride", [(2, 64, 64, 56, 1), (4, 128, 128, 28, 1), (8, 256, 256, 14, 1), (16, 512, 512, 7, 1),
                                                 (2, 128, 128, 56, 2), (4, 256, 256, 28, 2), (4, 512, 512, 14, 2), (3, 64, 192, 9, 1),
                                                 # several tiles per CTA (persistent tile loop, both TMEM accumulators in use)
                                                 (40, 64, 64, 56, 1), (100, 128, 128, 56, 2),
                                                 # the halo-band kernel (conv_halo.cu): weights resident / streamed, one and
                                                 # two channel blocks, a frame height whose last tile is mostly padding
                                                 (60, 128, 128, 28, 1), (50, 64, 128, 28, 1), (40, 128, 64, 56, 1), (90, 64, 64, 30, 1),
                                                 # CTA pairs in the halo kernel: an odd tile count (the last pair's second
                                                 # CTA has no frame), two output-channel tiles
                                                 (61, 128, 128, 28, 1), (45, 128, 256, 28, 1)])
def test_conv3x3_implicit_gemm(F, Cin, Cout, H, stride):
    """TMA-im2col implicit GEMM vs torch conv2d in fp32 on the same bf16-rounded operands; forward, fused statistics,
    and the stride-1 input gradient (rotated weights through the same kernel)."""
    from mvfnet_b200 import ops
    torch.backends.cudnn.allow_tf32 = True       # the fp32 reference conv is only a yard-stick at 1e-2; keep it fast
    g = torch.Generator(device="cuda").manual_seed(F + Cin + H)
    x = torch.randn(F, Cin, H, H, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)).bfloat16()
    y, sums = ops.conv3x3_raw(x, w.permute(0, 2, 3, 1).contiguous(), stride, stats=True)
    ref = torch.nn.functional.conv2d(x.float(), w.float(), None, stride, 1)
    assert y.shape == ref.shape
    err = rel(y, ref)
    if err >= 1e-2:      # diagnose an off-by-one in the im2col base coordinates before failing
        for dh in (-1, 0, 1):
            for dw in (-1, 0, 1):
                shifted = torch.roll(ref, shifts=(dh, dw), dims=(2, 3))
                print("shift", dh, dw, rel(y[:, :, 2:-2, 2:-2], shifted[:, :, 2:-2, 2:-2]))
    assert err < 1e-2
    assert rel(sums[0], y.float().sum((0, 2, 3))) < 1e-3 or float((sums[0] - y.float().sum((0, 2, 3))).abs().max()) < 1e-2 * float(y.float().abs().sum((0, 2, 3)).max())
    assert rel(sums[1], (y.float() ** 2).sum((0, 2, 3))) < 1e-3
    # autograd: dx (ours for stride 1, library for stride 2) and dw vs torch fp32
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    gy = torch.randn_like(ref)
    torch.nn.functional.conv2d(xr, wr, None, stride, 1).backward(gy)
    xo = x.clone().requires_grad_(True)
    wo = w.float().clone().requires_grad_(True)
    ops.conv3x3(xo, wo, stride).backward(gy.bfloat16().contiguous(memory_format=torch.channels_last))
    assert rel(xo.grad, xr.grad) < 2e-2
    assert rel(wo.grad, wr.grad) < 2e-2


@pytest.mark.parametrize("M,N,K,K0", [(4704, 512, 2048, 0), (50176, 256, 1024, 128), (100352, 64, 256, 0), (12544, 2048, 512, 0),
                                      (1000, 64, 64, 0), (6272, 256, 512, 64), (777, 128, 192, 0)])
def test_wgrad_mn_major_gemm(M, N, K, K0):
    """dW = dY^T X on the tensor cores with MN-major operands and split-K atomics vs an fp32 matmul."""
    from mvfnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    dy = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    x0 = torch.randn(M, K0, device="cuda", generator=g).bfloat16() if K0 else None
    dw = ops.gemm_wgrad(dy, x1, x0=x0, k0=K0)
    xcat = x1.float() if x0 is None else torch.cat([x0.float(), x1[:, K0:].float()], dim=1)
    ref = dy.float().t() @ xcat
    assert dw.dtype == torch.float32 and dw.shape == ref.shape
    assert rel(dw, ref) < 2e-3


@pytest.mark.parametrize("F,Cin,Cout,H", [(4, 256, 512, 56), (8, 512, 1024, 28), (8, 1024, 2048, 14)])
def test_strided_1x1_downsample(F, Cin, Cout, H):
    """make_res_layer's stride-2 1x1 convolution: im2col (1x1 window) forward + weight gradient, GEMM + scatter dgrad."""
    from mvfnet_b200 import ops
    torch.backends.cudnn.allow_tf32 = True
    g = torch.Generator(device="cuda").manual_seed(F + Cin)
    x = torch.randn(F, Cin, H, H, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) / Cin ** 0.5).bfloat16()
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, wr, None, 2)
    gy = torch.randn_like(ref)
    ref.backward(gy)
    xo, wo = x.clone().requires_grad_(True), w.float().clone().requires_grad_(True)
    y, sums = ops.conv1x1_strided(xo, wo, 2, True)
    y.backward(gy.bfloat16().contiguous(memory_format=torch.channels_last))
    assert rel(y.detach(), ref.detach()) < 1e-2
    assert rel(sums[1], (y.detach().float() ** 2).sum((0, 2, 3))) < 1e-3
    assert rel(xo.grad, xr.grad) < 2e-2
    assert rel(wo.grad, wr.grad) < 2e-2
    # the pixels the stride skips are written by the same epilogue (dx is allocated uninitialised): exact zeros
    assert not xo.grad[:, :, 1::2, :].any() and not xo.grad[:, :, :, 1::2].any()


# ------------------------------------------------------------------------------------------------ stem
@pytest.mark.parametrize("F,H,W", [(2, 224, 226), (3, 64, 66), (1, 30, 32), (2, 224, 224), (3, 62, 40), (2, 33, 16)])
def test_stem_conv_vs_torch(F, H, W):
    """conv1 of the ResNet stem (im2col + tcgen05 GEMM, backbones/resnet.py:424) against torch's fp32 convolution of the
    same bf16-rounded operands: output, the BatchNorm sums of the GEMM epilogue, and the weight gradient."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(7 + F + H)
    x = torch.randn((F, 3, H, W), generator=g).cuda()          # W % 8 == 0: row-staged im2col kernel, else the gather
    w = (torch.randn((64, 3, 7, 7), generator=g) * 0.1).cuda().requires_grad_(True)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert ops.stem_eligible(x, conv)
        y, sums = ops.stem_conv(x, w, True)
    xr, wr = x.bfloat16().float(), w.detach().bfloat16().float().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, wr, stride=2, padding=3)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    scale = ref.abs().max().item()
    assert (y.float() - ref).abs().max().item() < 1e-2 * scale
    yf = y.float()
    assert torch.allclose(sums[0], yf.sum(dim=(0, 2, 3)), rtol=2e-3, atol=2e-3 * scale * yf[:, 0].numel() ** 0.5)
    assert torch.allclose(sums[1], (yf * yf).sum(dim=(0, 2, 3)), rtol=2e-3)
    gy = torch.randn(ref.shape, generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    ref.backward(gy.float())
    assert (w.grad - wr.grad).abs().max().item() < 1e-2 * wr.grad.abs().max().item()


@pytest.mark.parametrize("F,C,H,W", [(2, 64, 112, 112), (3, 16, 9, 14), (1, 8, 5, 5)])
def test_maxpool3x3s2_vs_torch(F, C, H, W):
    """MaxPool2d(3, 2, 1) forward and backward against ATen on inputs full of ties (post-ReLU zeros, repeated values):
    the recorded arg-max must follow ATen's first-maximum rule, so both directions are bit-exact."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(11 + C + H)
    x = torch.randint(-3, 4, (F, C, H, W), generator=g).float().clamp_min(0) * 0.5
    x = x.cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    pool = torch.nn.MaxPool2d(3, 2, 1)
    assert ops.maxpool_eligible(x, pool)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    ya = ops.maxpool3x3s2(xa)
    yb = pool(xb)
    assert ya.shape == yb.shape and torch.equal(ya, yb)
    gy = torch.randn(yb.shape, generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    ya.backward(gy)
    yb.backward(gy)
    assert torch.equal(xa.grad, xb.grad)


@pytest.mark.parametrize("F,C,H,W,training", [(4, 64, 112, 112, True), (3, 64, 16, 20, True), (2, 32, 8, 8, True),
                                               (5, 128, 14, 14, True), (3, 64, 16, 20, False)])
def test_bn_relu_maxpool_fused_vs_separate_and_torch(F, C, H, W, training):
    """norm1 + ReLU + maxpool in one pass (bn_relu_maxpool_fwd / _bwd): the pooled output equals the three-launch path
    (bn_apply, then maxpool3x3s2 of the rounded activation) BIT FOR BIT, channels with a negative gamma included (the
    maximum of a decreasing function is taken at the window's minimum); running statistics equal to rounding; input / gamma / beta
    gradients match both the separate launches and torch's fp32 BatchNorm2d -> ReLU -> MaxPool2d on the same bf16 input."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(F * 1000 + C + H)
    x = (torch.randn((F, C, H, W), generator=g) * 1.5 + 0.3).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    gamma = torch.randn(C, generator=g) * 0.8 + 0.5            # a good share negative
    assert (gamma < 0).any() and (gamma > 0).any()
    beta = torch.randn(C, generator=g) * 0.3
    pool = torch.nn.MaxPool2d(3, 2, 1)

    def make_bn():
        bn = torch.nn.BatchNorm2d(C).cuda()
        with torch.no_grad():
            bn.weight.copy_(gamma); bn.bias.copy_(beta)
            bn.running_mean.copy_(torch.linspace(-0.2, 0.4, C)); bn.running_var.copy_(torch.linspace(0.5, 2.0, C))
        bn.train(training)
        return bn
    bn_a, bn_b, bn_c = make_bn(), make_bn(), make_bn()
    assert ops.bn_relu_maxpool_eligible(x, bn_a, pool)
    xa, xb, xc = (x.clone().requires_grad_(True) for _ in range(3))
    # the same statistics for both paths (bn_stats adds its partial sums with atomics: two runs differ in the last bit)
    sums = torch.stack([x.float().sum((0, 2, 3)), (x.float() ** 2).sum((0, 2, 3))]).contiguous() if training else None
    launches = _lib_launches()
    ya = ops.bn_relu_maxpool(xa, bn_a, sums=sums)
    assert _lib_launches() - launches == 1
    yb = ops.maxpool3x3s2(ops.bn_act(xb, bn_b, relu=True, sums=sums))
    yc = pool(torch.relu(bn_c(xc.float())))
    assert torch.equal(ya, yb)
    assert rel(ya, yc) < 6e-3
    assert rel(bn_a.running_mean, bn_b.running_mean) < 1e-6 and rel(bn_a.running_var, bn_b.running_var) < 1e-6
    assert rel(bn_a.running_var, bn_c.running_var) < 1e-4
    gy = torch.randn(yc.shape, generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    ya.backward(gy); yb.backward(gy); yc.backward(gy.float())
    print("dgamma vs fp32 %.2e, vs separate %.2e; dbeta vs fp32 %.2e, vs separate %.2e; dx vs fp32 %.2e (separate launches: %.2e)" % (
        rel(bn_a.weight.grad, bn_c.weight.grad), rel(bn_a.weight.grad, bn_b.weight.grad), rel(bn_a.bias.grad, bn_c.bias.grad), rel(bn_a.bias.grad, bn_b.bias.grad), rel_l2(xa.grad, xc.grad), rel_l2(xb.grad, xc.grad)))
    # The fused pass reproduces the fp32 reference's gradients to rounding.  The separate launches do not: where a window's
    # largest inputs are different bf16 numbers whose activations round to the SAME bf16 number (frequent for a small
    # |gamma|), max-pooling the rounded activation sends the gradient to the earliest of them -- as ATen's bf16 pipeline
    # does -- while the fused pass and the fp32 reference send it to the largest input: 2-3 % of dx's energy, up to 9 %
    # of a dgamma entry at 112 x 112.
    assert rel_l2(xa.grad, xc.grad) < 1e-3 and rel_l2(xa.grad, xb.grad) < 5e-2
    assert rel(bn_a.weight.grad, bn_c.weight.grad) < 1e-4 and rel(bn_a.bias.grad, bn_c.bias.grad) < 1e-4
    assert rel(bn_a.weight.grad, bn_b.weight.grad) < 0.15 and rel(bn_a.bias.grad, bn_b.bias.grad) < 1e-2
    assert rel_l2(xa.grad, xc.grad) <= rel_l2(xb.grad, xc.grad) + 1e-6


def test_maxpool3x3s2_negative_values_and_nans():
    """Signed inputs (the packed maximum must order negatives correctly) with a few NaNs: ATen's rule `(val > maxval) ||
    isnan(val)` makes a NaN win its windows and the LAST NaN of a window take the gradient."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn((3, 64, 30, 30), generator=g)
    x.view(-1)[torch.randint(0, x.numel(), (200,), generator=g)] = float("nan")
    x = x.cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    pool = torch.nn.MaxPool2d(3, 2, 1)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    ya = ops.maxpool3x3s2(xa)
    yb = pool(xb)
    assert torch.equal(torch.isnan(ya), torch.isnan(yb)) and torch.isnan(yb).any()
    assert torch.equal(torch.nan_to_num(ya.float(), nan=7.0), torch.nan_to_num(yb.float(), nan=7.0))
    gy = torch.randn(yb.shape, generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    ya.backward(gy)
    yb.backward(gy)
    assert torch.equal(xa.grad, xb.grad)


@pytest.mark.parametrize("M,N,K", [(300, 64, 64), (12544, 256, 64), (1000, 128, 256), (100000, 256, 64), (70001, 192, 128),
                                   (50000, 512, 512)])
def test_gemm_add_epilogue(M, N, K):
    """conv1x1_gemm_add: out = A B^T + addend (the fused gradient sum at a Bottleneck's input), M tail included."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g).cuda().bfloat16()
    b = (torch.randn((N, K), generator=g) * 0.1).cuda().bfloat16()
    r = torch.randn((M, N), generator=g).cuda().bfloat16()
    out, _, _ = ops.gemm_tn(a, b, add=r)
    ref = a.float() @ b.float().t() + r.float()
    assert (out.float() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    plain, _, _ = ops.gemm_tn(a, b)
    assert not torch.equal(out, plain)


@pytest.mark.parametrize("M,N,K,col0", [(70001, 256, 64, 32), (30000, 1024, 256, 128), (9000, 2048, 512, 256)])
def test_gemm_add_cols_epilogue(M, N, K, col0):
    """conv1x1_gemm_add_cols: the addend applies to columns >= col0 only (MVF blocks: the slab's gradient arrives through
    mvf_bwd_add); many tiles per CTA so the two addend buffers of the TMA-prefetch path cycle."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(M + N + K + col0)
    a = torch.randn((M, K), generator=g).cuda().bfloat16()
    b = (torch.randn((N, K), generator=g) * 0.1).cuda().bfloat16()
    r = torch.randn((M, N), generator=g).cuda().bfloat16()
    out, _, _ = ops.gemm_tn(a, b, add=r, add_col0=col0)
    ref = a.float() @ b.float().t()
    ref[:, col0:] += r.float()[:, col0:]
    assert (out.float() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K,relu", [(50000, 256, 64, True), (20000, 512, 128, True), (20000, 1024, 256, False),
                                        (6000, 2048, 512, True)])
def test_gemm_bnact_residual_epilogue(M, N, K, relu):
    """conv1x1_gemm_bnact: out = [relu]((A B^T) * scale + shift + res) -- conv3 + eval bn3 + identity + ReLU of the
    inference path, against fp32 torch."""
    from mvfnet_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g).cuda().bfloat16()
    b = (torch.randn((N, K), generator=g) * 0.1).cuda().bfloat16()
    r = torch.randn((M, N), generator=g).cuda().bfloat16()
    scale = (0.5 + torch.rand(N, generator=g)).cuda()
    shift = torch.randn(N, generator=g).cuda()
    out = ops.gemm_bnact(a, b, scale, shift, relu, res=r)
    ref = (a.float() @ b.float().t()) * scale + shift + r.float()
    if relu:
        ref = ref.clamp_min(0)
    assert (out.float() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    nores = ops.gemm_bnact(a, b, scale, shift, relu)
    ref2 = (a.float() @ b.float().t()) * scale + shift
    if relu:
        ref2 = ref2.clamp_min(0)
    assert (nores.float() - ref2).abs().max().item() < 1e-2 * ref2.abs().max().item()


def test_conv_halo_matches_im2col_path_bit_for_bit_in_structure():
    """The halo-band kernel and the im2col kernel compute the same sums in a different order: outputs agree to bf16
    rounding, the fused statistics to 1e-3, and the option that disables the halo kernel is honoured."""
    from mvfnet_b200 import ops, _lib
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(48, 64, 56, 56, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) / 24).bfloat16().permute(0, 2, 3, 1).contiguous()
    launches = _lib.launch_count()
    y1, s1 = ops.conv3x3_raw(x, w, 1, stats=True)
    assert _lib.launch_count() == launches + 1
    _lib.set_option(_lib.OPT_CONV_HALO_OFF, 1)
    try:
        y2, s2 = ops.conv3x3_raw(x, w, 1, stats=True)
    finally:
        _lib.set_option(_lib.OPT_CONV_HALO_OFF, 0)
    assert rel(y1, y2) < 4e-3 and not torch.isnan(y1.float()).any()
    assert rel(s1[0], s2[0]) < 2e-3 and rel(s1[1], s2[1]) < 1e-3


@pytest.mark.parametrize("F,Cin,Cout,H", [(61, 128, 128, 28), (45, 128, 256, 28), (50, 64, 128, 28), (160, 128, 128, 14)])
def test_conv_halo_cta_pair_matches_single_cta(F, Cin, Cout, H):
    """The halo kernel's CTA-pair mode (streamed weights, half of each weight tile per CTA, M = 256 MMAs issued by the
    leader) computes every output with the same products in the same order as the single-CTA launch: bit-identical
    outputs; the fused statistics differ only by the order of the atomics."""
    from mvfnet_b200 import ops, _lib
    g = torch.Generator(device="cuda").manual_seed(F + Cin + Cout)
    x = torch.randn(F, Cin, H, H, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)).bfloat16().permute(0, 2, 3, 1).contiguous()
    y1, s1 = ops.conv3x3_raw(x, w, 1, stats=True)
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 1)
    try:
        y2, s2 = ops.conv3x3_raw(x, w, 1, stats=True)
    finally:
        _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 0)
    assert torch.equal(y1, y2) and not torch.isnan(y1.float()).any()
    assert rel(s1[0], s2[0]) < 1e-4 or float((s1[0] - s2[0]).abs().max()) < 1e-3 * float(y1.float().abs().sum((0, 2, 3)).max())
    assert rel(s1[1], s2[1]) < 1e-5
    ref = torch.nn.functional.conv2d(x.float(), w.permute(0, 3, 1, 2).float(), None, 1, 1)
    assert rel(y1, ref) < 1e-2


@pytest.mark.parametrize("M,N,K,stats", [(9500, 512, 512, True), (25088, 1024, 512, True), (31000, 256, 1024, False), (9500, 768, 576, True)])
def test_gemm_cta_pair_matches_single_cta(M, N, K, stats):
    """Wide-N GEMMs launch as CTA pairs (tcgen05 cta_group::2: one M = 256 MMA over two CTAs, each holding half of the
    weight tile; odd tile counts: the second CTA of the last pair works on rows past M, which TMA zero-fills and clips).
    Same products, same accumulation order per output: bit-identical to the single-CTA launch, and both match fp32 torch."""
    from mvfnet_b200 import ops, _lib
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g).cuda().bfloat16()
    b = (torch.randn((N, K), generator=g) * 0.05).cuda().bfloat16()
    out, cs, cq = ops.gemm_tn(a, b, stats=stats)
    _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 1)
    try:
        ref, rs, rq = ops.gemm_tn(a, b, stats=stats)
    finally:
        _lib.set_option(_lib.OPT_GEMM_PAIR_OFF, 0)
    assert torch.equal(out, ref)
    f32 = a.float() @ b.float().t()
    assert (out.float() - f32).abs().max().item() < 1e-2 * f32.abs().max().item()
    if stats:
        assert rel(cs, rs) < 1e-4 and rel(cq, rq) < 1e-5
