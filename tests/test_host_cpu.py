"""CPU-side tests: the C-ABI library loads and exports every symbol include/mvf_b200.h declares, the host
mirror of the reference interface (registry / config / model structure / state_dict keys / error behaviour),
and the N>1 data-parallel gradient path on gloo with world_size 2."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT

REF_CFG = "/root/reference/configs/MVFNet/K400/mvf_kinetics400_2d_rgb_r50_dense.py"


def test_library_exports_every_declared_symbol():
    from mvfnet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mvf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([a-z_0-9]+)\s*\(", hdr)) - {"defined"}
    names = {n for n in names if n.startswith(("mvf_", "mvfb_", "conv", "bn_", "sgd_", "gemm_", "stem_", "maxpool", "copy_"))}
    assert {"mvf_fwd", "mvf_bwd", "mvf_b200_version", "mvf_b200_last_error", "conv1x1_gemm", "conv1x1_gemm_add",
            "stem_im2col", "maxpool3x3s2_fwd", "maxpool3x3s2_bwd"} <= names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), "libmvf_b200.so does not export %s" % n
    assert _lib.lib().mvf_b200_version() == 100 or _lib.lib().mvf_b200_version() > 100


def test_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    from mvfnet_b200 import _lib
    L = _lib.lib()
    d = _lib.MvfDesc(N=1, T=4, C=8, Cs=0, H=3, W=3, dtype=0, layout=0, mode=2, use_hs=1, training=0, eps=1e-5, momentum=0.1)
    rc = L.mvf_fwd(ctypes.byref(d), None, None, 0, None, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == 1 and b"Cs" in L.mvf_b200_last_error()
    d.Cs = 4
    d.dtype = 7
    rc = L.mvf_fwd(ctypes.byref(d), None, None, 0, None, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == 1 and b"dtype" in L.mvf_b200_last_error()
    with pytest.raises(_lib.MvfB200Error):
        _lib.check(rc, "mvf_fwd")


def test_stem_and_workspace_host_logic_without_gpu():
    """Host-side logic of the entry points added with the generation-4 forward and the stem: argument validation
    (no CUDA call is made for a rejected call) and the workspace the train-mode forward asks for -- one row of
    2*Cg {value, epoch} words (8 bytes each) per CTA of a grid of at most one CTA per SM, for every R50 slab shape."""
    from mvfnet_b200 import _lib, ops
    L = ops._L()
    assert L.stem_im2col(None, None, 4, 224, 224, None) == 1 and b"stem_im2col" in L.mvf_b200_last_error()
    assert L.maxpool3x3s2_fwd(None, None, None, 4, 112, 112, 64, None) == 1
    assert L.maxpool3x3s2_bwd(None, None, None, 4, 112, 112, 60, None) == 1      # C % 8 != 0 is rejected too
    g = ops.GemmDesc()
    g.M, g.N, g.K, g.K0, g.lda0, g.lda1, g.ldb, g.ldd = 128, 64, 64, 0, 0, 64, 64, 64
    assert L.conv1x1_gemm_add(ctypes.byref(g), None, None, None, None, 64, None, None) == 1
    for (C, H, Cs) in [(512, 28, 64), (1024, 14, 128), (2048, 7, 256)]:
        d = _lib.MvfDesc(N=64, T=8, C=C, Cs=Cs, H=H, W=H, dtype=_lib.MVFB_BF16, layout=_lib.MVFB_NHWC, mode=2, use_hs=1,
                         training=1, eps=1e-5, momentum=0.1)
        nbytes = L.mvf_fwd_workspace_bytes(ctypes.byref(d))
        assert nbytes >= 2 * Cs * 8                                   # at least one CTA row per channel group
        assert nbytes <= 148 * 2 * 64 * 8 + 16 * 8 * Cs + 4096       # never more than one 64-channel row per SM (+ generic)


def test_fused_stem_tail_argument_errors_without_gpu():
    """bn_relu_maxpool_fwd / _bwd (norm1 + ReLU + max-pool in one pass) reject bad calls before any CUDA call: null
    tensors, a population that is not F*H*W, a channel count whose 8-channel vectors do not divide the 256-thread CTA,
    a backward over odd sizes (its 2 x 2 input blocks need even H and W)."""
    from mvfnet_b200 import ops
    L = ops._L()
    d = ops.BnDesc()
    d.M, d.C, d.relu, d.training, d.eps, d.momentum = 4 * 112 * 112, 64, 1, 1, 1e-5, 0.1
    nul = [None] * 10
    assert L.bn_relu_maxpool_fwd(ctypes.byref(d), None, 4, 112, 112, *nul) == 1 and b"bn_relu_maxpool_fwd" in L.mvf_b200_last_error()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    ten = [p] * 10
    d.M = 5 * 112 * 112                                                  # M != F*H*W
    assert L.bn_relu_maxpool_fwd(ctypes.byref(d), p, 4, 112, 112, *ten) == 3 and b"M = F*H*W" in L.mvf_b200_last_error()
    d.M, d.C = 4 * 112 * 112, 24                                         # 3 vectors per pixel do not divide 256
    assert L.bn_relu_maxpool_fwd(ctypes.byref(d), p, 4, 112, 112, *ten) == 3
    d.C, d.relu = 64, 0                                                  # the fusion IS the ReLU
    assert L.bn_relu_maxpool_fwd(ctypes.byref(d), p, 4, 112, 112, *ten) == 3
    d.relu, d.M = 1, 4 * 111 * 112
    eight = [p] * 8
    assert L.bn_relu_maxpool_bwd(ctypes.byref(d), p, p, p, 4, 111, 112, *eight) == 3 and b"even H and W" in L.mvf_b200_last_error()
    assert L.bn_relu_maxpool_bwd(ctypes.byref(d), None, p, p, 4, 112, 112, *eight) == 1


def test_registry_contract():
    from mvfnet_b200 import Registry, build_from_cfg, RECOGNIZERS, BACKBONES, HEADS
    assert "Recognizer2D" in RECOGNIZERS.module_dict and "ResNet" in BACKBONES.module_dict and "TSNClsHead" in HEADS.module_dict
    r = Registry("thing")

    @r.register_module
    class A:
        def __init__(self, x, y=2):
            self.x, self.y = x, y

    with pytest.raises(KeyError):
        r.register_module(A)
    with pytest.raises(TypeError):
        r.register_module(lambda: 0)
    cfg = dict(type="A", x=1)
    a = build_from_cfg(cfg, r, dict(y=5))
    assert (a.x, a.y) == (1, 5) and cfg == dict(type="A", x=1)
    assert build_from_cfg(dict(type=A, x=3), r).x == 3
    with pytest.raises(KeyError):
        build_from_cfg(dict(type="B"), r)
    with pytest.raises(TypeError):
        build_from_cfg(dict(type=3), r)


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference configs not on this box")
@pytest.mark.parametrize("name,depth", [("r50", 50), ("r101", 101)])
def test_unmodified_reference_config_builds_same_structure(name, depth):
    """configs/MVFNet/K400/*.py drop in: same state_dict keys / parameter count as the reference model."""
    from mvfnet_b200 import Config, build_recognizer, MVF
    cfg = Config.fromfile(REF_CFG.replace("r50", name))
    assert cfg.model.module_cfg.alpha == 0.125 and cfg.data.videos_per_gpu == 12
    cfg.model.backbone.pretrained = None
    m = build_recognizer(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    z = np.load(GOLDEN + "/structure.npz")
    assert list(m.state_dict().keys()) == [str(k) for k in z["keys_" + name]]
    assert sum(p.numel() for p in m.parameters()) == int(z["params_" + name])
    n_mvf = sum(isinstance(x, MVF) for x in m.modules())
    assert n_mvf == {50: 9, 101: 26}[depth]
    assert "type" not in cfg.model.module_cfg          # popped like recognizer2d.py:52


def test_mvf_module_attributes_and_bypass():
    from mvfnet_b200 import MVF
    net = torch.nn.Conv2d(16, 4, 1, bias=False)
    m = MVF(net, n_segment=4, in_channels=16, alpha=0.125, share=False, mode="TH")
    assert m.num_shift_channel == 2 and m.split_sizes == [2, 14] and m.net is net and m.n_segment == 4
    assert hasattr(m, "h_conv") and not hasattr(m, "w_conv") and m.mode == "TH" and m.use_hs is True
    assert tuple(m.shift_conv.weight.shape) == (2, 1, 3, 1, 1) and tuple(m.h_conv.weight.shape) == (2, 1, 1, 3, 1)
    assert sorted(k for k in m.state_dict()) == sorted(
        ["net.weight", "shift_conv.weight", "h_conv.weight", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var",
         "bn.num_batches_tracked"])
    ms = MVF(net, 4, 16, alpha=0.5, share=True, mode="THW")
    assert not hasattr(ms, "h_conv") and not hasattr(ms, "w_conv")
    # Cs == 0: straight to self.net (MVF.py:108) -- the only CPU-executable path
    m0 = MVF(net, 4, 16, alpha=0.0)
    x = torch.randn(8, 16, 3, 3)
    assert torch.equal(m0(x), net(x))
    with pytest.raises(RuntimeError):
        m(x)                                            # no CPU fallback for the fused kernels
    with pytest.raises(ValueError):
        MVF(net, 4, 16, alpha=0.5, mode="HW")


def test_resnet_train_mode_semantics():
    from mvfnet_b200 import ResNet
    from torch.nn.modules.batchnorm import _BatchNorm
    r = ResNet(50, norm_eval=True, out_indices=(3,))
    r.train()
    assert all(not m.training for m in r.modules() if isinstance(m, _BatchNorm))
    r = ResNet(50, norm_eval=False, frozen_stages=1, out_indices=(3,))
    r.train()
    assert not r.bn1.training and not r.conv1.weight.requires_grad
    assert all(not p.requires_grad for p in r.layer1.parameters()) and all(p.requires_grad for p in r.layer2.parameters())
    with pytest.raises(KeyError):
        ResNet(18)


def test_checkpoint_roundtrip_strips_module_prefix(tmp_path):
    from mvfnet_b200 import Bottleneck
    from mvfnet_b200.checkpoint import load_checkpoint, save_checkpoint
    a, b = Bottleneck(16, 4), Bottleneck(16, 4)
    f = str(tmp_path / "c.pth")
    torch.save({"state_dict": {"module." + k: v for k, v in a.state_dict().items()}}, f)
    load_checkpoint(b, f, map_location="cpu", strict=True)
    assert all(torch.equal(a.state_dict()[k], b.state_dict()[k]) for k in a.state_dict())
    save_checkpoint(a, f, meta=dict(epoch=3))
    ck = torch.load(f)
    assert set(ck) == {"meta", "state_dict"} and ck["meta"]["epoch"] == 3


# ---------------------------------------------------------------- N>1 path on gloo
def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from mvfnet_b200.dist import DistOptimizerHook, FlatGrads, MMDistributedDataParallel, allreduce_grads, init_dist
    init_dist("pytorch", backend="gloo")
    torch.manual_seed(100 + rank)                      # different initial weights per rank ...
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ddp = MMDistributedDataParallel(net)               # ... made identical by the construction-time broadcast
    w0 = [p.detach().clone() for p in net.parameters()]
    flat = FlatGrads(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    xs, ys = x_all[rank::world], y_all[rank::world]    # DistributedSampler-style stride by rank

    class Runner:
        pass

    r = Runner()
    r.model, r.optimizer, r.flat_grads = ddp, opt, flat
    hook = DistOptimizerHook(grad_clip=dict(max_norm=40, norm_type=2))
    for _ in range(2):
        r.outputs = {"loss": torch.nn.functional.mse_loss(ddp(xs), ys)}
        hook.after_train_iter(r)
    # reference-style (non-flat) path must give the same averaged gradients
    net.zero_grad(set_to_none=True)
    torch.nn.functional.mse_loss(net(xs), ys).backward()
    allreduce_grads(net.parameters())
    out[rank] = dict(w0=w0, w=[p.detach().clone() for p in net.parameters()],
                     g=[p.grad.clone() for p in net.parameters()], numel=flat.numel())
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    for p, q in zip(a["w0"], b["w0"]):
        assert torch.equal(p, q)                       # broadcast from rank 0
    for p, q in zip(a["w"], b["w"]):
        assert torch.allclose(p, q, atol=1e-7)         # replicas stay in lock step
    # single-process equivalent: full batch of 8, mean loss == average of the two half-batch gradients
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    for _ in range(2):
        opt.zero_grad()
        torch.nn.functional.mse_loss(net(x_all), y_all).backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 40)
        opt.step()
    for p, q in zip(net.parameters(), a["w"]):
        assert torch.allclose(p, q, atol=1e-6)
    opt.zero_grad()
    torch.nn.functional.mse_loss(net(x_all), y_all).backward()
    for p, q in zip(net.parameters(), a["g"]):
        assert torch.allclose(p.grad, q, atol=1e-6)
    assert a["numel"] == sum(p.numel() for p in net.parameters())


def _mvf_desc(N, T, C, Cs, H, dtype=1, layout=1, training=1):
    from mvfnet_b200 import _lib
    d = _lib.MvfDesc()
    d.N, d.T, d.C, d.Cs, d.H, d.W = N, T, C, Cs, H, H
    d.dtype, d.layout, d.mode, d.use_hs, d.training = dtype, layout, 2, 1, training
    d.eps, d.momentum = 1e-5, 0.1
    return d


def test_kernel_tier_plan_is_pinned_without_gpu():
    """Which kernel tier serves which shape is part of the contract (host-only query mvf_b200_plan): every R50 / R101
    slab shape at 224 and 256 px with the configs' frame counts runs on the sweep kernels, forward and backward; other
    frame counts step down to the stream tier, fp32 / NCHW / ragged shapes to the generic kernels.  The GPU tests assert
    that mvf_b200_last_kernel() agrees with this plan after every call."""
    from mvfnet_b200 import _lib
    model_shapes = [(512, 28, 64), (1024, 14, 128), (2048, 7, 256), (512, 32, 64), (1024, 16, 128), (2048, 8, 256)]
    for C, H, Cs in model_shapes:
        for T in (4, 8, 16):
            for N in (1, 12, 160):
                for training in (0, 1):
                    d = _mvf_desc(N, T, C, Cs, H, training=training)
                    assert _lib.plan(d) == "sweep", (C, H, T, N, training)
                    assert _lib.plan(d, True) in EXPECTED_BWD_TIER[(C, H)], (C, H, T, N, training, _lib.plan(d, True))
        assert _lib.plan(_mvf_desc(12, 5, C, Cs, H)) == "stream"
        assert _lib.plan(_mvf_desc(12, 8, C, Cs, H, dtype=0)) == "generic"          # fp32 parity path
        assert _lib.plan(_mvf_desc(12, 8, C, Cs, H, layout=0)) == "generic"         # NCHW
    assert _lib.plan(_mvf_desc(2, 8, 24, 3, 5)) == "generic"                        # ragged channel count
    bad = _mvf_desc(2, 8, 24, 30, 5)                                                # Cs > C
    assert _lib.plan(bad) == ""


# backward tier per (C, H); updated together with the kernels (csrc/api.cu::mvf_bwd)
EXPECTED_BWD_TIER = {(512, 28): ("sweep",), (1024, 14): ("sweep",), (2048, 7): ("sweep",), (512, 32): ("generic",),
                     (1024, 16): ("sweep",), (2048, 8): ("sweep",)}      # 32 x 32 (256 px layer3.0) is inference-only


def test_force_option_and_last_kernel_without_gpu():
    from mvfnet_b200 import _lib
    L = _lib.lib()
    assert _lib.last_kernel() in ("", "sweep", "stream", "ring", "generic")
    d = _mvf_desc(12, 8, 1024, 128, 14)
    try:
        for tier in ("stream", "ring", "generic"):
            _lib.set_option(_lib.OPT_FORCE_FWD, _lib.KERNELS[tier])
            assert _lib.plan(d) == tier
        _lib.set_option(_lib.OPT_FORCE_FWD, _lib.KERNELS["sweep"])
        assert _lib.plan(_mvf_desc(12, 5, 1024, 128, 14)) == ""                     # a forced tier never falls through
    finally:
        _lib.set_option(_lib.OPT_FORCE_FWD, 0)
    assert L.mvf_b200_set_option(99, 1) != 0 and b"unknown option" in L.mvf_b200_last_error()


def test_result_io_and_scoring_vs_reference_golden(tmp_path):
    """mvfnet_b200/results.py (default.pkl dump, top-k / mean-class accuracy, weighted score fusion) against vectors
    produced by the unmodified reference's codes/core/evaluation/accuracy.py (oracle/make_golden_results.py)."""
    import numpy as np
    from conftest import load_cases
    from mvfnet_b200 import results as R
    cases = load_cases("results_cases.npz")
    assert set(cases) == {"small", "k400"}
    for name, c in cases.items():
        scores, scores2 = c["scores"].astype(np.float64), c["scores2"].astype(np.float64)
        labels = [int(v) for v in c["labels"]]
        assert np.allclose(R.top_k_accuracy(scores, labels, k=(1, 5)), c["top"], atol=1e-12)
        assert abs(R.mean_class_accuracy(scores, labels) - float(c["mca"])) < 1e-12
        fused = R.fuse_scores([scores, scores2], [1.0, 0.5])
        assert np.allclose(R.top_k_accuracy(fused, labels, k=(1, 5)), c["fused_top"], atol=1e-12)
        if "fused" in c:
            assert np.allclose(fused, c["fused"], atol=1e-12)
            assert np.allclose(R.softmax(scores), c["softmax"], atol=1e-12)
        # default.pkl round trip: per-video (1, classes) outputs -> one (videos, classes) array
        out = tmp_path / (name + ".pkl")
        stacked = R.dump_results([row[None, :] for row in scores], str(out))
        assert stacked.shape == scores.shape and np.array_equal(R.load_results(str(out)), scores)
        out2 = tmp_path / (name + "_2.pkl")
        R.dump_results(list(scores2), str(out2))
        fused2, report = R.fuse_result_files([str(out), str(out2)], [1.0, 0.5], labels=labels)
        assert np.allclose(fused2, fused) and np.allclose([report["top_k"][1], report["top_k"][5]], c["fused_top"])
    import pytest
    with pytest.raises(ValueError):
        R.dump_results([np.zeros((1, 4))], str(tmp_path / "scores.json"))
    assert R.top_k_accuracy(np.array([[0.1, 0.9, 0.0]]), [[0, 2]], k=(1, 2)) == [0.0, 1.0]     # multi-label sample
