#!/usr/bin/env python
"""bench.py -- MVFNet-R50 8x8 (T=8 frames, 224x224 synthetic clips) training step on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step = one pass of the hot path over one batch of B clips per GPU: Recognizer2D forward (ResNet-50
bottleneck stack with the fused MVF CUDA kernels spliced into layer3/4), loss, backward, (N>1: the ONE
gradient all-reduce over NCCL/NVLink), grad-clip + SGD-nesterov step (reference recipe,
configs/MVFNet/K400/mvf_kinetics400_2d_rgb_r50_dense.py:152-160; core/dist_utils.py:59-67).

Prints ONE JSON line.  `value` = clips/s with the batch resident in HBM; `e2e` = the same through the
public API `model(img_group, label)` with the float32 (B,T,3,224,224) batch in pinned HOST memory (H2D copy
and a D2H read of the loss inside the timed region, every step).  `roofline` = the fused MVF forward
kernel's achieved algorithmic HBM GB/s (2*E*s bytes per launch), timed with CUDA events on
the launching stream inside the timed region.  `cpu_baseline` / `--impl reference` = the reference's own
CPU PyTorch path (oracle/mvfnet_ref.py port, pinned to the reference by tests/golden) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "clips/sec (fwd+bwd) MVFNet-R50 8x8 224px"
T_FRAMES, PX, DEPTH = 8, 224, 50
# dram__bytes_read.sum + dram__bytes_write.sum of the train-mode mvf_sweep_kernel launch over its algorithmic bytes, from
# the `ncu --set full` capture of THIS build named below (per launch, like `achieved`); None = not captured
MVF_FWD_TRAFFIC_RATIO = (80.195328 + 17.586176) / 128.45
MVF_FWD_TRAFFIC_SOURCE = ("profiles/r02_mvf_sweep_fwd_ncu.csv (ncu --set full, mvf_sweep_kernel<2, 8, 1>, B = 160 clips, 14x14 slab of "
                          "32.1 M elements: dram read 80.2 MB + write 17.6 MB per launch against 128.5 MB algorithmic -- the second sweep "
                          "reads the slab from L2 and part of the output is still in the 126 MB L2 when the launch ends); scaled to "
                          "this run's mean algorithmic bytes per launch")


def model_cfg(depth=DEPTH, t=T_FRAMES, dropout=0.5):
    """configs/MVFNet/K400/mvf_kinetics400_2d_rgb_r50_dense.py:20-48 with pretrained=None."""
    return dict(
        type="Recognizer2D",
        backbone=dict(type="ResNet", pretrained=None, depth=depth, out_indices=(3,), norm_eval=False,
                      partial_norm=False, norm_cfg=dict(type="BN", requires_grad=True)),
        cls_head=dict(type="TSNClsHead", spatial_size=-1, spatial_type="avg", with_avg_pool=False,
                      temporal_feature_size=1, spatial_feature_size=1, dropout_ratio=dropout, in_channels=2048,
                      init_std=0.01, num_classes=400),
        module_cfg=dict(type="MVF", n_segment=t, alpha=0.125, mvf_freq=(0, 0, 1, 1), mode="THW"))


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU reference
class CpuReference:
    """The reference's own CPU path for this workload: R50 8x8, 224 px, fp32, train forward+backward on all host
    threads, B clips per step (oracle/mvfnet_ref.py: the port of the reference modules that tests/golden pins to
    the unmodified reference; test infrastructure, used here only as the timed CPU baseline)."""

    def __init__(self, batch=1):
        import torch
        from oracle.mvfnet_ref import RefModel, synth_state_dict
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.batch = batch
        self.m = RefModel(synth_state_dict(0, depth=DEPTH, n_segment=T_FRAMES), depth=DEPTH, n_segment=T_FRAMES,
                          dropout_ratio=0.5)
        self.m.training = True
        g = torch.Generator().manual_seed(0)
        self.img = torch.randn((batch, T_FRAMES, 3, PX, PX), generator=g)
        self.label = torch.randint(0, 400, (batch, 1), generator=g)

    def step(self):
        t0 = time.perf_counter()
        for p in self.m.parameters():
            p.grad = None
        loss, _ = self.m.forward_train(self.img, self.label)
        loss.backward()
        return time.perf_counter() - t0

    def describe(self, times):
        med = statistics.median(times)
        return {"value": self.batch / med, "unit": "clips/s", "cores": self.threads, "kind": "port",
                "sample": "%d steps of B=%d clip(s) R50 8x8 224px fp32 train fwd+bwd (median %.0f ms/step), "
                          "oracle/mvfnet_ref.py port of the reference modules, torch %d threads"
                          % (len(times), self.batch, med * 1e3, self.threads)}


def cpu_baseline(budget_s=15.0):
    ref = CpuReference(1)
    ref.step()                                                        # warm-up (oneDNN primitive creation)
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 100):
        times.append(ref.step())
    return ref.describe(times)


def run_reference(args):
    """`--impl reference`: K timed steps (after W warm-ups) of the reference CPU path, one clip per step."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    ref = CpuReference(1)
    for _ in range(args.warmup):
        ref.step()
    times = [ref.step() for _ in range(args.steps)]
    base = ref.describe(times)
    total = sum(times)
    value = ref.batch * args.steps / total
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MVFNet-R50 8x8 (T=8) 224x224 synthetic, train forward+backward, the reference's "
                                   "CPU PyTorch path on the host cores, bounded sample: B=1 clip per step"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- GPU bar
def gpu_bar(dev, world, rank, batches, steps=6, warmup=4):
    """The north-star bar: the UNMODIFIED reference model (baseline/_ref, codes/models/modules/MVF.py:104-138 inside
    codes/models/backbones/resnet.py:208-244, built from the config's model dict) on stock PyTorch / cuDNN on THIS
    GPU: `.cuda()`, cudnn.benchmark (r50_dense.py:180), torch.autocast(bfloat16), weights in channels_last (the faster
    of channels_last / contiguous is kept), and the same step as our arm: forward + backward (+ the reference's own
    `allreduce_grads` when world > 1, core/dist_utils.py:38-49) + clip(40) + SGD-nesterov.  None of this library's
    kernels is on that path.  Returns {"B<clips>": clips/s, ...} with the variant used."""
    import contextlib
    import io
    import torch
    import torch.distributed as dist
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "MVFNet", "codes")):
        return {"unavailable": "baseline/_ref is missing (python tools/install_reference.py in the build container)"}
    sys.path[:0] = [os.path.join(ref_root, "mmcv_stub"), os.path.join(ref_root, "MVFNet")]
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from codes.models import build_recognizer as ref_build
            from codes.core.dist_utils import allreduce_grads as ref_allreduce
    except Exception as e:                                       # pragma: no cover
        return {"unavailable": "reference import failed: %r" % (e,)}
    torch.backends.cudnn.benchmark = True

    def build(channels_last):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            m = ref_build(model_cfg(DEPTH, T_FRAMES), None, None).to(dev).train()
        if channels_last:
            for p_ in m.parameters():
                if p_.dim() == 4:
                    p_.data = p_.data.contiguous(memory_format=torch.channels_last)
        return m

    def measure(m, B):
        opt = torch.optim.SGD(m.parameters(), lr=0.015, momentum=0.9, weight_decay=1e-4, nesterov=True)
        params = [p_ for p_ in m.parameters() if p_.requires_grad]
        g = torch.Generator().manual_seed(2000 + rank)
        img = torch.randn((B, T_FRAMES, 3, PX, PX), generator=g).to(dev)
        label = torch.randint(0, 400, (B, 1), generator=g).to(dev)

        def step():
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = m(img, label)["loss_cls"]
            loss.backward()
            if world > 1:
                ref_allreduce(m.parameters(), True, -1)
            torch.nn.utils.clip_grad_norm_(params, max_norm=40, norm_type=2)
            opt.step()

        for _ in range(warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return world * B * steps / (ms / 1e3)

    out = {"how": "unmodified reference model (baseline/_ref) on torch %s / cuDNN %s, autocast bf16, cudnn.benchmark, "
                  "fwd+bwd+clip+SGD, %d timed steps after %d warm-ups, CUDA events, max over ranks"
                  % (torch.__version__, torch.backends.cudnn.version(), steps, warmup)}
    variants = {}
    for cl in (True, False):
        try:
            m = build(cl)
            variants["channels_last" if cl else "contiguous"] = (measure(m, batches[0]), m)
        except Exception as e:
            variants["channels_last" if cl else "contiguous"] = (0.0, None)
            out["error_%s" % ("channels_last" if cl else "contiguous")] = repr(e)[:200]
            torch.cuda.empty_cache()
    best = max(variants, key=lambda k: variants[k][0])
    out["variant"] = best
    out["variants_B%d" % batches[0]] = {k: v[0] for k, v in variants.items()}
    m = variants[best][1]
    for k, v in variants.items():
        if k != best and v[1] is not None:
            del v
    if m is None:
        out["unavailable"] = "the reference model did not run on this GPU"
        return out
    out["B%d" % batches[0]] = variants[best][0]
    variants.clear()
    torch.cuda.empty_cache()
    for B in batches[1:]:
        try:
            out["B%d" % B] = measure(m, B)
        except torch.OutOfMemoryError:
            out["B%d" % B] = None
            out["note_B%d" % B] = "out of memory on 180 GB"
            torch.cuda.empty_cache()
            if world > 1:
                break
    return out


# --------------------------------------------------------------------------------------- MVF kernels in isolation
def mvf_isolated(dev, B, peak, iters=8):
    """The fused MVF kernels alone on the dominant slab (layer3.1-5 / layer4.0: C = 1024, 14 x 14, Cs = 128, T = 8) at the
    bench's clip count, L2 flushed between launches, CUDA events on the launching stream: achieved algorithmic GB/s
    (forward 2*E*s, backward 3*E*s) of the eval-mode forward (one sweep), the train-mode forward (statistics sweep +
    exchange + apply sweep in one cooperative launch) and the train-mode backward."""
    import torch
    from mvfnet_b200 import MVF
    from mvfnet_b200 import mvf as mm
    from mvfnet_b200.mvf import mvf_slab_forward
    T, C, H, Cs = 8, 1024, 14, 128
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    x = torch.randn(B * T, H, H, C, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
    g = torch.randn(B * T, H, H, C, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
    E = B * T * Cs * H * H
    out = {"slab": "C=1024 14x14 Cs=128 T=8, %d clips, L2 flushed" % B}
    # what ANY kernel moving these bytes achieves at this size, timed the same way: a plain contiguous device copy of as many
    # bytes (ATen's copy kernel; measurement only -- tools/ubench/slab_copy.cu shows the slab's strided layout costs the
    # same).  Launch + ramp + tail cost ~6 us per launch, so a 128 MB transfer cannot reach the 4 GB copy's rate that
    # `peak` records (profiles/r02_mvf_fwd_ceiling_experiments.txt)
    slab_src = torch.empty(E, dtype=torch.bfloat16, device=dev).normal_()
    slab_dst = torch.empty_like(slab_src)
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slab_dst.copy_(slab_src)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    copy_us = sorted(ts[2:])[len(ts[2:]) // 2]
    out["plain_copy_same_bytes"] = {"us": copy_us, "achieved": 2 * E * 2 / copy_us / 1e3, "frac": 2 * E * 2 / copy_us / 1e3 / peak,
                                    "what": "contiguous device-to-device copy of E*s bytes (2*E*s moved), same timing"}
    for training in (False, True):
        m = MVF(torch.nn.Identity(), T, C, alpha=0.125).to(dev).train(training)
        cfg = m._cfg()
        wt, wh, ww = (w.detach().float().contiguous() for w in m._taps())
        ts = []
        for i in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mvf_slab_forward(x.detach(), cfg, wt, wh, ww, m.bn.weight.detach(), m.bn.bias.detach(), m.bn.running_mean,
                             m.bn.running_var, out="slab")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts[2:])[len(ts[2:]) // 2]
        key = "fwd_train" if training else "fwd_eval"
        out[key] = {"us": us, "achieved": 2 * E * 2 / us / 1e3, "frac": 2 * E * 2 / us / 1e3 / peak,
                    "frac_of_plain_copy": copy_us / us}
        if training:
            y = m.fuse(x)
            mm.timing_begin()
            for i in range(iters + 2):
                flush.zero_()
                x.grad = None
                y.backward(g, retain_graph=True)
            torch.cuda.synchronize()
            rec = sorted(s0.elapsed_time(s1) * 1e3 for k, b, s0, s1, _ in mm.timing_end()[2:] if k == "mvf_bwd")
            us = rec[len(rec) // 2]
            out["bwd_train"] = {"us": us, "achieved": 3 * E * 2 / us / 1e3, "frac": 3 * E * 2 / us / 1e3 / peak}
    return out


# --------------------------------------------------------------------------------------- the other BASELINE.json configs
def light_train_config(dev, world, rank, depth, t, B, steps=6, warmup=3):
    """configs[2] (R50 16x4) / configs[3] (R101 8x8): the same training step as the headline (uint8 frames, bf16,
    forward + backward + all-reduce + clip + SGD), a few steps, resident inputs -> whole-job clips/s."""
    import gc
    import torch
    import torch.distributed as dist
    from mvfnet_b200 import build_recognizer
    from mvfnet_b200.tail import FlatSGD, preprocess_frames
    from mvfnet_b200.utils import to_channels_last
    torch.manual_seed(0)
    model = to_channels_last(build_recognizer(model_cfg(depth, t), None, None).to(dev)).train()
    if world > 1:
        for v in model.state_dict().values():
            dist.broadcast(v, 0)
    opt = FlatSGD(model.parameters(), lr=0.015, momentum=0.9, weight_decay=1e-4, nesterov=True, max_norm=40)
    g = torch.Generator().manual_seed(4000 + rank)
    img = torch.randint(0, 256, (B, t, PX, PX, 3), generator=g, dtype=torch.uint8).to(dev)
    lbl = torch.randint(0, 400, (B, 1), generator=g).to(dev)

    def step():
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(preprocess_frames(img), lbl)["loss_cls"]
        loss.backward()
        opt.step(world)

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    n_params = sum(p.numel() for p in model.parameters())
    del model, opt, img, lbl
    gc.collect()
    torch.cuda.empty_cache()
    return {"value": world * B * steps / (ms / 1e3), "unit": "clips/s", "ms_per_step": ms / steps, "clips_per_gpu": B,
            "steps": steps, "allreduce_mb": round(n_params * 4 / 1e6, 1) if world > 1 else 0.0}


def inference_config(dev, world, rank, videos=12, px=256, clips=30):
    """configs[4]: R50 8x8, 256 x 256, 3 crops x 10 clips per video, fcn_testing, eval-mode BatchNorm (test_recognizer.py:72-77,
    recognizer2d.py:151-179, tsn_clshead.py:99-117): one video (30 clips = 240 frames) per rank per step, no collective.
    `graph`: the captured CUDA graph of the eval network (mvfnet_b200/infer.py); `eager`: the same kernels launched from
    Python; `e2e`: uint8 frames from pinned host memory + the (1, 400) scores read back, per video.  videos/s is the
    whole job (all ranks)."""
    import gc
    import torch
    import torch.distributed as dist
    from mvfnet_b200 import build_recognizer, _lib
    from mvfnet_b200.infer import GraphedInference
    from mvfnet_b200.utils import to_channels_last
    cfg = model_cfg(50, 8, dropout=0.5)
    cfg["fcn_testing"] = True
    cfg["cls_head"]["fcn_testing"] = True
    torch.manual_seed(0)
    model = to_channels_last(build_recognizer(cfg, None, dict(average_clips="prob")).to(dev)).eval()
    g = torch.Generator(device=dev).manual_seed(7)
    with torch.no_grad():                                            # non-trivial running statistics (they fold into the GEMMs)
        for m_ in model.modules():
            if isinstance(m_, torch.nn.modules.batchnorm._BatchNorm):
                m_.running_mean.normal_(0, 0.1, generator=g)
                m_.running_var.uniform_(0.5, 1.5, generator=g)
    gh = torch.Generator().manual_seed(6000 + rank)
    host = [torch.randint(0, 256, (1, clips * 8, px, px, 3), generator=gh, dtype=torch.uint8).pin_memory() for _ in range(2)]
    devv = [h.to(dev) for h in host]
    eng = GraphedInference(model, devv[0], uint8_input=True)

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    l0 = _lib.launch_count()
    eng(devv[0])
    launches = _lib.launch_count() - l0                               # a replay launches nothing from the host ...
    ms_graph = timed(lambda i: eng(devv[i % 2]), videos)
    eng._forward(devv[0])                                             # untimed: the caching allocator's first eager pass after capture
    ms_eager = timed(lambda i: eng._forward(devv[i % 2]), max(3, videos // 3)) / max(3, videos // 3) * videos
    ms_e2e = timed(lambda i: eng(host[i % 2]).float().cpu(), videos)
    out = {"workload": "MVFNet-R50 8x8 %dx%d, %d clips (%d frames) per video, fcn_testing, eval BatchNorm folded into the "
                       "convolution epilogues, uint8 frames normalised on the GPU" % (px, px, clips, clips * 8),
           "videos_per_s": world * videos / (ms_graph / 1e3), "clips_per_s": world * videos * clips / (ms_graph / 1e3),
           "ms_per_video": ms_graph / videos, "eager_videos_per_s": world * videos / (ms_eager / 1e3),
           "e2e_videos_per_s": world * videos / (ms_e2e / 1e3), "h2d_bytes_per_video": host[0].numel(),
           "d2h_bytes_per_video": 1600, "kernels_in_graph": "see profiles/", "videos_timed": videos}
    del eng, model, devv, host
    gc.collect()
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mvfnet_b200 import build_recognizer, _lib
    from mvfnet_b200 import mvf as mvf_mod
    from mvfnet_b200.dist import init_dist
    from mvfnet_b200.tail import FlatSGD, preprocess_frames
    from mvfnet_b200.utils import to_channels_last

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        init_dist("pytorch", backend="nccl")
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True                         # cfg: cudnn_benchmark = True (r50_dense.py:180)
    B = args.batch
    torch.manual_seed(0)
    model = to_channels_last(build_recognizer(model_cfg(DEPTH, T_FRAMES), None, None).to(dev)).train()
    if world > 1:                                                  # MMDistributedDataParallel: broadcast once
        for t in model.state_dict().values():
            dist.broadcast(t, 0)
    # optimizer + grad_clip of the recipe (r50_dense.py:152-154) on flat buffers: one all-reduce, one fused update
    opt = FlatSGD(model.parameters(), lr=0.015, momentum=0.9, weight_decay=1e-4, nesterov=True, max_norm=40)
    u8 = args.input == "u8"

    def make_batches(b, seed):
        g = torch.Generator().manual_seed(seed + rank)             # each rank owns different clips
        if u8:      # decoded frames as the pipeline's FrameSelector delivers them: uint8 HWC (B, T, H, W, 3)
            himg = [torch.randint(0, 256, (b, T_FRAMES, PX, PX, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
        else:       # the reference's wire format: float32 (B, T, 3, H, W), already normalised on the CPU
            himg = [torch.randn((b, T_FRAMES, 3, PX, PX), generator=g).pin_memory() for _ in range(2)]
        hlbl = [torch.randint(0, 400, (b, 1), generator=g).pin_memory() for _ in range(2)]
        return himg, hlbl

    host_img, host_lbl = make_batches(B, 1000)
    dev_img = [h.to(dev) for h in host_img]
    dev_lbl = [h.to(dev) for h in host_lbl]

    def train_step(img, label):
        """DistOptimizerHook.after_train_iter's order (core/dist_utils.py:59-67): zero_grad, forward, backward, ONE
        all-reduce, clip, SGD step -- the last three inside FlatSGD.step()."""
        opt.zero_grad()
        if u8:
            img = preprocess_frames(img)                             # Normalize + FormatShape on the GPU
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(img, label)["loss_cls"]
        loss.backward()
        opt.step(world)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(imgs, lbls, n):
        """n steps between two CUDA events on the compute stream, barrier + synchronize on both sides, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            train_step(imgs[i % 2], lbls[i % 2])
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- resident-input timing (value) + live per-family kernel timing (roofline, roofline_by_family)
    for i in range(args.warmup):
        train_step(dev_img[i % 2], dev_lbl[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    mvf_mod.timing_begin()
    ms = timed(dev_img, dev_lbl, args.steps)
    launches = _lib.launch_count() - launches0
    timing = mvf_mod.timing_end()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)
    # the timed steps must have been real training steps: a finite cross-entropy near ln(400) on random labels, a finite
    # gradient norm (a wrong BatchNorm statistic or a skipped kernel shows up here as inf / nan / 1e7)
    loss_check = float(train_step(dev_img[0], dev_lbl[0]).float().item())
    gnorm_check = float(opt.grad_norm.item())
    if not (0.0 < loss_check < 30.0) or not (0.0 < gnorm_check < 1e6):
        raise SystemExit("bench.py: the training step is numerically broken (loss %r, grad norm %r)" % (loss_check, gnorm_check))

    if args.kernels_only:
        if rank == 0:
            print(json.dumps({"kernels_only": True, "value": value, "ms_per_step": ms / args.steps,
                              "gpu_launches": int(launches)}))
        return

    # ---- end-to-end timing through the public API with host batches (double-buffered H2D on a side stream)
    copy_stream = torch.cuda.Stream(device=dev)
    slots_img = [torch.empty_like(dev_img[0]) for _ in range(2)]
    slots_lbl = [torch.empty_like(dev_lbl[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            slots_img[s].copy_(host_img[s], non_blocking=True)
            slots_lbl[s].copy_(host_lbl[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(2):
            freed[s].record()
        prefetch(0)
        losses = []
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            loss = train_step(slots_img[s], slots_lbl[s])
            freed[s].record()
            losses.append(loss.detach().float().cpu())             # D2H read of the step's result, every step
        return losses

    e2e_loop(max(2, min(args.warmup, 3)))
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps)
    t1.record()
    barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = host_img[0].numel() * host_img[0].element_size() + host_lbl[0].numel() * 8
    peak_gb = round(torch.cuda.max_memory_allocated(dev) / 2**30, 1)
    del slots_img, slots_lbl, dev_img, dev_lbl, host_img, host_lbl

    # ---- batch sweep of OUR arm (resident inputs): the reference recipe's B = 12 (r50_dense.py:122) and B = 64
    sweep = {"B%d" % B: value}
    for b in [int(v) for v in args.sweep.split(",") if v]:
        if b == B:
            continue
        himg, hlbl = make_batches(b, 3000)
        dimg, dlbl = [h.to(dev) for h in himg], [h.to(dev) for h in hlbl]
        for i in range(3):
            train_step(dimg[i % 2], dlbl[i % 2])
        n = max(6, min(args.steps, 20))
        sweep["B%d" % b] = world * b * n / (timed(dimg, dlbl, n) / 1e3)
        del himg, hlbl, dimg, dlbl

    # ---- the same step captured ONCE into a CUDA graph and replayed (mvfnet_b200/graph.py): what the step costs when the
    # host is out of the way -- decisive at the recipe's B = 12, where Python + autograd cannot issue ~620 launches as
    # fast as the GPU retires them
    import gc
    graphed = None
    if not args.no_graph and (world == 1 or os.environ.get("MVFB_GRAPH_DDP") == "1"):
        from mvfnet_b200.graph import GraphedTrainStep
        graphed = {}
        for b in [B] + [int(v) for v in args.sweep.split(",") if v and int(v) != B]:
            try:
                gc.collect()
                torch.cuda.empty_cache()
                himg, hlbl = make_batches(b, 5000)
                dimg, dlbl = [h.to(dev) for h in himg], [h.to(dev) for h in hlbl]
                step = GraphedTrainStep(model, opt, dimg[0], dlbl[0], world=world, uint8_input=u8)
                n = max(6, min(args.steps, 20))
                for i in range(2):
                    step(dimg[i % 2], dlbl[i % 2])
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for i in range(n):
                    step(dimg[i % 2], dlbl[i % 2])
                g1.record()
                barrier()
                graphed["B%d" % b] = world * b * n / (max_over_ranks(g0.elapsed_time(g1)) / 1e3)
                del step, himg, hlbl, dimg, dlbl
            except Exception as e:                               # capture is an optimisation, never a requirement
                graphed["B%d" % b] = None
                graphed["error_B%d" % b] = repr(e)[:300]
                break
    del model, opt
    gc.collect()
    torch.cuda.empty_cache()
    # ---- BASELINE.json configs[2..4], driver-visible beside the headline: short runs of the same code paths
    others = None
    if not args.no_other_configs and (DEPTH, T_FRAMES) == (50, 8):
        others = {"R50_16x4_train (configs[2])": light_train_config(dev, world, rank, 50, 16, 64),
                  "R101_8x8_train (configs[3])": light_train_config(dev, world, rank, 101, 8, 80),
                  "R50_8x8_256px_fcn_test (configs[4])": inference_config(dev, world, rank)}

    # ---- the GPU bar: the unmodified reference on PyTorch / cuDNN on the same GPU(s), same run
    bar = None
    if not args.no_gpu_bar:
        bar = gpu_bar(dev, world, rank, [12, 64] + ([B] if B not in (12, 64) else []))
        if "B%d" % B in bar and bar.get("B%d" % B):
            bar["ratio_equal_B"] = value / bar["B%d" % B]
        for b in (12, 64):
            if bar.get("B%d" % b) and sweep.get("B%d" % b):
                bar["ratio_B%d" % b] = sweep["B%d" % b] / bar["B%d" % b]
            if bar.get("B%d" % b) and graphed and graphed.get("B%d" % b):
                bar["ratio_B%d_cuda_graph" % b] = graphed["B%d" % b] / bar["B%d" % b]
        best_ref = max([v for k, v in bar.items() if k.startswith("B") and isinstance(v, float)], default=None)
        if best_ref:
            bar["value"], bar["unit"] = best_ref, "clips/s"
            bar["ratio"] = value / best_ref                         # our headline over the reference's best batch size

    if rank != 0:
        return
    # ---- roofline of the hot-path kernel the north star names (fused MVF forward), HBM-bound, and of every family
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_tf, peak_src = 1400.0, "fallback 6.65 TB/s / 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    peak = 6650.0
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak, peak_tf = float(pk["hbm_gbs"]), float(pk.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy) / bf16_tflops_sustained"
    fams = {}
    for kind, nbytes, s0, s1, flops in timing:
        f = fams.setdefault(kind, [0, 0.0, 0.0, 0.0])
        f[0] += 1
        f[1] += s0.elapsed_time(s1)
        f[2] += nbytes
        f[3] += flops
    by_family, timed_ms = {}, 0.0
    for kind, (cnt, fms, nbytes, flops) in sorted(fams.items(), key=lambda kv: -kv[1][1]):
        gbs = nbytes / (fms * 1e-3) / 1e9
        row = {"launches_per_step": cnt / args.steps, "ms_per_step": fms / args.steps,
               "share_of_step": fms / ms, "achieved_gbs": gbs, "frac_hbm": gbs / peak}
        if flops:
            tf = flops / (fms * 1e-3) / 1e12
            row["achieved_tflops"], row["frac_tensor"] = tf, tf / peak_tf
        by_family[kind] = row
        timed_ms += fms
    by_family["other (ATen / NCCL / gaps, not bracketed)"] = {"ms_per_step": (ms - timed_ms) / args.steps,
                                                              "share_of_step": 1.0 - timed_ms / ms}
    f_fwd, f_bwd = fams.get("mvf_fwd"), fams.get("mvf_bwd")
    a_f = f_fwd[2] / (f_fwd[1] * 1e-3) / 1e9 if f_fwd else None
    a_b = f_bwd[2] / (f_bwd[1] * 1e-3) / 1e9 if f_bwd else None
    roofline = {"bound": "hbm", "kernel": "mvf_fwd = mvf_sweep_kernel (fused T/H/W stencil + train-mode BN3d + hardswish, one cooperative launch)", "achieved": a_f,
                "peak": peak, "unit": "GB/s", "frac": (a_f / peak) if a_f else None,
                "traffic": MVF_FWD_TRAFFIC_RATIO * f_fwd[2] / f_fwd[0] if (f_fwd and MVF_FWD_TRAFFIC_RATIO) else None,
                "traffic_source": MVF_FWD_TRAFFIC_SOURCE,
                "peak_source": peak_src, "launches_timed": f_fwd[0] if f_fwd else 0,
                "avg_launch_us": f_fwd[1] / f_fwd[0] * 1e3 if f_fwd else None,
                "algorithmic_bytes": "2*E*s per launch (E = B*T*Cs*H*W slab elements, s = 2 B bf16), summed over "
                                     "the 9 MVF modules of R50",
                "mvf_bwd": {"achieved": a_b, "frac": (a_b / peak) if a_b else None,
                            "avg_launch_us": f_bwd[1] / f_bwd[0] * 1e3 if f_bwd else None,
                            "launches_timed": f_bwd[0] if f_bwd else 0, "algorithmic_bytes": "3*E*s"}}
    try:
        roofline["isolated"] = mvf_isolated(dev, B, peak)
    except Exception as e:                                       # informational only
        roofline["isolated"] = {"error": repr(e)[:200]}
    cpu = cpu_baseline(args.cpu_seconds) if world == 1 else None
    line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "MVFNet-R%d %dx%d (T=%d) 224x224 synthetic, bf16, train forward+backward+"
                                   "clip+SGD step, B=%d clips per GPU%s" % (DEPTH, T_FRAMES, 64 // T_FRAMES, T_FRAMES, B,
                                   " (BASELINE.json configs[1])" if (DEPTH, T_FRAMES) == (50, 8) else ""),
                       "clips_per_gpu": B, "frames_per_gpu": B * T_FRAMES, "parallelism": "dp%d" % world,
                       "input": ("uint8 (B,T,H,W,3) decoded frames, Normalize + FormatShape on the GPU (preprocess_u8)" if u8
                                 else "float32 (B,T,3,H,W) normalised on the host (the reference's wire format)"),
                       "l2": "no flush: one step streams >10 GB of activations, far above the 126 MB L2",
                       "peak_hbm_gb": peak_gb},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "sanity": {"loss_after_timed_steps": loss_check, "grad_norm": gnorm_check},
            "clocks": clocks, "roofline": roofline, "roofline_by_family": by_family,
            "sweep": sweep}
    if graphed is not None:
        line["cuda_graph_step"] = dict(graphed, how="the identical step (uint8 frames -> loss -> backward -> clip -> SGD) captured once, "
                                       "replayed per batch incl. the device-side copy of the batch into the graph's input")
    if bar is not None:
        line["gpu_bar"] = bar
    if others is not None:
        line["other_configs"] = others
    if (DEPTH, T_FRAMES) == (50, 8):
        # per-layer max(tensor time, min HBM traffic time) bound of the conv stack, fwd+bwd
        bound = 4680.0 * world
        line["conv_roofline"] = {"bound_clips_per_s": bound, "frac": value / bound,
                                 "how": "BASELINE.md: sum over conv layers of max(2*MACs/1383 TF/s, min bf16 bytes/6453 GB/s), x3 for fwd+bwd"}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def main():
    global DEPTH, T_FRAMES, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=160,
                    help="clips per GPU (measured on B200, final build: 128 -> 1425, 160 -> 1471, 192 -> 1475 clips/s)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--depth", type=int, default=DEPTH, help="ResNet depth (50: BASELINE configs[1]; 101: configs[3])")
    ap.add_argument("--frames", type=int, default=T_FRAMES, help="frames per clip T (8: configs[1]; 16: configs[2])")
    ap.add_argument("--input", default="u8", choices=["u8", "f32"],
                    help="u8: decoded uint8 frames, normalised on the GPU; f32: the reference's float32 wire format")
    ap.add_argument("--sweep", default="12,64", help="extra clips-per-GPU sizes our arm is also timed at (resident inputs)")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay measurement of the step")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of BASELINE.json configs[2..4]")
    ap.add_argument("--no-gpu-bar", action="store_true", help="skip the reference-on-PyTorch/cuDNN arm (gpu_bar)")
    ap.add_argument("--kernels-only", action="store_true",
                    help="profiling runs (ncu): skip the e2e loop and the cpu_baseline sample")
    args = ap.parse_args()
    if (args.depth, args.frames) != (DEPTH, T_FRAMES):
        DEPTH, T_FRAMES = args.depth, args.frames
        METRIC = "clips/sec (fwd+bwd) MVFNet-R%d %dx%d 224px" % (DEPTH, T_FRAMES, 64 // T_FRAMES)
    if args.impl == "ours" and not args.kernels_only:
        args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
