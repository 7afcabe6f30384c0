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


# --------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mvfnet_b200 import build_recognizer, _lib
    from mvfnet_b200 import mvf as mvf_mod
    from mvfnet_b200.dist import FlatGrads, init_dist
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
    flat = FlatGrads(model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.015, momentum=0.9, weight_decay=1e-4, nesterov=True)
    params = [p for p in model.parameters() if p.requires_grad]

    g = torch.Generator().manual_seed(1000 + rank)                 # each rank owns different clips
    host_img = [torch.randn((B, T_FRAMES, 3, PX, PX), generator=g).pin_memory() for _ in range(2)]
    host_lbl = [torch.randint(0, 400, (B, 1), generator=g).pin_memory() for _ in range(2)]
    dev_img = [h.to(dev) for h in host_img]
    dev_lbl = [h.to(dev) for h in host_lbl]

    def train_step(img, label):
        flat.zero_()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(img, label)["loss_cls"]
        loss.backward()
        if world > 1:
            flat.gather()
            flat.allreduce_()
        torch.nn.utils.clip_grad_norm_(params, max_norm=40, norm_type=2)
        opt.step()
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

    # ---- resident-input timing (value) + live MVF kernel timing (roofline)
    for i in range(args.warmup):
        train_step(dev_img[i % 2], dev_lbl[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    mvf_mod.timing_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        train_step(dev_img[i % 2], dev_lbl[i % 2])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    timing = mvf_mod.timing_end()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

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
    h2d = host_img[0].numel() * 4 + host_lbl[0].numel() * 8

    if rank != 0:
        return
    # ---- roofline of the dominant hot-path kernel (fused MVF forward), HBM-bound
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    fwd = [(b, s.elapsed_time(e)) for k, b, s, e in timing if k == "mvf_fwd"]
    bwd = [(b, s.elapsed_time(e)) for k, b, s, e in timing if k == "mvf_bwd"]

    def agg(rows):
        if not rows:
            return None, None, 0
        tot_b, tot_ms = sum(r[0] for r in rows), sum(r[1] for r in rows)
        return tot_b / (tot_ms * 1e-3) / 1e9, tot_ms / len(rows) * 1e3, len(rows)

    a_f, us_f, n_f = agg(fwd)
    a_b, us_b, n_b = agg(bwd)
    roofline = {"bound": "hbm", "kernel": "mvf_fwd = mvf_sweep_kernel (fused T/H/W stencil + train-mode BN3d + hardswish, one cooperative launch)", "achieved": a_f,
                "peak": peak, "unit": "GB/s", "frac": (a_f / peak) if a_f else None,
                # ncu --set full (profiles/r01_mvf_v4_sweep_ncu_full.csv, train-mode launch, 14x14 slab, 128 clips):
                # dram read 53.1 MB + write 6.0 MB per launch against 102.8 MB algorithmic -- x is read once (the second
                # sweep is served by L2) and most of the result slab is still dirty in L2 when the kernel ends, the
                # consumer GEMM reads it from there.  Scaled to this run's mean launch.
                "traffic": (0.575 * sum(r[0] for r in fwd) / len(fwd)) if fwd else None,
                "peak_source": peak_src, "launches_timed": n_f, "avg_launch_us": us_f,
                "algorithmic_bytes": "2*E*s per launch (E = B*T*Cs*H*W slab elements, s = 2 B bf16), summed over "
                                     "the 9 MVF modules of R50",
                "mvf_bwd": {"achieved": a_b, "frac": (a_b / peak) if a_b else None, "avg_launch_us": us_b,
                            "launches_timed": n_b, "algorithmic_bytes": "3*E*s"}}
    cpu = cpu_baseline(args.cpu_seconds) if world == 1 else None
    line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "MVFNet-R%d %dx%d (T=%d) 224x224 synthetic, bf16, train forward+backward+"
                                   "clip+SGD step, B=%d clips per GPU%s" % (DEPTH, T_FRAMES, 64 // T_FRAMES, T_FRAMES, B,
                                   " (BASELINE.json configs[1])" if (DEPTH, T_FRAMES) == (50, 8) else ""),
                       "clips_per_gpu": B, "frames_per_gpu": B * T_FRAMES, "parallelism": "dp%d" % world,
                       "l2": "no flush: one step streams >10 GB of activations, far above the 126 MB L2",
                       "peak_hbm_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 1)},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
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
