"""The whole training step as ONE CUDA graph.

`DistOptimizerHook.after_train_iter`'s sequence (core/dist_utils.py:59-67: zero_grad, forward, backward, all-reduce, clip,
SGD step) is ~620 kernel launches for R50 with no data-dependent host decision in between.  At the reference recipe's
batch (videos_per_gpu = 12, r50_dense.py:122) the GPU finishes them faster than Python + autograd can issue them; captured
once for a fixed batch shape and replayed, the step is bound by the kernels again.  Everything the step needs is
capture-safe by construction: the library never allocates or synchronises, the cooperative MVF launches tag their grid
exchange with a device-resident replay counter (csrc/mvf_sweep.cu), the dropout seed has a device-resident part
(csrc/tail.cu), `FlatSGD.step` computes the clip coefficient on the device.
"""
from __future__ import annotations

import torch

from . import tail


class GraphedTrainStep:
    """`step = GraphedTrainStep(model, optimizer, img_u8, label)`; `loss = step(img_u8, label)` replays the captured
    forward + backward + optimizer step on new data of the same shape.  `img_u8`: (B, T, H, W, 3) uint8 frames (or the
    float (B, T, 3, H, W) wire format with `uint8_input=False`).  `optimizer` is a `tail.FlatSGD`.  The returned loss is a
    view of the graph's static output: read it before the next call."""

    def __init__(self, model, optimizer, img, label, world=1, uint8_input=True, dtype=torch.bfloat16, warmup=3):
        self.model, self.opt, self.world = model, optimizer, world
        self.uint8_input, self.dtype = uint8_input, dtype
        dev = img.device
        self.static_img, self.static_lbl = img.clone(), label.clone()
        self.seed = torch.zeros((), dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                              # cuDNN autotuning, allocator and cache warm-up
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._step()

    def _step(self):
        prev, tail.DEVICE_SEED = tail.DEVICE_SEED, self.seed
        try:
            self.seed.add_(1)
            self.opt.zero_grad()
            img = tail.preprocess_frames(self.static_img) if self.uint8_input else self.static_img
            with torch.autocast("cuda", dtype=self.dtype):
                loss = self.model(img, self.static_lbl)["loss_cls"]
            loss.backward()
            self.opt.step(self.world)
            return loss.detach()
        finally:
            tail.DEVICE_SEED = prev

    def __call__(self, img, label, non_blocking=True):
        self.static_img.copy_(img, non_blocking=non_blocking)
        self.static_lbl.copy_(label, non_blocking=non_blocking)
        self.graph.replay()
        return self.static_loss
