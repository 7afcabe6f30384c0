"""Inference engine for `Recognizer2D.forward_test` (codes/models/recognizers/recognizer2d.py:151-179 as driven by
test_recognizer.py:72-77: model.eval(), optional fcn_testing, average_clips).

The eval-mode network is a fixed sequence of this library's kernels (every BatchNorm / ReLU / residual add folded into
the epilogue of the convolution before it, mvfnet_b200/resnet.py::Bottleneck._infer): ~80 launches for R50 with no
host-side decisions in between.  `GraphedInference` captures that sequence ONCE into a CUDA graph for a fixed input
shape and replays it per video -- launch latency and the Python dispatch disappear, which is what bounds the
reference's 30-clip test videos (240 frames of 256 x 256) on a GPU this fast.
"""
from __future__ import annotations

import torch


class GraphedInference:
    """`engine = GraphedInference(model, example)`; `scores = engine(img_group)` with img_group shaped like `example`
    ((1, clips*T, 3, H, W) float / bf16, or (1, clips*T, H, W, 3) uint8 decoded frames when `uint8_input`).
    Returns the (1, num_classes) averaged clip scores as a CUDA tensor (a view of the graph's static output: consume it
    before the next call)."""

    def __init__(self, model, example, uint8_input=False, dtype=torch.bfloat16, warmup=2):
        if not example.is_cuda:
            raise RuntimeError("GraphedInference needs CUDA tensors (the product has no CPU path)")
        self.model = model.eval()
        self.uint8_input = uint8_input
        self.dtype = dtype
        self.static_in = example.clone()
        stream = torch.cuda.Stream(device=example.device)
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            for _ in range(warmup):                                # allocator warm-up, cuDNN-free: only our kernels run
                self._forward(self.static_in)
        torch.cuda.current_stream().wait_stream(stream)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._forward(self.static_in)

    def _forward(self, img):
        with torch.no_grad(), torch.autocast("cuda", dtype=self.dtype):
            if self.uint8_input:
                from .tail import preprocess_frames
                img = preprocess_frames(img)
            return self.model(img, None, return_loss=False, return_numpy=False)

    def __call__(self, img_group, non_blocking=True):
        self.static_in.copy_(img_group, non_blocking=non_blocking)
        self.graph.replay()
        return self.static_out
