"""Data-parallel plumbing: one process per GPU, ONE gradient all-reduce per step over NVLink (NCCL).

Mirrors codes/core/dist_utils.py:15-92 (`init_dist`, `allreduce_grads`, `DistOptimizerHook`) and
core/parallel/distributed.py:11-62 (`MMDistributedDataParallel`: broadcast at construction, no grad
hooks).  B200 re-design of the collective: the reference flattens every gradient into a fresh buffer,
all-reduces it, divides and copies back (three extra passes over 97 MB for R50).  Here the gradients
LIVE in one persistent flat fp32 buffer (`FlatGrads`: every `param.grad` is a view), so the step is a
single in-place `all_reduce` on that buffer plus a scale; `torch.distributed` does the plumbing.
"""
import os

import torch
import torch.distributed as dist
import torch.nn as nn


def init_dist(launcher='pytorch', backend='nccl', **kwargs):
    """env:// rendezvous, one CUDA device per local rank (dist_utils.py:70-92)."""
    if launcher != 'pytorch':
        raise ValueError('Invalid launcher type: {} (only "pytorch"/torchrun is supported)'.format(launcher))
    rank = int(os.environ['RANK'])
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank % max(torch.cuda.device_count(), 1))))
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, **kwargs)


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class FlatGrads:
    """One persistent flat gradient buffer per dtype, in parameter order (the reference's bucket-by-type order,
    dist_utils.py:20-27).

    Per step: `zero_()` drops the previous gradients (`param.grad = None`, so autograd's AccumulateGrad simply
    adopts the freshly computed tensor instead of launching an add per parameter -- ncu showed ~450 such launches
    per step when .grad was pre-populated); after backward `gather()` packs all gradients into the flat buffer
    with one multi-tensor copy and re-points every `param.grad` at its slice, so the all-reduce, the norm clip and
    the optimizer all work in place on flat memory; `allreduce_()` is the single collective of the step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.buffers, self.groups, self.views = {}, {}, {}
        for p in self.params:
            self.groups.setdefault((p.dtype, p.device), []).append(p)
        for key, ps in self.groups.items():
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=key[0], device=key[1])
            views, off = [], 0
            for p in ps:
                views.append(flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
            self.buffers[key], self.views[key] = flat, views

    def restride(self, params):
        """Give every flat view its parameter's own (dense) strides -- FlatSGD moves channels_last weights into flat
        memory in their own element order, and its elementwise update must pair gradient i with parameter i."""
        for key, ps in self.groups.items():
            flat, views, off = self.buffers[key], [], 0
            for p in ps:
                dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
                v = flat[off:off + p.numel()]
                views.append(v.as_strided(p.shape, p.stride()) if dense else v.view(p.shape))
                off += p.numel()
            self.views[key] = views

    def zero_(self):
        """Replaces optimizer.zero_grad(): no kernel at all."""
        from . import ops
        ops.new_backward_pass()
        for p in self.params:
            p.grad = None

    def gather(self):
        """Pack the gradients autograd produced into the flat buffers; params without a gradient contribute 0."""
        for key, ps in self.groups.items():
            views = self.views[key]
            src, dst = [], []
            for p, v in zip(ps, views):
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    src.append(p.grad)
                    dst.append(v)
            if src:
                torch._foreach_copy_(dst, src)
            for p, v in zip(ps, views):
                p.grad = v

    def numel(self):
        return sum(f.numel() for f in self.buffers.values())

    def allreduce_(self, average=True):
        """The one collective of the training step (dist_utils.py:29-32): sum over ranks, / world."""
        rank, world = get_dist_info()
        if world == 1:
            return
        for flat in self.buffers.values():
            dist.all_reduce(flat)
            if average:
                flat.div_(world)


def _allreduce_coalesced(tensors, world_size, bucket_size_mb=-1):
    """Fallback for gradients that do not live in a FlatGrads buffer: one flat all-reduce per dtype."""
    by_type = {}
    for t in tensors:
        by_type.setdefault(t.type(), []).append(t)
    for bucket in by_type.values():
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat)
        flat.div_(world_size)
        off = 0
        for t in bucket:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def allreduce_grads(params, coalesce=True, bucket_size_mb=-1):
    """Average gradients over all ranks (dist_utils.py:38-49).  `params` may be a FlatGrads."""
    if isinstance(params, FlatGrads):
        params.allreduce_()
        return
    grads = [p.grad.data for p in params if p.requires_grad and p.grad is not None]
    world_size = dist.get_world_size()
    if coalesce:
        _allreduce_coalesced(grads, world_size, bucket_size_mb)
    else:
        for t in grads:
            dist.all_reduce(t.div_(world_size))


class MMDistributedDataParallel(nn.Module):
    """Broadcast parameters and buffers from rank 0 once, then just call the module
    (core/parallel/distributed.py:11-62: no gradient hooks -- the optimizer hook all-reduces)."""

    def __init__(self, module, dim=0, broadcast_buffers=True, bucket_cap_mb=25):
        super().__init__()
        self.module, self.dim, self.broadcast_buffers = module, dim, broadcast_buffers
        self.broadcast_bucket_size = bucket_cap_mb * 1024 * 1024
        self._sync_params()

    def _sync_params(self):
        rank, world = get_dist_info()
        if world == 1:
            return
        tensors = list(self.module.state_dict().values())
        if not self.broadcast_buffers:
            names = {n for n, _ in self.module.named_buffers()}
            tensors = [v for k, v in self.module.state_dict().items() if k not in names]
        for t in tensors:
            dist.broadcast(t, 0)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)


class DistOptimizerHook:
    """zero_grad -> backward -> one all-reduce -> clip -> step (dist_utils.py:52-67)."""

    def __init__(self, grad_clip=None, coalesce=True, bucket_size_mb=-1):
        self.grad_clip, self.coalesce, self.bucket_size_mb = grad_clip, coalesce, bucket_size_mb

    def clip_grads(self, params):
        params = [p for p in params if p.requires_grad and p.grad is not None]
        return torch.nn.utils.clip_grad_norm_(params, **self.grad_clip)

    def after_train_iter(self, runner):
        flat = getattr(runner, 'flat_grads', None)
        if flat is not None:
            flat.zero_()
        else:
            runner.optimizer.zero_grad()
        runner.outputs['loss'].backward()
        if flat is not None:
            if get_dist_info()[1] > 1:
                flat.gather()
                flat.allreduce_()
        elif get_dist_info()[1] > 1:
            allreduce_grads(runner.model.parameters(), self.coalesce, self.bucket_size_mb)
        if self.grad_clip is not None:
            self.clip_grads(runner.model.parameters())
        runner.optimizer.step()
