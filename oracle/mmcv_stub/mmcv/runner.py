"""mmcv.runner names the reference imports (never exercised by the golden generator)."""
import torch.distributed as dist


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class Hook:
    pass


class OptimizerHook(Hook):
    def __init__(self, grad_clip=None):
        self.grad_clip = grad_clip


class DistSamplerSeedHook(Hook):
    pass


class Runner:
    def __init__(self, *a, **k):
        raise NotImplementedError("stub")


def obj_from_dict(info, parent=None, default_args=None):
    args = dict(info)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_type = getattr(parent, obj_type)
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    return obj_type(**args)
