"""Checkpoint I/O with the reference's key contract (codes/utils/checkpoint.py:178-265): a file is
either a bare state_dict or {'meta', 'state_dict', 'optimizer'}; a leading 'module.' is stripped;
loading is non-strict by default and reports what did not match."""
import time
from collections import OrderedDict

import torch


def load_state_dict(module, state_dict, strict=False, logger=None):
    own = module.state_dict()
    unexpected, shape_mismatch = [], []
    for k, v in state_dict.items():
        if k not in own:
            unexpected.append(k)
        elif tuple(own[k].shape) != tuple(v.shape):
            shape_mismatch.append('%s: checkpoint %s vs model %s' % (k, tuple(v.shape), tuple(own[k].shape)))
        else:
            own[k].copy_(v)
    missing = [k for k in own if k not in state_dict]
    msgs = []
    if unexpected:
        msgs.append('unexpected key in source state_dict: ' + ', '.join(unexpected))
    if missing:
        msgs.append('missing keys in source state_dict: ' + ', '.join(missing))
    msgs += shape_mismatch
    if msgs:
        text = 'The model and loaded state dict do not match exactly\n' + '\n'.join(msgs)
        if strict:
            raise RuntimeError(text)
        (logger.warning if logger is not None else print)(text)
    return missing, unexpected


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    ckpt = torch.load(filename, map_location=map_location)
    if isinstance(ckpt, OrderedDict) or (isinstance(ckpt, dict) and 'state_dict' not in ckpt):
        sd = ckpt
    elif isinstance(ckpt, dict):
        sd = ckpt['state_dict']
    else:
        raise RuntimeError('No state_dict found in checkpoint file {}'.format(filename))
    if sd and next(iter(sd)).startswith('module.'):
        sd = OrderedDict((k[7:], v) for k, v in sd.items())
    target = model.module if hasattr(model, 'module') else model
    with torch.no_grad():
        load_state_dict(target, sd, strict, logger)
    return ckpt


def save_checkpoint(model, filename, optimizer=None, meta=None):
    meta = dict(meta or {})
    meta.setdefault('time', time.asctime())
    target = model.module if hasattr(model, 'module') else model
    ckpt = {'meta': meta, 'state_dict': OrderedDict((k, v.cpu()) for k, v in target.state_dict().items())}
    if optimizer is not None:
        ckpt['optimizer'] = optimizer.state_dict()
    torch.save(ckpt, filename)
