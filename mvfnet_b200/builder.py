"""Registries of the model graph and their builders (API of codes/models/builder.py:6-47)."""
import torch.nn as nn

from .registry import Registry, build_from_cfg

RECOGNIZERS, BACKBONES, HEADS = Registry('recognizer'), Registry('backbone'), Registry('head')
SEGMENTAL_CONSENSUSES = Registry('segmental_consensus')


def build(cfg, registry, default_args=None):
    """A list of configs becomes an nn.Sequential of the built modules."""
    if isinstance(cfg, list):
        return nn.Sequential(*(build_from_cfg(c, registry, default_args) for c in cfg))
    return build_from_cfg(cfg, registry, default_args)


def build_recognizer(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, RECOGNIZERS, dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_backbone(cfg):
    return build(cfg, BACKBONES)


def build_head(cfg):
    return build(cfg, HEADS)


def build_segmental_consensus(cfg):
    return build(cfg, SEGMENTAL_CONSENSUSES)
