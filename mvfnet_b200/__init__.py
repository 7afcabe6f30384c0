"""mvfnet_b200: B200-native MVF module + ResNet bottleneck hot path of whwu95/MVFNet.

Python here is the thin host-side mirror of the reference's operator interface; every kernel on the path
lives in libmvf_b200.so (csrc/, C ABI in include/mvf_b200.h).
"""
from .builder import BACKBONES, HEADS, RECOGNIZERS, build_backbone, build_head, build_recognizer  # noqa: F401
from .registry import Registry, build_from_cfg  # noqa: F401
from .mvf import MVF, make_multi_view_fusion  # noqa: F401
from .resnet import Bottleneck, ResNet, make_res_layer  # noqa: F401
from .heads import TSNClsHead, SimpleConsensus  # noqa: F401
from .recognizer import BaseRecognizer, Recognizer2D  # noqa: F401
from .config import Config  # noqa: F401

__version__ = "0.1.0"
