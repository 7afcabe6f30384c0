"""Minimal stand-in for the mmcv==0.4.3 surface that whwu95/MVFNet touches at import time.

TEST INFRASTRUCTURE ONLY.  mmcv 0.4.3 (reference requirements.txt:2) is an un-vendored,
un-installable dependency that contributes no arithmetic to the MVF / bottleneck path: only
weight-init helpers, `is_str`, and Runner/Hook base-class *names*.  This stub lets
`oracle/make_golden.py` import the unmodified reference from /root/reference in the build
container to generate the golden vectors under tests/golden/.  It is never imported by the
product package.
"""
import os
from . import cnn, runner, parallel  # noqa: F401

__version__ = "0.4.3-stub"


def is_str(x):
    return isinstance(x, str)


def is_tuple_of(seq, expected_type):
    return isinstance(seq, tuple) and all(isinstance(s, expected_type) for s in seq)


def is_list_of(seq, expected_type):
    return isinstance(seq, list) and all(isinstance(s, expected_type) for s in seq)


def mkdir_or_exist(dir_name, mode=0o777):
    if dir_name:
        os.makedirs(os.path.expanduser(dir_name), mode=mode, exist_ok=True)


class ProgressBar:
    def __init__(self, task_num=0, *a, **k):
        self.task_num = task_num

    def update(self):
        pass


class Config(dict):
    @staticmethod
    def fromfile(path):
        raise NotImplementedError("stub: use mvfnet_b200.config.Config")
