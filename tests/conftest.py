import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # kernel A/B work: run the suite against an experimental build (tools/build_variant.py); the product itself only
    # ever loads mvfnet_b200/libmvf_b200.so
    alt = os.environ.get("MVFB_TEST_LIB")
    if alt:
        from mvfnet_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(alt)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_cases(fname):
    """npz with keys 'case/field' -> {case: {field: array}}"""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
    cases = {}
    for k in z.files:
        c, f = k.split("/", 1)
        cases.setdefault(c, {})[f] = z[k]
    return cases


@pytest.fixture(autouse=True)
def _exact_fp32_library_math():
    """fp32 parity tests compare against an fp64 reference: keep cuDNN / cuBLAS out of TF32."""
    import torch
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
