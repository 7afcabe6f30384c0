"""ctypes binding of libmvf_b200.so (the C ABI declared in include/mvf_b200.h).

The shared library is the product: there is no PyTorch / CPU fallback for anything it exports.  If it
has not been built (`python -m mvfnet_b200.build`) importing this module's `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmvf_b200.so")

MVFB_F32, MVFB_BF16 = 0, 1
MVFB_NCHW, MVFB_NHWC = 0, 1
MODES = {"T": 0, "TH": 1, "THW": 2}
# mvf_b200_set_option keys / kernel tiers (include/mvf_b200.h)
OPT_FORCE_FWD, OPT_FORCE_BWD, OPT_SWEEP_DEBUG, OPT_CONV_HALO_OFF, OPT_GEMM_PAIR_OFF = 0, 1, 2, 3, 4
KERNELS = {"auto": 0, "sweep": 1, "stream": 2, "ring": 3, "generic": 4}


class MvfDesc(C.Structure):
    """mvfb_mvf_desc (include/mvf_b200.h)."""
    _fields_ = [("N", C.c_int), ("T", C.c_int), ("C", C.c_int), ("Cs", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("dtype", C.c_int), ("layout", C.c_int), ("mode", C.c_int), ("use_hs", C.c_int),
                ("training", C.c_int), ("eps", C.c_float), ("momentum", C.c_float)]


_lib = None

_VP, _LL, _SZ, _FP = C.c_void_p, C.c_longlong, C.c_size_t, C.c_void_p


def _declare(l):
    l.mvf_b200_version.restype = C.c_int
    l.mvf_b200_last_error.restype = C.c_char_p
    l.mvf_b200_launch_count.restype = C.c_ulonglong
    l.mvf_fwd_workspace_bytes.restype = C.c_size_t
    l.mvf_fwd_workspace_bytes.argtypes = [C.POINTER(MvfDesc)]
    l.mvf_bwd_workspace_bytes.restype = C.c_size_t
    l.mvf_bwd_workspace_bytes.argtypes = [C.POINTER(MvfDesc)]
    l.mvf_fwd.restype = C.c_int
    l.mvf_fwd.argtypes = [C.POINTER(MvfDesc), _VP, _VP, _LL, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _VP, _SZ, _VP]
    l.mvf_bwd.restype = C.c_int
    l.mvf_bwd.argtypes = [C.POINTER(MvfDesc), _VP, _LL, _VP, _VP, _LL, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP,
                          _FP, _FP, _FP, _FP, _FP, _VP, _SZ, _VP]
    l.mvf_bwd_add.restype = C.c_int
    l.mvf_bwd_add.argtypes = l.mvf_bwd.argtypes[:-1] + [_VP, _VP]
    l.mvf_b200_last_kernel.restype = C.c_char_p
    l.mvf_b200_plan.restype = C.c_char_p
    l.mvf_b200_plan.argtypes = [C.POINTER(MvfDesc), C.c_int]
    l.mvf_b200_set_option.restype = C.c_int
    l.mvf_b200_set_option.argtypes = [C.c_int, C.c_int]


def lib():
    """The loaded library; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmvf_b200.so is not built: run `python -m mvfnet_b200.build` (needs nvcc). "
                "mvfnet_b200 has no CPU / PyTorch fallback for its CUDA kernels.")
        l = C.CDLL(LIB_PATH)
        _declare(l)
        _lib = l
    return _lib


class MvfB200Error(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().mvf_b200_last_error().decode("utf-8", "replace")
        raise MvfB200Error("%s failed with code %d: %s" % (what or "libmvf_b200 call", rc, msg))


def launch_count() -> int:
    return int(lib().mvf_b200_launch_count())


def last_kernel() -> str:
    """Kernel tier ("sweep" / "stream" / "ring" / "generic") that served this thread's last mvf_fwd / mvf_bwd."""
    return lib().mvf_b200_last_kernel().decode()


def plan(desc, backward=False) -> str:
    """Tier mvf_fwd / mvf_bwd selects for a descriptor from its shape alone (host-only)."""
    return lib().mvf_b200_plan(C.byref(desc), int(backward)).decode()


def set_option(key: int, value: int) -> None:
    check(lib().mvf_b200_set_option(key, value), "mvf_b200_set_option")


class force_kernel:
    """Context manager for tests / tools: only the named tier may serve mvf_fwd (`fwd=`) / mvf_bwd (`bwd=`)."""

    def __init__(self, fwd="auto", bwd="auto"):
        self.fwd, self.bwd = KERNELS[fwd], KERNELS[bwd]

    def __enter__(self):
        set_option(OPT_FORCE_FWD, self.fwd)
        set_option(OPT_FORCE_BWD, self.bwd)
        return self

    def __exit__(self, *exc):
        set_option(OPT_FORCE_FWD, 0)
        set_option(OPT_FORCE_BWD, 0)
        return False


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
