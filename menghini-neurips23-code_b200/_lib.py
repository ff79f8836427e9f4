"""ctypes binding of libgripb200.so (the C ABI declared in include/gripb200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present the
binding raises, it never routes around the CUDA path.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgripb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gripb200.h")


class GripB200Error(RuntimeError):
    pass


_lib = None


def declared_symbols(header: str = HEADER_PATH):
    """Names of every function include/gripb200.h declares (used by the export test)."""
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", src)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GripB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    sig = {
        "gb_create": (c_int, [ctypes.POINTER(P), c_int]),
        "gb_destroy": (c_int, [P]),
        "gb_last_error": (c_char_p, [P]),
        "gb_launch_count": (c_uint64, [P]),
        "gb_version": (c_char_p, []),
        "gb_gemm_f16": (c_int, [P, P, c_int, P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int,
                                c_int, c_int, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Raw device/host address of a torch tensor (None → NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """One gb_ctx per (process, device)."""

    _by_device = {}

    def __init__(self, device: int = 0):
        self.lib = load()
        h = c_void_p()
        rc = self.lib.gb_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise GripB200Error(
                f"gb_create(device={device}) failed with status {rc}: libgripb200 needs an sm_100 "
                "(B200) GPU; there is no CPU fallback"
            )
        self.h = h
        self.device = int(device)

    @classmethod
    def get(cls, device: int = 0) -> "Context":
        ctx = cls._by_device.get(device)
        if ctx is None:
            ctx = cls._by_device[device] = Context(device)
        return ctx

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            msg = self.lib.gb_last_error(self.h)
            raise GripB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    @property
    def launches(self) -> int:
        return int(self.lib.gb_launch_count(self.h))

    # ---- op level -------------------------------------------------------------------------
    def gemm(self, A, W, bias=None, resid=None, out=None, act=0, out_f32=False):
        """out = epi(A @ W.T); A [M,K] fp16, W [N,K] fp16 (nn.Linear layout)."""
        import torch

        M, K = A.shape
        N = W.shape[0]
        if out is None:
            out = torch.empty(M, N, device=A.device, dtype=torch.float32 if out_f32 else torch.float16)
        rc = self.lib.gb_gemm_f16(
            self.h, ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), ptr(resid),
            0 if resid is None else resid.stride(0), ptr(out), out.stride(0), M, N, K, int(act),
            int(bool(out_f32)), stream_ptr())
        self.check(rc, "gb_gemm_f16")
        return out
