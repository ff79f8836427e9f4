"""ctypes binding of libgripb200.so (the C ABI declared in include/gripb200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present the
binding raises, it never routes around the CUDA path.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
# GRIPB200_LIB names another build of the same library (A/B measurements of build-time switches)
LIB_PATH = os.environ.get("GRIPB200_LIB") or os.path.join(_HERE, "libgripb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gripb200.h")


class GripB200Error(RuntimeError):
    pass


P = c_void_p


class ProfileStats(Structure):
    _fields_ = [("launches", c_uint64), ("ms", ctypes.c_double), ("work", ctypes.c_double)]


class ProfileLaunch(Structure):
    _fields_ = [("kind", c_int), ("m", c_int), ("n", c_int), ("k", c_int), ("ms", ctypes.c_double),
                ("work", ctypes.c_double)]


class BlockWeights(Structure):
    _fields_ = [(n, P) for n in (
        "ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_g", "ln2_b", "w_fc", "b_fc",
        "w_proj", "b_proj", "w_qkv_t", "w_o_t", "w_fc_t", "w_proj_t", "s_qkv", "s_fc")]


class VitWeights(Structure):
    _fields_ = [("width", c_int), ("layers", c_int), ("heads", c_int), ("out_dim", c_int),
                ("conv_w", P), ("cls", P), ("pos", P), ("ln_pre_g", P), ("ln_pre_b", P),
                ("ln_post_g", P), ("ln_post_b", P), ("proj_t", P), ("proj", P),
                ("blocks", POINTER(BlockWeights))]


class TextWeights(Structure):
    _fields_ = [("width", c_int), ("layers", c_int), ("heads", c_int), ("out_dim", c_int),
                ("ctx_len", c_int), ("vocab", c_int),
                ("tok_emb", P), ("pos", P), ("ln_final_g", P), ("ln_final_b", P), ("proj_t", P),
                ("proj", P), ("blocks", POINTER(BlockWeights))]


_SIGNATURES = {
    "gb_create": (c_int, [POINTER(P), c_int]),
    "gb_destroy": (c_int, [P]),
    "gb_last_error": (c_char_p, [P]),
    "gb_launch_count": (c_uint64, [P]),
    "gb_version": (c_char_p, []),
    "gb_set_sm_limit": (c_int, [P, c_int]),
    "gb_profile_begin": (c_int, [P]),
    "gb_profile_end": (c_int, [P, POINTER(ProfileStats), c_int]),
    "gb_profile_launches": (c_int, [P, POINTER(ProfileLaunch), c_int]),
    "gb_gemm_f16": (c_int, [P, P, c_int, P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int,
                            c_int, c_int, P]),
    "gb_layernorm_f16": (c_int, [P, P, c_int, P, c_int, P, P, P, c_int, c_int, c_int, c_int, P]),
    "gb_l2norm512": (c_int, [P, P, P, P, c_int, P]),
    "gb_checksum128": (c_int, [P, P, c_int, c_size_t, P, P]),
    "gb_resize_tmp_bytes": (c_size_t, [c_int, c_int]),
    "gb_resize_bicubic_crop_u8": (c_int, [P, P, P, c_int, c_int, c_int, P, P, c_int, c_int, P, P, c_int, c_int, c_int,
                                          c_int, P, P, P, P]),
    "gb_attention_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "gb_attention_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "gb_vit_set_weights": (c_int, [P, POINTER(VitWeights)]),
    "gb_text_set_weights": (c_int, [P, POINTER(TextWeights)]),
    "gb_tape_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gb_vit_forward": (c_int, [P, P, c_int, P, c_int, c_int, P, P, P, P]),
    "gb_vit_backward_prefix": (c_int, [P, P, P, c_int, c_int, P, P, P]),
    "gb_text_forward": (c_int, [P, P, c_int, P, P, c_int, c_int, c_int, P, P, P, P]),
    "gb_text_backward_prefix": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P]),
    "gb_sim_softmax_argmax": (c_int, [P, P, P, c_float, c_int, c_int, c_int, P, P, P, P]),
    "gb_leaderboard_state_bytes": (c_size_t, [c_int, c_int]),
    "gb_leaderboard_init": (c_int, [P, P, c_int, c_int, P]),
    "gb_leaderboard_update": (c_int, [P, P, c_int, c_int, P, P, P, c_int, c_int, c_int, c_int, P]),
    "gb_leaderboard_export": (c_int, [P, P, c_int, c_int, P, P, P, P]),
    "gb_pseudolabel_scan": (c_int, [P, P, P, P, c_float, c_int, c_int, c_int, c_int, c_int, P, P,
                                    P, P, P]),
    "gb_workspace_generation": (c_uint64, [P]),
    "gb_ce_text_grad": (c_int, [P, P, P, P, P, c_float, c_int, c_int, P, P, P, P]),
    "gb_ce_image_grad": (c_int, [P, P, P, P, P, c_float, c_int, c_int, P, P, P, P, P]),
    "gb_sgd_step": (c_int, [P, P, P, P, ctypes.c_longlong, c_float, P, c_float, c_float, c_int, P]),
    "gb_warmup_cosine_lr": (ctypes.c_double, [ctypes.c_double, c_int, c_int, c_int]),
}

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
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Raw device/host address of a torch tensor (None → NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """The current torch stream of `device` (an Engine / Leaderboard passes its own device: the caller's
    current device may be another GPU)."""
    import torch

    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


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
                "(B200) GPU; there is no CPU fallback")
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

    def set_sm_limit(self, sms: int):
        """Persistent GEMM grids use at most `sms` SMs (0 = all)."""
        self.check(self.lib.gb_set_sm_limit(self.h, int(sms)), "gb_set_sm_limit")

    def profile_begin(self):
        self.check(self.lib.gb_profile_begin(self.h), "gb_profile_begin")

    def profile_launches(self, cap=65536):
        """Per-launch records (kind, m, n, k, ms, work) since profile_begin; call before profile_end."""
        arr = (ProfileLaunch * cap)()
        n = self.lib.gb_profile_launches(self.h, arr, cap)
        if n < 0:
            self.check(n, "gb_profile_launches")
        return [(a.kind, a.m, a.n, a.k, a.ms, a.work) for a in arr[:min(n, cap)]]

    def profile_end(self):
        """[(launches, ms, work)] for kind 0 (GEMM, FLOPs) and kind 1 (sim kernel, bytes)."""
        arr = (ProfileStats * 2)()
        self.check(self.lib.gb_profile_end(self.h, arr, 2), "gb_profile_end")
        return [(int(a.launches), float(a.ms), float(a.work)) for a in arr]

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
            int(bool(out_f32)), stream_ptr(self.device))
        self.check(rc, "gb_gemm_f16")
        return out

    def layernorm(self, x, gamma, beta, row_idx=None, in_row_mul=1, rows=None, out_f32=False):
        import torch

        D = x.shape[-1]
        rows = (x.shape[0] if row_idx is None else row_idx.numel()) if rows is None else rows
        y = torch.empty(rows, D, device=x.device, dtype=torch.float32 if out_f32 else torch.float16)
        rc = self.lib.gb_layernorm_f16(self.h, ptr(x), x.stride(0), ptr(row_idx), in_row_mul,
                                       ptr(gamma), ptr(beta), ptr(y), D, rows, D, int(out_f32),
                                       stream_ptr(self.device))
        self.check(rc, "gb_layernorm_f16")
        return y

    def l2norm512(self, x, want16=True, want32=False):
        import torch

        rows = x.shape[0]
        y16 = torch.empty(rows, 512, device=x.device, dtype=torch.float16) if want16 else None
        y32 = torch.empty(rows, 512, device=x.device, dtype=torch.float32) if want32 else None
        self.check(self.lib.gb_l2norm512(self.h, ptr(x), ptr(y16), ptr(y32), rows, stream_ptr(self.device)),
                   "gb_l2norm512")
        return y16, y32

    def attention_fwd(self, qkv, B, L, D, causal):
        import torch

        out = torch.empty(B * L, D, device=qkv.device, dtype=torch.float16)
        self.check(self.lib.gb_attention_fwd(self.h, ptr(qkv), ptr(out), B, L, D, int(causal),
                                             stream_ptr(self.device)), "gb_attention_fwd")
        return out

    def attention_bwd(self, qkv, dout, B, L, D, causal):
        import torch

        dqkv = torch.empty_like(qkv)
        self.check(self.lib.gb_attention_bwd(self.h, ptr(qkv), ptr(dout), ptr(dqkv), B, L, D,
                                             int(causal), stream_ptr(self.device)), "gb_attention_bwd")
        return dqkv
