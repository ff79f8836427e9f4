"""Stall breakdown of the tcgen05 attention backward (needs a -DGB_ATTN_STALLS build named by GRIPB200_LIB; run under gpurun):
    make -C menghini-neurips23-code_b200/csrc dbg && GRIPB200_LIB=$PWD/menghini-neurips23-code_b200/libgripb200_dbg.so python tools/gpu_attn_stalls.py"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("menghini-neurips23-code_b200")
ctx = pkg.Context.get(0)
lib = ctx.lib
lib.gb_debug_attn_stalls.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
lib.gb_debug_attn_stalls.restype = ctypes.c_int
names = ["mma:operands", "mma:sdp_free", "mma:pds_ready", "mma:out_free", "mma:total", "sm:sdp_full", "sm:pds_free", "sm:total",
         "out:out_full", "out:store_read", "out:total", "prod:empty", "items", "sm:tmem_ld", "sm:math", "sm:store+fence"]
for B, L, D, causal in ((1024, 50, 768, 0), (512, 66, 768, 0), (861, 66, 768, 0)):
    qkv = torch.randn(B * L, 3 * D, device="cuda").half()
    dout = torch.randn(B * L, D, device="cuda").half()
    for _ in range(3):
        ctx.attention_bwd(qkv, dout, B, L, D, causal)
    buf = (ctypes.c_ulonglong * 16)()
    assert lib.gb_debug_attn_stalls(buf, 1) == 0
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ctx.attention_bwd(qkv, dout, B, L, D, causal)
    e1.record()
    torch.cuda.synchronize()
    lib.gb_debug_attn_stalls(buf, 1)
    v = list(buf)
    items = max(v[12], 1)
    print(f"attn bwd B={B} L={L}: {e0.elapsed_time(e1) / n * 1e3:.1f} us; clocks per item: " +
          "  ".join(f"{nm} {v[i] / items:.0f}" for i, nm in enumerate(names) if i != 12))
