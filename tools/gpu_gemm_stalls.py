"""Stall breakdown of the CTA-pair GEMM (needs a -DGB_GEMM_STALLS build; run under gpurun).
usage: python tools/gpu_gemm_stalls.py"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
pkg = importlib.import_module(PKG)
ctx = pkg.Context.get(0)
lib = ctx.lib
lib.gb_debug_gemm_stalls.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
lib.gb_debug_gemm_stalls.restype = ctypes.c_int


def run(M, N, K, bias=False, act=0, resid=False, tag="", n=20):
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda") if bias else None
    r = torch.randn(M, N, device="cuda").half() if resid else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        ctx.gemm(A, W, b, r, out=out, act=act)
    buf = (ctypes.c_ulonglong * 16)()
    assert lib.gb_debug_gemm_stalls(buf, 1) == 0, "not a GB_GEMM_STALLS build"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ctx.gemm(A, W, b, r, out=out, act=act)
    e1.record()
    torch.cuda.synchronize()
    lib.gb_debug_gemm_stalls(buf, 1)
    v = list(buf)
    us = e0.elapsed_time(e1) / n * 1e3
    tot = max(v[2], 1)
    print(f"{tag:26s} M={M} N={N} K={K}: {us:7.1f} us {2.0 * M * N * K / us / 1e6:6.0f} TF/s | MMA thread: "
          f"operand wait {100 * v[0] / tot:4.1f}%  accumulator wait {100 * v[1] / tot:4.1f}%  "
          f"| epilogue warp 4: idle {100 * v[4] / max(v[5], 1):4.1f}%", flush=True)


M = 51200
if len(sys.argv) > 1 and sys.argv[1] == "iso":
    for K in (64,):
        run(9472, 3072, K, tag=f"iso plain K={K}")          # 37 row pairs x 12 = 444 tiles = 6 per cluster
        run(9472, 3072, K, bias=True, act=1, tag=f"iso gelu K={K}")
        run(9472, 768, K, bias=True, resid=True, tag=f"iso res K={K}")
    run(M, 3072, 768, tag="fc plain")
    run(M, 3072, 768, bias=True, act=1, tag="fc bias+gelu")
    sys.exit(0)
run(M, 3072, 768, tag="fc plain")
run(M, 3072, 768, bias=True, act=1, tag="fc bias+gelu")
run(M, 2304, 768, tag="qkv plain")
run(M, 2304, 768, bias=True, tag="qkv bias")
run(M, 768, 768, tag="o plain")
run(M, 768, 768, bias=True, resid=True, tag="o bias+res")
run(M, 768, 3072, tag="cproj plain")
run(M, 768, 3072, bias=True, resid=True, tag="cproj bias+res")
run(8192, 8192, 8192, tag="square", n=5)
# epilogue in (near) isolation: one or two k-blocks per tile, output small enough to stay in L2
for K in (64, 128, 768):
    run(9472, 3072, K, tag=f"iso plain K={K}")          # 37 row pairs x 12 = 444 tiles = 6 per cluster
    run(9472, 3072, K, bias=True, act=1, tag=f"iso gelu K={K}")
    run(9472, 768, K, bias=True, resid=True, tag=f"iso res K={K}")
