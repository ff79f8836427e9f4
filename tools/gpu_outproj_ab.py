"""What the residual operand costs the out-proj GEMM (M = 94 700 rows = 1894 images x 50 tokens, N = K = 768):
plain / bias / bias + residual (separate buffer, in place), next to c_fc for the box's calibration.
usage: python tools/gpu_outproj_ab.py [M]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("menghini-neurips23-code_b200")
ctx = pkg.Context.get(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 94700


def run(N, K, bias=False, act=0, resid=None, tag="", n=40):
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda") if bias else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    r = None
    if resid == "sep":
        r = torch.randn(M, N, device="cuda").half()
    elif resid == "inplace":
        out.normal_()
        r = out
    for _ in range(5):
        ctx.gemm(A, W, b, r, out=out, act=act)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        ctx.gemm(A, W, b, r, out=out, act=act)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{tag:24s} M={M} N={N} K={K}: {us:7.1f} us {2.0 * M * N * K / us / 1e6:6.0f} TF/s", flush=True)


for rep in range(2):
    run(768, 768, tag="o plain")
    run(768, 768, bias=True, tag="o bias")
    run(768, 768, bias=True, resid="sep", tag="o bias+res")
    run(768, 768, bias=True, resid="inplace", tag="o bias+res in place")
    run(3072, 768, tag="fc plain")
    run(3072, 768, bias=True, act=1, tag="fc bias+gelu")
    run(768, 3072, tag="cproj plain")
    run(768, 3072, bias=True, resid="sep", tag="cproj bias+res")
