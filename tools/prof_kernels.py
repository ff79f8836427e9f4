"""Kernel-level timing / profiling driver (run under gpurun, optionally under ncu).
usage: python tools/prof_kernels.py [sim|gemm|attn|all]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
pkg = importlib.import_module(PKG)
lib_mod = importlib.import_module(PKG + "._lib")
ctx = pkg.Context.get(0)
ptr, sp = lib_mod.ptr, lib_mod.stream_ptr
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if what in ("sim", "all"):
    for N, C in ((1 << 20, 100), (1 << 20, 10), (1 << 20, 45)):
        F = torch.nn.functional.normalize(torch.randn(N, 512, device="cuda"), dim=1).half()
        T = torch.nn.functional.normalize(torch.randn(C, 512, device="cuda"), dim=1).half()
        pred = torch.empty(N, device="cuda", dtype=torch.int32)
        pp = torch.empty(N, device="cuda", dtype=torch.float32)
        for mode in (0, 1):
            ms = timeit(lambda: ctx.check(ctx.lib.gb_sim_softmax_argmax(ctx.h, ptr(F), ptr(T), 100.0, N, C, mode,
                                                                        ptr(pred), ptr(pp), None, sp()), "sim"))
            print(f"sim N={N} C={C} mode={mode}: {ms*1e3:.1f} us  {N*1032/ms/1e6:.0f} GB/s", flush=True)
        del F, T

if what in ("gemm", "all"):
    M = 51200
    for (N, K, bias, act, resid, name) in ((2304, 768, 1, 0, 0, "qkv"), (768, 768, 1, 0, 1, "o+res"),
                                           (3072, 768, 1, 1, 0, "fc+gelu"), (768, 3072, 1, 0, 1, "proj+res"),
                                           (768, 3072, 0, 0, 0, "plain")):
        A = torch.randn(M, K, device="cuda").half()
        W = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
        b = torch.randn(N, device="cuda") if bias else None
        r = torch.randn(M, N, device="cuda").half() if resid else None
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        ms = timeit(lambda: ctx.gemm(A, W, b, r, out=out, act=act))
        ms_ref = timeit(lambda: torch.matmul(A, W.t()))
        print(f"gemm {name} M={M} N={N} K={K}: {ms*1e3:.1f} us = {2*M*N*K/ms/1e9:.0f} TF/s | cuBLAS plain "
              f"{ms_ref*1e3:.1f} us = {2*M*N*K/ms_ref/1e9:.0f} TF/s", flush=True)

if what in ("attn", "all"):
    for B, L, D, causal in ((1024, 50, 768, 0), (1024, 66, 768, 0)):
        qkv = torch.randn(B * L, 3 * D, device="cuda").half()
        ms = timeit(lambda: ctx.attention_fwd(qkv, B, L, D, causal))
        print(f"attn fwd B={B} L={L}: {ms*1e3:.1f} us  ({B*L*D*2*4/ms/1e6:.0f} GB/s algorithmic)", flush=True)

if what in ("vit", "all"):
    eng_mod = importlib.import_module(PKG + ".engine")
    synthetic = importlib.import_module(PKG + ".synthetic")
    sd = synthetic.synthetic_state_dict(1234)
    img = torch.randn(1024, 3, 224, 224, device="cuda")
    for fold in (False, True, False, True):
        eng = eng_mod.Engine(sd, "cuda:0", fold_ln=fold)
        ms = timeit(lambda: eng.vit_forward(img, None, want_feat=True, want_featn=True), n=10, warm=3)
        ctx.profile_begin()
        eng.vit_forward(img, None, want_feat=True, want_featn=True)
        recs = ctx.profile_launches()
        ctx.profile_end()
        shapes = {}
        for kind, m, n, k, lms, work in recs:
            if kind == 0:
                e = shapes.setdefault((m, n, k), [0, 0.0])
                e[0] += 1; e[1] += lms
        print(f"vit_forward B=1024 fold_ln={fold}: {ms:.3f} ms  ({1024/ms*1e3:.0f} img/s)  " +
              "  ".join(f"{s}: {v[1]/v[0]*1e3:.0f}us" for s, v in sorted(shapes.items())), flush=True)
        del eng

if what in ("attnbwd", "all"):
    for B, L, D, causal in ((512, 66, 768, 0), (1024, 50, 768, 0)):
        qkv = torch.randn(B * L, 3 * D, device="cuda").half()
        dout = torch.randn(B * L, D, device="cuda").half()
        ms = timeit(lambda: ctx.attention_bwd(qkv, dout, B, L, D, causal))
        print(f"attn bwd B={B} L={L}: {ms*1e3:.1f} us", flush=True)
