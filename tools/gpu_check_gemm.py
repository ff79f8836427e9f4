"""Bring-up check for the tcgen05 GEMM on a real B200 (run under gpurun).

usage: python tools/gpu_check_gemm.py <case>      case ∈ small | shapes | epi | perf
"""
import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "gb_lib", os.path.join(ROOT, "menghini-neurips23-code_b200", "_lib.py"))
gb_lib = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gb_lib)


def ref_gemm(A, W, bias=None, resid=None, act=0):
    y = A.float() @ W.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    if resid is not None:
        y = y + resid.float()
    return y


def check(ctx, M, N, K, bias=False, resid=False, act=0, out_f32=False, inplace=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).half()
    b = torch.randn(N, device="cuda", generator=g) * 0.1 if bias else None
    r = torch.randn(M, N, device="cuda", generator=g).half() if resid else None
    want = ref_gemm(A, W, b, r, act)
    if inplace:
        out = r.clone()
        ctx.gemm(A, W, b, out, out=out, act=act)
    else:
        out = ctx.gemm(A, W, b, r, act=act, out_f32=out_f32)
    torch.cuda.synchronize()
    err = (out.float() - want).abs().max().item()
    scale = want.abs().max().item()
    tol = 2e-3 * scale if not out_f32 else 1e-4 * scale + 1e-5
    ok = err <= tol
    print(f"  M={M} N={N} K={K} bias={bias} resid={resid} act={act} f32={out_f32} inplace={inplace}: "
          f"max_err={err:.3e} (scale {scale:.2f}) {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "small"
    ctx = gb_lib.Context(0)
    print(gb_lib.load().gb_version().decode(), torch.cuda.get_device_name(0), flush=True)
    ok = True
    if case == "small":
        ok &= check(ctx, 128, 128, 64, out_f32=True)
        ok &= check(ctx, 128, 128, 128, out_f32=True)
        ok &= check(ctx, 128, 128, 768, out_f32=True)
        ok &= check(ctx, 128, 256, 768)
        ok &= check(ctx, 256, 768, 768)
    elif case == "shapes":
        for (M, N, K) in [(50, 768, 768), (77, 512, 512), (800, 2304, 768), (3300, 3072, 768),
                          (3300, 768, 3072), (51200, 2304, 768), (20000, 3072, 768),
                          (7700, 1536, 512), (7700, 2048, 512), (7700, 512, 2048), (1, 512, 768),
                          (130, 128, 64), (19999, 768, 3072)]:
            ok &= check(ctx, M, N, K, out_f32=True)
            ok &= check(ctx, M, N, K)
    elif case == "epi":
        ok &= check(ctx, 1000, 768, 768, bias=True)
        ok &= check(ctx, 1000, 3072, 768, bias=True, act=1)
        ok &= check(ctx, 1000, 768, 3072, bias=True, resid=True)
        ok &= check(ctx, 1000, 768, 3072, bias=True, resid=True, inplace=True)
        ok &= check(ctx, 30000, 768, 3072, bias=True, resid=True, inplace=True)
        ok &= check(ctx, 30000, 3072, 768, bias=True, act=1)
        ok &= check(ctx, 333, 512, 768, out_f32=True)
    elif case == "perf":
        for (M, N, K) in [(51200, 2304, 768), (51200, 768, 768), (51200, 3072, 768),
                          (51200, 768, 3072), (8192, 8192, 8192), (78848, 1536, 512)]:
            A = torch.randn(M, K, device="cuda").half()
            W = torch.randn(N, K, device="cuda").half()
            out = torch.empty(M, N, device="cuda", dtype=torch.float16)
            for _ in range(3):
                ctx.gemm(A, W, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                ctx.gemm(A, W, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            t0 = time.perf_counter()
            for _ in range(3):
                ref = A @ W.t()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                ref = A @ W.t()
            e1.record()
            torch.cuda.synchronize()
            ms_ref = e0.elapsed_time(e1) / n
            fl = 2.0 * M * N * K
            print(f"  perf M={M} N={N} K={K}: ours {ms:.3f} ms = {fl / ms / 1e9:.1f} TF/s | "
                  f"cuBLAS {ms_ref:.3f} ms = {fl / ms_ref / 1e9:.1f} TF/s", flush=True)
    print("RESULT", case, "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
