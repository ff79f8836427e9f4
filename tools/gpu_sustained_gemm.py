"""Sustained (power-capped) GEMM rates, ours vs cuBLAS, with NVML clock/power (run under gpurun).
usage: python tools/gpu_sustained_gemm.py [seconds-per-variant]"""
import importlib
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
pkg = importlib.import_module(PKG)
ctx = pkg.Context.get(0)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5

import pynvml

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def sustained(name, fn, flop):
    samples, stop = [], threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            stop.wait(0.005)

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = threading.Thread(target=sampler, daemon=True)
    t.start()
    t0 = time.perf_counter()
    # first half heats up, second half is measured
    n = 0
    while time.perf_counter() - t0 < secs / 2:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t1 = time.perf_counter()
    while time.perf_counter() - t1 < secs / 2:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    t.join()
    ms = e0.elapsed_time(e1) / n
    body = samples[len(samples) // 2:]
    clk = sorted(s[0] for s in body)[len(body) // 2]
    pw = sorted(s[1] for s in body)[len(body) // 2]
    util = flop / (ms * 1e-3) / (clk * 1e6 * 148 * 8192)
    print(f"{name:46s} {ms * 1e3:8.1f} us {flop / ms / 1e9:7.0f} TF/s  clk {clk} MHz  {pw:5.0f} W  "
          f"tensor util/clk {100 * util:4.1f}%", flush=True)
    time.sleep(0.5)


def ours(M, N, K, bias=False, act=0, resid=False, tag=""):
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda") if bias else None
    r = torch.randn(M, N, device="cuda").half() if resid else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    sustained(f"ours {tag} M={M} N={N} K={K}", lambda: ctx.gemm(A, W, b, r, out=out, act=act), 2.0 * M * N * K)


def cublas(M, N, K, dtype=torch.float16):
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dtype)
    out = torch.empty(M, N, device="cuda", dtype=dtype)
    sustained(f"cuBLAS {str(dtype)[6:]} M={M} N={N} K={K}", lambda: torch.matmul(A, W.t(), out=out), 2.0 * M * N * K)


M = 51200
only_ours = len(sys.argv) > 2 and sys.argv[2] == "ours"
if only_ours:
    def cublas(*a, **k):
        pass
cublas(8192, 8192, 8192, torch.bfloat16)
cublas(8192, 8192, 8192)
for (N, K) in ((3072, 768), (2304, 768), (768, 768), (768, 3072)):
    cublas(M, N, K)
    ours(M, N, K, tag="plain")
ours(M, 2304, 768, bias=True, tag="bias")
ours(M, 3072, 768, bias=True, act=1, tag="bias+gelu")
ours(M, 768, 768, bias=True, resid=True, tag="bias+res")
ours(M, 768, 3072, bias=True, resid=True, tag="bias+res")
ours(8192, 8192, 8192, tag="plain")
