"""Sustained-load diagnostic (run under gpurun): the ViT forward at B=1024 for a few seconds with NVML
clock / power sampled every 5 ms, then per-GEMM-shape in-loop timings from the ctx profiler.
usage: python tools/gpu_power_diag.py [seconds] [batch]"""
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
eng_mod = importlib.import_module(PKG + ".engine")
synthetic = importlib.import_module(PKG + ".synthetic")

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024

import pynvml

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []
stop = threading.Event()


def sampler():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        stop.wait(0.005)


ctx = pkg.Context.get(0)
eng = eng_mod.Engine(synthetic.synthetic_state_dict(1234), "cuda:0")
img = torch.randn(B, 3, 224, 224, device="cuda")
for _ in range(3):
    eng.vit_forward(img, None, want_feat=True, want_featn=True)
torch.cuda.synchronize()
t = threading.Thread(target=sampler, daemon=True)
t.start()
t0 = time.perf_counter()
n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.perf_counter() - t0 < secs:
    for _ in range(5):
        eng.vit_forward(img, None, want_feat=True, want_featn=True)
    n += 5
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
# per-shape timings while still hot
ctx.profile_begin()
for _ in range(3):
    eng.vit_forward(img, None, want_feat=True, want_featn=True)
recs = ctx.profile_launches()
ctx.profile_end()
stop.set()
t.join()
ms = e0.elapsed_time(e1) / n
flop = 2 * 49 * 3072 * 768 + 12 * (24 * 50 * 768 ** 2 + 4 * 50 ** 2 * 768) + 2 * 768 * 512
print(f"vit fwd B={B}: {ms:.3f} ms/iter over {n} iters = {B / ms * 1e3:.0f} img/s = "
      f"{B * flop / ms / 1e9:.0f} TFLOP/s")
body = samples[len(samples) // 4:]
clk = sorted(s[1] for s in body)
pw = sorted(s[2] for s in body)
reasons = 0
for s in body:
    reasons |= s[3]
print(f"nvml: {len(body)} samples, sm clock median {clk[len(clk) // 2]} MHz (p10 {clk[len(clk) // 10]}, "
      f"p90 {clk[9 * len(clk) // 10]}), power median {pw[len(pw) // 2]:.0f} W (max {pw[-1]:.0f}), reasons 0x{reasons:x}")
shapes = {}
for kind, m, nn, k, lms, work in recs:
    if kind == 0:
        e = shapes.setdefault((m, nn, k), [0, 0.0, 0.0])
        e[0] += 1; e[1] += lms; e[2] += work
tot = sum(v[1] for v in shapes.values())
for (m, nn, k), (cnt, tms, w) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
    print(f"  gemm M={m} N={nn} K={k}: {cnt} launches, avg {tms / cnt * 1e3:.1f} us, {w / tms / 1e9:.0f} TFLOP/s, "
          f"{100 * tms / tot:.1f}% of GEMM time")
print(f"  all GEMMs: {tot / 3:.3f} ms per forward")
