"""One GEMM configuration in a loop, for ncu captures (run under gpurun).
usage: python tools/prof_one_gemm.py M N K bias act resid [iters]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("menghini-neurips23-code_b200")
ctx = pkg.Context.get(0)
M, N, K, bias, act, resid = (int(x) for x in sys.argv[1:7])
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 8
A = (torch.randn(M, K, device="cuda") * 0.5).half()
W = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
b = torch.randn(N, device="cuda") if bias else None
r = torch.randn(M, N, device="cuda").half() if resid else None
out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(iters):
    ctx.gemm(A, W, b, r, out=out, act=act)
torch.cuda.synchronize()
print("done")
