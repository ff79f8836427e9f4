"""Per-launch picture of one fused pool scan (run under `ncu --metrics gpu__time_duration.sum`):
    python tools/gpu_scan_diag.py [N] [C] [k]"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
eng = importlib.import_module("menghini-neurips23-code_b200.engine")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
C = int(sys.argv[2]) if len(sys.argv) > 2 else 100
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
F = torch.nn.functional.normalize(torch.randn(N, 512, device=dev, generator=g), dim=1).half()
T = torch.nn.functional.normalize(torch.randn(C, 512, device=dev, generator=g), dim=1).half()
rk = torch.randperm(N, generator=torch.Generator().manual_seed(9)).to(torch.int32).to(dev)
for i in range(3):
    lb = eng.Leaderboard(C, k, dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    lb.scan(F, T, 100.0, rank=rk)
    t1.record()
    torch.cuda.synchronize()
    hdr = lb.state[:32].view(torch.int32).cpu().tolist()
    print(f"scan N={N} C={C} k={k}: {t0.elapsed_time(t1):.3f} ms   events {hdr[3]} waits {hdr[4]} flagged {hdr[5]} "
          f"spill-admits {hdr[6]} replay-kclocks {hdr[7]}")
