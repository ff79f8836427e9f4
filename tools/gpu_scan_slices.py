"""Where the fused pool scan's time goes along the pool: the scan is fed the pool in the geometric slices
gb_pseudolabel_scan itself uses (4096, 8192, … 262144 rows), one timed call per slice (rank[] is global, so the
boards evolve exactly as in the one-call scan), with the replay kernel's counters (GB_LB_DIAG=1: events / waits /
flagged / spill admissions; GB_LB_DIAG=2: kilo-clocks of the phases) read from the state header after each.
usage: [GB_LB_DIAG=1|2] python tools/gpu_scan_slices.py [N] [C] [k]"""
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
for rep in range(2):
    lb = eng.Leaderboard(C, k, dev)
    r0, chunk, rows = 0, 4096, []
    prev = [0] * 8
    while r0 < N:
        r1 = min(r0 + chunk, N)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        lb.scan(F[r0:r1], T, 100.0, idx0=r0, rank=rk)
        t1.record()
        torch.cuda.synchronize()
        hdr = lb.state[:32].view(torch.int32).cpu().tolist()
        rows.append((r0, r1, t0.elapsed_time(t1), [hdr[i] - prev[i] for i in range(3, 8)]))
        prev = hdr[:8]
        r0 = r1
        if chunk < (1 << 18):
            chunk *= 2
    if rep:
        for r0, r1, ms, d in rows:
            print(f"rows [{r0:8d},{r1:8d}): {ms * 1e3:8.1f} us   hdr3..7 deltas {d}")
        print(f"sum of slices {sum(r[2] for r in rows):.3f} ms")
lb = eng.Leaderboard(C, k, dev)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
lb.scan(F, T, 100.0, rank=rk)
t1.record()
torch.cuda.synchronize()
print(f"one call: {t0.elapsed_time(t1):.3f} ms")
