"""Chain steps of the parallel replay: SM clock of every published verdict (GB_LB_DIAG=5, hook gb_debug_lb_ts).
usage: GB_LB_DIAG=5 python tools/gpu_scan_ts.py [rows]"""
import ctypes, importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
eng = importlib.import_module("menghini-neurips23-code_b200.engine")
N, C, k = (int(sys.argv[1]) if len(sys.argv) > 1 else 12288), 100, 16
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
F = torch.nn.functional.normalize(torch.randn(1 << 20, 512, device=dev, generator=g), dim=1).half()[:N].contiguous()
T = torch.nn.functional.normalize(torch.randn(C, 512, device=dev, generator=g), dim=1).half()
rk = torch.randperm(1 << 20, generator=torch.Generator().manual_seed(9)).to(torch.int32).to(dev)
for rep in range(2):
    lb = eng.Leaderboard(C, k, dev)
    lb.scan(F, T, 100.0, rank=rk)
torch.cuda.synchronize()
lib = lb.lib
buf = (ctypes.c_longlong * N)()
lib.gb_debug_lb_ts.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert lib.gb_debug_lb_ts(buf, N) == 0
ts = np.array(buf[:], dtype=np.int64)
for lo, hi in ((0, 1024), (1024, 4096), (4096, 12288)):
    if hi > N:
        break
    t = ts[lo:hi]
    nz = np.nonzero(t)[0]
    d = np.diff(t[nz])
    d = d[(d > 0) & (d < 10_000_000)]
    print(f"rows [{lo},{hi}): {len(nz)} published, span {int(t[nz].max() - t[nz].min())} clocks; step median {np.median(d):.0f} "
          f"mean {d.mean():.0f} p90 {np.percentile(d, 90):.0f} p99 {np.percentile(d, 99):.0f} max {d.max()}; "
          f"steps > 2000: {(d > 2000).sum()} carrying {d[d > 2000].sum()} clocks")
t = ts[:min(N, 4096)]
print("first 96 steps:", np.diff(t[:97]).tolist())
print("rows 2048..2144 steps:", np.diff(t[2048:2145]).tolist())
