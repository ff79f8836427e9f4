"""ViT forward at B=1024 a few times, for ncu captures of the tower's kernels in context (run under gpurun).
usage: python tools/prof_vit.py [iters] [batch]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
eng_mod = importlib.import_module(PKG + ".engine")
synthetic = importlib.import_module(PKG + ".synthetic")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
eng = eng_mod.Engine(synthetic.synthetic_state_dict(1234), "cuda:0")
img = torch.randint(0, 256, (B, 3, 224, 224), dtype=torch.uint8, device="cuda")
for _ in range(iters):
    eng.vit_forward(img, None, want_feat=True, want_featn=True)
torch.cuda.synchronize()
print("done")
