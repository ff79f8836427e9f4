"""The UPT training step of bench.py's `upt_step` extra (configs[4]: 4+4 coupled prompts, C = 100) a few times, for launch
lists / ncu captures (run under gpurun):  ncu --metrics gpu__time_duration.sum … python tools/gpu_upt_step.py [steps]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
clip = importlib.import_module(PKG + ".clip")
models = importlib.import_module(PKG + ".models")
training = importlib.import_module(PKG + ".training")
synthetic = importlib.import_module(PKG + ".synthetic")
Engine = importlib.import_module(PKG + ".engine").Engine
bench = importlib.import_module("bench")

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
model, _ = clip.load("ViT-B/32", dev, state_dict=synthetic.synthetic_state_dict(1234))
Pt = Pv = 4
C = 100
B = Engine.wave_aligned_batch(1024, L=50 + Pv, sms=148)
gp = torch.Generator().manual_seed(3)
classes = bench.make_classes(C, seed=5)
cie = models.CustomImageEncoder(model.visual)
cte = models.CustomTextEncoder(model, dev, torch.float32)
torch.manual_seed(4)
upt = models.UPTModel((0.02 * torch.randn(1, Pt, 512, generator=gp)).to(dev),
                      ((768 ** -0.5) * torch.randn(1, Pv, 768, generator=gp)).to(dev), None, cie, cte, classes, 128,
                      device=dev, dtype=torch.float32)
opt = torch.optim.SGD(upt.parameters(), lr=1e-4)
ustep = training.UPTStep(upt, opt)
img = torch.randint(0, 256, (B, 3, 224, 224), dtype=torch.uint8).to(dev)
lab = torch.randint(0, C, (B,)).to(dev)
for _ in range(2):
    ustep.step(img, lab)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ustep.step(img, lab)
e1.record()
torch.cuda.synchronize()
print(f"UPT step B={B} C={C}: {e0.elapsed_time(e1) / steps:.3f} ms")
