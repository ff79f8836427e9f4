"""CoOp step latency on cached image features: the reference-shaped autograd loop vs training.CoOpStep
(SURVEY §8f N1).  usage: python tools/gpu_coop_step.py [batch] [classes]"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
clip = importlib.import_module(PKG + ".clip")
models = importlib.import_module(PKG + ".models")
training = importlib.import_module(PKG + ".training")
synthetic = importlib.import_module(PKG + ".synthetic")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model, _ = clip.load("ViT-B/32", "cuda:0", state_dict=synthetic.synthetic_state_dict(1234))
classes = [f"class number {j}" for j in range(C)]
cte = models.CustomTextEncoder(model, "cuda:0", torch.float16)
imfn16 = torch.nn.functional.normalize(torch.randn(B, 512, device="cuda"), dim=1).half()
labels = (torch.arange(B) % C).cuda()
scale = model.logit_scale.exp().float()


def timeit(fn, n=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


a = models.TextPrefixModel((0.02 * torch.randn(1, 16, 512)).cuda(), cte, classes, device="cuda:0")
opt = torch.optim.SGD([a.prefix], lr=1e-4, momentum=0.9, weight_decay=0.1)


def loop_step():
    tf = a(classes)
    tfn = tf / tf.norm(dim=-1, keepdim=True)
    loss = torch.nn.functional.cross_entropy(scale * imfn16.float() @ tfn.t(), labels)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


b = models.TextPrefixModel((0.02 * torch.randn(1, 16, 512)).cuda(), cte, classes, device="cuda:0")
fused = training.CoOpStep(b, lr=1e-4, weight_decay=0.1, momentum=0.9, warmup_epochs=5, epochs=150)
l0 = model.engine.ctx.launches
loop_step()
l1 = model.engine.ctx.launches
fused.step(imfn16, labels)
l2 = model.engine.ctx.launches
ms_loop = timeit(loop_step)
ms_fused = timeit(lambda: fused.step(imfn16, labels))
print(f"CoOp step on cached features, B={B} C={C} P=16: autograd loop {ms_loop:.3f} ms/step ({l1 - l0} library launches "
      f"+ torch glue), CoOpStep {ms_fused:.3f} ms/step ({l2 - l1} launches, no torch kernels)")
