"""Ad-hoc diagnostics on the GPU box (not part of the test suite)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
from oracle import clip_ref, leaderboard_ref, synth  # noqa: E402

clip = importlib.import_module(PKG + ".clip")
eng_mod = importlib.import_module(PKG + ".engine")
model, _ = clip.load("ViT-B/32", "cuda:0", state_dict=clip_ref.synth_state_dict(seed=1234))
eng = model.engine

img = synth.images(37, seed=3).cuda()
a = eng.vit_forward(img, None, want_feat=True, want_featn=True)
b = eng.vit_forward(img, None, want_feat=True, want_featn=True)
print("run-to-run feat equal:", torch.equal(a[0], b[0]), "featn equal:", torch.equal(a[1], b[1]))

# tape comparison: sample 0 of B=37 vs B=1
M37, M1, D = 37 * 50, 50, 768
_, _, t37 = eng.vit_forward(img, None, tape=True)
_, _, t1 = eng.vit_forward(img[:1], None, tape=True)
t37 = t37.view(torch.float16)
t1 = t1.view(torch.float16)
names = ["x0", "qkv", "x1", "f"]
widths = [D, 3 * D, D, 4 * D]
for l in range(12):
    off37 = l * 9 * M37 * D
    off1 = l * 9 * M1 * D
    for n, w in zip(names, widths):
        s37 = t37[off37:off37 + M37 * w].view(M37, w)[:50]
        s1 = t1[off1:off1 + M1 * w].view(M1, w)
        d = (s37.float() - s1.float()).abs().max().item()
        if d != 0:
            print(f"layer {l} {n}: max diff {d:.3e}  nonzero rows {((s37 != s1).any(1)).nonzero().flatten()[:10].tolist()}")
        off37 += M37 * w
        off1 += M1 * w
    if l == 0:
        print("layer 0 checked")
print("tape compare done")

# small-N fused scan vs oracle
for N, C, k in ((40, 4, 3), (100, 4, 3), (130, 10, 2)):
    f, t = synth.pool(N, C, peaked=0.0)
    F, T = f.half().cuda(), t.half().cuda()
    rank_np = synth.path_ranks(N)
    rank = torch.from_numpy(rank_np).to(torch.int32).cuda()
    pred, pp, probs = eng.sim_softmax_argmax(F, T, 100.0, want_probs=True)
    want = leaderboard_ref.leaderboard(probs.cpu().numpy(), pred.cpu().numpy(), k, rank_np)
    lb = eng_mod.Leaderboard(C, k, "cuda:0")
    p2, pp2, _ = lb.scan(F, T, 100.0, rank=rank)
    lb2 = eng_mod.Leaderboard(C, k, "cuda:0")
    p3, pp3, pr3 = lb2.scan(F, T, 100.0, rank=rank, want_probs=True)
    lb3 = eng_mod.Leaderboard(C, k, "cuda:0")
    lb3.update(probs, pred, rank, prefilter=False)
    print(N, C, k, "fused==oracle", lb.result() == want, "fused+probs==oracle", lb2.result() == want,
          "update==oracle", lb3.result() == want, "pred eq", torch.equal(pred, p2), torch.equal(pp, pp2),
          "probs eq", torch.equal(pr3, probs))
    if lb.result() != want:
        print(" want", want)
        print(" got ", lb.result())
