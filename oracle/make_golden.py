"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN MODULES (read from
/root/reference, never copied) on top of the restated `clip` (oracle/clip_ref.py).

    python oracle/make_golden.py [--reference /root/reference] [--out tests/golden]

Runs only where /root/reference is mounted (the authoring container).  The fixtures it writes are
committed; tests never need the reference again.  What is executed from the reference:
  models/clip_encoders.py   CustomTextEncoder, CustomImageEncoder, TextEncoder, ImageEncoder
  models/prompts_models.py  TextPrefixModel, ImagePrefixModel, UPTModel
  utils/clip_pseudolabels.py compute_pseudo_labels  (on real PNG files, with a stub clip_model that
                            returns prescribed logits — pins the leaderboard state machine)
Inputs are regenerated from seeds by oracle/synth.py, so only outputs (and tiny inputs) are stored.
"""
from __future__ import annotations

import argparse
import importlib.util
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import clip_ref, leaderboard_ref, synth  # noqa: E402


def load_ref_module(ref_root, rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref_root, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_towers(ref_root, out_dir):
    enc = load_ref_module(ref_root, "models/clip_encoders.py", "ref_clip_encoders")
    pm = load_ref_module(ref_root, "models/prompts_models.py", "ref_prompts_models")
    model = clip_ref.build_model(seed=1234)
    out = {}
    with torch.no_grad():
        img = synth.images(2, seed=0)
        # V3 / ImageEncoder: plain encode_image (P = 0)
        out["img_feat_p0"] = enc.ImageEncoder(model)(img).numpy()
        cie = enc.CustomImageEncoder(model.visual)
        for p in (4, 16):
            out[f"img_feat_p{p}"] = cie(img, synth.image_prefix(p)).numpy()
        # [1,P,768] prefix form (UPT passes this shape)
        out["img_feat_p4_3d"] = cie(img, synth.image_prefix(4)[None]).numpy()
        classes = [" ".join(c.split("_")) for c in synth.class_names(5, seed=1)]
        cte = enc.CustomTextEncoder(model, "cpu", torch.float32)
        for p in (4, 16):
            out[f"txt_feat_p{p}"] = cte(synth.text_prefix(p), classes).numpy()
        # T2 / TextEncoder: encode_text on template prompts (utils/clip_pseudolabels.py:24-25 —
        # string CONCAT keeps the literal "{}" of the template in the prompt)
        template = "a photo of a {}"
        prompts = [f"{template}{c}" for c in classes]
        ids = clip_ref.tokenize(prompts)
        out["txt_ids_zeroshot"] = ids.numpy()
        out["txt_feat_zeroshot"] = enc.TextEncoder(model)(ids).numpy()
        li, _ = model(img, ids)
        out["logits_zeroshot"] = li.numpy()
        # M1/M2 wrappers return the un-normalised encoder output
        m1 = pm.TextPrefixModel(synth.text_prefix(16), cte, classes)
        out["m1_out"] = m1(classes).numpy()
        m2 = pm.ImagePrefixModel(synth.image_prefix(16), cie)
        out["m2_out"] = m2(img).numpy()
        # M3 UPT: seeded init of the trainable coupling head
        torch.manual_seed(4)
        upt = pm.UPTModel(synth.text_prefix(4, seed=2), synth.image_prefix(4, seed=3)[None], None,
                          cie, cte, classes, 128, device="cpu", dtype=torch.float32)
        t_out, v_out = upt(img, classes)
        out["upt_text"] = t_out.numpy()
        out["upt_visual"] = v_out.numpy()
        for k, v in upt.state_dict().items():
            if k.startswith(("proj_", "transformer.")):
                out["upt_sd." + k] = v.numpy()
    # CoOp gradient w.r.t. the prefix through the reference modules (autograd, fp32):
    prefix = synth.text_prefix(16).clone().requires_grad_(True)
    feats = cte(prefix, classes)
    feats_n = feats / feats.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        imf = model.encode_image(synth.images(4, seed=0))
        imf = imf / imf.norm(dim=-1, keepdim=True)
    logits = model.logit_scale.exp().detach() * imf @ feats_n.t()
    labels = torch.tensor([0, 1, 2, 3])
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    out["coop_loss"] = loss.detach().numpy()
    out["coop_grad_prefix"] = prefix.grad.numpy()
    # VPT gradient w.r.t. the image prefix
    vp = synth.image_prefix(16).clone().requires_grad_(True)
    vf = cie(synth.images(2, seed=0), vp)
    vf = vf / vf.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        tf = model.encode_text(ids)
        tf = tf / tf.norm(dim=-1, keepdim=True)
    vloss = torch.nn.functional.cross_entropy(model.logit_scale.exp().detach() * vf @ tf.t(),
                                              torch.tensor([1, 3]))
    vloss.backward()
    out["vpt_loss"] = vloss.detach().numpy()
    out["vpt_grad_prefix"] = vp.grad.numpy()
    np.savez_compressed(os.path.join(out_dir, "towers_vitb32_seed1234.npz"), **out)
    print("towers:", {k: v.shape for k, v in out.items() if not k.startswith("upt_sd.")})


class _StubClip:
    """clip_model(img, text) → prescribed logits for the image whose index is encoded in `img`."""

    def __init__(self, logits):
        self.logits = logits

    def __call__(self, img, text):
        i = int(img.reshape(-1)[0].item())
        li = self.logits[i:i + 1]
        return li, li.t()


def _index_transform(img):
    r, g, b = img.getpixel((0, 0))
    return torch.tensor([float(r * 65536 + g * 256 + b)])


class _DS:
    pass


def leaderboard_cases():
    """(name, logits[N,C] fp32, k).  Adversarial cases of SURVEY.md §8c / Appendix A."""
    cases = []
    rng = np.random.RandomState(11)
    # generic random: k small → boards fill, sort, threshold drop, spill
    cases.append(("rand_n200_c5_k4", rng.randn(200, 5).astype(np.float32) * 2, 4))
    cases.append(("rand_n300_c10_k16", rng.randn(300, 10).astype(np.float32) * 3, 16))
    cases.append(("rand_n64_c3_k1", rng.randn(64, 3).astype(np.float32), 1))
    # never full: k ≥ arrivals
    cases.append(("neverfull_n20_c4_k50", rng.randn(20, 4).astype(np.float32), 50))
    # exact ties: logits drawn from a 3-value set → identical prob rows, path ordering decides
    cases.append(("ties_n120_c4_k5", rng.choice([0.0, 1.0, 2.0], size=(120, 4)).astype(np.float32), 5))
    # all images identical → strict '<' never admits after fill
    cases.append(("const_n40_c3_k4", np.tile(np.array([[1.0, 0.5, 0.2]], np.float32), (40, 1)), 4))
    # unsorted-fill rejection + threshold drop: ascending then descending confidence in class 0
    up = np.linspace(0.1, 3.0, 30)
    lg = np.zeros((60, 2), np.float32)
    lg[:30, 0] = up
    lg[30:, 0] = up[::-1]
    cases.append(("ramp_n60_c2_k6", lg, 6))
    # one dominant class: other boards are seeded purely by spill
    dom = rng.randn(150, 6).astype(np.float32)
    dom[:, 2] += 4.0
    cases.append(("dominant_n150_c6_k8", dom, 8))
    # k == 10000000 branch
    cases.append(("all_n50_c7", rng.randn(50, 7).astype(np.float32), leaderboard_ref.ALL_UNLABELED_K))
    return cases


def golden_leaderboard(ref_root, out_dir):
    from PIL import Image

    pl = load_ref_module(ref_root, "utils/clip_pseudolabels.py", "ref_clip_pseudolabels")
    out = {}
    names = []
    for name, logits_np, k in leaderboard_cases():
        n, c = logits_np.shape
        logits = torch.from_numpy(logits_np)
        classnames = [f"class_{j}" for j in range(c)]
        # label ids deliberately not the identity and not monotone in j
        label_to_idx = {cn: (7 * j + 3) % (c + 5) + (100 if j % 2 else 0) for j, cn in enumerate(classnames)}
        rng = np.random.RandomState(123)
        tags = rng.permutation(n)
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for i in range(n):
                p = os.path.join(td, f"img_{tags[i]:05d}.png")
                Image.new("RGB", (1, 1), ((i >> 16) & 255, (i >> 8) & 255, i & 255)).save(p)
                paths.append(p)
            ds = _DS()
            ds.filepaths = list(paths)
            ds.labels = [0] * n
            ds = pl.compute_pseudo_labels(k, "a photo of a {}", ds, classnames, _index_transform,
                                          _StubClip(logits), label_to_idx, "cpu",
                                          os.path.join(td, "cache.pickle"))
            idx_of = {p: i for i, p in enumerate(paths)}
            got_idx = np.array([idx_of[p] for p in ds.filepaths], np.int64)
            got_lab = np.array(ds.labels, np.int64)
        probs = torch.softmax(logits, dim=-1)  # what the reference computes per row (:62)
        rows = torch.stack([torch.softmax(logits[i:i + 1], dim=-1)[0] for i in range(n)])
        assert torch.equal(probs, rows)
        pred = torch.argmax(probs, dim=1).numpy()
        # path tie-break: rank of the path in ascending string order
        order = sorted(range(n), key=lambda i: paths[i])
        rank = np.empty(n, np.int64)
        rank[order] = np.arange(n)
        class_ids = np.array([label_to_idx[cn] for cn in classnames], np.int64)
        # the restatement must reproduce the reference on these inputs
        r_idx, r_lab = leaderboard_ref.leaderboard(probs.numpy(), pred, k, rank, class_ids.tolist())
        assert r_idx == got_idx.tolist() and r_lab == got_lab.tolist(), name
        out[f"{name}.logits"] = logits_np
        out[f"{name}.probs"] = probs.numpy()
        out[f"{name}.pred"] = pred.astype(np.int64)
        out[f"{name}.k"] = np.int64(k)
        out[f"{name}.rank"] = rank
        out[f"{name}.class_ids"] = class_ids
        out[f"{name}.out_idx"] = got_idx
        out[f"{name}.out_lab"] = got_lab
        names.append(name)
        print(f"leaderboard {name}: {len(got_idx)} entries, restatement == reference")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(out_dir, "leaderboard_cases.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    if not os.path.isdir(args.reference):
        sys.exit(f"{args.reference} not present: golden vectors can only be regenerated where the "
                 "reference is mounted")
    os.makedirs(args.out, exist_ok=True)
    clip_ref.install_as_clip()
    torch.set_num_threads(os.cpu_count())
    golden_leaderboard(args.reference, args.out)
    golden_towers(args.reference, args.out)


if __name__ == "__main__":
    main()
