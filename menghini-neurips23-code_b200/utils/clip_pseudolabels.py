"""B200 pool scan behind the reference's `utils.pseudolabel_top_k` / `compute_pseudo_labels`
(same names, arguments, cache-file name and pickle schema as utils/clip_pseudolabels.py of the
reference; citations are to that file).

What changes is the execution, not the result: the reference runs `clip_model(img, text)` once per
image (batch 1, text tower re-encoded every time, one device sync per Python comparison, :55-101);
here the prompts are encoded once, images are encoded in batches, and similarity → softmax → arg-max →
per-class leaderboard run as one fused device pass (gb_pseudolabel_scan) that reproduces the
reference's sequential leaderboard decisions exactly.
"""
from __future__ import annotations

import logging
import os
import pickle

import torch

from .. import clip as _clip
from ..engine import Leaderboard

log = logging.getLogger(__name__)

ALL_UNLABELED_K = 10000000  # :27
ENCODE_BATCH = None  # None: Engine.wave_aligned_batch — the largest chunk ≤ 2048 that fills whole GEMM waves


def path_ranks(paths):
    """rank[i] = position of paths[i] among the sorted path strings (ties keep arrival order): the
    integer image of the `(prob, path)` tuple ordering Python's sorted() uses at :78-82."""
    order = sorted(range(len(paths)), key=lambda i: paths[i])
    rank = [0] * len(paths)
    r = -1
    prev = None
    for i in order:
        if paths[i] != prev:
            r += 1
            prev = paths[i]
        rank[i] = r
    return torch.tensor(rank, dtype=torch.int32)


def encode_pool(clip_model, filepaths, transform, device, batch=ENCODE_BATCH, loader=None, prefix=None):
    """Unit-norm fp16 image features [N,512] of the whole pool, encoded in batches.

    The reference decodes and encodes one image at a time (:55-61).  Here a background thread decodes chunk
    i+1 (PIL releases the GIL) into pinned memory while the device encodes chunk i; when `transform` is this
    package's CLIP transform its `raw_u8` twin is used — resize + centre crop on the host, ToTensor +
    Normalize on the device, bit-identical features.  `prefix` ([P,768] / [1,P,768]): visual prompt
    rows every image is encoded with (the VPT / UPT strategies' assign_pseudo_labels, visual_fpl.py:262-268)."""
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    eng = clip_model.engine
    if prefix is not None:
        prefix = prefix.detach().reshape(-1, prefix.shape[-1]).float()
    if not batch:
        batch = type(eng).wave_aligned_batch(2048, L=50 + (0 if prefix is None else prefix.shape[0]), sms=148)
    feats = torch.empty(len(filepaths), 512, device=eng.device, dtype=torch.float16)
    tf = getattr(transform, "raw_u8", None) or transform

    def load(chunk):
        if loader is not None:
            imgs = loader(chunk)
        else:
            imgs = torch.stack([tf(Image.open(p).convert("RGB")) for p in chunk])
        return imgs.pin_memory() if imgs.device.type == "cpu" else imgs

    starts = list(range(0, len(filepaths), batch))
    with ThreadPoolExecutor(max_workers=1) as pool:
        fut = pool.submit(load, filepaths[starts[0]:starts[0] + batch]) if starts else None
        for i, s in enumerate(starts):
            imgs = fut.result()
            if i + 1 < len(starts):
                fut = pool.submit(load, filepaths[starts[i + 1]:starts[i + 1] + batch])
            imgs = imgs.to(eng.device, non_blocking=True)
            _, fn, _ = eng.vit_forward(imgs, prefix, want_feat=False, want_featn=True)
            feats[s:s + imgs.shape[0]] = fn
    return feats


def scan_features(engine, feats16, protos16, k, paths, class_ids, mode=0, scale=None):
    """Leaderboard over precomputed unit-norm features.  Returns (image indices, labels) in the
    reference's output order (:103-109)."""
    n, c = feats16.shape[0], protos16.shape[0]
    scale = engine.logit_scale_exp if scale is None else scale
    if k == ALL_UNLABELED_K:  # :27-44 — label every image with its arg-max
        pred, _, _ = engine.sim_softmax_argmax(feats16, protos16, scale, mode)
        pred = pred.cpu().tolist()
        return list(range(n)), [class_ids[j] for j in pred]
    rank = path_ranks(paths).to(engine.device)
    board = Leaderboard(c, k, engine.device)
    board.scan(feats16, protos16, scale, mode=mode, idx0=0, rank=rank)
    return board.result(class_ids)


def compute_pseudo_labels(k, template, dataset, classnames, transform, clip_model, label_to_idx,
                          device, filename, loader=None):
    # :24 — string CONCAT, not .format(): the literal "{}" of the template stays in the prompt
    prompts = [f"{template}{' '.join(i.split('_'))}" for i in classnames]
    text = _clip.tokenize(prompts)
    class_ids = [label_to_idx[cn] for cn in classnames]
    if len(set(class_ids)) != len(class_ids):
        raise ValueError("label_to_idx maps two class names to one id")
    eng = clip_model.engine
    with torch.no_grad():
        _, protos, _ = eng.text_forward(text, None, want_feat=False, want_featn=True)
        feats = encode_pool(clip_model, dataset.filepaths, transform, device, loader=loader)
        log.info("Compute %s pseudo-labeles", "all" if k == ALL_UNLABELED_K else k)
        idx, labels = scan_features(eng, feats, protos, k, dataset.filepaths, class_ids, mode=0)
    new_imgs = [dataset.filepaths[i] for i in idx]
    dataset.filepaths = new_imgs
    dataset.labels = labels
    with open(filename, "wb") as f:  # :114-115
        pickle.dump({"filepaths": new_imgs, "labels": labels}, f)
    return dataset


def pseudolabel_top_k(config, data_name, k, template, dataset, classnames, transform, clip_model,
                      label_to_idx, device, vis_encoder, split_seed):
    filename = (f"pseudolabels/{data_name}_{vis_encoder.replace('/', '')}_{config.LEARNING_PARADIGM}_"
                f"{config.MODEL}_{k}_pseudolabels_split_{split_seed}.pickle")  # :134
    if os.path.exists(filename):
        with open(filename, "rb") as f:
            pseudolabels = pickle.load(f)
        dataset.filepaths = pseudolabels["filepaths"]
        dataset.labels = pseudolabels["labels"]
        return dataset
    return compute_pseudo_labels(k, template, dataset, classnames, transform, clip_model,
                                 label_to_idx, device, filename)
