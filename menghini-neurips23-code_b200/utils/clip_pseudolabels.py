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


# Features of a pool under the FROZEN image tower (no visual prompt rows) depend on nothing but the weights, the
# transform and the files: GRIP's textual strategies call assign_pseudo_labels once per iteration on the same pool
# (methods/semi_supervised_learning/textual_fpl.py:168-191, STEP_QUANTILE 10 → ten passes), and the reference decodes
# and encodes every image again each time.  The cache key carries (path, mtime, size) of every file, so a pool whose
# files changed is encoded again; GRIPB200_POOL_CACHE=0 switches it off.
# The cache lives on the engine (it dies with the weights it belongs to) and keeps a reference to the transform it was
# filled under.
_POOL_CACHE_MAX = 4


def _pool_cache_key(filepaths):
    if os.environ.get("GRIPB200_POOL_CACHE") == "0":
        return None
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    try:
        for pth in filepaths:
            st = os.stat(pth)
            h.update(f"{pth}\0{st.st_mtime_ns}\0{st.st_size}\n".encode())
    except OSError:
        return None
    return (len(filepaths), h.hexdigest())


def encode_pool(clip_model, filepaths, transform, device, batch=ENCODE_BATCH, loader=None, prefix=None,
                workers=None, device_resize=None):
    """Unit-norm fp16 image features [N,512] of the whole pool, encoded in batches.

    The reference decodes and encodes one image at a time on one thread (:55-61; data/dataset.py:64-79 even applies
    the transform three times per sample).  Here `workers` host threads (default: every core; PIL's decoders and
    resampling release the GIL) decode + resize + crop the images of chunk i+1 straight into a pinned staging
    buffer while the device encodes chunk i; two staging buffers alternate, the H2D copy is asynchronous.  When
    `transform` is this package's CLIP transform its `raw_u8` twin is used — resize + centre crop on the host,
    ToTensor + Normalize on the device: a quarter of the staging and PCIe bytes, bit-identical features
    (SURVEY §8f N2).  With that transform the resize and the crop move to the device as well (`device_resize`, default
    on; $GRIPB200_DEVICE_RESIZE=0 keeps them on the host): Pillow's bicubic resampler is integer arithmetic on 22-bit
    fixed-point weights, restated bit for bit in utils/pil_resample.py + csrc/resize.cu, so the host only DECODES
    (≈2 ms of the ≈6 ms a 512×384 JPEG costs per core) and the image tower still sees exactly the reference's pixels;
    on this route the decoders are `workers` forked PROCESSES writing into shared page-locked arenas (threads share the
    GIL for everything PIL does in Python and stop scaling at ≈3 cores; $GRIPB200_DECODE_PROCESSES=0 keeps threads).
    `prefix` ([P,768] / [1,P,768]): visual prompt rows every image is encoded with (the VPT / UPT
    strategies' assign_pseudo_labels, visual_fpl.py:262-268).  `loader(chunk) -> tensor` replaces decoding."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    eng = clip_model.engine
    if prefix is not None:
        prefix = prefix.detach().reshape(-1, prefix.shape[-1]).float()
    if not batch:
        batch = type(eng).wave_aligned_batch(2048, L=50 + (0 if prefix is None else prefix.shape[0]), sms=148)
    n = len(filepaths)
    key = _pool_cache_key(filepaths) if (prefix is None and loader is None and n > 0) else None
    if key is not None:
        key = key + (id(transform),)   # the entry keeps the transform alive, so the id cannot be recycled
    cache = eng.__dict__.setdefault("_pool_cache", {}) if key is not None else None
    if key is not None and key in cache and cache[key][0] is transform:
        log.info(f"[encode_pool] {n} pool features from the cache (frozen image tower, unchanged files)")
        return cache[key][1]
    feats = torch.empty(n, 512, device=eng.device, dtype=torch.float16)
    if n == 0:
        return feats
    tf = getattr(transform, "raw_u8", None) or transform
    if not workers:      # this rank's share of the cores this process may run on
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        workers = cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    workers = max(1, min(int(workers), 64))
    if device_resize is None:
        device_resize = os.environ.get("GRIPB200_DEVICE_RESIZE", "1") != "0"
    # only this package's own CLIP transform is known to be "Pillow bicubic resize → centre crop" and nothing else
    device_resize = bool(device_resize) and loader is None and getattr(tf, "is_clip_preprocess_u8", False)
    resizer = None
    if device_resize:
        import numpy as np

        from .pil_resample import get_resizer
        # forked decoders pay off from a few hundred files on; smaller pools are decoded by threads
        procs = workers if (os.environ.get("GRIPB200_DECODE_PROCESSES", "1") != "0" and n >= 256) else 0
        resizer = get_resizer(eng, procs)

    def decode(path):
        return tf(Image.open(path).convert("RGB"))

    if device_resize:       # chunks sized for the pinned arenas (full-size decoded pixels), not for GEMM waves: this
        batch = min(batch, 512)   # route is bound by the host's JPEG decoding either way

    starts = list(range(0, n, batch))
    staging = [None, None]       # pinned, allocated once the first decoded image tells shape and dtype
    consumed = [None, None]      # event: the device has finished reading staging[b]

    with ThreadPoolExecutor(max_workers=workers) as dec_pool, ThreadPoolExecutor(max_workers=1) as chunk_pool:

        def load(ci):
            chunk = filepaths[starts[ci]:starts[ci] + batch]
            if loader is not None:
                imgs = loader(chunk)
                return imgs.pin_memory() if imgs.device.type == "cpu" and not imgs.is_pinned() else imgs
            if device_resize:            # decoded RGB pixels of whatever size → pinned arena; resize + crop on the device
                slot = ci % 2
                resizer.begin(slot)
                if resizer.processes:    # worker processes decode straight into the shared, page-locked arena
                    return resizer.stage_paths(slot, chunk)
                return list(dec_pool.map(
                    lambda pth: resizer.put(slot, np.array(Image.open(pth).convert("RGB"), dtype=np.uint8)), chunk))
            b = ci % 2
            first = decode(chunk[0])
            if staging[b] is None or staging[b].dtype != first.dtype or staging[b].shape[1:] != first.shape:
                staging[b] = torch.empty((batch,) + tuple(first.shape), dtype=first.dtype).pin_memory()
            if consumed[b] is not None:
                consumed[b].synchronize()
            buf = staging[b]
            buf[0].copy_(first)

            def put(i):
                buf[i].copy_(decode(chunk[i]))

            list(dec_pool.map(put, range(1, len(chunk))))
            return buf[:len(chunk)]

        fut = chunk_pool.submit(load, 0)
        for i, s in enumerate(starts):
            imgs = fut.result()
            if i + 1 < len(starts):
                fut = chunk_pool.submit(load, i + 1)
            if device_resize:
                dev = resizer.flush(i % 2, imgs, torch.empty(len(imgs), 3, 224, 224, dtype=torch.uint8, device=eng.device))
            else:
                dev = imgs.to(eng.device, non_blocking=True)
                if loader is None:
                    consumed[i % 2] = torch.cuda.Event()
                    consumed[i % 2].record(torch.cuda.current_stream(eng.device))
            _, fn, _ = eng.vit_forward(dev, prefix, want_feat=False, want_featn=True)
            feats[s:s + dev.shape[0]] = fn
    if key is not None:
        while len(cache) >= _POOL_CACHE_MAX:
            cache.pop(next(iter(cache)))
        cache[key] = (transform, feats)
    return feats


def prob_dtype():
    """"fp32" (default) or "fp16" ($GRIPB200_PROB_DTYPE).  The reference on CUDA holds `clip_model` in fp16, so what its
    loop compares are fp16 probabilities of fp16 logits (`logit_scale.exp() * image_features @ text_features.t()` with
    fp16 operands rounds the logits to fp16 — 1/16 apart near 100 — and `.softmax()` returns fp16; :59-62): they saturate
    to 1.0 and tie, and ties are resolved by the strict `<` and the path order.  "fp16" reproduces THAT arithmetic — the
    same two torch operations on the fp16 unit features — and replays the leaderboard exactly on those numbers; the
    default keeps the fused kernel's fp32 soft-max, which is what the reference's CPU path (fp32 model) computes."""
    v = os.environ.get("GRIPB200_PROB_DTYPE", "fp32").lower()
    if v not in ("fp32", "fp16"):
        raise ValueError(f"GRIPB200_PROB_DTYPE must be fp32 or fp16, not {v!r}")
    return v


def scan_features(engine, feats16, protos16, k, paths, class_ids, mode=0, scale=None, probs="auto"):
    """Leaderboard over precomputed unit-norm features.  Returns (image indices, labels) in the
    reference's output order (:103-109).  `probs`: "fp32" | "fp16" | "auto" (= prob_dtype())."""
    n, c = feats16.shape[0], protos16.shape[0]
    scale = engine.logit_scale_exp if scale is None else scale
    if probs == "auto":
        probs = prob_dtype()
    if probs == "fp16":
        # the reference's CUDA arithmetic, operation for operation (third-party CLIP.forward + :62-64 / textual_fpl.py:228)
        logits = (torch.tensor(scale, device=feats16.device) * feats16) @ protos16.t()      # fp16
        p16 = logits.softmax(dim=-1)                                                        # fp16
        pred = (p16 if mode == 0 else logits).argmax(dim=-1).to(torch.int32)
        if k == ALL_UNLABELED_K:
            return list(range(n)), [class_ids[j] for j in pred.cpu().tolist()]
        board = Leaderboard(c, k, engine.device)
        board.update(p16.float().contiguous(), pred, path_ranks(paths).to(engine.device), prefilter=True)
        return board.result(class_ids)
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
