"""Stand-in for the third-party `clip` package on the reference's import seam (`import clip`,
`from clip import clip`, `clip.model.Transformer`): same entry points, B200 engine underneath.

    clip.load("ViT-B/32", device)  → (model, preprocess)     methods/clip_baseline.py:39-41
    clip.tokenize(list[str])       → LongTensor [n, 77]       models/clip_encoders.py:41,60

Weights: there is no network in this environment, so `load` takes them from (in this order) the
`state_dict=` argument, the file named by $GRIPB200_CLIP_WEIGHTS (a torch-saved state_dict in
openai/CLIP layout, e.g. `torch.jit.load('ViT-B-32.pt').state_dict()`), or — when
$GRIPB200_SYNTHETIC_SEED is set — seeded random-init weights of the ViT-B/32 architecture.
Tokeniser: `clip.tokenize` is openai/CLIP's byte-pair tokenizer (simple_tokenizer.py) when
$GRIPB200_BPE_VOCAB names its vocabulary file `bpe_simple_vocab_16e6.txt.gz` (third-party data, not
available offline); without it a deterministic word-hash tokenizer with the same framing
([SOT] … [EOT], zero padded, EOT = arg-max id) is used and says so once.
"""
from __future__ import annotations

import os
import warnings

import torch

from . import model  # noqa: F401  (clip.model.Transformer)
from .model import CLIP, Transformer, build_model  # noqa: F401
from .._lib import GripB200Error
from ..synthetic import synthetic_state_dict

SOT, EOT = 49406, 49407
_warned = False
_real_weights = False  # set by load() when the weights came from a file of trained parameters


def available_models():
    return ["ViT-B/32"]


def _preprocess(raw: bool = False):
    """CLIP's transform: Resize(224, bicubic) → CenterCrop → RGB → ToTensor → Normalize.
    raw=True stops after the crop and returns the uint8 [3,224,224] pixels: the image tower applies
    ToTensor + Normalize on the device (bit-identical features, a quarter of the host→device bytes)."""
    import numpy as np
    from PIL import Image

    mean = torch.tensor((0.48145466, 0.4578275, 0.40821073)).view(3, 1, 1)
    std = torch.tensor((0.26862954, 0.26130258, 0.27577711)).view(3, 1, 1)

    def transform(img):
        # torchvision Resize(224): the short side becomes exactly 224, the long side int(224·long/short)
        # (truncation); CenterCrop(224): origin int(round((dim − 224) / 2.0)) (Python's round)
        w, h = img.size
        if w <= h:
            nw, nh = 224, int(224 * h / w)
        else:
            nw, nh = int(224 * w / h), 224
        if (nw, nh) != (w, h):
            img = img.resize((nw, nh), Image.BICUBIC)
        l, t = int(round((nw - 224) / 2.0)), int(round((nh - 224) / 2.0))
        img = img.crop((l, t, l + 224, t + 224)).convert("RGB")
        x = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1)
        if raw:
            return x.contiguous()
        return (x.float() / 255.0 - mean) / std

    transform.is_clip_preprocess_u8 = bool(raw)   # utils.encode_pool may run exactly this resize + crop on the device
    if not raw:
        # the same transform without its ToTensor / Normalize tail: batched pool encoders use it and let the
        # device normalise (bit-identical features, a quarter of the host→device bytes)
        transform.raw_u8 = _preprocess(raw=True)
    return transform


def preprocess_u8():
    """The transform of `load()` without its ToTensor/Normalize tail (see `_preprocess`)."""
    return _preprocess(raw=True)


def normalize_u8(x: torch.Tensor) -> torch.Tensor:
    """Host restatement of what the device does with uint8 pixels [..., 3, 224, 224] (ToTensor + Normalize)."""
    mean = torch.tensor((0.48145466, 0.4578275, 0.40821073)).view(3, 1, 1)
    std = torch.tensor((0.26862954, 0.26130258, 0.27577711)).view(3, 1, 1)
    return (x.float() / 255.0 - mean) / std


def load(name="ViT-B/32", device="cuda", jit=False, state_dict=None, download_root=None):
    if name.replace("/", "").replace("-", "").lower() != "vitb32":
        raise GripB200Error(f"only ViT-B/32 is built for B200 (got {name})")
    if state_dict is None:
        path = os.environ.get("GRIPB200_CLIP_WEIGHTS")
        seed = os.environ.get("GRIPB200_SYNTHETIC_SEED")
        if path:
            global _real_weights
            state_dict = torch.load(path, map_location="cpu")
            _real_weights = True
        elif seed is not None:
            state_dict = synthetic_state_dict(int(seed))
        else:
            raise GripB200Error(
                "no CLIP weights: pass state_dict=, or set GRIPB200_CLIP_WEIGHTS to a saved "
                "state_dict, or GRIPB200_SYNTHETIC_SEED for random-init weights")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise GripB200Error("the B200 CLIP runs on CUDA devices only; there is no CPU fallback")
    return build_model(state_dict, dev), _preprocess()


def _word_id(word: str) -> int:
    h = 2166136261
    for ch in word.encode("utf-8"):
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return 1000 + h % 39000


def tokenize(texts, context_length: int = 77, truncate: bool = False):
    global _warned
    if isinstance(texts, str):
        texts = [texts]
    vocab = os.environ.get("GRIPB200_BPE_VOCAB")
    if vocab:
        return _tokenize_bpe(texts, context_length, truncate, vocab)
    if _real_weights and os.environ.get("GRIPB200_ALLOW_HASH_TOKENIZER") != "1":
        # hash ids bear no relation to a trained token embedding: everything would run and mean nothing
        raise GripB200Error(
            "clip.tokenize: real CLIP weights were loaded (GRIPB200_CLIP_WEIGHTS) but no BPE vocabulary is "
            "configured; set GRIPB200_BPE_VOCAB to bpe_simple_vocab_16e6.txt.gz (or GRIPB200_ALLOW_HASH_TOKENIZER=1 "
            "to opt into the word-hash tokenizer)")
    if not _warned:
        warnings.warn("clip.tokenize: BPE vocabulary unavailable offline, using the word-hash tokenizer")
        _warned = True
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        toks = [SOT] + [_word_id(w) for w in text.lower().split()] + [EOT]
        if len(toks) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            toks = toks[:context_length]
            toks[-1] = EOT
        out[i, :len(toks)] = torch.tensor(toks)
    return out


_bpe = {}


def _tokenize_bpe(texts, context_length, truncate, vocab_path):
    """openai/CLIP clip.tokenize: [SOT] + BPE(text) + [EOT], zero padded to context_length."""
    from .simple_tokenizer import SimpleTokenizer

    tok = _bpe.get(vocab_path)
    if tok is None:
        tok = _bpe[vocab_path] = SimpleTokenizer(vocab_path)
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        toks = [sot] + tok.encode(text) + [eot]
        if len(toks) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            toks = toks[:context_length]
            toks[-1] = eot
        out[i, :len(toks)] = torch.tensor(toks)
    return out


# `from clip import clip` (models/clip_encoders.py:7) expects a sub-module with load/tokenize
import sys as _sys
clip = _sys.modules[__name__]
