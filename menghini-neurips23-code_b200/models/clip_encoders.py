"""B200 implementations behind the reference's encoder classes (same names, constructor and forward
signatures as models/clip_encoders.py of BatsResearch/menghini-neurips23-code; citations below are to
that file).  The arithmetic runs in libgripb200 through `clip_model.engine`; these classes only do the
host-side work the reference does in Python: building the "X X … X <class>" prompt strings,
tokenising, and handing the learnable prompt rows to the towers."""
from __future__ import annotations

import logging

import torch
import torch.nn as nn

from .. import clip as _clip
from .. import engine as _engine
from .._lib import GripB200Error

log = logging.getLogger(__name__)


class TextEncoder(nn.Module):
    """CLIP text encoder (:13-22)."""

    def __init__(self, clip_model):
        super().__init__()
        self.clip_model = clip_model

    def forward(self, text):
        return self.clip_model.encode_text(text)


class CustomTextEncoder(nn.Module):
    """Text tower with learnable prompt rows (:25-90): the embeddings of the P placeholder tokens
    (rows 1..P) are overwritten by `class_embeddings` before the positional embedding is added."""

    def __init__(self, clip_model, device, dtype):
        super().__init__()
        self.dtype = dtype
        self.clip_model = clip_model
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.token_embedding = clip_model.token_embedding
        self.device = device
        self._ids_cache = {}

    def tokenize(self, text):
        return torch.cat([_clip.tokenize(tok) for tok in text])

    def _prompt_ids(self, n_prefix, classes):
        # :54-60 — the reference re-tokenises on every forward; the ids only depend on (P, classes)
        key = (n_prefix, tuple(classes))
        ids = self._ids_cache.get(key)
        if ids is None:
            prompts = [" ".join([" ".join(["X"] * n_prefix).strip(), c]) for c in classes]
            ids = _clip.tokenize(prompts)
            self._ids_cache = {key: ids}
        return ids

    def forward(self, class_embeddings, classes, enable_pos_emb=True):
        if not enable_pos_emb:
            raise GripB200Error("enable_pos_emb=False is not built (no caller in the reference uses it)")
        if class_embeddings.dim() != 3 or class_embeddings.shape[0] != 1:
            raise GripB200Error(
                f"class_embeddings must be [1, P, 512] (one shared prompt, :67); got {tuple(class_embeddings.shape)}")
        ids = self._prompt_ids(class_embeddings.shape[1], classes)
        return self.clip_model.encode_text(ids, prefix=class_embeddings[0])


class ImageEncoder(nn.Module):
    """CLIP image encoder (:93-102)."""

    def __init__(self, clip_model):
        super().__init__()
        self.clip_model = clip_model

    def forward(self, text):
        return self.clip_model.encode_image(text)


class CustomVisionTransformer(nn.Module):
    """Image tower with prompt rows inserted between CLS and the patches (:105-194); the rows get no
    positional embedding (:148-155) and go through ln_pre with everything else (:157)."""

    def __init__(self, vision_transformer):
        super().__init__()
        self.input_resolution = vision_transformer.input_resolution
        self.output_dim = vision_transformer.output_dim
        self.conv1 = vision_transformer.conv1
        self.class_embedding = vision_transformer.class_embedding
        self.positional_embedding = vision_transformer.positional_embedding
        self.ln_pre = vision_transformer.ln_pre
        self.transformer = vision_transformer.transformer
        self.ln_post = vision_transformer.ln_post
        self.proj = vision_transformer.proj
        self._vt = vision_transformer

    def forward(self, x, image_prefix, pos_emb=True, deep_embs=None):
        if deep_embs is not None:
            # :166-184 dereferences attributes that do not exist (self.visual, self.mvlpt_model); every
            # shipped config sets VPT_DEEP: False
            raise GripB200Error("deep visual prompts are unreachable in the reference and not built")
        if not pos_emb:
            raise GripB200Error("pos_emb=False is not built (no caller in the reference uses it)")
        if image_prefix.dim() == 3:
            if image_prefix.shape[0] != 1:
                raise GripB200Error("image_prefix must be [P,768] or [1,P,768]")
            image_prefix = image_prefix[0]
        return self._vt(x, image_prefix)


class CustomImageEncoder(nn.Module):
    """CLIP image encoder with prompt rows (:198-208)."""

    def __init__(self, visual):
        super().__init__()
        self.visual = CustomVisionTransformer(visual)
        self.dtype = self.visual.conv1.weight.dtype

    def forward(self, image, prefix, deep_embds=None):
        # the reference casts both to the conv dtype (:207); the engine takes fp32 or fp16 pixels and
        # keeps the prompt rows in fp32 (≥ the reference's precision)
        return self.visual(image, prefix, deep_embs=deep_embds)
