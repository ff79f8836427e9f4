"""Module tree of CLIP ViT-B/32 with the attribute and state_dict layout of the third-party
`clip.model` the reference re-wires (models/clip_encoders.py:29-38,106-119), whose tower forwards
run on the B200 engine (libgripb200) instead of torch ops.

The parameters live here in `clip.load`'s CUDA convention (fp16 conv/linear/attention/projection
weights, fp32 LayerNorm and embeddings) so a released checkpoint's state_dict loads unchanged and
callers that poke at sub-modules (`visual.conv1.weight.dtype`, `logit_scale`, `text_projection`,
`token_embedding`) keep working.  `Transformer.forward` is a plain torch implementation: it serves
the 1-layer, 128-wide prompt-coupling transformer of UPT (models/prompts_models.py:116-119, sequence
length 2 — not a hot path); the two towers never go through it.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from .. import engine as _engine
from .._lib import GripB200Error


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return super().forward(x.type(torch.float32)).type(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)),
                                              ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def forward(self, x):
        mask = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=mask)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask)
                                         for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(width))
        self.positional_embedding = nn.Parameter(torch.zeros((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(torch.zeros(width, output_dim))
        self._owner = None  # set by CLIP: the engine lives on the full model

    def forward(self, x, prefix=None):
        clip_model = self._owner() if self._owner is not None else None
        if clip_model is None:
            raise GripB200Error("VisionTransformer is not attached to a B200 CLIP model")
        return clip_model._encode_image(x, prefix)


class CLIP(nn.Module):
    def __init__(self, embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
                 vision_patch_size=32, context_length=77, vocab_size=49408, transformer_width=512,
                 transformer_heads=8, transformer_layers=12):
        super().__init__()
        if (embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
                context_length, transformer_width, transformer_heads, transformer_layers) != (
                512, 224, 12, 768, 32, 77, 512, 8, 12):
            raise GripB200Error("only the ViT-B/32 geometry is built for B200")
        self.context_length = context_length
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width,
                                        vision_layers, vision_width // 64, embed_dim)
        mask = torch.empty(context_length, context_length).fill_(float("-inf")).triu_(1)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads, mask)
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.zeros(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.zeros(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        import weakref
        self.visual._owner = weakref.ref(self)
        self._engine = None

    # -- engine --------------------------------------------------------------------------------
    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    @property
    def engine(self) -> "_engine.Engine":
        if self._engine is None:
            dev = self.visual.conv1.weight.device
            self._engine = _engine.Engine(self.state_dict(), dev)
        return self._engine

    # The engine produces fp32 features (fp16 storage / fp32 accumulation inside the towers).  The reference's
    # CUDA model returns them in the model dtype, fp16 (clip.load's convention); a caller that needs that
    # rounding — e.g. to mix them with fp16 tensors of its own — sets `clip_model.feature_dtype = torch.float16`
    # (or GRIPB200_FEATURE_DTYPE=fp16 before clip.load); the default keeps the extra precision.
    feature_dtype = torch.float32

    def _feat_dtype(self, t):
        return t if t.dtype == self.feature_dtype else t.to(self.feature_dtype)

    def _encode_image(self, image, prefix=None):
        image = image.to(self.visual.conv1.weight.device)
        if prefix is None:
            return self._feat_dtype(self.engine.frozen_image_features(image))
        return self._feat_dtype(_engine.vit_with_prefix(self.engine, image, prefix))

    def encode_image(self, image):
        return self._encode_image(image, None)

    def encode_text(self, text, prefix=None):
        if prefix is None:
            return self._feat_dtype(self.engine.text_forward(text, None)[0])
        return self._feat_dtype(_engine.text_with_prefix(self.engine, text, prefix))

    def forward(self, image, text):
        eng = self.engine
        image = image.to(eng.device)
        _, fi, _ = eng.vit_forward(image, None, want_feat=False, want_featn=True)
        _, ft, _ = eng.text_forward(text, None, want_feat=False, want_featn=True)
        # logits of every (image, prompt) pair; tiny next to the towers
        logits = eng.logit_scale_exp * (fi.float() @ ft.float().t())
        return logits, logits.t()


def convert_weights(model: nn.Module):
    """clip.model.convert_weights: fp16 for conv/linear/attention weights+biases and projections."""
    def _c(l):
        if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            l.weight.data = l.weight.data.half()
            if l.bias is not None:
                l.bias.data = l.bias.data.half()
        if isinstance(l, nn.MultiheadAttention):
            for a in ("in_proj_weight", "q_proj_weight", "k_proj_weight", "v_proj_weight",
                      "in_proj_bias", "bias_k", "bias_v"):
                t = getattr(l, a)
                if t is not None:
                    t.data = t.data.half()
        for n in ("text_projection", "proj"):
            if hasattr(l, n):
                a = getattr(l, n)
                if a is not None:
                    a.data = a.data.half()
    model.apply(_c)


def build_model(state_dict, device="cuda:0"):
    """state_dict in openai/CLIP's layout → frozen eval model on `device` in clip.load's fp16 layout."""
    model = CLIP()
    sd = {k: v for k, v in state_dict.items()
          if k not in ("input_resolution", "context_length", "vocab_size")}
    model.load_state_dict(sd)
    convert_weights(model)
    model = model.to(device).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    import os
    if os.environ.get("GRIPB200_FEATURE_DTYPE", "").lower() in ("fp16", "float16", "half"):
        model.feature_dtype = torch.float16
    return model
