"""Modules owning the trainable prompt parameters — same classes, constructor arguments, attribute
names and return values as models/prompts_models.py of the reference (citations are to that file).
Gradients reach `prefix` / the UPT head through the engine's prompt-only backward."""
from __future__ import annotations

import logging

import torch
from torch import nn

from .. import clip as _clip

log = logging.getLogger(__name__)


class TextPrefixModel(nn.Module):
    """:10-36 — forward(classes) returns the UN-normalised text features (norm_out is dead code)."""

    def __init__(self, initial_prefix, text_encoder, classes, temperature=0.07, device="cpu"):
        super().__init__()
        self.device = device
        self.initialized_prefix = initial_prefix
        self.classes = classes
        self.prefix = nn.Parameter(initial_prefix)
        self.text_encoder = text_encoder

    def forward(self, classes):
        return self.text_encoder(self.prefix, classes)


class ImagePrefixModel(nn.Module):
    """:39-61 — forward(x) returns the UN-normalised image features."""

    def __init__(self, initial_prefix, image_encoder, temperature=0.07, device="cpu"):
        super().__init__()
        self.device = device
        self.initialized_prefix = initial_prefix
        self.prefix = nn.Parameter(initial_prefix)
        self.image_encoder = image_encoder

    def forward(self, x):
        return self.image_encoder(x, self.prefix)


class UPTModel(nn.Module):
    """:64-153 — CoOp and VPT prompts coupled through Linear → 1-layer/1-head transformer → Linear.
    The coupling head is 0.5 M parameters acting on 2×4 tokens; it stays a torch module (its
    autograd graph continues into the two towers' prompt-gradient kernels)."""

    def __init__(self, coop_embeddings, vpt_embeddings, vpt_embeddings_deep, image_encoder,
                 text_encoder, classes, dim_transformer, temperature=0.07, device="cpu",
                 dtype=torch.float32):
        super().__init__()
        self.device = device
        self.classes = classes
        self.temperature = temperature
        self.dtype = dtype
        self.coop_embeddings = nn.Parameter(coop_embeddings)
        self.vpt_embeddings = nn.Parameter(vpt_embeddings)
        self.coop_length, self.coop_dim = self.coop_embeddings.size()[1], self.coop_embeddings.size()[2]
        self.vpt_length, self.vpt_dim = self.vpt_embeddings.size()[1], self.vpt_embeddings.size()[2]
        self.vpt_embeddings_deep = (nn.Parameter(vpt_embeddings_deep)
                                    if vpt_embeddings_deep is not None else None)
        self.proj_coop_pre = nn.Linear(self.coop_dim, dim_transformer, dtype=self.dtype).to(self.device)
        self.proj_coop_post = nn.Linear(dim_transformer, self.coop_dim, dtype=self.dtype).to(self.device)
        self.proj_vpt_pre = nn.Linear(self.vpt_dim, dim_transformer, dtype=self.dtype).to(self.device)
        self.proj_vpt_post = nn.Linear(dim_transformer, self.vpt_dim, dtype=self.dtype).to(self.device)
        self.transformer = _clip.model.Transformer(width=dim_transformer, layers=1, heads=1).to(self.device)
        self.image_encoder = image_encoder
        self.text_encoder = text_encoder

    def prompt_embeddings(self):
        """The coupled prompts (coop_embs [1,Pt,512], vpt_embs [1,Pv,768]) the two towers receive: :131-145."""
        coop_embds = self.proj_coop_pre(self.coop_embeddings).to(self.device)          # :131-132
        vpt_embds = self.proj_vpt_pre(self.vpt_embeddings).to(self.device)             # :135
        # :138 — dim 0 (coop | vpt) is what the transformer treats as the sequence
        prompt_seq = torch.cat((coop_embds, vpt_embds), dim=0).to(torch.float32)
        output_seq = self.transformer(prompt_seq).to(torch.float16)                    # :141 (hard-coded)
        n = len(self.coop_embeddings)
        coop_embs = self.proj_coop_post(output_seq[:n].to(self.dtype)).reshape(-1, self.coop_length, self.coop_dim)
        vpt_embs = self.proj_vpt_post(output_seq[n:].to(self.dtype)).reshape(-1, self.vpt_length, self.vpt_dim)
        return coop_embs, vpt_embs

    def forward(self, x, classes):
        coop_embs, vpt_embs = self.prompt_embeddings()
        text_out = self.text_encoder(coop_embs, classes)                               # :148
        visual_out = self.image_encoder(x, vpt_embs)                                   # :150
        return text_out, visual_out
