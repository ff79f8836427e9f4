"""CPU oracle: the reference's prompt-injecting forwards, restated on top of oracle/clip_ref.py.

TEST INFRASTRUCTURE ONLY (see oracle/clip_ref.py header).  /root/reference does not exist on the GPU
box, so the functions the parity tests, smoke() and bench.py's CPU baseline need are restated here,
each citing the reference lines it follows; tests/test_oracle_golden.py pins them against
tests/golden/towers_vitb32_seed1234.npz, which was produced by the reference's own classes.
"""
from __future__ import annotations

import torch

from . import clip_ref


def text_forward(model: clip_ref.CLIP, class_embeddings: torch.Tensor, classes):
    """CustomTextEncoder.forward — models/clip_encoders.py:43-90 (CPU branch: fp32 transformer)."""
    n_prefix = class_embeddings.size()[1]
    prompts = [" ".join([" ".join(["X"] * n_prefix).strip(), c]) for c in classes]      # :54-57
    token_ids = clip_ref.tokenize(prompts)                                                # :60
    text_embedding = model.token_embedding(token_ids)                                     # :63
    text_embedding[:, 1:(class_embeddings[0].size()[0] + 1), :] = class_embeddings        # :67
    x = text_embedding.type(model.dtype) + model.positional_embedding.type(model.dtype)   # :69-74
    x = x.permute(1, 0, 2)
    x = model.transformer(x.float())                                                      # :84
    x = x.permute(1, 0, 2)
    x = model.ln_final(x)                                                                 # :85
    return x[torch.arange(x.shape[0]), token_ids.argmax(dim=-1)] @ model.text_projection  # :86-89


def image_forward(model: clip_ref.CLIP, image: torch.Tensor, image_prefix: torch.Tensor):
    """CustomImageEncoder.forward → CustomVisionTransformer.forward — models/clip_encoders.py:123-194,
    206-208 (deep_embs is None in every shipped config)."""
    v = model.visual
    x = v.conv1(image.type(model.dtype))                                                  # :131
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)                            # :132-133
    cls = v.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype)
    x = torch.cat([cls, x], dim=1)                                                        # :135-144
    x = x + v.positional_embedding.to(x.dtype)                                            # :146
    image_prefix = image_prefix.type(model.dtype).expand(x.shape[0], -1, -1)              # :148
    x = torch.cat([x[:, :1, :], image_prefix, x[:, 1:, :]], dim=1)                        # :150-155
    x = v.ln_pre(x)                                                                       # :157
    x = v.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)                                # :163-187
    x = v.ln_post(x[:, 0, :])                                                             # :189
    return x @ v.proj                                                                     # :191-192


def coop_step(model, prefix, classes, images, labels, lr=None):
    """One CoOp training step — methods/semi_supervised_learning/textual_prompt.py:94-135 with the
    default cross-entropy loss: text features with the learnable prefix, frozen image features,
    cosine logits, CE, backward to the prefix, optional SGD update.  Returns (loss, grad, logits)."""
    prefix = prefix.detach().clone().requires_grad_(True)
    text_features = text_forward(model, prefix, classes)
    text_features = text_features / text_features.norm(dim=-1, keepdim=True)              # :98
    with torch.no_grad():
        image_features = model.encode_image(images)                                       # :100
        image_features = image_features / image_features.norm(dim=-1, keepdim=True)       # :101-103
    logits = model.logit_scale.exp().detach() * image_features @ text_features.t()        # :106-107
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()                                                                       # :131
    if lr is not None:
        with torch.no_grad():
            prefix -= lr * prefix.grad
    return loss.detach(), prefix.grad.detach(), logits.detach()


class UPTHead(torch.nn.Module):
    """The trainable part of UPTModel — models/prompts_models.py:64-153 — restated: CoOp and VPT
    prompts, four Linear projections and the 1-layer / 1-head prompt-coupling transformer."""

    def __init__(self, coop_embeddings, vpt_embeddings, dim_transformer=128):
        super().__init__()
        self.coop_embeddings = torch.nn.Parameter(coop_embeddings)                    # :88
        self.vpt_embeddings = torch.nn.Parameter(vpt_embeddings)                      # :89
        self.coop_length, self.coop_dim = coop_embeddings.size()[1], coop_embeddings.size()[2]
        self.vpt_length, self.vpt_dim = vpt_embeddings.size()[1], vpt_embeddings.size()[2]
        self.proj_coop_pre = torch.nn.Linear(self.coop_dim, dim_transformer)          # :99-114
        self.proj_coop_post = torch.nn.Linear(dim_transformer, self.coop_dim)
        self.proj_vpt_pre = torch.nn.Linear(self.vpt_dim, dim_transformer)
        self.proj_vpt_post = torch.nn.Linear(dim_transformer, self.vpt_dim)
        self.transformer = clip_ref.Transformer(width=dim_transformer, layers=1, heads=1)  # :116-119

    def prompts(self):
        coop = self.proj_coop_pre(self.coop_embeddings)                                # :131-132
        vpt = self.proj_vpt_pre(self.vpt_embeddings)                                   # :135
        seq = torch.cat((coop, vpt), dim=0).to(torch.float32)                          # :138
        out = self.transformer(seq).to(torch.float16)                                  # :141 (hard-coded)
        n = len(self.coop_embeddings)
        coop_embs = self.proj_coop_post(out[:n].to(torch.float32)).reshape(-1, self.coop_length, self.coop_dim)
        vpt_embs = self.proj_vpt_post(out[n:].to(torch.float32)).reshape(-1, self.vpt_length, self.vpt_dim)
        return coop_embs, vpt_embs                                                     # :144-145


def upt_forward(model, head: UPTHead, images, classes):
    """UPTModel.forward — :129-153: (text features, image features), both un-normalised."""
    coop_embs, vpt_embs = head.prompts()
    return text_forward(model, coop_embs, classes), image_forward(model, images, vpt_embs)
