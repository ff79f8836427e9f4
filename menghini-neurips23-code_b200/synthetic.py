"""Seeded random-init weights of the CLIP ViT-B/32 architecture (no checkpoint can be downloaded
here).  CLIP-style init scales for both towers, non-trivial biases and LayerNorm affine parameters.
Keys follow openai/CLIP's `CLIP.state_dict()`."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


def synthetic_state_dict(seed: int = 1234):
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = OrderedDict()

    def tower(prefix, width, layers):
        attn_std = width ** -0.5
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        fc_std = (2 * width) ** -0.5
        for i in range(layers):
            p = f"{prefix}resblocks.{i}."
            sd[p + "attn.in_proj_weight"] = rn(3 * width, width, std=attn_std)
            sd[p + "attn.in_proj_bias"] = rn(3 * width, std=0.02)
            sd[p + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
            sd[p + "attn.out_proj.bias"] = rn(width, std=0.02)
            sd[p + "ln_1.weight"] = 1.0 + rn(width, std=0.1)
            sd[p + "ln_1.bias"] = rn(width, std=0.1)
            sd[p + "mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std)
            sd[p + "mlp.c_fc.bias"] = rn(4 * width, std=0.02)
            sd[p + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
            sd[p + "mlp.c_proj.bias"] = rn(width, std=0.02)
            sd[p + "ln_2.weight"] = 1.0 + rn(width, std=0.1)
            sd[p + "ln_2.bias"] = rn(width, std=0.1)

    sd["visual.conv1.weight"] = rn(768, 3, 32, 32, std=3072 ** -0.5)
    sd["visual.class_embedding"] = rn(768, std=768 ** -0.5)
    sd["visual.positional_embedding"] = rn(50, 768, std=768 ** -0.5)
    sd["visual.ln_pre.weight"] = 1.0 + rn(768, std=0.1)
    sd["visual.ln_pre.bias"] = rn(768, std=0.1)
    tower("visual.transformer.", 768, 12)
    sd["visual.ln_post.weight"] = 1.0 + rn(768, std=0.1)
    sd["visual.ln_post.bias"] = rn(768, std=0.1)
    sd["visual.proj"] = rn(768, 512, std=768 ** -0.5)
    sd["token_embedding.weight"] = rn(49408, 512, std=0.02)
    sd["positional_embedding"] = rn(77, 512, std=0.01)
    tower("transformer.", 512, 12)
    sd["ln_final.weight"] = 1.0 + rn(512, std=0.1)
    sd["ln_final.bias"] = rn(512, std=0.1)
    sd["text_projection"] = rn(512, 512, std=512 ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(100.0))
    return sd
