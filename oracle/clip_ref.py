"""CPU oracle: restatement of the third-party `clip` package the reference's hot path runs on.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Why it exists: the reference (BatsResearch/menghini-neurips23-code) has no arithmetic of its own on
this path — `models/clip_encoders.py:29-38,106-119` re-wires sub-modules of `openai/CLIP`
(`requirements.txt:2`, `git+https://github.com/openai/CLIP.git`, UN-PINNED and not vendored under
/root/reference).  This file restates the published algorithm of that dependency's
`clip/model.py` + `clip/clip.py` (public v1.0 source; class and attribute names kept so the
reference's own modules run on top of it unchanged when it is installed as `sys.modules['clip']`,
see `install_as_clip()`).

Parity status: **unpinned by the reference** (it ships no tests, golden vectors or fixtures for this
path).  The restatement is pinned instead by an independent implementation of the same
published model — Hugging Face `transformers.CLIPModel` with `hidden_act="quick_gelu"` —
in tests/test_oracle_clip.py, and the reference's own `models/*.py` / `utils/clip_pseudolabels.py`
are executed on top of it by oracle/make_golden.py to produce tests/golden/*.npz.

Reference call sites this module serves:
  clip.load              methods/clip_baseline.py:39-41
  clip.tokenize          models/clip_encoders.py:41,60 ; utils/clip_pseudolabels.py:25
  clip.model.Transformer models/prompts_models.py:116-119
  clip_model(img, text)  utils/clip_pseudolabels.py:35-37,59-61
  encode_image/_text     methods/semi_supervised_learning/textual_prompt.py:100, visual_prompt.py:117
"""
from __future__ import annotations

import math
import sys
import types
from collections import OrderedDict
from typing import List, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

# ViT-B/32 hyper-parameters of the released checkpoint (openai/CLIP model card).
VITB32 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
              vision_patch_size=32, context_length=77, vocab_size=49408, transformer_width=512,
              transformer_heads=8, transformer_layers=12)
SOT, EOT = 49406, 49407


class LayerNorm(nn.LayerNorm):
    """clip/model.py LayerNorm: computes in fp32 whatever the storage dtype, eps 1e-5."""

    def forward(self, x: torch.Tensor):
        orig_type = x.dtype
        ret = super().forward(x.type(torch.float32))
        return ret.type(orig_type)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    """x = x + attn(ln_1(x)); x = x + c_proj(QuickGELU(c_fc(ln_2(x)))).  Input layout LND."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x: torch.Tensor):
        self.attn_mask = (self.attn_mask.to(dtype=x.dtype, device=x.device)
                          if self.attn_mask is not None else None)
        return self.attn(x, x, x, need_weights=False, attn_mask=self.attn_mask)[0]

    def forward(self, x: torch.Tensor):
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(
            *[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int,
                 output_dim: int):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(
            scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor):
        x = self.conv1(x)  # [B, width, grid, grid]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # [B, grid², width]
        cls = self.class_embedding.to(x.dtype) + torch.zeros(
            x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1)
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)  # NLD -> LND
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        if self.proj is not None:
            x = x @ self.proj
        return x


class CLIP(nn.Module):
    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: int, vision_width: int,
                 vision_patch_size: int, context_length: int, vocab_size: int,
                 transformer_width: int, transformer_heads: int, transformer_layers: int):
        super().__init__()
        self.context_length = context_length
        vision_heads = vision_width // 64
        self.visual = VisionTransformer(image_resolution, vision_patch_size, vision_width,
                                        vision_layers, vision_heads, embed_dim)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    def encode_text(self, text):
        x = self.token_embedding(text).type(self.dtype)  # [C, ctx, d]
        x = x + self.positional_embedding.type(self.dtype)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        # features of the EOT token = the highest id in each sequence
        x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection
        return x

    def forward(self, image, text):
        image_features = self.encode_image(image)
        text_features = self.encode_text(text)
        image_features = image_features / image_features.norm(dim=1, keepdim=True)
        text_features = text_features / text_features.norm(dim=1, keepdim=True)
        logit_scale = self.logit_scale.exp()
        logits_per_image = logit_scale * image_features @ text_features.t()
        logits_per_text = logits_per_image.t()
        return logits_per_image, logits_per_text


# --------------------------------------------------------------------------------------------------
# Synthetic, seeded weights (there is no network: no released checkpoint).
# --------------------------------------------------------------------------------------------------
def synth_state_dict(cfg: dict = VITB32, seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    """fp32 state_dict with CLIP-style init scales for BOTH towers, plus non-trivial biases and
    LayerNorm affine parameters so every term of the forward is exercised.  Deterministic in
    `seed` (torch CPU generator).  Keys are those of openai/CLIP's `CLIP.state_dict()`."""
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

    vw, ps = cfg["vision_width"], cfg["vision_patch_size"]
    grid = cfg["image_resolution"] // ps
    sd["visual.conv1.weight"] = rn(vw, 3, ps, ps, std=(3 * ps * ps) ** -0.5)
    sd["visual.class_embedding"] = rn(vw, std=vw ** -0.5)
    sd["visual.positional_embedding"] = rn(grid * grid + 1, vw, std=vw ** -0.5)
    sd["visual.ln_pre.weight"] = 1.0 + rn(vw, std=0.1)
    sd["visual.ln_pre.bias"] = rn(vw, std=0.1)
    tower("visual.transformer.", vw, cfg["vision_layers"])
    sd["visual.ln_post.weight"] = 1.0 + rn(vw, std=0.1)
    sd["visual.ln_post.bias"] = rn(vw, std=0.1)
    sd["visual.proj"] = rn(vw, cfg["embed_dim"], std=vw ** -0.5)
    tw = cfg["transformer_width"]
    sd["token_embedding.weight"] = rn(cfg["vocab_size"], tw, std=0.02)
    sd["positional_embedding"] = rn(cfg["context_length"], tw, std=0.01)
    tower("transformer.", tw, cfg["transformer_layers"])
    sd["ln_final.weight"] = 1.0 + rn(tw, std=0.1)
    sd["ln_final.bias"] = rn(tw, std=0.1)
    sd["text_projection"] = rn(tw, cfg["embed_dim"], std=tw ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(100.0))
    return sd


def round_fp16_(sd):
    """Round the tensors `clip.load` keeps in fp16 on CUDA (conv/linear/MHA weights+biases, proj,
    text_projection) through fp16, in place: the 'fp16-rounded fp32' oracle variant that isolates
    activation rounding from weight rounding."""
    for k, v in sd.items():
        if (k.endswith(("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias",
                        "c_fc.weight", "c_fc.bias", "c_proj.weight", "c_proj.bias", "conv1.weight"))
                or k in ("visual.proj", "text_projection")):
            sd[k] = v.half().float()
    return sd


def build_model(state_dict=None, cfg: dict = VITB32, seed: int = 1234) -> CLIP:
    """`clip.load(name, device='cpu')` equivalent: fp32 model in eval mode."""
    model = CLIP(**cfg)
    sd = synth_state_dict(cfg, seed) if state_dict is None else state_dict
    model.load_state_dict(sd)
    return model.float().eval()


# --------------------------------------------------------------------------------------------------
# Tokenizer stand-in.  The BPE vocabulary (bpe_simple_vocab_16e6.txt.gz) is not available offline,
# so words map to deterministic ids; the structure openai/CLIP's tokenize() guarantees is kept:
# [SOT] + tokens + [EOT] zero-padded to context_length, EOT being the arg-max id of the row
# (that is all models/clip_encoders.py:86-89 relies on).  "X" placeholders map to one fixed id
# (their embeddings are overwritten by the learned prefix, models/clip_encoders.py:67).
# --------------------------------------------------------------------------------------------------
def _word_id(word: str) -> int:
    h = 2166136261
    for ch in word.encode("utf-8"):
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return 1000 + h % 39000  # in [1000, 40000) — below SOT/EOT


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False):
    if isinstance(texts, str):
        texts = [texts]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        words = text.lower().split()
        tokens = [SOT] + [_word_id(w) for w in words] + [EOT]
        if len(tokens) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            tokens = tokens[:context_length]
            tokens[-1] = EOT
        result[i, :len(tokens)] = torch.tensor(tokens)
    return result


def _identity_transform(img):
    return img


def clip_transform():
    """openai/CLIP's `_transform(224)`: Resize(224, bicubic) → CenterCrop(224) → RGB → ToTensor → Normalize, through
    torchvision (what the third-party package itself uses)."""
    import torchvision.transforms as tv

    return tv.Compose([tv.Resize(224, interpolation=tv.InterpolationMode.BICUBIC), tv.CenterCrop(224),
                       lambda im: im.convert("RGB"), tv.ToTensor(),
                       tv.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


def load(name: str = "ViT-B/32", device="cpu", jit: bool = False, seed: int = 1234):
    """`clip.load` on CPU: (fp32 eval model, preprocess).  Only ViT-B/32 shapes exist offline."""
    if name.replace("/", "").replace("-", "").lower() != "vitb32":
        raise RuntimeError(f"oracle clip: only ViT-B/32 is restated (got {name})")
    if str(device) != "cpu":
        raise RuntimeError("oracle clip is the CPU path; the CUDA path is the product (libgripb200)")
    return build_model(seed=seed), _identity_transform


def install_as_clip():
    """Register this module as `clip`, `clip.clip` and `clip.model` so the reference's own
    `models/*.py` / `utils/clip_pseudolabels.py` import it (`from clip import clip`, `import clip`,
    `clip.model.Transformer`)."""
    me = sys.modules[__name__]
    top = types.ModuleType("clip")
    for n in ("load", "tokenize", "CLIP", "Transformer", "LayerNorm", "QuickGELU",
              "VisionTransformer", "ResidualAttentionBlock", "build_model"):
        setattr(top, n, getattr(me, n))
    top.clip = top  # `from clip import clip`
    top.model = me  # `clip.model.Transformer`
    top.available_models = lambda: ["ViT-B/32"]
    sys.modules["clip"] = top
    sys.modules["clip.clip"] = top
    sys.modules["clip.model"] = me
    return top
