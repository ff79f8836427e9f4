"""Pins oracle/clip_ref.py (the restated third-party `clip`) against an INDEPENDENT implementation
of the same published model: Hugging Face transformers' CLIPModel with quick_gelu.  The reference
ships no fixture for this path, so this is what stands behind the restatement."""
import numpy as np
import pytest
import torch

from oracle import clip_ref

CFG = dict(embed_dim=32, image_resolution=64, vision_layers=2, vision_width=128,
           vision_patch_size=32, context_length=77, vocab_size=49408, transformer_width=64,
           transformer_heads=2, transformer_layers=2)


def _to_hf(model, cfg):
    from transformers import CLIPConfig, CLIPModel

    hc = CLIPConfig(
        text_config=dict(hidden_size=cfg["transformer_width"], intermediate_size=4 * cfg["transformer_width"],
                         num_hidden_layers=cfg["transformer_layers"], num_attention_heads=cfg["transformer_heads"],
                         max_position_embeddings=77, vocab_size=cfg["vocab_size"], hidden_act="quick_gelu",
                         layer_norm_eps=1e-5, eos_token_id=clip_ref.EOT, bos_token_id=clip_ref.SOT, pad_token_id=0),
        vision_config=dict(hidden_size=cfg["vision_width"], intermediate_size=4 * cfg["vision_width"],
                           num_hidden_layers=cfg["vision_layers"], num_attention_heads=cfg["vision_width"] // 64,
                           image_size=cfg["image_resolution"], patch_size=cfg["vision_patch_size"],
                           hidden_act="quick_gelu", layer_norm_eps=1e-5),
        projection_dim=cfg["embed_dim"])
    hf = CLIPModel(hc).eval()
    sd = model.state_dict()
    new = {}

    def tower(src, dst, layers, width):
        for i in range(layers):
            s, d = f"{src}resblocks.{i}.", f"{dst}encoder.layers.{i}."
            w, b = sd[s + "attn.in_proj_weight"], sd[s + "attn.in_proj_bias"]
            for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
                new[d + f"self_attn.{n}.weight"] = w[j * width:(j + 1) * width]
                new[d + f"self_attn.{n}.bias"] = b[j * width:(j + 1) * width]
            new[d + "self_attn.out_proj.weight"] = sd[s + "attn.out_proj.weight"]
            new[d + "self_attn.out_proj.bias"] = sd[s + "attn.out_proj.bias"]
            for a, bname in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                             ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
                new[d + bname + ".weight"] = sd[s + a + ".weight"]
                new[d + bname + ".bias"] = sd[s + a + ".bias"]

    tower("visual.transformer.", "vision_model.", cfg["vision_layers"], cfg["vision_width"])
    tower("transformer.", "text_model.", cfg["transformer_layers"], cfg["transformer_width"])
    new["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
    new["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
    new["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
    new["vision_model.pre_layrnorm.weight"] = sd["visual.ln_pre.weight"]
    new["vision_model.pre_layrnorm.bias"] = sd["visual.ln_pre.bias"]
    new["vision_model.post_layernorm.weight"] = sd["visual.ln_post.weight"]
    new["vision_model.post_layernorm.bias"] = sd["visual.ln_post.bias"]
    new["visual_projection.weight"] = sd["visual.proj"].t()
    new["text_model.embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
    new["text_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    new["text_model.final_layer_norm.weight"] = sd["ln_final.weight"]
    new["text_model.final_layer_norm.bias"] = sd["ln_final.bias"]
    new["text_projection.weight"] = sd["text_projection"].t()
    new["logit_scale"] = sd["logit_scale"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return hf


@pytest.fixture(scope="module")
def pair():
    torch.manual_seed(0)
    model = clip_ref.build_model(clip_ref.synth_state_dict(CFG, seed=7), CFG)
    return model, _to_hf(model, CFG)


def test_image_tower_matches_hf(pair):
    model, hf = pair
    img = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = model.encode_image(img)
        b = hf.get_image_features(pixel_values=img)
        b = getattr(b, "pooler_output", b)
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-4), (a - b).abs().max()


def test_text_tower_matches_hf(pair):
    model, hf = pair
    ids = clip_ref.tokenize(["a photo of a forest", "x x x x river bank", "sea"])
    with torch.no_grad():
        a = model.encode_text(ids)
        b = hf.get_text_features(input_ids=ids, attention_mask=(ids != 0).long())
        b = getattr(b, "pooler_output", b)
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-4), (a - b).abs().max()


def test_logits_match_hf(pair):
    model, hf = pair
    img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    ids = clip_ref.tokenize(["annual crop land", "highway or road", "industrial buildings"])
    with torch.no_grad():
        li, lt = model(img, ids)
        out = hf(input_ids=ids, pixel_values=img, attention_mask=(ids != 0).long())
    assert torch.allclose(li, out.logits_per_image, atol=1e-4, rtol=1e-4)
    assert torch.equal(lt, li.t())


def test_tokenize_structure():
    ids = clip_ref.tokenize(["X X X X annual crop", "sea"])
    assert ids.shape == (2, 77) and ids.dtype == torch.long
    assert ids[0, 0] == clip_ref.SOT and ids[0, 7] == clip_ref.EOT and ids[0, 8:].sum() == 0
    assert ids.argmax(-1).tolist() == [7, 2]
    assert ids[0, 1] == ids[0, 4]  # the placeholder maps to one id
    with pytest.raises(RuntimeError):
        clip_ref.tokenize(" ".join(["w"] * 100))


def test_fp16_rounding_variant_changes_only_fp16_tensors():
    sd = clip_ref.synth_state_dict(CFG, seed=7)
    ref = {k: v.clone() for k, v in sd.items()}
    clip_ref.round_fp16_(sd)
    assert torch.equal(sd["visual.ln_pre.weight"], ref["visual.ln_pre.weight"])
    assert torch.equal(sd["token_embedding.weight"], ref["token_embedding.weight"])
    assert not torch.equal(sd["visual.proj"], ref["visual.proj"])
    assert torch.equal(sd["visual.proj"], ref["visual.proj"].half().float())
