"""CPU: the oracle against the committed golden vectors (which were produced by running the
reference's own modules, oracle/make_golden.py), plus host-side logic of the product."""
import importlib

import numpy as np
import pytest
import torch

from oracle import clip_ref, leaderboard_ref, prompt_ref, synth


def test_leaderboard_restatement_matches_reference_goldens(golden_dir):
    g = np.load(f"{golden_dir}/leaderboard_cases.npz")
    names = [str(n) for n in g["names"]]
    assert len(names) >= 9
    for name in names:
        idx, lab = leaderboard_ref.leaderboard(g[f"{name}.probs"], g[f"{name}.pred"], int(g[f"{name}.k"]),
                                               g[f"{name}.rank"], g[f"{name}.class_ids"].tolist())
        assert idx == g[f"{name}.out_idx"].tolist(), name
        assert lab == g[f"{name}.out_lab"].tolist(), name


def test_leaderboard_quirks_are_reproduced():
    # SURVEY Appendix A consequence 1+2: full-but-unsorted board rejects a true top-k member, and the
    # threshold DROPS at the first sort.
    probs = np.array([[0.9, 0.1], [0.6, 0.4], [0.7, 0.3],   # board 0 fills in arrival order: .9 .6 .7
                      [0.65, 0.35],                         # .65 ≤ last(.7) → rejected (beats .6!)
                      [0.8, 0.2],                           # admitted → sorted [.9 .8 .7], drops .6
                      [0.75, 0.25]], np.float32)            # .75 > last(.7) → [.9 .8 .75]
    idx, lab = leaderboard_ref.leaderboard(probs, [0] * 6, 3, list(range(6)))
    assert idx[:3] == [0, 4, 5]
    # spill: image 3 was offered to board 1 with p=0.35
    assert 3 in idx[3:]
    # strict '<': equal probability never replaces
    probs = np.tile(np.array([[0.7, 0.3]], np.float32), (5, 1))
    idx, _ = leaderboard_ref.leaderboard(probs, [0] * 5, 2, [4, 3, 2, 1, 0])
    assert idx[:2] == [0, 1]


def test_towers_golden_reproducible_from_seeds(golden_dir):
    """The oracle alone (no reference present) reproduces the golden features: inputs and weights are
    pure functions of their seeds."""
    g = np.load(f"{golden_dir}/towers_vitb32_seed1234.npz")
    model = clip_ref.build_model(seed=1234)
    with torch.no_grad():
        got = model.encode_image(synth.images(2, seed=0)).numpy()
        ids = torch.from_numpy(g["txt_ids_zeroshot"])
        got_t = model.encode_text(ids).numpy()
    assert np.abs(got - g["img_feat_p0"]).max() < 1e-4
    assert np.abs(got_t - g["txt_feat_zeroshot"]).max() < 1e-4
    classes = [" ".join(c.split("_")) for c in synth.class_names(5, seed=1)]
    prompts = [f"a photo of a {{}}{c}" for c in classes]
    assert torch.equal(clip_ref.tokenize(prompts), ids)


def test_product_tokenizer_and_synthetic_match_oracle():
    clip = importlib.import_module("menghini-neurips23-code_b200.clip")
    texts = ["X X X X annual crop land", "a photo of a {}sea lake", "River"]
    with pytest.warns(UserWarning):
        a = clip.tokenize(texts)
    assert torch.equal(a, clip_ref.tokenize(texts))


def test_path_ranks():
    U = importlib.import_module("menghini-neurips23-code_b200.utils")
    r = U.path_ranks(["b/2.png", "a/9.png", "b/10.png", "a/9.png"])
    assert r.tolist() == [2, 0, 1, 0]


def test_text_prompt_strings_follow_reference():
    """CustomTextEncoder builds 'X X … X <class>' (models/clip_encoders.py:54-57); the placeholder
    rows are 1..P and EOT is the arg-max id."""
    M = importlib.import_module("menghini-neurips23-code_b200.models")
    enc = M.CustomTextEncoder.__new__(M.CustomTextEncoder)
    torch.nn.Module.__init__(enc)
    enc._ids_cache = {}
    ids = enc._prompt_ids(4, ["annual crop land", "sea"])
    assert ids.shape == (2, 77)
    assert ids[0, 1:5].unique().numel() == 1 and ids[0, 5] != ids[0, 1]
    assert ids.argmax(-1).tolist() == [8, 6]


def test_prompt_forward_restatements_match_reference_goldens(golden_dir):
    g = np.load(f"{golden_dir}/towers_vitb32_seed1234.npz")
    model = clip_ref.build_model(seed=1234)
    classes = [" ".join(c.split("_")) for c in synth.class_names(5, seed=1)]
    with torch.no_grad():
        t16 = prompt_ref.text_forward(model, synth.text_prefix(16), classes).numpy()
        v4 = prompt_ref.image_forward(model, synth.images(2, seed=0), synth.image_prefix(4)).numpy()
    assert np.abs(t16 - g["txt_feat_p16"]).max() < 1e-4
    assert np.abs(v4 - g["img_feat_p4"]).max() < 1e-4
    loss, grad, _ = prompt_ref.coop_step(model, synth.text_prefix(16), classes, synth.images(4, seed=0),
                                         torch.tensor([0, 1, 2, 3]))
    assert abs(float(loss) - float(g["coop_loss"])) < 1e-4
    assert np.abs(grad.numpy() - g["coop_grad_prefix"]).max() < 1e-4


def test_upt_restatement_matches_reference_golden(golden_dir):
    g = np.load(f"{golden_dir}/towers_vitb32_seed1234.npz")
    model = clip_ref.build_model(seed=1234)
    classes = [" ".join(c.split("_")) for c in synth.class_names(5, seed=1)]
    head = prompt_ref.UPTHead(synth.text_prefix(4, seed=2), synth.image_prefix(4, seed=3)[None])
    sd = {k[len("upt_sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("upt_sd.")}
    head.load_state_dict(sd, strict=False)
    with torch.no_grad():
        t, v = prompt_ref.upt_forward(model, head, synth.images(2, seed=0), classes)
    assert np.abs(t.numpy() - g["upt_text"]).max() < 1e-4
    assert np.abs(v.numpy() - g["upt_visual"]).max() < 1e-4
