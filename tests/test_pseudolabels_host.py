"""Host logic of the ONE assign_pseudo_labels that replaces the nine copies of the reference (methods/pseudolabels.py;
reference: methods/{semi_supervised_learning,transductive_zsl,unsupervised_learning}/{textual,visual,multimodal}_fpl.py,
e.g. semi_supervised_learning/textual_fpl.py:195-283): which class list is scored per paradigm, which tower carries the
learned prompts per modality, the arg-max(logits) mode, and the mutation of the dataset — with the device calls stubbed."""
import importlib
import types

import pytest
import torch

PL = importlib.import_module("menghini-neurips23-code_b200.methods.pseudolabels")


class _Eng:
    device = torch.device("cpu")
    logit_scale_exp = 100.0


class _Clip:
    engine = _Eng()

    def encode_text(self, ids):
        self.text_calls = getattr(self, "text_calls", 0) + 1
        return torch.eye(ids.shape[0], 512)


def _strategy(module_name, modality, classes, unseen):
    s = type("TextualFPL", (), {"__module__": module_name})()
    s.classes, s.unseen_classes = classes, unseen
    s.config = types.SimpleNamespace(MODALITY=modality)
    s.clip_model = _Clip()
    s.device = "cpu"
    s.template = "a photo of a {}"
    s.transform = object()
    s.label_to_idx = {c: 100 + i for i, c in enumerate(classes)}
    if modality == "text":
        s.model = type("M", (), {"classes": None, "prefix": 0, "__call__": lambda self, cl: torch.eye(len(cl), 512)})()
    elif modality == "image":
        s.model = types.SimpleNamespace(prefix=torch.nn.Parameter(torch.zeros(16, 768)))
    else:
        s.model = types.SimpleNamespace(coop_embeddings=0,
                                        prompt_embeddings=lambda: (torch.zeros(1, 4, 512), torch.ones(1, 4, 768)),
                                        text_encoder=lambda emb, cl: torch.eye(len(cl), 512))
    return s


@pytest.mark.parametrize("module_name,expect_all", [("methods.semi_supervised_learning.textual_fpl", False),
                                                    ("methods.transductive_zsl.visual_fpl", False),
                                                    ("methods.unsupervised_learning.multimodal_fpl", True)])
@pytest.mark.parametrize("modality", ["text", "image", "multi"])
def test_assign_pseudo_labels_host_logic(monkeypatch, module_name, expect_all, modality):
    classes = ["a_b", "c", "d_e_f", "g"]
    unseen = ["c", "g"]
    s = _strategy(module_name, modality, classes, unseen)
    seen = {}

    def fake_encode_pool(clip_model, filepaths, transform, device, prefix=None, **kw):
        seen["prefix"] = prefix
        seen["paths"] = list(filepaths)
        return torch.zeros(len(filepaths), 512, dtype=torch.float16)

    def fake_scan(engine, feats, protos, k, paths, class_ids, mode=0, **kw):
        seen.update(k=k, class_ids=list(class_ids), mode=mode, protos=protos)
        return [2, 0], [class_ids[-1], class_ids[0]]

    monkeypatch.setattr(PL, "encode_pool", fake_encode_pool)
    monkeypatch.setattr(PL, "scan_features", fake_scan)
    monkeypatch.setattr(PL._clip, "tokenize", lambda prompts: torch.zeros(len(prompts), 77, dtype=torch.long))
    ds = types.SimpleNamespace(filepaths=["p0", "p1", "p2"], labels=[9, 9, 9], label_id=False)
    out = PL.assign_pseudo_labels(s, 7, ds)
    scored = classes if expect_all else unseen                      # self.classes under UL, self.unseen_classes elsewhere
    assert seen["class_ids"] == [s.label_to_idx[c] for c in scored]
    assert seen["k"] == 7 and seen["mode"] == 1                     # arg-max of the LOGITS (textual_fpl.py:228)
    assert seen["protos"].dtype == torch.float16 and tuple(seen["protos"].shape) == (len(scored), 512)
    assert torch.allclose(seen["protos"].float().norm(dim=-1), torch.ones(len(scored)), atol=1e-3)
    if modality == "text":
        assert seen["prefix"] is None                               # frozen image tower
    elif modality == "image":
        assert seen["prefix"] is s.model.prefix                     # the learned visual prompt rows
    else:
        assert torch.equal(seen["prefix"], torch.ones(1, 4, 768))   # the coupled visual prompt of the UPT head
    assert out is ds and ds.filepaths == ["p2", "p0"] and ds.labels == [seen["class_ids"][-1], seen["class_ids"][0]]
    assert ds.label_id is True


def test_assign_pseudo_labels_unwraps_a_ddp_wrapped_model(monkeypatch):
    """accelerator.prepare wraps the trained module (`.module`); the reference reaches through it
    (visual_fpl.py:256 `self.model.module.prefix` under DDP) — the prompt rows must be the inner module's."""
    s = _strategy("methods.transductive_zsl.visual_fpl", "image", ["a", "b"], ["b"])
    inner = s.model
    s.model = types.SimpleNamespace(module=inner)
    seen = {}
    monkeypatch.setattr(PL, "encode_pool", lambda cm, fp, tf, dev, prefix=None, **kw: seen.setdefault("prefix", prefix) is None
                        or torch.zeros(len(fp), 512, dtype=torch.float16))
    monkeypatch.setattr(PL, "scan_features", lambda *a, **kw: ([0], [a[5][0]]))
    monkeypatch.setattr(PL._clip, "tokenize", lambda prompts: torch.zeros(len(prompts), 77, dtype=torch.long))
    ds = types.SimpleNamespace(filepaths=["p0"], labels=[0], label_id=False)
    PL.assign_pseudo_labels(s, 1, ds)
    assert seen["prefix"] is inner.prefix and ds.labels == [101]


def test_install_without_the_reference_tree_patches_nothing():
    import sys
    if any(m.startswith("methods.") and "_fpl" in m for m in sys.modules):
        pytest.skip("reference strategies already imported in this process")
    if importlib.util.find_spec("methods") is not None:
        pytest.skip("a `methods` package is importable here")
    assert PL.install() == []


def test_prob_dtype_reads_and_validates_the_environment(monkeypatch):
    CP = importlib.import_module("menghini-neurips23-code_b200.utils.clip_pseudolabels")
    monkeypatch.delenv("GRIPB200_PROB_DTYPE", raising=False)
    assert CP.prob_dtype() == "fp32"                          # the reference's CPU arithmetic is the default
    monkeypatch.setenv("GRIPB200_PROB_DTYPE", "FP16")
    assert CP.prob_dtype() == "fp16"                          # the reference's CUDA arithmetic, on request
    monkeypatch.setenv("GRIPB200_PROB_DTYPE", "bf16")
    with pytest.raises(ValueError):
        CP.prob_dtype()


def test_pool_cache_key_follows_the_files(tmp_path, monkeypatch):
    """The pool-feature cache (GRIP re-encodes the same pool every iteration under the frozen tower,
    methods/semi_supervised_learning/pseudo_iterative.py:62-75) is keyed by path, mtime and size of EVERY file."""
    import os
    CP = importlib.import_module("menghini-neurips23-code_b200.utils.clip_pseudolabels")
    monkeypatch.delenv("GRIPB200_POOL_CACHE", raising=False)
    files = []
    for i in range(3):
        f = tmp_path / f"{i}.png"
        f.write_bytes(b"x" * (10 + i))
        files.append(str(f))
    k0 = CP._pool_cache_key(files)
    assert k0 is not None and k0[0] == 3 and CP._pool_cache_key(list(files)) == k0
    assert CP._pool_cache_key(files[::-1]) != k0              # the order of the pool is part of the result
    (tmp_path / "1.png").write_bytes(b"y" * 40)               # a file rewritten → another pool
    assert CP._pool_cache_key(files) != k0
    assert CP._pool_cache_key(files + [str(tmp_path / "missing.png")]) is None   # unreadable → no caching, no error here
    monkeypatch.setenv("GRIPB200_POOL_CACHE", "0")
    assert CP._pool_cache_key(files) is None
