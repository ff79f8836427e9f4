"""SURVEY §8f N4 — on-disk artifacts interchange with the reference, both ways (CPU; needs /root/reference mounted).

The reference's OWN writers (utils/compute_metrics.py:58-171, executed from /root/reference, never copied) are fed
with what THIS package's classes hold after training, and the files they leave are read back into the REFERENCE's own
model classes (models/prompts_models.py on the restated clip) — and the other way round:
  * prompt pickle  `trained_prompts/{…}.pickle`  = `[ndarray]`            (save_parameters, text / image modality)
  * UPT artifacts  5 `torch.save`d state_dicts + 3 pickles                 (save_parameters, MODALITY == 'multi';
                   the parameter list is built exactly as multimodal_prompt.py:149-158 builds it)
  * pseudolabel cache `{"filepaths": [...], "labels": [...]}`              (save_pseudo_labels; the scan's own cache is
                   covered by test_pseudolabel_top_k_end_to_end)
  * predictions pickle, results JSONL                                      (save_predictions, store_results fed with
                   utils.predictions_frame / evaluate_predictions output)
"""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import importlib, json, os, pickle, sys, types
import numpy as np, torch
ROOT, REF, TMP = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"
sys.modules["accelerate"] = importlib.import_module(PKG + ".accelerate_shim")
from oracle import clip_ref, synth
clip_ref.install_as_clip()                       # `import clip` of the reference's modules → the CPU restatement
sys.path.insert(0, REF)
ref_metrics = importlib.import_module("utils.compute_metrics")       # the reference's writers, unmodified
ref_pm = importlib.import_module("models.prompts_models")            # the reference's model classes
b200_pm = importlib.import_module(PKG + ".models.prompts_models")
b200_utils = importlib.import_module(PKG + ".utils")
os.chdir(TMP)
for d in ("trained_prompts", "pseudolabels", "evaluation"):
    os.makedirs(d)


class Cfg:
    DATASET_NAME, LEARNING_PARADIGM, VIS_ENCODER, OPTIM_SEED, SPLIT_SEED = "EuroSAT", "ssl", "ViT-B/32", 1, 500


def cfg(model, modality):
    c = Cfg()
    c.MODEL, c.MODALITY = model, modality
    return c


enc = lambda *a, **k: None   # the encoders are not called here: only what the models OWN is written
classes = ["a", "b", "c"]

# ---- text / image prompts: [ndarray] pickle (textual_prompt.py:154-157 builds the list) ----
for model_name, modality, mk_b200, mk_ref, init in (
        ("textual_prompt", "text", b200_pm.TextPrefixModel, ref_pm.TextPrefixModel, synth.text_prefix(16)),
        ("visual_prompt", "image", lambda p, e, c: b200_pm.ImagePrefixModel(p, e), lambda p, e, c: ref_pm.ImagePrefixModel(p, e),
         synth.image_prefix(16))):
    m = mk_b200(init.clone(), enc, classes)
    with torch.no_grad():
        m.prefix.add_(0.25)                                         # "training"
    c = cfg(model_name, modality)
    ref_metrics.save_parameters([m.prefix.detach().cpu().numpy()], c)
    fn = f"trained_prompts/EuroSAT_ssl_{model_name}_ViT-B32_opt_1_spl_500.pickle"
    got = pickle.load(open(fn, "rb"))
    assert isinstance(got, list) and len(got) == 1 and got[0].dtype == np.float32
    r = mk_ref(torch.from_numpy(got[0]), enc, classes)              # the reference's class takes it as its prefix
    assert torch.equal(r.prefix.detach(), m.prefix.detach())
    ref_metrics.save_parameters([m.prefix.detach().cpu().numpy()], c, iteration=3)   # iterative strategies
    assert os.path.exists(f"trained_prompts/EuroSAT_ssl_{model_name}_ViT-B32_iter_3_opt_1_spl_500.pickle")

# ---- UPT: the list of multimodal_prompt.py:149-158, written by the reference, read into the reference's UPTModel ----
torch.manual_seed(4)
mine = b200_pm.UPTModel(synth.text_prefix(4, seed=2), synth.image_prefix(4, seed=3)[None], None, enc, enc, classes, 128,
                        device="cpu", dtype=torch.float32)
params = [mine.transformer.state_dict(), mine.proj_coop_pre.state_dict(), mine.proj_coop_post.state_dict(),
          mine.proj_vpt_pre.state_dict(), mine.proj_vpt_post.state_dict(),
          mine.coop_embeddings.detach().cpu().numpy(), None, mine.vpt_embeddings.detach().cpu().numpy()]
c = cfg("multimodal_prompt", "multi")
ref_metrics.save_parameters(params, c)
base = "trained_prompts/EuroSAT_ssl_multimodal_prompt_ViT-B32_opt_1_spl_500"
torch.manual_seed(99)
theirs = ref_pm.UPTModel(torch.from_numpy(pickle.load(open(base + "_coop_embeddings.pickle", "rb"))),
                         torch.from_numpy(pickle.load(open(base + "_vpt_embeddings.pickle", "rb"))),
                         pickle.load(open(base + "_deep_vpt.pickle", "rb")), enc, enc, classes, 128, device="cpu",
                         dtype=torch.float32)
for name in ("transformer", "proj_coop_pre", "proj_coop_post", "proj_vpt_pre", "proj_vpt_post"):
    sd = torch.load(f"{base}_{name}.pt")
    getattr(theirs, name).load_state_dict(sd, strict=True)          # same keys, same shapes
    back = getattr(theirs, name).state_dict()
    assert list(back) == list(getattr(mine, name).state_dict())
    for k in back:
        assert torch.equal(back[k], getattr(mine, name).state_dict()[k]), (name, k)
# … and the coupled prompts the two towers would receive are the same numbers from either class
with torch.no_grad():
    coop_m, vpt_m = mine.prompt_embeddings()

    class Spy:
        def __init__(self): self.got = None
        def __call__(self, *a): self.got = a; return torch.zeros(1)
    ts, vs = Spy(), Spy()
    theirs.text_encoder, theirs.image_encoder = ts, vs
    theirs(torch.zeros(1, 3, 224, 224), classes)
assert torch.allclose(ts.got[0].float(), coop_m.float(), atol=0, rtol=0)
assert torch.allclose(vs.got[1].float(), vpt_m.float(), atol=0, rtol=0)
# the other direction: the reference's freshly initialised head loads into this package's class
torch.manual_seed(7)
theirs2 = ref_pm.UPTModel(synth.text_prefix(4, seed=2), synth.image_prefix(4, seed=3)[None], None, enc, enc, classes, 128,
                          device="cpu", dtype=torch.float32)
for name in ("transformer", "proj_coop_pre", "proj_coop_post", "proj_vpt_pre", "proj_vpt_post"):
    getattr(mine, name).load_state_dict(getattr(theirs2, name).state_dict(), strict=True)

# ---- pseudolabel file of the iterative strategies, predictions pickle, results JSONL ----
ref_metrics.save_pseudo_labels(["x/1.png", "x/2.png"], [3, 1], cfg("grip_textual", "text"), 2)
d = pickle.load(open("pseudolabels/EuroSAT_ssl_grip_textual_ViT-B32_iter_2_opt_1_spl_500.pickle", "rb"))
assert d == {"filepaths": ["x/1.png", "x/2.png"], "labels": [3, 1]}
names = ["forest", "river", "lake"]
df = b200_utils.predictions_frame(["/p/a.png", "/p/b.png", "/p/c.png"], torch.tensor([2, 0, 1]), names)
ref_metrics.save_predictions(df, cfg("textual_prompt", "text"))
back = pickle.load(open("evaluation/EuroSAT_ssl_textual_prompt_ViT-B32_opt_1_spl_500.pickle", "rb"))
assert list(back.columns) == ["id", "class"] and back["class"].tolist() == ["lake", "forest", "river"]
c = cfg("textual_prompt", "text")
ref_metrics.store_results(c, (0.5,))
ref_metrics.store_results(c, (0.75,))
lines = [json.loads(l) for l in open("results_model_textual_prompt.json")]
assert [l["accuracy"] for l in lines] == [0.5, 0.75] and lines[0]["model"] == "textual_prompt"
print("formats ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "utils")), reason="the reference tree is not mounted")
def test_artifacts_interchange_with_the_reference(tmp_path):
    script = tmp_path / "interop.py"
    script.write_text(textwrap.dedent(_SCRIPT))
    work = tmp_path / "work"
    work.mkdir()
    r = subprocess.run([sys.executable, str(script), ROOT, REF, str(work)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "formats ok" in r.stdout
