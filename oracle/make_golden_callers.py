"""Generates tests/golden/callers_seed0.npz by running the REFERENCE'S OWN caller classes (imported from
/root/reference, never copied) — TextualPrompt, VisualPrompt, MultimodalPrompt, TextualFPL
(methods/semi_supervised_learning/*.py) and ClipBaseline (methods/clip_baseline.py) — on the restated CPU `clip`
(oracle/clip_ref.py, fp32), the reference's own `models/`, `utils/`, `data/`, and the two pieces the scrape lacks:
the re-created `training_strategies` and the `accelerate` stand-in of the product package.

    python oracle/make_golden_callers.py [--reference /root/reference] [--out tests/golden] [--check]

It also runs oracle/callers_ref.py's restatements of the same classes on the same inputs and REQUIRES identical
results — that is what pins the restatements (which are all a GPU box has) to the reference.
--check re-runs and compares with the committed file instead of writing it (tests/test_reference_callers.py).
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
PKG = "menghini-neurips23-code_b200"


def install_cpu_seam(ref_root):
    """sys.modules as the reference expects them, on the CPU oracle: `clip` = oracle/clip_ref.py, `accelerate` =
    the stand-in, `models` / `utils` / `data` / `methods` = the reference's own packages, plus the re-created
    methods.<paradigm>.training_strategies."""
    from oracle import clip_ref

    top = clip_ref.install_as_clip()
    top.load = lambda name="ViT-B/32", device="cpu", **kw: (clip_ref.build_model(seed=1234), clip_ref.clip_transform())
    sys.modules["accelerate"] = importlib.import_module(PKG + ".accelerate_shim")
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import utils.schedulers as ref_sched   # the reference's own utils package

    fixed = importlib.import_module(PKG + ".utils.schedulers")
    ref_sched.WarmupCosineSchedule = fixed.WarmupCosineSchedule   # verbose=True is rejected by torch ≥ 2.7
    ts = importlib.import_module(PKG + ".methods.training_strategies")
    for par in ("semi_supervised_learning", "transductive_zsl", "unsupervised_learning"):
        sys.modules.setdefault(f"methods.{par}.training_strategies", ts)
    return ts


def reference_strategies():
    import types

    ssl = importlib.import_module("methods.semi_supervised_learning")
    cb = importlib.import_module("methods.clip_baseline")
    return types.SimpleNamespace(TextualPrompt=ssl.TextualPrompt, VisualPrompt=ssl.VisualPrompt,
                                 MultimodalPrompt=ssl.MultimodalPrompt, TextualFPL=ssl.TextualFPL,
                                 ClipBaseline=cb.ClipBaseline)


def run(ref_root, which):
    from oracle import callers_ref

    ts = install_cpu_seam(ref_root)
    from utils import dataset_object

    torch.set_num_threads(os.cpu_count() or 1)
    results = {}
    for name, S in (("reference", reference_strategies()),
                    ("restated", callers_ref.build_ref_strategies(ts.TrainingStrategy))):
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)
            try:
                for d in ("pseudolabels", "trained_prompts", "evaluation", "logs"):
                    os.makedirs(d)
                results[name] = callers_ref.run_all(S, dataset_object("EuroSAT"), os.path.join(tmp, "data"), "cpu", which)
            finally:
                os.chdir(cwd)
    a, b = results["reference"], results["restated"]
    assert a.keys() == b.keys()
    for k in a:
        same = np.array_equal(a[k], b[k]) if a[k].dtype.kind in "US" else np.allclose(a[k], b[k], rtol=0, atol=0)
        assert same, f"restated caller differs from the reference's own class at {k}"
    return a


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--which", default="clip,textual,visual,multimodal,fpl")
    a = ap.parse_args()
    res = run(a.reference, tuple(a.which.split(",")))
    path = os.path.join(a.out, "callers_seed0.npz")
    np.savez_compressed(path, **res)
    print(f"wrote {path}: {sorted(res)}")
