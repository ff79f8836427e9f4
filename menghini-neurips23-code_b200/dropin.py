"""One-call installation on the reference's import seam.

    import importlib; importlib.import_module("menghini-neurips23-code_b200.dropin").install()

After install(), code written against the reference resolves
    import clip / from clip import clip / clip.model.Transformer      → the B200 `clip` stand-in
    from models import CustomTextEncoder, …, UPTModel                  → the B200 classes
    utils.pseudolabel_top_k / utils.clip_pseudolabels.compute_pseudo_labels
                                                                        → the fused pool scan
    from accelerate import Accelerator                                  → accelerate_shim (only if the real package is absent)
    methods.<paradigm>.training_strategies.TrainingStrategy            → the re-created base class (the file is
                                                                          missing from the reference)
    <Strategy>FPL.assign_pseudo_labels (nine copies)                    → methods.pseudolabels.assign_pseudo_labels
The reference's own `utils` package is only patched when it is importable (it pulls pandas / scipy /
accelerate); otherwise a minimal `utils` exposing the hot functions is registered.
"""
from __future__ import annotations

import importlib
import sys
import types

_PKG = __name__.rsplit(".", 1)[0]


def install(patch_reference_utils: bool = True, patch_strategies: bool = True):
    clip = importlib.import_module(_PKG + ".clip")
    models = importlib.import_module(_PKG + ".models")
    utils_b200 = importlib.import_module(_PKG + ".utils")
    sys.modules["clip"] = clip
    sys.modules["clip.clip"] = clip
    sys.modules["clip.model"] = importlib.import_module(_PKG + ".clip.model")
    sys.modules["models"] = models
    try:
        importlib.import_module("accelerate")
    except Exception:
        shim = importlib.import_module(_PKG + ".accelerate_shim")
        sys.modules["accelerate"] = shim
    ref_utils = None
    if patch_reference_utils:
        try:
            ref_utils = importlib.import_module("utils")
            if getattr(ref_utils, "__name__", "") != "utils" or not hasattr(ref_utils, "pseudolabel_top_k"):
                ref_utils = None
        except Exception:
            ref_utils = None
    if ref_utils is not None:
        ref_utils.pseudolabel_top_k = utils_b200.pseudolabel_top_k
        sub = sys.modules.get("utils.clip_pseudolabels")
        if sub is not None:
            sub.pseudolabel_top_k = utils_b200.pseudolabel_top_k
            sub.compute_pseudo_labels = utils_b200.compute_pseudo_labels
        # utils/schedulers.py:50-52 passes verbose=True to LambdaLR, which torch ≥ 2.7 rejects
        ref_utils.make_scheduler = utils_b200.make_scheduler
        sched = sys.modules.get("utils.schedulers")
        if sched is not None:
            sched.make_scheduler = utils_b200.make_scheduler
            sched.WarmupCosineSchedule = utils_b200.WarmupCosineSchedule
    else:
        shim = types.ModuleType("utils")
        for name in utils_b200.__all__:
            setattr(shim, name, getattr(utils_b200, name))
        shim.clip_pseudolabels = importlib.import_module(_PKG + ".utils.clip_pseudolabels")
        sys.modules["utils"] = shim
        sys.modules["utils.clip_pseudolabels"] = shim.clip_pseudolabels
    out = {"clip": clip, "models": models, "utils": sys.modules["utils"], "accelerate": sys.modules["accelerate"]}
    if patch_strategies:
        ts = importlib.import_module(_PKG + ".methods.training_strategies")
        pl = importlib.import_module(_PKG + ".methods.pseudolabels")
        for par in pl.PARADIGMS:   # found in sys.modules before the reference's package looks for the file
            sys.modules.setdefault(f"methods.{par}.training_strategies", ts)
        out["training_strategies"] = ts
        out["assign_pseudo_labels"] = pl.install()
    return out
