"""ONE `assign_pseudo_labels` for the nine copies in the reference
(methods/{semi_supervised_learning,transductive_zsl,unsupervised_learning}/{textual,visual,multimodal}_fpl.py, e.g.
semi_supervised_learning/textual_fpl.py:195-283, visual_fpl.py:239-328, multimodal_fpl.py:208-285).

The copies differ in three things only: which tower carries the learned prompts (config.MODALITY), which class
list is scored (`self.classes` under unsupervised_learning, `self.unseen_classes` elsewhere), and nothing else —
all of them run a batch-1 loop `encode → logits → softmax → argmax(logits) → leaderboard` over the pool (the
multimodal one re-running BOTH towers per image, multimodal_fpl.py:223).  Here the prompts are encoded once,
the pool goes through the image tower in batches (with the visual prompt rows where the strategy has them) and
similarity → soft-max → arg-max(logits) → exact leaderboard is the fused device scan (`gb_pseudolabel_scan`,
mode 1).  `install()` swaps it into every strategy class that is importable; the method signature, the
mutation of `unlabeled_data` (filepaths / labels / label_id = True) and the return value are the reference's.
"""
from __future__ import annotations

import importlib
import logging

import torch

from .. import clip as _clip
from ..utils.clip_pseudolabels import encode_pool, scan_features

log = logging.getLogger(__name__)

PARADIGMS = ("semi_supervised_learning", "transductive_zsl", "unsupervised_learning")
STRATEGIES = (("textual_fpl", "TextualFPL"), ("visual_fpl", "VisualFPL"), ("multimodal_fpl", "MultimodalFPL"))


def _unwrapped(model):
    return model.module if hasattr(model, "module") and not hasattr(model, "prefix") and not hasattr(
        model, "coop_embeddings") else model


def assign_pseudo_labels(self, k, unlabeled_data):
    """Drop-in body for `<Strategy>.assign_pseudo_labels(k, unlabeled_data)`."""
    paradigm_ul = ".unsupervised_learning." in (type(self).__module__ + ".")
    classes = self.classes if paradigm_ul else self.unseen_classes
    modality = self.config.MODALITY
    clip_model = self.clip_model
    eng = clip_model.engine
    log.info(f"[self.assign_pseudo_labels] Number of prompts: {len(classes)}")
    with torch.no_grad():
        prefix = None
        if modality == "text":                               # textual_fpl.py:203-205
            self.model.classes = classes
            text_features = self.model(self.model.classes)
        elif modality == "image":                            # visual_fpl.py:244-251
            prompts = [self.template.format(" ".join(i.split("_"))) for i in classes]
            text_features = clip_model.encode_text(_clip.tokenize(prompts).to(self.device))
            prefix = _unwrapped(self.model).prefix
        else:                                                # multimodal_fpl.py:218-226 — prompts do not depend on the image
            upt = _unwrapped(self.model)
            coop_embs, vpt_embs = upt.prompt_embeddings()
            text_features = upt.text_encoder(coop_embs, classes)
            prefix = vpt_embs
        tf = text_features.detach().float()
        protos = (tf / tf.norm(dim=-1, keepdim=True)).half().to(eng.device).contiguous()
        feats = encode_pool(clip_model, unlabeled_data.filepaths, self.transform, eng.device, prefix=prefix)
        class_ids = [self.label_to_idx[c] for c in classes]
        idx, labels = scan_features(eng, feats, protos, k, unlabeled_data.filepaths, class_ids, mode=1)
    unlabeled_data.filepaths = [unlabeled_data.filepaths[i] for i in idx]
    unlabeled_data.labels = labels
    unlabeled_data.label_id = True
    return unlabeled_data


def install():
    """Patches every importable strategy class; returns the list of patched 'module.Class' names."""
    done = []
    for par in PARADIGMS:
        for mod, cls in STRATEGIES:
            try:
                m = importlib.import_module(f"methods.{par}.{mod}")
            except Exception:  # the reference tree is not on sys.path (or lacks a dependency): nothing to patch
                continue
            klass = getattr(m, cls, None)
            if klass is not None:
                klass.assign_pseudo_labels = assign_pseudo_labels
                done.append(f"methods.{par}.{mod}.{cls}")
    return done
