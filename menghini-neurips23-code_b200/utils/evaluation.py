"""Evaluation path of the reference's prompt strategies on B200.

`test_predictions` / `evaluation` in every methods/*/*_prompt.py (e.g.
methods/semi_supervised_learning/textual_prompt.py:226-356) loop over a DataLoader, encode a batch, L2-normalise,
form `logit_scale.exp() * image_features @ text_features.t()`, take `argmax(dim=1)`, map indices to class names,
look every file name up with `test_files.index(img)` (quadratic in the pool) and build a DataFrame
{"id", "class"} with duplicates dropped (:256-294).  Here the pool is encoded in batches by the image tower and
similarity + arg-max run as the fused HBM pass (`gb_sim_softmax_argmax`, mode 1 = arg-max over the logits, first
index on ties like torch.argmax); the bookkeeping is linear.  Accuracy / harmonic mean stay the reference's own
`utils.compute_metrics.evaluate_predictions`, which consumes the DataFrame unchanged.
"""
from __future__ import annotations

import torch

from .clip_pseudolabels import encode_pool


def predict_features(engine, feats16: torch.Tensor, protos16: torch.Tensor, scale=None, want_logits: bool = False):
    """arg-max class index per row of unit-norm fp16 features [N,512] against unit-norm fp16 prompts [C,512]
    (int32 [N]); optionally also the probabilities the fused pass produced (fp32 [N,C])."""
    pred, _, probs = engine.sim_softmax_argmax(feats16, protos16, scale, mode=1, want_probs=want_logits)
    return (pred, probs) if want_logits else pred


def predictions_frame(filepaths, pred_idx, class_names):
    """The DataFrame `test_predictions` returns (:289-294): one row per distinct (file name, class)."""
    import pandas as pd

    ids = [f.split("/")[-1] for f in filepaths]
    pred = pred_idx.cpu().tolist() if torch.is_tensor(pred_idx) else list(pred_idx)
    df = pd.DataFrame({"id": ids, "class": [class_names[i] for i in pred]})
    df.drop_duplicates(subset=["id", "class"], inplace=True)
    return df


def test_predictions(clip_model, text_features, dataset, transform, class_names, loader=None, batch=None):
    """Batched replacement of the per-strategy `test_predictions` bodies: `text_features` are the caller's prompts
    [C,512] (any float dtype, normalised or not — they are normalised here as at :250-251)."""
    eng = clip_model.engine
    with torch.no_grad():
        tf = text_features.detach().float()
        protos = (tf / tf.norm(dim=-1, keepdim=True)).half().to(eng.device).contiguous()
        feats = encode_pool(clip_model, dataset.filepaths, transform, eng.device, batch=batch, loader=loader)
        pred = predict_features(eng, feats, protos)
    return predictions_frame(dataset.filepaths, pred, class_names)


test_predictions.__test__ = False  # not a pytest test
