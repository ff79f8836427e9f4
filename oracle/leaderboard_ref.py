"""CPU oracle: the reference's sequential per-class pseudolabel "leaderboard".

TEST INFRASTRUCTURE ONLY (see oracle/clip_ref.py header for who may import oracle/).

Restates, on a precomputed probability matrix, the control flow of
  utils/clip_pseudolabels.py:46-112          (compute_pseudo_labels, zero-shot prompts)
  methods/semi_supervised_learning/textual_fpl.py:208-283 and its 8 siblings (assign_pseudo_labels)
which is NOT a top-k: boards fill in arrival order, the admission test looks at the LAST list entry
(strict `<`), a successful admission sorts (descending by (p, path)) and truncates to k, and an
image rejected by its own board is offered to every other board with its probability for that class
(no break).  Pinned against the reference's own function by oracle/make_golden.py (the real
compute_pseudo_labels is executed on PNG files with a stub clip_model) → tests/golden/leaderboard_*.

Inputs are numbers, not files: `probs[i, j]` is what the reference calls `probs[0][j]` for image i,
`pred[i]` its arg-max (`argmax(probs)` in clip_pseudolabels.py:63, `argmax(logits)` in
textual_fpl.py:228 — the caller decides), `paths[i]` the tie-break key (image path).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

ALL_UNLABELED_K = 10000000  # utils/clip_pseudolabels.py:27 — "label every image with its argmax"


class Boards:
    """Resumable form of the same state machine: feed() consecutive index ranges in order (what a
    shard does after receiving the state from the shard before it)."""

    def __init__(self, c: int, k: int, class_ids: Sequence[int] = None):
        self.c, self.k = c, k
        self.class_ids = list(range(c)) if class_ids is None else list(class_ids)
        self.boards = {}
        for j in range(c):                                  # dict semantics of :49-51
            self.boards[self.class_ids[j]] = []

    def feed(self, probs: np.ndarray, pred: Sequence[int], paths: Sequence, idx0: int = 0):
        """probs/pred: rows of this range; paths[i] tie-break key of GLOBAL image i."""
        k, boards, cid = self.k, self.boards, self.class_ids

        def key(t):  # sorted(..., reverse=True) on (prob, path) tuples
            return (t[0], t[1])

        for r in range(probs.shape[0]):
            i = idx0 + r
            row = probs[r]
            j_star = int(pred[r])
            own = boards[cid[j_star]]
            p = row[j_star]
            if len(own) < k:                                   # :73-74
                own.append((p, paths[i], i))
            elif own[-1][0] < p:                               # :75-82
                boards[cid[j_star]] = sorted(own + [(p, paths[i], i)], key=key, reverse=True)[:k]
            else:                                              # :83-101 (order over j is immaterial)
                for j in range(self.c):
                    if j == j_star:
                        continue
                    b = boards[cid[j]]
                    if len(b) < k:
                        b.append((row[j], paths[i], i))
                    elif b[-1][0] < row[j]:
                        boards[cid[j]] = sorted(b + [(row[j], paths[i], i)], key=key, reverse=True)[:k]
        return self

    def result(self) -> Tuple[List[int], List[int]]:
        out_idx, out_lab = [], []
        for cid, b in self.boards.items():                     # :103-109
            out_idx += [t[2] for t in b]
            out_lab += [cid for _ in b]
        return out_idx, out_lab


def leaderboard(probs: np.ndarray, pred: Sequence[int], k: int, paths: Sequence,
                class_ids: Sequence[int] = None) -> Tuple[List[int], List[int]]:
    """Returns (image indices, labels) in the order the reference rebuilds the dataset
    (utils/clip_pseudolabels.py:103-109): boards in class insertion order, each in its current
    list order.  `class_ids[j]` is `label_to_idx[classnames[j]]` (default j)."""
    n, c = probs.shape
    if class_ids is None:
        class_ids = list(range(c))
    if k == ALL_UNLABELED_K:  # :27-44
        return list(range(n)), [class_ids[int(pred[i])] for i in range(n)]
    return Boards(c, k, class_ids).feed(probs, pred, paths).result()


def softmax_argmax(feat: np.ndarray, proto: np.ndarray, scale: float, dtype=np.float32):
    """logits = scale * feat @ proto.T on L2-normalised rows, softmax, arg-max — the arithmetic of
    CLIP.forward + `.softmax(dim=-1)` + `argmax` (utils/clip_pseudolabels.py:59-65) in `dtype`."""
    f = feat.astype(dtype)
    t = proto.astype(dtype)
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    t = t / np.linalg.norm(t, axis=1, keepdims=True)
    logits = dtype(scale) * (f @ t.T)
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m)
    probs = e / e.sum(axis=1, keepdims=True)
    return logits, probs, probs.argmax(axis=1)
