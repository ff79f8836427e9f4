"""CoOp training step with the glue on the device.

The reference's loop (methods/semi_supervised_learning/textual_prompt.py:63-159) re-encodes every image batch
with the frozen tower on every epoch (:99-103), re-tokenises the prompts on every batch
(models/clip_encoders.py:54-60), and leaves normalisation, logits, cross-entropy, backward and the SGD update to
~40 small framework kernels per step.  `CoOpStep` keeps the step on the device: frozen image features come from
a cache (`utils.encode_pool`, exact because the reference passes no augmentations, main_SSL.py:152-153), prompt ids
are cached per class list, and text tower → `gb_ce_text_grad` → prompt-only backward → `gb_sgd_step` run back to
back on the caller's stream.  The learning rate follows utils/schedulers.py:36-65, stepped per epoch
(`update_scheduler`, textual_prompt.py:152).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import dist as _dist


class CoOpStep:
    """`momentum` is the optimizer's (the reference's base trainer file is missing from the scrape, SURVEY F3;
    torch.optim.SGD with the YAML's LR / DECAY is what its call sites imply) — pass what your trainer uses."""

    def __init__(self, text_prefix_model, lr: float, weight_decay: float = 0.0, momentum: float = 0.0,
                 warmup_epochs: int = 0, epochs: int = 1, world: int = 1):
        self.model = text_prefix_model
        self.enc = text_prefix_model.text_encoder
        self.engine = self.enc.clip_model.engine
        self.base_lr, self.wd, self.mu = float(lr), float(weight_decay), float(momentum)
        self.warmup, self.epochs, self.epoch = int(warmup_epochs), int(epochs), 0
        self.world = world
        self.buf = torch.zeros_like(self.model.prefix.data, dtype=torch.float32)
        self.steps = 0

    @property
    def lr(self) -> float:
        return self.engine.warmup_cosine_lr(self.base_lr, self.warmup, self.epochs, self.epoch)

    def update_scheduler(self):
        self.epoch += 1

    def step(self, imfn16: torch.Tensor, labels: torch.Tensor, coef: Optional[torch.Tensor] = None,
             classes=None, want_pred: bool = False):
        """One optimisation step on a batch of cached unit image features.  Returns (loss [1] on the device,
        pred | None); nothing is read back to the host."""
        eng = self.engine
        classes = self.model.classes if classes is None else classes
        prefix = self.model.prefix
        P = prefix.shape[1]
        ids = self.enc._prompt_ids(P, classes)
        with torch.no_grad():
            text, _, saved = eng.text_forward(ids, prefix[0], tape=True)
            loss, dtext, pred = eng.ce_text_grad(imfn16, text, labels, coef, want_pred=want_pred)
            dprefix = eng.text_backward_prefix(dtext, P, saved)
            if self.world > 1:
                _dist.allreduce_mean_(dprefix)
            eng.sgd_step(prefix.data.view(-1), dprefix.view(-1), self.buf.view(-1), self.lr, self.mu, self.wd,
                         first_step=self.steps == 0)
        self.steps += 1
        return loss, pred
