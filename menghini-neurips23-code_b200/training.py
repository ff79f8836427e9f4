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


def fpl_coefficients(second_group: torch.Tensor, w_first: float = 1.0, w_second: float = 1.0) -> torch.Tensor:
    """Per-sample weights that turn the FPL strategies' two-term loss into one weighted sum.

    The reference computes `w_first * CE(batch[first]) + w_second * CE(batch[second])`, each CE the MEAN over its
    group and an empty group contributing 0 — SSL: first = labeled (not in `check_unlabeled`), w_first =
    `balance_param`, second = pseudolabeled, w_second = 1 (methods/semi_supervised_learning/textual_fpl.py:123-165);
    TRZSL: first = seen-class samples, w_first = 1, second = unseen-class samples, w_second = `balance_param`
    (methods/transductive_zsl/textual_fpl.py:117-147).  `second_group` is the boolean membership mask of the batch;
    the result feeds `CoOpStep.step(..., coef=...)` / `gb_ce_text_grad`."""
    second = second_group.bool()
    n2 = int(second.sum().item())
    n1 = int(second.numel()) - n2
    coef = torch.zeros(second.shape, dtype=torch.float32, device=second.device)
    if n1:
        coef[~second] = float(w_first) / n1
    if n2:
        coef[second] = float(w_second) / n2
    return coef


class CoOpStep:
    """`momentum` is the optimizer's (the reference's base trainer file is missing from the scrape, SURVEY F3;
    torch.optim.SGD with the YAML's LR / DECAY is what its call sites imply) — pass what your trainer uses.

    With `graph=True` (single process) the whole chain — text tower with tape, loss + gradient, prompt-only
    backward, SGD — is captured once per (class list, batch size) as a CUDA graph and replayed: the ~160 small
    launches of a step cost one submission.  The learning rate lives in a device scalar, so the captured graph
    follows the schedule; the graph is re-captured when the library re-allocates a scratch arena."""

    def __init__(self, text_prefix_model, lr: float, weight_decay: float = 0.0, momentum: float = 0.0,
                 warmup_epochs: int = 0, epochs: int = 1, world: int = 1, graph: bool = True):
        self.model = text_prefix_model
        self.enc = text_prefix_model.text_encoder
        self.engine = self.enc.clip_model.engine
        self.base_lr, self.wd, self.mu = float(lr), float(weight_decay), float(momentum)
        self.warmup, self.epochs, self.epoch = int(warmup_epochs), int(epochs), 0
        self.world = world
        self.use_graph = bool(graph) and world == 1
        dev = self.engine.device
        self.buf = torch.zeros_like(self.model.prefix.data, dtype=torch.float32)  # zero ⇒ first step: buf = g
        self.lr_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self.steps = 0
        self._key = None
        self._graph = None
        self._gen = None

    @property
    def lr(self) -> float:
        return self.engine.warmup_cosine_lr(self.base_lr, self.warmup, self.epochs, self.epoch)

    def update_scheduler(self):
        self.epoch += 1

    # static buffers of one (classes, batch) configuration
    def _prepare(self, classes, B, weighted):
        eng, dev = self.engine, self.engine.device
        P = self.model.prefix.shape[1]
        ids = self.enc._prompt_ids(P, classes).detach().cpu()
        eot = ids.argmax(dim=-1)
        self._Lt = max(int(eot.max().item()) + 1, P + 2)
        self._ids = ids.to(dev, torch.int32).contiguous()
        self._eot = eot.to(dev, torch.int32).contiguous()
        C = ids.shape[0]
        self._C, self._B, self._P = C, B, P
        f32 = dict(device=dev, dtype=torch.float32)
        self._text = torch.empty(C, 512, **f32)
        self._tape = torch.empty(eng.tape_bytes(C, self._Lt, 512), device=dev, dtype=torch.uint8)
        self._imfn = torch.empty(B, 512, device=dev, dtype=torch.float16)
        self._labels = torch.empty(B, device=dev, dtype=torch.int32)
        self._coef = torch.empty(B, **f32) if weighted else None
        self._dtext = torch.empty(C, 512, **f32)
        self._loss = torch.empty(1, **f32)
        self._pred = torch.empty(B, device=dev, dtype=torch.int32)
        self._dprefix = torch.empty(P, 512, **f32)
        self._graph = None

    def _chain(self):
        from ._lib import ptr, stream_ptr

        eng = self.engine
        lib, h, chk, st = eng.lib, eng.ctx.h, eng.ctx.check, stream_ptr(eng.device)
        prefix = self.model.prefix.data
        chk(lib.gb_text_forward(h, ptr(self._ids), self._ids.stride(0), ptr(self._eot), ptr(prefix), self._C,
                                self._P, self._Lt, ptr(self._text), None, ptr(self._tape), st), "gb_text_forward")
        chk(lib.gb_ce_text_grad(h, ptr(self._imfn), ptr(self._text), ptr(self._labels), ptr(self._coef),
                                eng.logit_scale_exp, self._B, self._C, ptr(self._dtext), ptr(self._loss),
                                ptr(self._pred), st), "gb_ce_text_grad")
        chk(lib.gb_text_backward_prefix(h, ptr(self._dtext), ptr(self._eot), self._C, self._P, self._Lt,
                                        ptr(self._tape), ptr(self._dprefix), st), "gb_text_backward_prefix")
        if self.world > 1:
            _dist.allreduce_mean_(self._dprefix)
        chk(lib.gb_sgd_step(h, ptr(prefix), ptr(self._dprefix), ptr(self.buf), prefix.numel(), 0.0,
                            ptr(self.lr_dev), self.mu, self.wd, 0, st), "gb_sgd_step")

    def step(self, imfn16: torch.Tensor, labels: torch.Tensor, coef: Optional[torch.Tensor] = None,
             classes=None, want_pred: bool = False):
        """One optimisation step on a batch of cached unit image features.  Returns (loss [1] on the device,
        pred | None) — views of buffers the next step overwrites; nothing is read back to the host."""
        eng = self.engine
        classes = self.model.classes if classes is None else classes
        prefix = self.model.prefix
        if prefix.dtype != torch.float32 or not prefix.data.is_contiguous():
            raise ValueError("CoOpStep needs a contiguous fp32 prefix parameter")
        key = (tuple(classes), int(imfn16.shape[0]), coef is not None, prefix.data_ptr())
        if key != self._key:
            self._prepare(classes, int(imfn16.shape[0]), coef is not None)
            self._key = key
        with torch.no_grad():
            self._imfn.copy_(imfn16)
            self._labels.copy_(labels)
            if coef is not None:
                self._coef.copy_(coef)
            self.lr_dev.fill_(self.lr)
            eng._bind()
            gen = int(eng.lib.gb_workspace_generation(eng.ctx.h))
            if self._graph is not None and gen == self._gen:
                self._graph.replay()
            else:
                self._chain()   # eager: sizes the library's scratch arenas, and IS this step
                self._graph = None
                if self.use_graph:
                    self._gen = int(eng.lib.gb_workspace_generation(eng.ctx.h))
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):   # capture only: nothing executes here
                        self._chain()
                    self._graph = g
        self.steps += 1
        return self._loss, (self._pred if want_pred else None)


class VPTStep:
    """Visual prompt tuning step (methods/semi_supervised_learning/visual_prompt.py:115-145) with the glue on the
    device: image tower with tape → `gb_ce_image_grad` (cosine-logit CE + gradient w.r.t. the image features) →
    prompt-only backward → `gb_sgd_step`.  `text_features` are the frozen prompts of the epoch ([C,512], any
    scale: they are normalised on the device, :115-118); nothing is read back to the host."""

    def __init__(self, image_prefix_model, text_features: torch.Tensor, lr: float, weight_decay: float = 0.0,
                 momentum: float = 0.0, warmup_epochs: int = 0, epochs: int = 1, world: int = 1):
        self.model = image_prefix_model
        self.engine = image_prefix_model.image_encoder.visual._vt._owner().engine
        self.text = text_features.detach().to(self.engine.device, torch.float32).contiguous()
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

    def step(self, img: torch.Tensor, labels: torch.Tensor, coef: Optional[torch.Tensor] = None,
             want_pred: bool = False):
        """One optimisation step on a batch of images (fp32 / fp16 normalised or uint8 pixels).  Returns
        (loss [1], image features fp32 [B,512], pred | None), all on the device."""
        eng = self.engine
        prefix = self.model.prefix
        if prefix.dtype != torch.float32 or not prefix.data.is_contiguous():
            raise ValueError("VPTStep needs a contiguous fp32 prefix parameter")
        with torch.no_grad():
            p2 = prefix.data.reshape(-1, 768)
            feat, _, tape = eng.vit_forward(img, p2, tape=True)
            loss, dimage, _, pred = eng.ce_image_grad(feat, self.text, labels, coef, want_pred=want_pred)
            dprefix = eng.vit_backward_prefix(dimage, p2, tape)
            del tape
            if self.world > 1:
                _dist.allreduce_mean_(dprefix)
            eng.sgd_step(prefix.data, dprefix, self.buf, self.lr, self.mu, self.wd, first_step=False)
        self.steps += 1
        return loss, feat, pred


class UPTStep:
    """Unified prompt tuning step (methods/semi_supervised_learning/multimodal_prompt.py:103-127): the 0.5 M
    parameter coupling head stays a torch module (its graph is 2×4 tokens wide); both towers run with tape, the
    loss and BOTH feature gradients come from one `gb_ce_image_grad` call, the two prompt-only backward passes
    hand d loss / d coop_embs and d loss / d vpt_embs to the head's autograd graph, and `optimizer` (the caller's
    torch optimiser over the head and the two prompt tensors) takes the step."""

    def __init__(self, upt_model, optimizer, world: int = 1):
        self.model = upt_model
        self.opt = optimizer
        self.enc = upt_model.text_encoder
        self.engine = self.enc.clip_model.engine
        self.world = world

    def step(self, img, labels, classes=None, coef=None, want_pred: bool = False):
        eng = self.engine
        classes = self.model.classes if classes is None else classes
        coop_embs, vpt_embs = self.model.prompt_embeddings()
        Pt = coop_embs.shape[1]
        ids = self.enc._prompt_ids(Pt, classes)
        with torch.no_grad():
            c2 = coop_embs.detach().reshape(-1, 512).float()
            v2 = vpt_embs.detach().reshape(-1, 768).float()
            tfeat, _, tsaved = eng.text_forward(ids, c2, tape=True)
            ifeat, _, tape = eng.vit_forward(img, v2, tape=True)
            loss, dimage, dtext, pred = eng.ce_image_grad(ifeat, tfeat, labels, coef, want_dtext=True,
                                                          want_pred=want_pred)
            dcoop = eng.text_backward_prefix(dtext, Pt, tsaved)
            dvpt = eng.vit_backward_prefix(dimage, v2, tape)
            del tape, tsaved
        self.opt.zero_grad(set_to_none=True)
        torch.autograd.backward([coop_embs, vpt_embs],
                                [dcoop.reshape(coop_embs.shape).to(coop_embs.dtype),
                                 dvpt.reshape(vpt_embs.shape).to(vpt_embs.dtype)])
        if self.world > 1:
            for p in self.model.parameters():
                if p.grad is not None:
                    _dist.allreduce_mean_(p.grad)
        self.opt.step()
        return loss, (tfeat, ifeat), pred
