"""`TrainingStrategy`: the base class every strategy of the reference derives from
(`from methods.semi_supervised_learning import TrainingStrategy`, textual_prompt.py:15 and siblings).  The file
is MISSING from the reference scrape (methods/semi_supervised_learning/__init__.py:1 imports it); this is a
re-creation from its call sites only (contract: SURVEY.md §3.5) — what the subclasses call, read and override:

    __init__(config, label_to_idx, classes, seen_classes, unseen_classes, device)   textual_prompt.py:44-46
    declare_custom_encoder / initialize_prompts_parameters / define_model           :56-61, 298
    define_loss_function(logits, labs[, paths]) / loss_func                         :125, textual_fpl.py:150
    backpropagate / update_scheduler / unwrap_model                                 :135, 152, 154
    train / fixed_iterative_train / grip_train → (best_val_accuracy, best_prompt)   main_SSL.py:210-396
    attributes: clip_model, transform, template, text_encoder, image_encoder, model, training_model,
                val_unseen_files, val_unseen_labs

One implementation serves the three paradigms (semi_supervised_learning, transductive_zsl,
unsupervised_learning): everything paradigm specific lives in the reference's own subclasses.
"""
from __future__ import annotations

import copy
import logging
import math
import sys

import numpy as np
import torch
from torch import nn

import clip
from accelerate import Accelerator
from models import (CustomImageEncoder, CustomTextEncoder, ImagePrefixModel, TextPrefixModel, UPTModel)
from utils import make_scheduler, seed_worker

accelerator = Accelerator()
log = logging.getLogger(__name__)


class TrainingStrategy(object):
    def __init__(self, config, label_to_idx, classes, seen_classes, unseen_classes, device):
        self.config = config
        self.classes, self.seen_classes, self.unseen_classes = classes, seen_classes, unseen_classes
        self.label_to_idx = label_to_idx
        self.device = device
        # cf. methods/clip_baseline.py:39-42
        self.clip_model, self.transform = clip.load(self.config.VIS_ENCODER, device=self.device)
        self.template = self.config.PROMPT_TEMPLATE
        self.val_unseen_files = None   # textual_prompt.py:181 reads it before any FPL subclass sets it
        self.val_unseen_labs = None
        self.balance_param = 1.0
        self._gen = torch.Generator().manual_seed(int(getattr(config, "OPTIM_SEED", 0)))

    # ---- model pieces ------------------------------------------------------------------------------------
    def declare_custom_encoder(self):
        """MODALITY text → CustomTextEncoder, image → CustomImageEncoder, multi → both; the backbone is frozen."""
        dtype = torch.float16 if torch.cuda.is_available() else torch.float32
        for p in self.clip_model.parameters():
            p.requires_grad = False
        if self.config.MODALITY in ("text", "multi"):
            self.text_encoder = CustomTextEncoder(self.clip_model, self.device, dtype).to(self.device)
        if self.config.MODALITY in ("image", "multi"):
            self.image_encoder = CustomImageEncoder(self.clip_model.visual).to(self.device)

    def _normal(self, *shape):
        c = self.config
        return torch.normal(float(c.MEAN_INIT), float(c.VAR_INIT), size=shape, generator=self._gen)

    def initialize_prompts_parameters(self):
        c = self.config
        if c.MODALITY == "text":       # [1, P, 512]: models/clip_encoders.py:55,67 index dim 1 / [0]
            self.initial_prefix = self._normal(1, c.PREFIX_SIZE, self.clip_model.token_embedding.embedding_dim)
        elif c.MODALITY == "image":    # [P, 768]: expanded over the batch at models/clip_encoders.py:148
            width = self.clip_model.visual.class_embedding.size()[0]
            if getattr(c, "VIS_PREFIX_INIT", "normal") == "normal":
                self.vis_initial_prefix = self._normal(c.PREFIX_SIZE, width)
            else:
                self.vis_initial_prefix = torch.rand(c.PREFIX_SIZE, width, generator=self._gen) * 2 - 1
        else:                          # UPT: models/prompts_models.py:88-92
            dt = getattr(self, "dtype", torch.float32)   # multimodal_prompt.py:47 — fp16 on CUDA
            self.coop_embeddings = self._normal(1, c.TEXT_PREFIX_SIZE, self.clip_model.token_embedding.embedding_dim).to(dt)
            self.vpt_embeddings = self._normal(1, c.VISION_PREFIX_SIZE, self.clip_model.visual.class_embedding.size()[0]).to(dt)
            self.vpt_embeddings_deep = None   # VPT_DEEP: False in every shipped config

    def define_model(self, classes=None):
        c = self.config
        if c.MODALITY == "text":
            self.model = TextPrefixModel(self.initial_prefix.clone(), self.text_encoder,
                                         self.classes if classes is None else classes, device=self.device)
        elif c.MODALITY == "image":
            self.model = ImagePrefixModel(self.vis_initial_prefix.clone(), self.image_encoder, device=self.device)
        else:
            dtype = getattr(self, "dtype", torch.float32)
            self.model = UPTModel(self.coop_embeddings.clone(), self.vpt_embeddings.clone(), self.vpt_embeddings_deep,
                                  self.image_encoder, self.text_encoder, self.classes if classes is None else classes,
                                  c.TRANSFORMER_DIM, device=self.device, dtype=dtype)
        self.model = self.model.to(self.device)
        for p in self.model.parameters():
            p.requires_grad = False
        trainable = [p for n, p in self.model.named_parameters()
                     if not n.startswith(("image_encoder.", "text_encoder."))]
        for p in trainable:
            p.requires_grad = True
        if getattr(c, "OPTIM", "SGD") == "SGD":
            self.optimizer = torch.optim.SGD(trainable, lr=c.LR, weight_decay=c.DECAY, momentum=0.9)
        else:
            self.optimizer = torch.optim.Adam(trainable, lr=c.LR, weight_decay=c.DECAY)
        self.scheduler = make_scheduler(self.optimizer, c)
        self.loss_func = nn.CrossEntropyLoss()
        self.training_model = self.model

    # ---- step pieces the subclasses' _train_epoch calls ----------------------------------------------------
    def define_loss_function(self, logits, labs, paths=None):
        return self.loss_func(logits, labs)

    def backpropagate(self):
        self.optimizer.step()
        self.optimizer.zero_grad()

    def update_scheduler(self):
        self.scheduler.step()

    def unwrap_model(self):
        return accelerator.unwrap_model(self.model)

    def create_training_dataset(self, train_data, unlabeled_data=None):
        return train_data   # the prompt baselines train on the labeled data alone; the FPL strategies override

    # ---- training loops ------------------------------------------------------------------------------------
    def train(self, train_data, val_data, unlabeled_data=None, only_unlabelled=False, only_seen=False,
              pseudo_labeled=None):
        """→ (best validation accuracy, parameters of that epoch).  `pseudo_labeled` (iterations > 1 of the
        iterative strategies) is a dataset already labeled by `get_pseudo_labels`."""
        if pseudo_labeled is not None:
            train_data = self._merge_pseudo_labels(train_data, pseudo_labeled)
        elif unlabeled_data is not None:
            # some copies mutate train_data in place and return nothing (semi_supervised_learning/visual_fpl.py:54-114)
            train_data = self.create_training_dataset(train_data, unlabeled_data) or train_data
        if only_seen:
            self.define_model(self.seen_classes)
        elif only_unlabelled:
            self.define_model(self.unseen_classes)
        else:
            self.define_model(self.classes)
        if self.val_unseen_files is not None and val_data is not None:
            seen_labs = [self.label_to_idx[l] for l in val_data.labels] if not val_data.label_id else list(val_data.labels)
            val_data.filepaths = list(self.val_unseen_files) + list(val_data.filepaths)
            val_data.labels = [int(l) for l in self.val_unseen_labs] + seen_labs
            val_data.label_id = True
        train_data.transform = self.transform
        g = torch.Generator().manual_seed(0)
        train_loader = torch.utils.data.DataLoader(train_data, batch_size=self.config.BATCH_SIZE, shuffle=True,
                                                   worker_init_fn=seed_worker, generator=g)
        val_loader = None
        if val_data is not None and len(val_data.filepaths) > 0:
            val_data.transform = self.transform
            val_loader = torch.utils.data.DataLoader(val_data, batch_size=self.config.BATCH_SIZE)
        self.model, self.optimizer, train_loader, val_loader = accelerator.prepare(
            self.model, self.optimizer, train_loader, val_loader)
        self.training_model = self.model
        best_val_accuracy, best_prompt, loss = -1.0, None, None
        accum = getattr(self.config, "ACCUMULATION_ITER", 1)
        for epoch in range(self.config.EPOCHS):
            log.info(f"Run Epoch {epoch}")
            total_loss = 0
            loss, total_loss, epoch_parameters = self._train_epoch(loss, total_loss, train_loader, accum, epoch,
                                                                   only_unlabelled=only_unlabelled, only_seen=only_seen)
            log.info(f"Loss Epoch {epoch}: {total_loss / max(1, len(train_loader))}")
            accelerator.free_memory()
            if val_loader is not None:
                val_accuracy = float(self._run_validation(val_loader, only_unlabelled, only_seen))
                if val_accuracy > best_val_accuracy:
                    best_val_accuracy, best_prompt = val_accuracy, copy.deepcopy(epoch_parameters)
            else:
                best_val_accuracy, best_prompt = None, epoch_parameters
        return best_val_accuracy, best_prompt

    def _merge_pseudo_labels(self, train_data, pseudo_labeled):
        """Runs the subclass's own create_training_dataset (validation split, balance_param, concatenation —
        paradigm specific, e.g. textual_fpl.py:84-121) on pseudolabels that already exist: the module-level
        `pseudolabel_top_k` it calls is answered with them for the duration of the call."""
        mod = sys.modules[type(self).__module__]
        fpl_mod = next((sys.modules[k.__module__] for k in type(self).__mro__
                        if hasattr(sys.modules.get(k.__module__), "pseudolabel_top_k")), mod)
        saved = fpl_mod.pseudolabel_top_k
        fpl_mod.pseudolabel_top_k = lambda *a, **k: pseudo_labeled
        try:
            return self.create_training_dataset(train_data, pseudo_labeled) or train_data
        finally:
            fpl_mod.pseudolabel_top_k = saved

    def _iterate(self, train_data, val_data, unlabeled_data, only_seen, grow):
        """GRIP / fixed iterative refresh (schedule as in pseudo_iterative.py:62-75,113-125):
        num_iter = int(100 / STEP_QUANTILE) rounds; every round re-initialises the prompts, trains on labeled +
        pseudolabeled data and relabels the pool with the trained prompts (`get_pseudo_labels`); GRIP grows
        N_PSEUDOSHOTS by num_samples / n_unseen per round, capped at floor(N / n_unseen)."""
        from utils import save_parameters, save_pseudo_labels

        c = self.config
        num_iter = int(100 / c.STEP_QUANTILE)
        n_unseen = len(self.unseen_classes)
        n_pool = len(unlabeled_data.filepaths)
        num_samples = int(n_pool / num_iter)
        if grow:
            n_per_class = int(num_samples / n_unseen)
            c.N_PSEUDOSHOTS = n_per_class if n_per_class * n_unseen <= n_pool else math.floor(n_pool / n_unseen)
        original_train, original_val, original_pool = (copy.deepcopy(train_data), copy.deepcopy(val_data),
                                                       copy.deepcopy(unlabeled_data))
        pseudo, val_accuracy, optimal_prompt = None, None, None
        for niter in range(1, num_iter + 1):
            log.info(f"Iteration {niter}: {c.N_PSEUDOSHOTS} pseudolabels per class")
            train_data, val_data = copy.deepcopy(original_train), copy.deepcopy(original_val)
            self.val_unseen_files = self.val_unseen_labs = None
            self.initialize_prompts_parameters()
            val_accuracy, optimal_prompt = self.train(train_data, val_data, copy.deepcopy(original_pool),
                                                      only_seen=only_seen, pseudo_labeled=pseudo)
            if accelerator.is_local_main_process:
                save_parameters(optimal_prompt, c, iteration=niter)
            if niter == num_iter:
                break
            if grow:
                n_per_class = int((niter + 1) * num_samples / n_unseen)
                c.N_PSEUDOSHOTS = n_per_class if n_per_class * n_unseen <= n_pool else math.floor(n_pool / n_unseen)
            pseudo = self.get_pseudo_labels(copy.deepcopy(original_pool))
            if accelerator.is_local_main_process:
                save_pseudo_labels(pseudo.filepaths, pseudo.labels, c, niter)
        return val_accuracy, optimal_prompt

    def fixed_iterative_train(self, train_data, val_data, unlabeled_data, only_seen=False):
        return self._iterate(train_data, val_data, unlabeled_data, only_seen, grow=False)

    def grip_train(self, train_data, val_data, unlabeled_data, only_seen=False):
        return self._iterate(train_data, val_data, unlabeled_data, only_seen, grow=True)
