"""CPU/GPU-neutral restatement of the reference's CALLERS of the hot path, plus the synthetic scenario the
caller tests run.

TEST INFRASTRUCTURE ONLY (see oracle/clip_ref.py header for who may import oracle/).

/root/reference does not exist on the GPU box, so the loops that drive the seam there are restated here, each
citing the reference lines it follows (methods/semi_supervised_learning/*.py, methods/clip_baseline.py).  They are
pinned to the reference's own classes by oracle/make_golden_callers.py: in the authoring container it runs the REAL
`TextualPrompt`, `VisualPrompt`, `MultimodalPrompt`, `TextualFPL` and `ClipBaseline` (imported from
/root/reference, never copied) on top of the restated CPU `clip` and the re-created `training_strategies`, runs
these restatements on the same inputs, requires identical results, and commits them as
tests/golden/callers_seed0.npz.  The -m gpu tests then run the restatements (and, where a box has
/root/reference, the real classes as well) through the B200 seam and compare with those goldens.

Everything here is written against the SEAM names (`import clip`, `from accelerate import Accelerator`,
`methods.semi_supervised_learning.training_strategies`), so the same code runs on the oracle (CPU) and on the
product (B200) depending on what has been registered in sys.modules.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

CLASSES = ["AnnualCrop", "Forest", "Highway", "Pasture", "River", "SeaLake"]
TEMPLATE = "a centered satellite photo of {}."   # data/dataset_prompts.py:2 style (with the '{}' placeholder)


# ---- scenario ---------------------------------------------------------------------------------------------
def make_pngs(root, n_per_class=6, seed=0, classes=CLASSES):
    """EuroSAT-style tree root/<Class>/<Class>_<i>.png (data/dataset.py:128) of 224×224 class-tinted noise images.
    PNG is lossless, so every box regenerates exactly the same pixels from the seed."""
    from PIL import Image

    rng = np.random.default_rng(seed)
    files = []
    for ci, c in enumerate(classes):
        os.makedirs(os.path.join(root, c), exist_ok=True)
        tint = rng.integers(40, 216, size=3)
        for i in range(n_per_class):
            low = rng.integers(0, 256, size=(7, 7, 3)).astype(np.float32)
            img = np.kron(low, np.ones((32, 32, 1), dtype=np.float32))
            img = 0.55 * img + 0.35 * tint + 0.10 * rng.integers(0, 256, size=(224, 224, 3))
            name = f"{c}_{i}.png"
            Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(os.path.join(root, c, name))
            files.append((name, c))
    return files


def make_config(model, modality, **over):
    """The keys the strategies read (Appendix C of SURVEY.md; methods_config/*.yml)."""
    cfg = dict(DATASET_NAME="EuroSAT", MODEL=model, MODALITY=modality, VIS_ENCODER="ViT-B/32",
               LEARNING_PARADIGM="ssl", PROMPT_TEMPLATE=TEMPLATE, SPLIT_SEED=500, OPTIM_SEED=1,
               PREFIX_SIZE=16, TEXT_PREFIX_SIZE=4, VISION_PREFIX_SIZE=4, TRANSFORMER_DIM=128, VPT_DEEP=False,
               VIS_PREFIX_INIT="normal", MEAN_INIT=0, VAR_INIT=0.02, N_LABEL=2, N_PSEUDOSHOTS=2, STEP_QUANTILE=50,
               validation_seed=0, ratio_train_val=0.8, BATCH_SIZE=8, EPOCHS=3, SCHEDULER="cosine", WARMUP_EPOCHS=1,
               WARMUP_LR=0.0001, ACCUMULATION_ITER=1, OPTIM="SGD", LR=0.02, DECAY=0.1, STEP_SIZE=1)
    cfg.update(over)
    return types.SimpleNamespace(**cfg)


def scenario(root, dataset_cls, n_per_class=6, seed=0):
    """Labeled train / val, unlabeled pool and test datasets over the synthetic tree."""
    files = make_pngs(root, n_per_class, seed)
    label_to_idx = {c: i for i, c in enumerate(CLASSES)}
    by_class = {c: [f for f, cc in files if cc == c] for c in CLASSES}
    tr, va, un, te = [], [], [], []
    for c in CLASSES:
        fs = by_class[c]
        tr += [(fs[0], c), (fs[1], c)]
        va += [(fs[2], c)]
        un += [(f, c) for f in fs[3:5]]
        te += [(f, c) for f in fs[5:]]

    def ds(items, labeled, train=True):
        return dataset_cls([f for f, _ in items], root, transform=None, augmentations=None, train=train,
                           labels=[c for _, c in items] if labeled else None, label_map=label_to_idx)

    return dict(label_to_idx=label_to_idx, train=ds(tr, True), val=ds(va, True), unlabeled=ds(un, False),
                test=ds(te, False, train=False), unlabeled_names=[f for f, _ in un],
                test_labels=[c for _, c in te], root=root)


class EuroSATRef(torch.utils.data.Dataset):
    """data/dataset.py:12-90 (CustomDataset) + :93-128 (EuroSAT): items are (img, aug1, aug2[, label], file name);
    without augmentations the three images are the same transform applied three times (:64-79)."""

    def __init__(self, filepaths, root, transform, augmentations=None, train=True, labels=None, label_id=False,
                 label_map=None, class_folder=False, original_filepaths=None):
        self.train = train
        self.filepaths = [f"{root}/{f.split('_')[0]}/{f}" for f in filepaths]          # :128
        self.transform = transform
        self.aug1_transform = self.aug2_transform = None
        self.labels, self.label_id, self.label_map = labels, label_id, label_map

    def __len__(self):
        return len(self.filepaths)

    def __getitem__(self, index):
        from PIL import Image

        img = Image.open(self.filepaths[index]).convert("RGB")
        aug_1, aug_2 = self.transform(img), self.transform(img)
        img = self.transform(img)
        name = self.filepaths[index].split("/")[-1]
        if self.labels is not None:
            label = int(self.labels[index]) if self.label_id else int(self.label_map[self.labels[index]])
            return img, aug_1, aug_2, label, name
        return img, aug_1, aug_2, name


# ---- restated callers --------------------------------------------------------------------------------------
def build_ref_strategies(TrainingStrategy):
    """The reference's strategy classes, restated on top of whichever `TrainingStrategy` the seam provides.
    Returned as a namespace: TextualPrompt, VisualPrompt, MultimodalPrompt, TextualFPL, ClipBaseline."""
    import clip
    import pandas as pd
    from accelerate import Accelerator
    from PIL import Image

    accelerator = Accelerator()
    cuda = torch.cuda.is_available()

    def _classes_of(model):   # textual_prompt.py:86-97 — DDP wrapper on CUDA, bare module on CPU
        return model.module.classes if cuda else model.classes

    class TextualPrompt(TrainingStrategy):
        """methods/semi_supervised_learning/textual_prompt.py"""

        def __init__(self, config, label_to_idx, classes, seen_classes, unseen_classes, device):   # :30-61
            super().__init__(config, label_to_idx, classes, seen_classes, unseen_classes, device)
            seen_to_idx = {c: idx for idx, c in enumerate(self.seen_classes)}
            self.idx_to_real = {seen_to_idx[c]: self.label_to_idx[c] for c in self.seen_classes}
            self.real_to_idx = {self.label_to_idx[c]: seen_to_idx[c] for c in self.seen_classes}
            self.declare_custom_encoder()
            self.initialize_prompts_parameters()

        def _train_epoch(self, loss, total_loss, train_loader, accum_iter, epoch, only_unlabelled=False,
                         only_seen=False):                                                          # :63-159
            predictions, labels = [], []
            for i, (img, _, _, label, img_path) in enumerate(train_loader):
                text_features = self.model(_classes_of(self.model))                                 # :94-97
                text_features = text_features / text_features.norm(dim=-1, keepdim=True)
                with torch.no_grad():
                    image_features = self.clip_model.encode_image(img)                              # :100
                    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()     # :106-107
                idx_preds = torch.argmax(logits, dim=1)
                names = self.seen_classes if only_seen else self.classes
                predictions += [names[j.item()] for j in idx_preds]
                labels += [self.classes[j.item()] for j in label]
                if only_seen:
                    labs = torch.tensor([self.real_to_idx[l.item()] for l in label]).to(self.device)
                else:
                    labs = torch.tensor([l.item() for l in label]).to(self.device)
                loss = self.define_loss_function(logits, labs, img_path)                            # :125
                total_loss += loss.item()
                accelerator.wait_for_everyone()
                loss = loss / accum_iter
                accelerator.backward(loss)                                                          # :131
                if ((i + 1) % accum_iter == 0) or (i + 1 == len(train_loader)):
                    self.backpropagate()                                                            # :134-135
            self.update_scheduler()                                                                 # :152
            return loss, total_loss, [self.unwrap_model().prefix.detach().cpu().numpy()]            # :154-159

        def _run_validation(self, val_loader, only_unlabelled=False, only_seen=False):              # :161-224
            predictions, labels = [], []
            for img, _, _, label, img_path in val_loader:
                cl = self.classes if self.val_unseen_files is not None else self.seen_classes
                if cuda:
                    self.model.module.classes = cl
                else:
                    self.model.classes = cl
                text_features = self.model(cl)
                text_features = text_features / text_features.norm(dim=-1, keepdim=True)
                with torch.no_grad():
                    image_features = self.clip_model.encode_image(img)
                    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                predictions += [cl[j.item()] for j in torch.argmax(logits, dim=1)]
                labels += [self.classes[j.item()] for j in label]
            p = torch.tensor([self.label_to_idx[x] for x in predictions][: len(val_loader.dataset)])
            t = torch.tensor([self.label_to_idx[x] for x in labels][: len(val_loader.dataset)])
            return torch.sum(p == t) / len(p)

        def test_predictions(self, data, standard_zsl=False):                                       # :226-296
            data.transform = self.transform
            test_loader = torch.utils.data.DataLoader(data, batch_size=self.config.BATCH_SIZE)
            self.model, test_loader = accelerator.prepare(self.model, test_loader)
            self.model.classes = self.unseen_classes if standard_zsl else self.classes
            text_features = self.model(self.model.classes)
            text_features = text_features / text_features.norm(dim=-1, keepdim=True)
            test_files = [f.split("/")[-1] for f in test_loader.dataset.filepaths]
            predictions, images = [], []
            for img, _, _, img_path in test_loader:
                with torch.no_grad():
                    image_features = self.clip_model.encode_image(img)
                    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                predictions += [self.model.classes[j] for j in torch.argmax(logits, dim=1)]
                images += list(img_path)
            df = pd.DataFrame({"id": [test_files[test_files.index(i)] for i in images], "class": predictions})
            df.drop_duplicates(subset=["id", "class"], inplace=True)
            return df

    class VisualPrompt(TrainingStrategy):
        """methods/semi_supervised_learning/visual_prompt.py"""

        def __init__(self, config, label_to_idx, classes, seen_classes, unseen_classes, device):   # :24-53
            super().__init__(config, label_to_idx, classes, seen_classes, unseen_classes, device)
            self.declare_custom_encoder()
            self.initialize_prompts_parameters()

        def define_textual_prompts(self, only_unlabelled=False, validation=False):                  # :55-65
            return [self.template.format(" ".join(i.split("_"))) for i in self.seen_classes]

        def _text_features(self, prompts):                                                          # :115-118
            with torch.no_grad():
                text_features = self.clip_model.encode_text(clip.tokenize(prompts).to(self.device))
                return text_features / text_features.norm(dim=-1, keepdim=True)

        def _train_epoch(self, loss, total_loss, train_loader, accum_iter, epoch, only_unlabelled=False,
                         only_seen=False):                                                          # :87-171
            text_features = self._text_features(self.define_textual_prompts(only_unlabelled))
            for i, (img, _, _, label, img_path) in enumerate(train_loader):
                image_features = self.training_model(img)                                           # :123
                image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                labs = torch.tensor([self.seen_classes.index(self.classes[l.item()]) for l in label]).to(self.device)
                loss = self.define_loss_function(logits, labs, img_path)
                total_loss += loss.item()
                accelerator.wait_for_everyone()
                loss = loss / accum_iter
                accelerator.backward(loss)
                if ((i + 1) % accum_iter == 0) or (i + 1 == len(train_loader)):
                    self.backpropagate()
            self.update_scheduler()
            return loss, total_loss, [self.unwrap_model().prefix.detach().cpu().numpy()]

        def _run_validation(self, val_loader, only_unlabelled=False, only_seen=False):              # :173-231
            text_features = self._text_features(
                self.define_textual_prompts(only_unlabelled, validation=self.val_unseen_files is None))
            predictions, labels = [], []
            for img, _, _, label, img_path in val_loader:
                image_features = self.training_model(img)
                image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                names = self.classes if self.val_unseen_files is not None else self.seen_classes
                predictions += [names[j.item()] for j in torch.argmax(logits, dim=1)]
                labels += [self.classes[j.item()] for j in label]
            p = torch.tensor([self.label_to_idx[x] for x in predictions])
            t = torch.tensor([self.label_to_idx[x] for x in labels])
            return torch.sum(p == t) / len(p)

        def test_predictions(self, data, standard_zsl=False):                                       # :233-310
            data.transform = self.transform
            test_loader = torch.utils.data.DataLoader(data, batch_size=self.config.BATCH_SIZE)
            self.model, test_loader = accelerator.prepare(self.model, test_loader)
            names = self.unseen_classes if standard_zsl else self.classes
            text_features = self._text_features([self.template.format(" ".join(i.split("_"))) for i in names])
            predictions, images = [], []
            for img, _, _, img_path in test_loader:
                with torch.no_grad():
                    image_features = self.model(img)
                    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                predictions += [names[j] for j in torch.argmax(logits, dim=1)]
                images += list(img_path)
            df = pd.DataFrame({"id": images, "class": predictions})
            df.drop_duplicates(subset=["id", "class"], inplace=True)
            return df

    class MultimodalPrompt(TrainingStrategy):
        """methods/semi_supervised_learning/multimodal_prompt.py"""

        def __init__(self, config, label_to_idx, classes, seen_classes, unseen_classes, device):   # :23-52
            super().__init__(config, label_to_idx, classes, seen_classes, unseen_classes, device)
            self.dtype = torch.float16 if cuda else torch.float32                                 # :47
            # (test hook, not in the reference: UPT_DTYPE keeps the head's parameters in fp32 on CUDA so that the
            # trajectory can be compared with the fp32 CPU oracle; fp16 parameters lose updates below 2^-11 relative)
            self.dtype = getattr(config, "UPT_DTYPE", self.dtype)
            self.declare_custom_encoder()
            self.initialize_prompts_parameters()

        def _train_epoch(self, loss, total_loss, train_loader, accum_iter, epoch, only_unlabelled=False,
                         only_seen=False):                                                          # :73-166
            classes = self.unseen_classes if only_unlabelled else self.seen_classes if only_seen else self.classes
            for i, (img, _, _, label, img_path) in enumerate(train_loader):
                text_features, image_features = self.model(img, classes)                            # :105
                text_features = text_features / text_features.norm(dim=-1, keepdim=True)
                image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                labs = torch.tensor([self.seen_classes.index(self.classes[l.item()]) for l in label]).to(self.device)
                loss = self.define_loss_function(logits, labs, img_path)
                total_loss += loss.item()
                accelerator.wait_for_everyone()
                loss = loss / accum_iter
                accelerator.backward(loss)
                if ((i + 1) % accum_iter == 0) or (i + 1 == len(train_loader)):
                    self.backpropagate()
            self.update_scheduler()
            m = self.unwrap_model()                                                                 # :149-164
            return loss, total_loss, [m.transformer.state_dict(), m.proj_coop_pre.state_dict(),
                                      m.proj_coop_post.state_dict(), m.proj_vpt_pre.state_dict(),
                                      m.proj_vpt_post.state_dict(), m.coop_embeddings.detach().cpu().numpy(),
                                      None if m.vpt_embeddings_deep is None else
                                      m.vpt_embeddings_deep.detach().cpu().numpy(),
                                      m.vpt_embeddings.detach().cpu().numpy()]

        def _run_validation(self, val_loader, only_unlabelled=False, only_seen=False):              # :168-218
            classes = self.seen_classes if self.val_unseen_files is None else self.classes
            predictions, labels = [], []
            for img, _, _, label, img_path in val_loader:
                text_features, image_features = self.model(img, classes)
                text_features = text_features / text_features.norm(dim=-1, keepdim=True)
                image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                predictions += [classes[j.item()] for j in torch.argmax(logits, dim=1)]
                labels += [self.classes[j.item()] for j in label]
            p = torch.tensor([self.label_to_idx[x] for x in predictions])
            t = torch.tensor([self.label_to_idx[x] for x in labels])
            return torch.sum(p == t) / len(p)

    class TextualFPL(TextualPrompt):
        """methods/semi_supervised_learning/textual_fpl.py (the SSL variant)"""

        def __init__(self, config, label_to_idx, data_folder, unlabeled_files, classes, seen_classes,
                     unseen_classes, device):                                                       # :30-56
            super().__init__(config, label_to_idx, classes, seen_classes, unseen_classes, device)
            self.data_folder = data_folder
            self.check_unlabeled = unlabeled_files

        def create_training_dataset(self, train_data, unlabeled_data=None):                         # :58-121
            from utils import pseudolabel_top_k

            c = self.config
            ds = pseudolabel_top_k(c, c.DATASET_NAME, c.N_PSEUDOSHOTS, c.PROMPT_TEMPLATE, unlabeled_data,
                                   self.unseen_classes, self.transform, self.clip_model, self.label_to_idx,
                                   self.device, c.VIS_ENCODER, c.SPLIT_SEED)
            unseen_imgs, unseen_labs = ds.filepaths, ds.labels
            if c.N_PSEUDOSHOTS >= 10:                                                               # :88-103
                np.random.seed(c.validation_seed)
                tr = np.random.choice(range(len(unseen_imgs)), size=int(len(unseen_imgs) * c.ratio_train_val),
                                      replace=False)
                va = list(set(range(len(unseen_imgs))).difference(set(tr)))
                self.val_unseen_files = np.array(unseen_imgs)[va]
                self.val_unseen_labs = np.array(unseen_labs)[va]
                unseen_imgs, unseen_labs = list(np.array(unseen_imgs)[tr]), list(np.array(unseen_labs)[tr])
            else:
                self.val_unseen_files = self.val_unseen_labs = None
            seen_imgs = train_data.filepaths
            seen_labs = [self.label_to_idx[l] for l in train_data.labels]
            self.balance_param = len(unseen_imgs) / len(seen_imgs)                                   # :115
            train_data.filepaths = list(unseen_imgs) + list(seen_imgs)
            train_data.labels = list(unseen_labs) + list(seen_labs)
            train_data.label_id = True
            return train_data

        def define_loss_function(self, logits, labs, paths):                                        # :123-128
            return (self.balance_param * self.cross_entropy(logits, labs, paths, False)
                    + self.cross_entropy(logits, labs, paths, True))

        def cross_entropy(self, logits, labels, paths, unlabeled=True):                             # :130-165
            samples = [i for i in range(len(paths)) if (paths[i] in self.check_unlabeled) == unlabeled]
            return self.loss_func(logits[samples], labels[samples]) if samples else 0

        def assign_pseudo_labels(self, k, unlabeled_data):                                          # :195-283
            self.model.classes = self.unseen_classes
            text_features = self.model(self.model.classes)
            text_features = text_features / text_features.norm(dim=-1, keepdim=True)
            boards = {self.label_to_idx[c]: [] for c in self.unseen_classes}
            for img_path in unlabeled_data.filepaths:
                img = torch.unsqueeze(self.transform(Image.open(img_path).convert("RGB")), 0).to(self.device)
                with torch.no_grad():
                    image_features = self.clip_model.encode_image(img)
                    image_features = image_features / image_features.norm(dim=-1, keepdim=True)
                logits = self.clip_model.logit_scale.exp() * image_features @ text_features.t()
                probs = logits.softmax(dim=-1)
                pred_id = torch.argmax(logits, dim=1).item()
                pred = self.label_to_idx[self.unseen_classes[pred_id]]
                score = probs[0][pred_id]
                if len(boards[pred]) < k:
                    boards[pred].append((score, img_path))
                elif boards[pred][-1][0] < score:
                    boards[pred] = sorted(boards[pred] + [(score, img_path)], reverse=True)[:k]
                else:
                    for j in range(len(self.unseen_classes)):
                        if j == pred_id:
                            continue
                        cj = self.label_to_idx[self.unseen_classes[j]]
                        if len(boards[cj]) < k:
                            boards[cj].append((probs[0][j], img_path))
                        elif boards[cj][-1][0] < probs[0][j]:
                            boards[cj] = sorted(boards[cj] + [(probs[0][j], img_path)], reverse=True)[:k]
            unlabeled_data.filepaths = [t[1] for b in boards.values() for t in b]
            unlabeled_data.labels = [cid for cid, b in boards.items() for _ in b]
            unlabeled_data.label_id = True
            return unlabeled_data

    class ClipBaseline(object):
        """methods/clip_baseline.py"""

        def __init__(self, config, label_to_idx, classes, seen_classes, unseen_classes, device):   # :18-42
            self.config, self.classes, self.label_to_idx, self.device = config, classes, label_to_idx, device
            self.seen_classes, self.unseen_classes = seen_classes, unseen_classes
            self.model, self.transform = clip.load(config.VIS_ENCODER, device=device)
            self.template = config.PROMPT_TEMPLATE

        def test_predictions(self, data):                                                           # :44-86
            data.transform = self.transform
            test_loader = torch.utils.data.DataLoader(data, batch_size=self.config.BATCH_SIZE)
            prompts = [self.template.format(" ".join(i.split("_"))) for i in self.classes]
            text = clip.tokenize(prompts).to(self.device)
            predictions, images, prob_preds = [], [], []
            for img, _, _, img_path in test_loader:
                with torch.no_grad():
                    logits_per_image, _ = self.model(img.to(self.device), text.to(self.device))
                    idx_preds = torch.argmax(logits_per_image.softmax(dim=-1), dim=1)
                    predictions += [self.classes[j] for j in idx_preds]
                    images += list(img_path)
                    prob_preds += [logits_per_image]
            prob_preds = torch.cat(prob_preds, axis=0).detach().to("cpu")
            return pd.DataFrame({"id": images, "class": predictions}), images, predictions, prob_preds

    return types.SimpleNamespace(TextualPrompt=TextualPrompt, VisualPrompt=VisualPrompt,
                                 MultimodalPrompt=MultimodalPrompt, TextualFPL=TextualFPL, ClipBaseline=ClipBaseline)


# ---- the runs the golden file records ----------------------------------------------------------------------
def run_all(S, dataset_cls, root, device, which=("clip", "textual", "visual", "multimodal", "fpl")):
    """Drives the strategy classes in namespace `S` (the reference's own, or build_ref_strategies') through
    training, validation, evaluation and pseudolabel assignment on the synthetic scenario; returns a flat dict of
    numpy arrays (what tests/golden/callers_seed0.npz stores).  Must be run with cwd = a scratch directory that
    has pseudolabels/ (the pseudolabel cache is written relative to cwd, utils/clip_pseudolabels.py:134)."""
    out = {}
    kw = dict(classes=CLASSES, seen_classes=CLASSES, unseen_classes=CLASSES, device=device)   # SSL: main_SSL.py:74-75

    def frame(df):
        return np.array([f"{i}|{c}" for i, c in zip(df["id"], df["class"])])

    def trainer(cls, model, modality, extra=(), **cfg):
        sc = scenario(root, dataset_cls)
        config = make_config(model, modality, **cfg)
        torch.manual_seed(config.OPTIM_SEED)   # main_SSL.py:493-500 seeds torch before the strategy is built
        strat = cls(config, sc["label_to_idx"], *extra, **kw) if not extra else cls(
            config, sc["label_to_idx"], extra[0], unlabeled_files=sc["unlabeled_names"], **kw)
        return sc, config, strat

    if "clip" in which:                                   # BASELINE configs[0]: zero-shot, B = 8
        sc = scenario(root, dataset_cls)
        cb = S.ClipBaseline(make_config("clip_baseline", "text"), sc["label_to_idx"], **kw)
        df, images, predictions, logits = cb.test_predictions(sc["test"])
        out["clip.pred"], out["clip.logits"] = frame(df), logits.float().numpy()
    if "textual" in which:                                # configs[1]: CoOp
        sc, config, st = trainer(S.TextualPrompt, "textual_prompt", "text")
        torch.manual_seed(config.OPTIM_SEED)   # the UPT head's Linear / transformer init draws from the global RNG
        acc, best = st.train(sc["train"], sc["val"], only_seen=True)
        out["textual.val_acc"], out["textual.best_prefix"] = np.float32(acc), np.asarray(best[0])
        out["textual.final_prefix"] = st.unwrap_model().prefix.detach().float().cpu().numpy()
        out["textual.test"] = frame(st.test_predictions(sc["test"], standard_zsl=True))
    if "visual" in which:                                 # configs[2]: VPT
        sc, config, st = trainer(S.VisualPrompt, "visual_prompt", "image")
        torch.manual_seed(config.OPTIM_SEED)   # the UPT head's Linear / transformer init draws from the global RNG
        acc, best = st.train(sc["train"], sc["val"], only_seen=True)
        out["visual.val_acc"] = np.float32(acc)
        out["visual.final_prefix"] = st.unwrap_model().prefix.detach().float().cpu().numpy()
        out["visual.test"] = frame(st.test_predictions(sc["test"], standard_zsl=True))
    if "multimodal" in which:                             # configs[4]: UPT
        sc, config, st = trainer(S.MultimodalPrompt, "multimodal_prompt", "multi", LR=0.01, UPT_DTYPE=torch.float32)
        torch.manual_seed(config.OPTIM_SEED)   # the UPT head's Linear / transformer init draws from the global RNG
        acc, best = st.train(sc["train"], sc["val"], only_seen=True)
        m = st.unwrap_model()
        out["multimodal.val_acc"] = np.float32(acc)
        out["multimodal.coop"] = m.coop_embeddings.detach().float().cpu().numpy()
        out["multimodal.vpt"] = m.vpt_embeddings.detach().float().cpu().numpy()
        out["multimodal.proj_coop_pre.weight"] = m.proj_coop_pre.weight.detach().float().cpu().numpy()
    if "fpl" in which:                                    # configs[3]-style: pseudolabels + FPL loss + assign_pseudo_labels
        sc, config, st = trainer(S.TextualFPL, "textual_fpl", "text", extra=(root,), EPOCHS=2)
        torch.manual_seed(config.OPTIM_SEED)   # the UPT head's Linear / transformer init draws from the global RNG
        acc, best = st.train(sc["train"], sc["val"], sc["unlabeled"], only_seen=False)
        out["fpl.balance"] = np.float32(st.balance_param)
        out["fpl.final_prefix"] = st.unwrap_model().prefix.detach().float().cpu().numpy()
        import copy
        pool = copy.deepcopy(scenario(root, dataset_cls)["unlabeled"])
        pool.transform = st.transform
        res = st.assign_pseudo_labels(2, pool)
        out["fpl.assign"] = np.array([f"{os.path.basename(f)}|{l}" for f, l in zip(res.filepaths, res.labels)])
    return out
