"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header
declares, and refuses to run without a B200 (no CPU fallback)."""
import ctypes
import importlib
import os
import subprocess

import pytest
import torch


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load()
    names = pkg.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_binding_covers_header(pkg):
    lib_mod = importlib.import_module("menghini-neurips23-code_b200._lib")
    assert set(lib_mod._SIGNATURES) == set(pkg.declared_symbols())


def test_version_and_sizes_without_gpu(pkg):
    lib = pkg.load()
    assert b"sm_100a" in lib.gb_version()
    # pure host arithmetic entry points
    assert lib.gb_tape_bytes(2, 50, 768, 12) == (12 * 9 + 1) * 2 * 50 * 768 * 2
    assert lib.gb_tape_bytes(0, 50, 768, 12) == 0
    assert lib.gb_leaderboard_state_bytes(10, 16) == (8 + 30 + 2 * 160 + 2 * 17) * 4
    assert lib.gb_leaderboard_state_bytes(0, 16) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(pkg):
    lib = pkg.load()
    h = ctypes.c_void_p()
    assert lib.gb_create(ctypes.byref(h), 0) == -4  # GB_ERR_NO_DEVICE
    with pytest.raises(pkg.GripB200Error):
        pkg.Context(0)
    clip = importlib.import_module("menghini-neurips23-code_b200.clip")
    with pytest.raises(pkg.GripB200Error):
        clip.load("ViT-B/32", "cpu", state_dict={})
    eng = importlib.import_module("menghini-neurips23-code_b200.engine")
    with pytest.raises(pkg.GripB200Error):
        eng.Leaderboard(4, 2, "cpu")


def test_product_does_not_import_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "menghini-neurips23-code_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), os.path.join(dp, f)


def test_dropin_install_registers_the_reference_import_seam():
    """`import clip`, `from clip import clip`, `clip.model.Transformer`, `from models import …` and
    `utils.pseudolabel_top_k` resolve to the B200 implementations (run in a subprocess: it edits
    sys.modules)."""
    import sys

    code = (
        "import importlib, sys; sys.path.insert(0, %r)\n"
        "importlib.import_module('menghini-neurips23-code_b200.dropin').install(patch_reference_utils=False)\n"
        "import clip\n"
        "from clip import clip as c2\n"
        "from models import (CustomImageEncoder, CustomTextEncoder, ImageEncoder, TextEncoder,\n"
        "                    ImagePrefixModel, TextPrefixModel, UPTModel)\n"
        "from utils import pseudolabel_top_k\n"
        "import utils.clip_pseudolabels as cp\n"
        "assert c2 is clip and hasattr(clip, 'load') and hasattr(clip, 'tokenize')\n"
        "t = clip.model.Transformer(width=128, layers=1, heads=1)\n"
        "import torch; assert t(torch.zeros(2, 4, 128)).shape == (2, 4, 128)\n"
        "assert 'menghini' in CustomTextEncoder.__module__ and 'menghini' in pseudolabel_top_k.__module__\n"
        "assert cp.compute_pseudo_labels.__module__ == pseudolabel_top_k.__module__\n"
        "print('ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_warmup_cosine_lr_is_the_reference_rule():
    """gb_warmup_cosine_lr restates utils/schedulers.py:54-65 (WarmupCosineSchedule.lr_lambda, cycles 0.5):
    same numbers as the Python formula for every epoch of the shipped configs (WARMUP_EPOCHS 5, EPOCHS 150)."""
    import importlib
    import math

    lib = importlib.import_module("menghini-neurips23-code_b200").load()

    def lr_lambda(step, warmup_steps, t_total, cycles=0.5):
        if step < warmup_steps:
            return float(step) / float(max(1.0, warmup_steps))
        progress = float(step - warmup_steps) / float(max(1, t_total - warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(cycles) * 2.0 * progress)))

    for warm, total in ((5, 150), (0, 10), (1, 2), (5, 5)):
        for step in range(total + 3):
            got = lib.gb_warmup_cosine_lr(0.1, warm, total, step)
            assert abs(got - 0.1 * lr_lambda(step, warm, total)) <= 1e-15, (warm, total, step, got)


def test_wave_aligned_batch_fills_whole_gemm_waves():
    """Engine.wave_aligned_batch: the returned batch makes every tower GEMM's tile count (256-row blocks x
    N/256 column tiles, N in {768, 2304, 3072}) a multiple of the CTA pairs in use, and is the largest such."""
    import importlib

    Engine = importlib.import_module("menghini-neurips23-code_b200.engine").Engine
    for sms, L in ((144, 50), (148, 50), (144, 66), (148, 66)):
        b = Engine.wave_aligned_batch(1024, L=L, sms=sms)
        pairs = sms // 2
        blocks = -(-(b * L) // 256)
        assert all((blocks * n) % pairs == 0 for n in (3, 9, 12)), (sms, L, b)
        assert all((-(-(c * L) // 256) * 3) % pairs != 0 for c in range(b + 1, 1025)), (sms, L, b)
    assert Engine.wave_aligned_batch(1024, L=50, sms=144) == 983


def test_predictions_frame_matches_the_reference_bookkeeping():
    """utils.predictions_frame = the tail of the reference's test_predictions
    (methods/semi_supervised_learning/textual_prompt.py:256-294): ids are file names, classes are names, rows with
    the same (id, class) are dropped, first occurrence order is kept."""
    import importlib

    import pandas as pd

    utils = importlib.import_module("menghini-neurips23-code_b200.utils")
    classes = ["forest", "river", "highway"]
    paths = ["/d/a/img1.jpg", "/d/b/img2.jpg", "/d/c/img1.jpg", "/d/a/img3.jpg", "/d/z/img2.jpg"]
    pred = [2, 0, 2, 1, 1]
    got = utils.predictions_frame(paths, pred, classes)
    # what the reference builds (:289-294)
    want = pd.DataFrame({"id": [p.split("/")[-1] for p in paths], "class": [classes[i] for i in pred]})
    want.drop_duplicates(subset=["id", "class"], inplace=True)
    pd.testing.assert_frame_equal(got, want)
    assert list(got["id"]) == ["img1.jpg", "img2.jpg", "img3.jpg", "img2.jpg"]


def test_fpl_coefficients_reproduce_the_two_term_loss():
    """training.fpl_coefficients: Σ coef_i · CE_i equals the reference's `w1 * mean CE(first) + w2 * mean CE(second)`
    (methods/semi_supervised_learning/textual_fpl.py:123-165, methods/transductive_zsl/textual_fpl.py:117-147),
    including batches where one group is empty."""
    import importlib

    import torch

    training = importlib.import_module("menghini-neurips23-code_b200.training")
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(16, 7, generator=g) * 3
    labels = torch.randint(0, 7, (16,), generator=g)
    ce = torch.nn.functional.cross_entropy
    for mask in (torch.rand(16, generator=g) < 0.4, torch.zeros(16, dtype=torch.bool), torch.ones(16, dtype=torch.bool)):
        for w1, w2 in ((2.5, 1.0), (1.0, 0.3)):
            first = w1 * ce(logits[~mask], labels[~mask]) if (~mask).any() else 0
            second = w2 * ce(logits[mask], labels[mask]) if mask.any() else 0
            want = first + second
            coef = training.fpl_coefficients(mask, w1, w2)
            got = (coef * ce(logits, labels, reduction="none")).sum()
            assert abs(float(got) - float(want)) <= 1e-6 * max(1.0, abs(float(want)))
