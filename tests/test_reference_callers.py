"""The reference's OWN caller classes on the seam (CPU side of SURVEY §8b): ClipBaseline.test_predictions
(methods/clip_baseline.py:44-86) and TextualPrompt.train → _train_epoch / _run_validation / test_predictions
(methods/semi_supervised_learning/textual_prompt.py:63-296), imported unmodified from /root/reference, run on top of
the re-created training_strategies + the accelerate stand-in and reproduce tests/golden/callers_seed0.npz; the
restatements in oracle/callers_ref.py (what the GPU box drives the B200 seam with) give identical results.
Needs /root/reference (the authoring container); the full five-strategy run is `python oracle/make_golden_callers.py`."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "methods")), reason="the reference tree is not mounted")
def test_real_reference_callers_reproduce_the_golden(golden_dir, tmp_path):
    # a fresh interpreter: the run re-binds `clip`, `models`, `utils`, `accelerate` in sys.modules
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); "
            f"import importlib, numpy as np; m = importlib.import_module('oracle.make_golden_callers'); "
            f"r = m.run({REF!r}, ('clip', 'textual')); np.savez({str(tmp_path / 'out.npz')!r}, **r)")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT, timeout=900)
    got, want = np.load(tmp_path / "out.npz"), np.load(os.path.join(golden_dir, "callers_seed0.npz"))
    for k in got.files:
        if got[k].dtype.kind in "US":
            assert np.array_equal(got[k], want[k]), k
        else:
            assert np.allclose(got[k], want[k], rtol=1e-4, atol=1e-5), k   # thread-count dependent fp32 sums


def test_training_strategies_contract():
    """The re-created base class exposes everything the reference's subclasses call or read."""
    src = open(os.path.join(ROOT, "menghini-neurips23-code_b200", "methods", "training_strategies.py")).read()
    for name in ("declare_custom_encoder", "initialize_prompts_parameters", "define_model", "define_loss_function",
                 "backpropagate", "update_scheduler", "unwrap_model", "train", "fixed_iterative_train", "grip_train",
                 "create_training_dataset"):
        assert f"def {name}(" in src, name
    for attr in ("self.clip_model", "self.transform", "self.template", "self.val_unseen_files", "self.loss_func",
                 "self.training_model", "self.text_encoder", "self.image_encoder"):
        assert attr in src, attr
    shim = importlib.import_module("menghini-neurips23-code_b200.accelerate_shim")
    acc = shim.Accelerator()
    for name in ("prepare", "backward", "wait_for_everyone", "gather", "unwrap_model", "free_memory",
                 "is_local_main_process"):
        assert hasattr(acc, name), name
