"""utils/pil_resample.py — the integer restatement of Pillow's bicubic resize + the CLIP crop — against Pillow itself
(CPU; bit-exact).  This is what allows the resize to move to the device (gb_resize_bicubic_crop_u8) without changing a
single input pixel of the image tower."""
import importlib

import numpy as np
import pytest
from PIL import Image

R = importlib.import_module("menghini-neurips23-code_b200.utils.pil_resample")
clip = importlib.import_module("menghini-neurips23-code_b200.clip")


@pytest.mark.parametrize("w,h", [(64, 64), (500, 375), (375, 500), (640, 480), (224, 224), (256, 256), (1000, 700),
                                 (28, 28), (300, 224), (224, 300), (513, 384), (225, 224), (224, 225), (97, 201),
                                 (1, 5), (1600, 1200), (223, 500)])
def test_resize_crop_matches_pillow_bit_for_bit(w, h):
    rng = np.random.RandomState(w * 1000 + h)
    # smooth structure + noise + saturated patches: exercises negative lobes and the clip to [0, 255]
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    img[: h // 3, : w // 2] = 255
    img[h // 2:, w // 3:] = (rng.rand(h - h // 2, w - w // 3, 3) > 0.5) * 255
    want = clip.preprocess_u8()(Image.fromarray(img)).numpy()
    got = R.resize_crop_np(img)
    assert got.shape == (3, 224, 224) and got.dtype == np.uint8
    assert np.array_equal(got, want), (np.abs(got.astype(int) - want.astype(int)).max(), (got != want).sum())


def test_geometry_is_torchvisions():
    assert R.clip_geometry(500, 375) == (298, 224, 37, 0)
    assert R.clip_geometry(375, 500) == (224, 298, 0, 37)
    assert R.clip_geometry(64, 64) == (224, 224, 0, 0)
    assert R.clip_geometry(640, 480) == (298, 224, 37, 0)
