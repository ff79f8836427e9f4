"""gb_resize_bicubic_crop_u8 on a real B200: Pillow's bicubic resize + CLIP's centre crop, bit for bit, and encode_pool
with the resize on the device against encode_pool with it on the host (same features, bit for bit)."""
import importlib
import types

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import clip_ref

pytestmark = pytest.mark.gpu
R = importlib.import_module("menghini-neurips23-code_b200.utils.pil_resample")
clip = importlib.import_module("menghini-neurips23-code_b200.clip")


@pytest.fixture(scope="module")
def resizer(pkg):
    eng = types.SimpleNamespace(device=torch.device("cuda", 0), ctx=pkg.Context.get(0))
    return R.DeviceResizer(eng, arena_bytes=24 << 20)


def _img(rng, w, h):
    a = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    a[: h // 3, : w // 2] = 255
    a[h // 2:, w // 3:] = (rng.rand(h - h // 2, w - w // 3, 3) > 0.5) * 255
    return a


def test_device_resize_matches_pillow_bit_for_bit(resizer):
    rng = np.random.RandomState(0)
    sizes = [(64, 64), (500, 375), (375, 500), (640, 480), (224, 224), (256, 256), (1000, 700), (28, 28), (300, 224),
             (224, 300), (513, 384), (225, 224), (97, 201), (1600, 1200), (500, 375), (64, 64), (64, 64), (500, 375)]
    arrays = [_img(rng, w, h) for w, h in sizes]
    tf = clip.preprocess_u8()
    want = torch.stack([tf(Image.fromarray(a)) for a in arrays])
    out = torch.zeros(len(arrays), 3, 224, 224, dtype=torch.uint8, device="cuda")
    resizer.run(arrays, out)       # 24 MB arena: what does not fit (the later images) goes through the overflow route
    got = out.cpu()
    for i, (w, h) in enumerate(sizes):
        assert torch.equal(got[i], want[i]), (w, h, (got[i] != want[i]).sum().item())
    # a second call reuses the arena and the cached coefficient tables
    out2 = torch.zeros_like(out)
    resizer.run(arrays[::-1], out2)
    assert torch.equal(out2.cpu(), want.flip(0))


def test_encode_pool_device_resize_is_bit_identical(pkg, tmp_path):
    U = importlib.import_module("menghini-neurips23-code_b200.utils")
    sd = clip_ref.synth_state_dict(seed=1234)
    model, transform = clip.load("ViT-B/32", "cuda:0", state_dict=sd)
    rng = np.random.RandomState(1)
    paths = []
    for i, (w, h) in enumerate([(96, 64), (500, 375), (64, 64), (375, 500), (64, 64), (300, 224), (640, 480), (224, 224),
                                (96, 64), (500, 375)]):
        p = tmp_path / (f"im_{i}.jpg" if i % 2 else f"im_{i}.png")
        Image.fromarray(_img(rng, w, h)).save(p)
        paths.append(str(p))
    grey = tmp_path / "grey.png"                         # a non-RGB file: convert("RGB") precedes the resize in both routes
    Image.fromarray(rng.randint(0, 256, (80, 120)).astype(np.uint8)).save(grey)
    paths.append(str(grey))
    host = U.encode_pool(model, paths, transform, "cuda:0", device_resize=False)
    model.engine.__dict__.pop("_pool_cache", None)
    dev = U.encode_pool(model, paths, transform, "cuda:0", device_resize=True, batch=4)     # small pool: decoder threads
    assert torch.equal(host, dev)
    model.engine.__dict__.pop("_pool_cache", None)
    many = paths * 24                                                                       # ≥ 256 files: forked decoders
    devp = U.encode_pool(model, many, transform, "cuda:0", device_resize=True, batch=100, workers=4)
    assert torch.equal(devp, host.repeat(24, 1))
    assert any(k[1] > 1 for k in R._RESIZERS)
    model.engine.__dict__.pop("_pool_cache", None)
    dev1 = U.encode_pool(model, paths, transform, "cuda:0", device_resize=True, batch=5, workers=1)   # one thread
    assert torch.equal(host, dev1)
    assert any(k[1] == 0 for k in R._RESIZERS)
    R.close_resizers()
