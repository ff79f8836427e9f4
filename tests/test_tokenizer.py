"""The restated openai/CLIP byte-pair tokenizer against Hugging Face's independent CLIP BPE
(Rust `tokenizers` backend) on a synthetic merge table — the real vocabulary file is not available
offline — plus the clip.tokenize framing with a vocabulary file."""
import gzip
import importlib
import random

import pytest
import torch


def _toy_merges(seed=0, n=400):
    """Merge table learnt greedily from a toy corpus, in the tokenizer's own symbol alphabet."""
    words = ("a photo of annual crop land forest herbaceous vegetation highway road industrial buildings "
             "pasture permanent residential river sea lake airplane airport baseball diamond beach bridge "
             "chaparral church cloud desert freeway golf course harbor island meadow mountain palace "
             "railway runway stadium terrace wetland boeing airbus x satellite centered the").split()
    rnd = random.Random(seed)
    corpus = {}
    for w in words:
        corpus[tuple(w[:-1]) + (w[-1] + "</w>",)] = rnd.randint(1, 20)
    merges = []
    for _ in range(n):
        counts = {}
        for sym, f in corpus.items():
            for a, b in zip(sym, sym[1:]):
                counts[(a, b)] = counts.get((a, b), 0) + f
        if not counts:
            break
        best = max(sorted(counts), key=lambda p: counts[p])
        merges.append(best)
        new = {}
        for sym, f in corpus.items():
            out, i = [], 0
            while i < len(sym):
                if i + 1 < len(sym) and (sym[i], sym[i + 1]) == best:
                    out.append(sym[i] + sym[i + 1]); i += 2
                else:
                    out.append(sym[i]); i += 1
            new[tuple(out)] = f
        corpus = new
    return merges


@pytest.fixture(scope="module")
def toks():
    st = importlib.import_module("menghini-neurips23-code_b200.clip.simple_tokenizer")
    merges = _toy_merges()
    mine = st.SimpleTokenizer(merges=merges)
    from transformers import CLIPTokenizer
    hf = CLIPTokenizer(vocab=dict(mine.encoder), merges=[tuple(m) for m in merges])
    return mine, hf


TEXTS = ["a photo of a forest", "X X X X annual crop land", "a photo of a {}sea lake", "Highway or Road",
         "permanent crop land 7", "baseball-diamond", "it's a river's bridge", "unseenword zzz", "airport 42"]


def test_bpe_matches_huggingface_clip_bpe(toks):
    mine, hf = toks
    for t in TEXTS:
        want = hf(t, add_special_tokens=False)["input_ids"]
        assert mine.encode(t) == want, t


def test_decode_round_trip(toks):
    mine, _ = toks
    for t in ("a photo of a forest", "annual crop land"):
        assert mine.decode(mine.encode(t)).strip() == t


def test_clip_tokenize_with_vocab_file(toks, tmp_path, monkeypatch):
    mine, _ = toks
    # a file in bpe_simple_vocab_16e6.txt.gz format: header line, then one merge per line
    merges = list(mine.bpe_ranks)
    path = tmp_path / "bpe_toy.txt.gz"
    with gzip.open(path, "wt", encoding="utf-8") as f:
        f.write("#version: toy\n" + "\n".join(" ".join(m) for m in merges) + "\n")
    clip = importlib.import_module("menghini-neurips23-code_b200.clip")
    monkeypatch.setenv("GRIPB200_BPE_VOCAB", str(path))
    ids = clip.tokenize(["X X annual crop land", "sea"])
    sot, eot = mine.encoder["<|startoftext|>"], mine.encoder["<|endoftext|>"]
    assert ids.shape == (2, 77) and ids.dtype == torch.long
    assert ids[0, 0] == sot and ids[1, 0] == sot
    n0 = len(mine.encode("X X annual crop land"))
    assert ids[0, 1:1 + n0].tolist() == mine.encode("X X annual crop land")
    assert ids[0, 1 + n0] == eot and ids[0, 2 + n0:].sum() == 0
    assert ids.argmax(-1).tolist() == [1 + n0, 1 + len(mine.encode("sea"))]  # EOT has the highest id
    with pytest.raises(RuntimeError):
        clip.tokenize(" ".join(["land"] * 100))
    assert clip.tokenize(" ".join(["land"] * 100), truncate=True)[0, -1] == eot


def test_preprocess_u8_is_the_transform_without_its_normalisation_tail():
    """clip.preprocess_u8() stops after resize + centre crop; normalize_u8 (the host restatement of what the
    device applies to uint8 pixels) then reproduces clip.load's transform bit for bit."""
    import importlib

    import numpy as np
    import torch
    from PIL import Image

    clip = importlib.import_module("menghini-neurips23-code_b200.clip")
    rng = np.random.RandomState(0)
    img = Image.fromarray(rng.randint(0, 256, (300, 260, 3), dtype=np.uint8))
    full = clip._preprocess()(img)
    raw = clip.preprocess_u8()(img)
    assert raw.dtype == torch.uint8 and tuple(raw.shape) == (3, 224, 224)
    assert torch.equal(clip.normalize_u8(raw), full)


@pytest.mark.parametrize("size", [(500, 375), (640, 480), (375, 500), (333, 517), (224, 224), (1024, 683),
                                  (225, 224), (300, 301)])
def test_preprocess_matches_torchvision_clip_transform(size):
    """clip.load()'s preprocess = openai/CLIP's torchvision pipeline Resize(224, bicubic) → CenterCrop(224) → RGB →
    ToTensor → Normalize, pixel for pixel (the long side is truncated, the crop origin rounded)."""
    import importlib

    import numpy as np
    from PIL import Image
    tv = pytest.importorskip("torchvision.transforms")

    clip = importlib.import_module("menghini-neurips23-code_b200.clip")
    rng = np.random.default_rng(size[0] * 1000 + size[1])
    img = Image.fromarray(rng.integers(0, 256, (size[1], size[0], 3), dtype=np.uint8))
    want = tv.Compose([
        tv.Resize(224, interpolation=tv.InterpolationMode.BICUBIC), tv.CenterCrop(224),
        lambda im: im.convert("RGB"), tv.ToTensor(),
        tv.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])(img)
    got = clip._preprocess()(img)
    assert got.shape == want.shape == (3, 224, 224)
    assert torch.allclose(got, want, atol=1e-6, rtol=0)
    assert torch.equal(clip.normalize_u8(clip.preprocess_u8()(img)), got)


def test_u8_normalise_fma_is_exhaustively_exact():
    """csrc/rowops.cu im2col_patch32_u8_kernel: fp16(fma(x, a_c, b_c)) with a_c = fp32(1/(255·std_c)),
    b_c = fp32(−mean_c/std_c) (constants formed in double) is, for every one of the 3 × 256 (channel, pixel value)
    pairs, the same fp16 number as torch's ToTensor (x/255) → Normalize ((· − mean)/std) in fp32 followed by the
    rounding to the fp16 GEMM operand."""
    import numpy as np
    import torch
    x = torch.arange(256, dtype=torch.float32)
    for mean, std in ((0.48145466, 0.26862954), (0.4578275, 0.26130258), (0.40821073, 0.27577711)):
        m, s = torch.tensor(mean, dtype=torch.float32), torch.tensor(std, dtype=torch.float32)
        want = ((x / 255.0 - m) / s).half()                       # torch's own operation order, then fp16
        a = np.float32(1.0 / (255.0 * float(s)))
        b = np.float32(-float(m) / float(s))
        # fp32 fma: the double product of two fp32 numbers is exact, the double sum is rounded once more to fp32
        got = (x.numpy().astype(np.float64) * np.float64(a) + np.float64(b)).astype(np.float32)
        assert np.array_equal(got.astype(np.float16).view(np.uint16), want.numpy().view(np.uint16)), (mean, std)
