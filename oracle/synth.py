"""Seeded synthetic inputs shared by the golden-vector generator, the parity tests and bench.py
(SURVEY.md §8d).  TEST INFRASTRUCTURE ONLY.  Everything is a pure function of its seed on the
torch CPU generator, so /root/reference never has to exist where the tests run."""
from __future__ import annotations

import numpy as np
import torch

WORDS = ["annual", "crop", "land", "forest", "herbaceous", "vegetation", "highway", "road",
         "industrial", "buildings", "pasture", "permanent", "residential", "river", "sea", "lake",
         "airplane", "airport", "baseball", "diamond", "beach", "bridge", "chaparral", "church",
         "cloud", "desert", "freeway", "golf", "course", "harbor", "island", "meadow", "mountain",
         "palace", "railway", "runway", "stadium", "terrace", "wetland", "boeing", "airbus"]


def class_names(c: int, seed: int = 1):
    """c synthetic class names of 1..4 words ('_'-joined like the reference's class files)."""
    rng = np.random.RandomState(seed)
    names = []
    for i in range(c):
        n = int(rng.randint(1, 5))
        ws = [WORDS[int(rng.randint(0, len(WORDS)))] for _ in range(n)]
        names.append("_".join(ws) + f"_{i}")
    return names


def images(b: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, 3, 224, 224, generator=g)


def text_prefix(p: int, seed: int = 2) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return 0.02 * torch.randn(1, p, 512, generator=g)


def image_prefix(p: int, seed: int = 3) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (768 ** -0.5) * torch.randn(p, 768, generator=g)


def pool(n: int, c: int, seed_f: int = 5, seed_t: int = 6, peaked: float = 0.0):
    """Unit-norm feature rows F[n,512] and prototypes T[c,512] (fp32).  `peaked` > 0 pulls each
    image toward a random prototype so arg-max classes are well separated."""
    gf = torch.Generator().manual_seed(seed_f)
    gt = torch.Generator().manual_seed(seed_t)
    t = torch.randn(c, 512, generator=gt)
    t = t / t.norm(dim=1, keepdim=True)
    f = torch.randn(n, 512, generator=gf)
    if peaked > 0:
        cls = torch.randint(0, c, (n,), generator=gf)
        f = f / f.norm(dim=1, keepdim=True) + peaked * t[cls]
    f = f / f.norm(dim=1, keepdim=True)
    return f, t


def path_ranks(n: int, seed: int = 7) -> np.ndarray:
    """Tie-break rank of each image = position of its path in ascending string order; a seeded
    permutation stands in for real file names."""
    return np.random.RandomState(seed).permutation(n).astype(np.int64)
