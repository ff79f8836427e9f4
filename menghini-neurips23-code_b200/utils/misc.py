"""Small host helpers the strategies import from `utils` (utils/utils.py:37-45): only used when the reference's
own `utils` package is not importable (stand-alone use of this package)."""
import random

import numpy as np
import torch


def seed_worker(worker_id):
    worker_seed = torch.initial_seed() % 2 ** 32
    np.random.seed(worker_seed)
    random.seed(worker_seed)


class Config(object):
    def __init__(self, config):
        for k, v in config.items():
            setattr(self, k, v)
