"""Minimal stand-in for `accelerate.Accelerator` (accelerate==0.15.0 in the reference's requirements.txt:1, absent
from this image): exactly the surface the reference's methods/ and utils/ touch — `prepare`, `backward`,
`wait_for_everyone`, `gather`, `unwrap_model`, `free_memory`, `is_local_main_process`, `device`.
One process per GPU: when torch.distributed is initialised `gather` is an all-gather and `backward` averages
the trainable gradients over the ranks (what DDP does for the 32 KiB prompt tensor); otherwise single process.
`dropin.install()` registers it as `accelerate` only when the real package is not importable."""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn


class _Prepared(nn.Module):
    """What `prepare` returns for a model on CUDA: the reference reads `model.module.classes` there
    (textual_prompt.py:86-97), i.e. it expects a DistributedDataParallel-style wrapper."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


class _DeviceLoader:
    """DataLoader whose tensor batch members arrive on the device (accelerate's prepared loader does that)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, device
        self.dataset = loader.dataset
        self.batch_size = loader.batch_size

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            yield tuple(b.to(self.device, non_blocking=True) if torch.is_tensor(b) else b for b in batch)


class Accelerator:
    def __init__(self, *args, **kwargs):
        self._models = []

    @property
    def device(self):
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    @property
    def num_processes(self):
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    @property
    def is_local_main_process(self):
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0

    is_main_process = is_local_main_process

    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, nn.Module) and not isinstance(o, _Prepared) and torch.cuda.is_available():
                o = _Prepared(o)
                self._models.append(o)
            elif isinstance(o, torch.utils.data.DataLoader):
                o = _DeviceLoader(o, self.device)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def unwrap_model(self, model):
        return model.module if isinstance(model, _Prepared) else model

    def backward(self, loss, **kwargs):
        loss.backward(**kwargs)
        if self.num_processes > 1:
            for m in self._models:
                for p in m.parameters():
                    if p.requires_grad and p.grad is not None:
                        dist.all_reduce(p.grad)
                        p.grad /= self.num_processes

    def wait_for_everyone(self):
        if self.num_processes > 1:
            dist.barrier()

    def gather(self, tensor):
        if self.num_processes == 1:
            return tensor
        parts = [torch.empty_like(tensor) for _ in range(self.num_processes)]
        dist.all_gather(parts, tensor.contiguous())
        return torch.cat(parts)

    def free_memory(self):
        self._models = [m for m in self._models if m is not None][-2:]

    def print(self, *args, **kwargs):
        if self.is_local_main_process:
            print(*args, **kwargs)
