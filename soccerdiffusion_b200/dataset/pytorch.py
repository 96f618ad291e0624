"""Normalizer (reference: soccer_diffusion/dataset/pytorch.py:401-414) on libsd_b200 kernels.

The SQLite dataset itself is out of scope (no dataset offline)."""
from __future__ import annotations

import torch

from soccerdiffusion_b200 import _lib, ops


class Normalizer:
    def __init__(self, mean: torch.Tensor, std: torch.Tensor):
        self.mean = mean
        self.std = std

    @classmethod
    def fit(cls, data: torch.Tensor):
        # one-off host-side statistics over <=1000 samples (train.py:108-110): plain torch reduction
        return cls(data.mean(dim=0), data.std(dim=0))

    def _apply(self, data: torch.Tensor, mode: int) -> torch.Tensor:
        _lib.require_cuda(data)
        J = data.shape[-1]
        mean = self.mean.to(device=data.device, dtype=torch.float32).reshape(-1)
        std = self.std.to(device=data.device, dtype=torch.float32).reshape(-1)
        if mean.numel() != J or std.numel() != J:
            raise RuntimeError(f"Normalizer expects per-joint mean/std of length {J}")
        x = data.float().contiguous()
        out = torch.empty_like(x)
        ops.affine_joints(x, mean.contiguous(), std.contiguous(), out, mode)
        return out

    def normalize(self, data: torch.Tensor):
        return self._apply(data, 0)

    def denormalize(self, data: torch.Tensor):
        return self._apply(data, 1)
