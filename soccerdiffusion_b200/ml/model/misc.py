"""StepToken and PositionalEncoding (reference: soccer_diffusion/ml/model/misc.py:6-65)."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from soccerdiffusion_b200 import _lib, ops


def positional_encoding_table(d_model: int, max_len: int) -> torch.Tensor:
    """Host-side constant table, same float32 op order as misc.py:51-56."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def step_token_frequencies(dim: int) -> torch.Tensor:
    """misc.py:31-32 evaluated once on the host with the same torch ops (int64 arange x numpy
    float64 scalar -> float32), so the 32-entry table is bit-identical to the reference's."""
    half_dim = dim // 4
    return torch.exp(torch.arange(half_dim) * -np.log(10000) / (half_dim - 1))


class _StepTokenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, steps, freqs, token, dim):
        B = steps.shape[0]
        out = torch.empty((B, 1, dim), device=token.device, dtype=torch.float32)
        ops.step_token(steps, freqs, token, out, dim, B, dim)
        ctx.dim, ctx.B = dim, B
        ctx.save_for_backward(token)
        return out

    @staticmethod
    def backward(ctx, dout):
        (token,) = ctx.saved_tensors
        dtoken = torch.zeros_like(token)
        dout = dout.contiguous()
        ops.step_token_bwd(dout.data_ptr(), ctx.dim, ctx.B, ctx.dim, dtoken)
        return None, None, dtoken, None


def normalize_steps(steps: torch.Tensor, device) -> torch.Tensor:
    """int64 (train.py:210, ros.py:306) or float (distill.py:194) steps; other dtypes are converted the
    way ``steps[:, None] * emb`` would promote them."""
    if steps.dtype in (torch.int64, torch.float32):
        s = steps
    elif steps.dtype.is_floating_point:
        s = steps.float()
    else:
        s = steps.long()
    return s.to(device).contiguous()


class StepToken(nn.Module):
    """[sin(t f_k) | cos(t f_k) | learnable(dim/2)] -> (B,1,dim)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.token = nn.Parameter(torch.randn(1, dim // 2))
        self.register_buffer("freqs", step_token_frequencies(dim), persistent=False)

    def forward(self, steps: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(self.token)
        return _StepTokenFn.apply(normalize_steps(steps, self.token.device), self.freqs, self.token, self.dim)


class PositionalEncoding(nn.Module):
    """x + pe[:, :S]; ``pe`` is a non-persistent buffer (not in the state_dict), shape (1,max_len,d)."""

    def __init__(self, d_model, max_len):
        super().__init__()
        self.register_buffer("pe", positional_encoding_table(d_model, max_len).unsqueeze(0), persistent=False)

    def table(self, S: int) -> torch.Tensor:
        return self.pe[0, :S]

    def forward(self, x):
        # standalone use only (the stacks fuse PE into the embedding GEMM epilogue)
        _lib.require_cuda(x)
        B, S, d = x.shape
        out = x.contiguous().clone()
        ops.copy_rows(self.pe.data_ptr(), 0, d, out.data_ptr(), S * d, d, B, S, d, accumulate=True)
        return out
