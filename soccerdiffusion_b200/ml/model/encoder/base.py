"""BaseEncoder (reference: soccer_diffusion/ml/model/encoder/base.py:7-53)."""
from __future__ import annotations

import torch
from torch import nn

from soccerdiffusion_b200 import _lib, runtime
from soccerdiffusion_b200.functional import EncoderStackFn
from soccerdiffusion_b200.ml.model.misc import PositionalEncoding
from soccerdiffusion_b200.ml.model.params import EncoderLayerParams, LayerStack


class BaseEncoder(nn.Module):
    """Conv1d(in->d, kernel=stride=patch) + PE + ``num_layers`` pre-LN encoder layers (GELU, FFN width d)."""

    def __init__(
        self, input_dim: int, patch_size: int, hidden_dim: int, num_layers: int, num_heads: int, max_seq_len: int
    ):
        super().__init__()
        if hidden_dim % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        # nn.Conv1d is kept as the parameter holder (weight (d,in,p), bias (d)) with torch's init
        self.embedding = nn.Conv1d(input_dim, hidden_dim, kernel_size=patch_size, stride=patch_size)
        self.positional_encoding = PositionalEncoding(hidden_dim, max_seq_len)
        self.transformer_encoder = LayerStack([EncoderLayerParams(hidden_dim, hidden_dim) for _ in range(num_layers)])
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.hidden_dim = hidden_dim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, S, input_dim) -> (B, S // patch, hidden_dim)"""
        w = self.embedding.weight
        _lib.require_cuda(x, w)
        B, S, cin = x.shape
        p, d = self.patch_size, self.hidden_dim
        S2 = (S - p) // p + 1
        if x.dtype != torch.float32:
            x = x.float()
        # token j sees x[b, j*p:(j+1)*p, :] flattened (k, c); the conv kernel is permuted to match
        x2 = x[:, : S2 * p].reshape(B * S2, p * cin)
        w2 = w.permute(0, 2, 1).reshape(d, p * cin)
        if S2 > self.positional_encoding.pe.shape[1]:
            raise RuntimeError("sequence longer than max_seq_len of the positional encoding")
        cfg = runtime.make_cfg(self.training)
        return EncoderStackFn.apply(cfg, B, S2, self.num_heads, self.positional_encoding.table(S2).contiguous(),
                                    x2.contiguous(), w2.contiguous(), self.embedding.bias,
                                    *self.transformer_encoder.tensors())
