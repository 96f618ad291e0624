"""DiffusionActionGenerator (reference: soccer_diffusion/ml/model/decoder.py:6-54)."""
from __future__ import annotations

import torch
from torch import nn

from soccerdiffusion_b200 import _lib, runtime
from soccerdiffusion_b200.functional import DenoiserFn
from soccerdiffusion_b200.ml.model.misc import PositionalEncoding
from soccerdiffusion_b200.ml.model.params import DecoderLayerParams, LayerStack


class DiffusionActionGenerator(nn.Module):
    """Linear(J->d) + PE + pre-LN decoder layers cross-attending the context + Linear(d->J)."""

    def __init__(self, num_joints, hidden_dim, num_layers, num_heads, max_seq_len):
        super().__init__()
        if hidden_dim % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        self.embedding = nn.Linear(num_joints, hidden_dim)
        self.positional_encoding = PositionalEncoding(hidden_dim, max_seq_len)
        self.transformer_decoder = LayerStack([DecoderLayerParams(hidden_dim, hidden_dim) for _ in range(num_layers)])
        self.fc_out = nn.Linear(hidden_dim, num_joints)
        self.num_heads = num_heads
        self.hidden_dim = hidden_dim
        self.num_joints = num_joints
        self.max_seq_len = max_seq_len

    def forward(self, x, context):
        """x (B,T,J) noisy actions; context (B,M,d) memory -> (B,T,J) predicted noise."""
        _lib.require_cuda(x, context, self.embedding.weight)
        B, T, J = x.shape
        Mm = context.shape[1]
        if T > self.positional_encoding.pe.shape[1]:
            raise RuntimeError("trajectory longer than max_seq_len of the positional encoding")
        cfg = runtime.make_cfg(self.training)
        return DenoiserFn.apply(cfg, B, T, Mm, self.num_heads, self.positional_encoding.table(T).contiguous(),
                                x.float(), context.float(), self.embedding.weight, self.embedding.bias,
                                self.fc_out.weight, self.fc_out.bias, *self.transformer_decoder.tensors())
