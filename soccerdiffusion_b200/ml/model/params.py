"""Parameter containers with the reference's state_dict names and torch's default initialisation.

The reference builds ``nn.TransformerEncoder/Decoder`` (encoder/base.py:29-40, decoder.py:25-35); the
names below (``layers.{i}.self_attn.in_proj_weight`` ... SURVEY.md §8 N2) are the on-disk checkpoint
contract.  These modules only HOLD parameters: the arithmetic is in libsd_b200 (functional.py).
"""
from __future__ import annotations

import math

import torch
from torch import nn


class _Affine(nn.Module):
    """weight/bias holder (LayerNorm or Linear); never called."""

    def __init__(self, weight_shape, bias_shape, kind: str):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(weight_shape))
        self.bias = nn.Parameter(torch.empty(bias_shape))
        if kind == "norm":
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)
        else:  # nn.Linear.reset_parameters
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            bound = 1.0 / math.sqrt(weight_shape[1])
            nn.init.uniform_(self.bias, -bound, bound)


class AttentionParams(nn.Module):
    """nn.MultiheadAttention parameters: packed in_proj (xavier_uniform, zero bias), out_proj (zero bias)."""

    def __init__(self, d: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = _Affine((d, d), (d,), "linear")
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def tensors(self):
        return [self.in_proj_weight, self.in_proj_bias, self.out_proj.weight, self.out_proj.bias]


class EncoderLayerParams(nn.Module):
    def __init__(self, d: int, ff: int):
        super().__init__()
        self.self_attn = AttentionParams(d)
        self.linear1 = _Affine((ff, d), (ff,), "linear")
        self.linear2 = _Affine((d, ff), (d,), "linear")
        self.norm1 = _Affine((d,), (d,), "norm")
        self.norm2 = _Affine((d,), (d,), "norm")

    def tensors(self):
        """Order expected by functional.EncoderStackFn."""
        return [*self.self_attn.tensors(), self.linear1.weight, self.linear1.bias, self.linear2.weight,
                self.linear2.bias, self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias]


class DecoderLayerParams(nn.Module):
    def __init__(self, d: int, ff: int):
        super().__init__()
        self.self_attn = AttentionParams(d)
        self.multihead_attn = AttentionParams(d)
        self.linear1 = _Affine((ff, d), (ff,), "linear")
        self.linear2 = _Affine((d, ff), (d,), "linear")
        self.norm1 = _Affine((d,), (d,), "norm")
        self.norm2 = _Affine((d,), (d,), "norm")
        self.norm3 = _Affine((d,), (d,), "norm")

    def tensors(self):
        """Order expected by functional.DenoiserFn."""
        return [*self.self_attn.tensors(), *self.multihead_attn.tensors(), self.linear1.weight, self.linear1.bias,
                self.linear2.weight, self.linear2.bias, self.norm1.weight, self.norm1.bias, self.norm2.weight,
                self.norm2.bias, self.norm3.weight, self.norm3.bias]


class LayerStack(nn.Module):
    """Holds ``layers.{i}`` like nn.TransformerEncoder / nn.TransformerDecoder (no final norm)."""

    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def tensors(self):
        out = []
        for layer in self.layers:
            out.extend(layer.tensors())
        return out
