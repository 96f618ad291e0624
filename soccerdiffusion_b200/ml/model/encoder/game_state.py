"""GameStateEncoder (reference: soccer_diffusion/ml/model/encoder/game_state.py:7-27)."""
from enum import Enum

import torch
from torch import nn

from soccerdiffusion_b200 import _lib
from soccerdiffusion_b200.functional import GatherRowsFn


class RobotState(str, Enum):
    """The four robot states of the dataset schema (reference: dataset/models.py:13-25); only its
    length leaks into the hot path."""

    PLAYING = "PLAYING"
    POSITIONING = "POSITIONING"
    STOPPED = "STOPPED"
    UNKNOWN = "UNKNOWN"


class GameStateEncoder(nn.Module):
    def __init__(self, hidden_dim: int):
        super().__init__()
        self.embedding = nn.Embedding(len(RobotState), hidden_dim)  # parameter holder (N(0,1) init)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B,) int64 -> (B,1,d)"""
        w = self.embedding.weight
        _lib.require_cuda(x, w)
        return GatherRowsFn.apply(w, x.long().contiguous())
