"""Image encoders (reference: soccer_diffusion/ml/model/encoder/image.py:11-174).

In scope here: the per-frame token head (``Conv2d(512->32,1x1)`` replacing avgpool,
flatten, ``Linear(1568->d)``; or ``Linear(512->d)`` after the stock avgpool) and the transformer over
the frame sequence — both run on libsd_b200 GEMMs.  The convolutional / Swin *trunk* (a8', torchvision)
stays a cuDNN library call in this round (channels_last, bf16 autocast in bf16 mode) — it is row (f)-1
of the scope table.
"""
from __future__ import annotations

from enum import Enum

import torch
from torch import nn
from torchvision.models import resnet18, resnet50, swin_s, swin_t

from soccerdiffusion_b200 import _lib, ops, runtime
from soccerdiffusion_b200.functional import LinearFn
from soccerdiffusion_b200.ml.model.encoder.base import BaseEncoder

# ImageNet weights need a download (image.py:64,66).  Offline they are skipped unless a local
# torchvision cache holds them; set SD_B200_PRETRAINED_TRUNK=1 to request them like the reference.
import os

_PRETRAINED = os.environ.get("SD_B200_PRETRAINED_TRUNK", "0") == "1"


class ImageEncoderType(Enum):
    RESNET18 = "resnet18"
    RESNET50 = "resnet50"
    SWIN_TRANSFORMER_TINY = "swin_transformer_tiny"
    SWIN_TRANSFORMER_SMALL = "swin_transformer_small"


class SequenceEncoderType(Enum):
    TRANSFORMER = "transformer"
    NONE = "none"


import contextlib


@contextlib.contextmanager
def _trunk_autocast():
    """Arithmetic mode of the library trunk: bf16 autocast in bf16 mode; true fp32 (TF32 convolutions off —
    they cost ~1e-3) in the 1e-4 fp32 mode."""
    if runtime.get_precision() == ops.PREC_BF16:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yield
    else:
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            yield


class AbstractImageEncoder(nn.Module):
    encoder: nn.Module

    def features(self, images: torch.Tensor) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, F, 3, R, R) -> (B, F, d).  Besides the reference's normalised float32 frames, raw uint8 frames are accepted:
        the host preprocessing (ToDtype(scale) + Normalize, dataset/pytorch.py:198-204) then runs on the device, fused
        into the stem's packing kernel in bf16 mode."""
        _lib.require_cuda(x)
        images = x.reshape(-1, *x.shape[2:])
        tokens = self.tokens(images)
        return tokens.view(x.shape[0], x.shape[1], -1)


class ResNetImageEncoder(AbstractImageEncoder):
    def __init__(self, resnet_type: ImageEncoderType, hidden_dim: int, use_final_avgpool: bool, resolution: int):
        super().__init__()
        weights = "DEFAULT" if _PRETRAINED else None
        match resnet_type:
            case ImageEncoderType.RESNET18:
                self.encoder = resnet18(weights=weights)
            case ImageEncoderType.RESNET50:
                self.encoder = resnet50(weights=weights)
            case _:
                raise ValueError(f"Invalid ResNet type: {resnet_type}")
        self.use_final_avgpool = use_final_avgpool
        if use_final_avgpool:
            self.encoder.fc = nn.Linear(self.encoder.fc.in_features, hidden_dim)
        else:
            self.encoder.avgpool = nn.Conv2d(self.encoder.fc.in_features, 32, 1)
            self.encoder.fc = nn.Linear(ResNetImageEncoder.calculate_output_size(resolution) ** 2 * 32, hidden_dim)

    @staticmethod
    def calculate_output_size(resolution):
        resolution = (resolution - 7 + 2 * 3) // 2 + 1   # conv1
        resolution = (resolution - 3 + 2 * 1) // 2 + 1   # maxpool
        return resolution // 2 // 2 // 2                 # layer2..4

    def trunk(self, images: torch.Tensor) -> torch.Tensor:
        """torchvision/cuDNN conv stack -> (n, C, h, w), channels_last."""
        e = self.encoder
        if runtime.get_precision() == ops.PREC_BF16 and runtime.fused_trunk() and (self.training or not torch.is_grad_enabled()):
            from soccerdiffusion_b200.ml.model.encoder import trunk as _trunk

            if _trunk.supported(e):
                return _trunk.resnet_trunk_bf16(e, images)
        from soccerdiffusion_b200.ml.model.encoder.trunk import normalize_u8

        images = normalize_u8(images).contiguous(memory_format=torch.channels_last)   # raw uint8 frames: preprocess on device
        with _trunk_autocast():
            x = e.maxpool(e.relu(e.bn1(e.conv1(images))))
            x = e.layer4(e.layer3(e.layer2(e.layer1(x))))
        return x

    def tokens(self, images: torch.Tensor) -> torch.Tensor:
        e = self.encoder
        prec = runtime.get_precision()
        feat = self.trunk(images)
        n, c, h, w = feat.shape
        if self.use_final_avgpool:
            with _trunk_autocast():
                pooled = torch.flatten(e.avgpool(feat), 1)
            return LinearFn.apply(prec, pooled.float().contiguous(), e.fc.weight, e.fc.bias)
        # 1x1 conv == GEMM over NHWC pixels; flatten order of the reference is (c_out, pixel)
        pix = feat.permute(0, 2, 3, 1).float().contiguous().view(n * h * w, c)
        co = e.avgpool.weight.shape[0]
        y = LinearFn.apply(prec, pix, e.avgpool.weight.view(co, c), e.avgpool.bias)      # (n*hw, 32)
        y = y.view(n, h * w * co)                                                        # (pixel, c_out) order
        d = e.fc.weight.shape[0]
        if e.fc.weight.shape[1] != h * w * co:
            raise RuntimeError(f"image_resolution mismatch: fc expects {e.fc.weight.shape[1]} features, trunk gives {h*w*co}")
        w_fc = e.fc.weight.view(d, co, h * w).permute(0, 2, 1).reshape(d, h * w * co)   # re-ordered to match
        return LinearFn.apply(prec, y, w_fc.contiguous(), e.fc.bias)


class SwinTransformerImageEncoder(AbstractImageEncoder):
    def __init__(self, swin_type: ImageEncoderType, hidden_dim: int):
        super().__init__()
        match swin_type:
            case ImageEncoderType.SWIN_TRANSFORMER_TINY:
                self.encoder = swin_t()
            case ImageEncoderType.SWIN_TRANSFORMER_SMALL:
                self.encoder = swin_s()
            case _:
                raise ValueError(f"Invalid Swin Transformer type: {swin_type}")
        self.encoder.head = nn.Linear(self.encoder.head.in_features, hidden_dim)

    def tokens(self, images: torch.Tensor) -> torch.Tensor:
        from soccerdiffusion_b200.ml.model.encoder.trunk import normalize_u8

        e = self.encoder
        images = normalize_u8(images)
        with _trunk_autocast():
            x = e.flatten(e.avgpool(e.permute(e.norm(e.features(images)))))
        return LinearFn.apply(runtime.get_precision(), x.float().contiguous(), e.head.weight, e.head.bias)


class TransformerImageSequenceEncoder(nn.Module):
    def __init__(self, image_encoder: AbstractImageEncoder, hidden_dim: int, num_layers: int, max_seq_len: int):
        super().__init__()
        self.image_encoder = image_encoder
        self.transformer_encoder = BaseEncoder(
            input_dim=hidden_dim,
            patch_size=1,
            hidden_dim=hidden_dim,
            num_layers=num_layers,
            num_heads=8,
            max_seq_len=max_seq_len,
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.transformer_encoder(self.image_encoder(x))


def image_encoder_factory(
    encoder_type: ImageEncoderType, hidden_dim: int, use_final_avgpool: bool, resolution: int
) -> AbstractImageEncoder:
    if encoder_type in [ImageEncoderType.RESNET18, ImageEncoderType.RESNET50]:
        return ResNetImageEncoder(encoder_type, hidden_dim, use_final_avgpool, resolution)
    if encoder_type in [ImageEncoderType.SWIN_TRANSFORMER_TINY, ImageEncoderType.SWIN_TRANSFORMER_SMALL]:
        return SwinTransformerImageEncoder(encoder_type, hidden_dim)
    raise ValueError(f"Invalid image encoder type: {encoder_type}")


def image_sequence_encoder_factory(
    encoder_type: SequenceEncoderType,
    image_encoder_type: ImageEncoderType,
    hidden_dim: int,
    num_layers: int,
    max_seq_len: int,
    use_final_avgpool: bool,
    resolution: int,
):
    image_encoder = image_encoder_factory(image_encoder_type, hidden_dim, use_final_avgpool, resolution)
    match encoder_type:
        case SequenceEncoderType.TRANSFORMER:
            return TransformerImageSequenceEncoder(image_encoder, hidden_dim, num_layers, max_seq_len)
        case SequenceEncoderType.NONE:
            return image_encoder
        case _:
            raise ValueError(f"Invalid sequence encoder type: {encoder_type}")
