"""bf16 NHWC execution of the torchvision ResNet trunk with libsd_b200's fused BatchNorm(+residual)(+ReLU) and
max-pool kernels; the convolutions are cuDNN calls (``F.conv2d`` under bf16 autocast).

Semantics: torchvision ``ResNet._forward_impl`` up to ``layer4`` with ``BasicBlock`` / ``Bottleneck``
(reference call site ml/model/encoder/image.py:46-52, 55-73); BatchNorm in train mode uses batch statistics and
updates ``running_mean/var`` (momentum 0.1, unbiased variance) and ``num_batches_tracked`` like nn.BatchNorm2d.
Parameters are read from the torchvision modules (they stay the ``state_dict`` holders).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from soccerdiffusion_b200 import ops


def _cl(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


class FusedBNAct(torch.autograd.Function):
    """y = relu?(BN(x) (+ residual)) on bf16 channels_last tensors."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, relu: bool, training: bool, momentum: float,
                eps: float):
        x = _cl(x)
        N, C, H, W = x.shape
        R = N * H * W
        dev = x.device
        mean = torch.empty(C, device=dev, dtype=torch.float32)
        invstd = torch.empty(C, device=dev, dtype=torch.float32)
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        if training:
            ops.bn_stats(x, R, C, sums, eps, momentum, mean, invstd, running_mean, running_var)
        else:
            mean.copy_(running_mean)
            torch.rsqrt(running_var + eps, out=invstd)
        res = _cl(residual) if residual is not None else None
        y = torch.empty_like(x)
        ops.bn_apply(x, res, mean, invstd, gamma, beta, relu, y, R, C)
        ctx.save_for_backward(x, y if relu else None, mean, invstd, gamma)
        ctx.meta = (R, C, residual is not None, training)
        ctx.sums = sums
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, invstd, gamma = ctx.saved_tensors
        R, C, has_res, training = ctx.meta
        if not training:
            raise RuntimeError("FusedBNAct backward implements train-mode BatchNorm only")
        dy = _cl(dy)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        dgamma = torch.empty(C, device=x.device, dtype=torch.float32)
        dbeta = torch.empty(C, device=x.device, dtype=torch.float32)
        ops.bn_bwd(dy, y, x, mean, invstd, gamma, ctx.sums, dx, dres, dgamma, dbeta, R, C)
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None


class MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _cl(x)
        N, C, H, W = x.shape
        HO, WO = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, C, HO, WO), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        idx = torch.empty((N, HO, WO, C), device=x.device, dtype=torch.uint8)
        ops.maxpool_fwd(x, y, idx, N, H, W, C)
        ctx.save_for_backward(idx)
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, C, H, W = ctx.shape
        dy = _cl(dy)
        dx = torch.empty((N, C, H, W), device=dy.device, dtype=dy.dtype, memory_format=torch.channels_last)
        ops.maxpool_bwd(dy, idx, dx, N, H, W, C)
        return dx


def _bn(bn, x, residual=None, relu=True):
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return FusedBNAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, relu, bn.training,
                            float(momentum), float(bn.eps))


def _conv(conv, x):
    return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)


def supported(encoder) -> bool:
    from torchvision.models.resnet import BasicBlock, Bottleneck

    ok = all(isinstance(b, (BasicBlock, Bottleneck)) for layer in (encoder.layer1, encoder.layer2, encoder.layer3,
                                                                   encoder.layer4) for b in layer)
    mp = encoder.maxpool
    return ok and mp.kernel_size == 3 and mp.stride == 2 and mp.padding == 1 and isinstance(encoder.bn1, torch.nn.BatchNorm2d)


def resnet_trunk_bf16(encoder, images: torch.Tensor) -> torch.Tensor:
    """(n,3,R,R) -> (n,C,h,w) bf16 channels_last: conv1..layer4 of a torchvision ResNet."""
    from torchvision.models.resnet import BasicBlock

    # cuDNN autotuning for the (fixed) convolution shapes of the trunk; restored on exit
    with torch.backends.cudnn.flags(enabled=True, benchmark=True), torch.autocast("cuda", dtype=torch.bfloat16):
        x = _conv(encoder.conv1, images.to(dtype=torch.bfloat16, memory_format=torch.channels_last))
        x = _bn(encoder.bn1, x, None, True)
        x = MaxPool3x3s2.apply(x)
        for layer in (encoder.layer1, encoder.layer2, encoder.layer3, encoder.layer4):
            for blk in layer:
                identity = x
                if blk.downsample is not None:
                    identity = _bn(blk.downsample[1], _conv(blk.downsample[0], x), None, False)
                if isinstance(blk, BasicBlock):
                    out = _bn(blk.bn1, _conv(blk.conv1, x), None, True)
                    x = _bn(blk.bn2, _conv(blk.conv2, out), identity, True)
                else:
                    out = _bn(blk.bn1, _conv(blk.conv1, x), None, True)
                    out = _bn(blk.bn2, _conv(blk.conv2, out), None, True)
                    x = _bn(blk.bn3, _conv(blk.conv3, out), identity, True)
    return x
