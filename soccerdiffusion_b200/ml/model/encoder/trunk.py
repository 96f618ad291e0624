"""bf16 NHWC execution of the torchvision ResNet trunk with libsd_b200's fused BatchNorm(+residual)(+ReLU) and
max-pool kernels; the convolutions are cuDNN calls (``F.conv2d`` under bf16 autocast).

Semantics: torchvision ``ResNet._forward_impl`` up to ``layer4`` with ``BasicBlock`` / ``Bottleneck``
(reference call site ml/model/encoder/image.py:46-52, 55-73); BatchNorm in train mode uses batch statistics and
updates ``running_mean/var`` (momentum 0.1, unbiased variance) and ``num_batches_tracked`` like nn.BatchNorm2d.
Parameters are read from the torchvision modules (they stay the ``state_dict`` holders).
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from soccerdiffusion_b200 import ops, runtime

_USE_TC_STEM_WGRAD = os.environ.get("SD_B200_STEM_WGRAD", "tc") == "tc"
_USE_TC_STEM_FPROP = os.environ.get("SD_B200_STEM_FPROP", "tc") == "tc"


def _cl(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


class FusedBNAct(torch.autograd.Function):
    """y = relu?(BN(x) (+ residual)) on bf16 channels_last tensors."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, relu: bool, training: bool, momentum: float,
                eps: float, fork: bool = False):
        """``fork``: return the output TWICE (two tensors sharing storage).  A residual block's output feeds the next
        block's convolution and its identity path; handing each consumer its own output makes autograd deliver the two
        gradients separately, and the backward kernels sum them on the fly (sd_bn_bwd2_nhwc_bf16) instead of autograd
        running an elementwise add over the whole tensor first."""
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
        # 1 bit per element records where the ReLU passed: the backward reads R*C/8 bytes instead of y (2 B/element)
        mask = torch.empty(R * C // 8, device=dev, dtype=torch.uint8) if (relu and training) else None
        ops.bn_apply(x, res, mean, invstd, gamma, beta, relu, y, mask, R, C)
        ctx.save_for_backward(x, mask, mean, invstd, gamma)
        ctx.meta = (R, C, residual is not None, training)
        ctx.sums = sums
        if fork:
            ctx.set_materialize_grads(False)   # an unused twin yields None, not a zero tensor
            return y, y.view_as(y)
        return y

    @staticmethod
    def backward(ctx, dy, dy2=None):
        x, mask, mean, invstd, gamma = ctx.saved_tensors
        R, C, has_res, training = ctx.meta
        if not training:
            raise RuntimeError("FusedBNAct backward implements train-mode BatchNorm only")
        if dy is None:
            dy, dy2 = dy2, None
        if dy is None:
            return (None,) * 11
        dy = _cl(dy)
        dy2 = _cl(dy2) if dy2 is not None else None
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        dgamma = torch.empty(C, device=x.device, dtype=torch.float32)
        dbeta = torch.empty(C, device=x.device, dtype=torch.float32)
        ops.bn_bwd(dy, mask, x, mean, invstd, gamma, ctx.sums, dx, dres, dgamma, dbeta, R, C, dy2=dy2)
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None, None


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


class StemBNReLUPool(torch.autograd.Function):
    """maxpool3x3s2(relu(BN(x))) in one pass (torchvision ResNet stem after conv1)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training: bool, momentum: float, eps: float,
                presums=None):
        """``presums``: per-channel (sum x, sum x^2) already accumulated by the producer of x (the stem convolution's
        epilogue) — the statistics pass over x is skipped."""
        x = _cl(x)
        N, C, H, W = x.shape
        dev = x.device
        mean = torch.empty(C, device=dev, dtype=torch.float32)
        invstd = torch.empty(C, device=dev, dtype=torch.float32)
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        if training and presums is not None and presums.numel() == 2 * C:
            ops.bn_finalize(presums, N * H * W, C, eps, momentum, mean, invstd, running_mean, running_var)
        elif training:
            ops.bn_stats(x, N * H * W, C, sums, eps, momentum, mean, invstd, running_mean, running_var)
        else:
            mean.copy_(running_mean)
            torch.rsqrt(running_var + eps, out=invstd)
        HO, WO = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, C, HO, WO), device=dev, dtype=x.dtype, memory_format=torch.channels_last)
        idx = torch.empty((N, HO, WO, C), device=dev, dtype=torch.uint8)
        ops.stem_fwd(x, mean, invstd, gamma, beta, y, idx, N, H, W, C)
        ctx.save_for_backward(x, idx, mean, invstd, gamma, beta, y)   # y: the backward's reductions run in the pooled domain
        ctx.sums = sums
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, idx, mean, invstd, gamma, beta, y_pooled = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("StemBNReLUPool backward implements train-mode BatchNorm only")
        N, C, H, W = x.shape
        dy = _cl(dy)
        dgamma = torch.empty(C, device=x.device, dtype=torch.float32)
        dbeta = torch.empty(C, device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        if ops.stem_band_supported(H, W, C):
            # per-channel reductions over the POOLED tensors (dy, y: only arg-max pixels carry gradient), then one pass
            # that routes the pooled gradient to 2x2 pixel blocks, recomputes the ReLU mask from x and writes dx.  The
            # activated map's gradient is never materialised.
            ops.stem_bwd(dy, idx, x, mean, invstd, gamma, beta, ctx.sums, dx, dgamma, dbeta, N, H, W, C, y_pooled=y_pooled)
        else:
            # generic shapes: gradient w.r.t. the activated map, then BatchNorm backward with the recomputed ReLU mask
            dact = torch.empty_like(x)
            ops.maxpool_bwd(dy, idx, dact, N, H, W, C)
            ops.bn_bwd(dact, None, x, mean, invstd, gamma, ctx.sums, dx, None, dgamma, dbeta, N * H * W, C, beta_recompute=beta)
        return dx, dgamma, dbeta, None, None, None, None, None, None


def _bn(bn, x, residual=None, relu=True, fork=False):
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return FusedBNAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual, relu, bn.training,
                            float(momentum), float(bn.eps), fork)


def _conv(conv, x):
    if _trunk_conv_ok(conv, x):
        return TrunkConv.apply(x, conv.weight, conv.stride[0], conv.padding[0])
    if _own_wgrad_c64(conv, x):
        return Conv3x3C64.apply(x, conv.weight)
    return F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)


class DownsampleConv1x1S2(torch.autograd.Function):
    """The 1x1 stride-2 convolution of a ResNet stage's downsample path (torchvision resnet.py conv1x1(inplanes, planes,
    stride); image.py:55-73).  Forward and weight gradient stay library calls (0.05-0.15 ms each); the DATA gradient is
    libsd_b200's TMA-fed tcgen05 GEMM with a strided scatter epilogue (sd_conv1x1s2_dgrad_bf16): cuDNN serves this shape with
    a legacy sm80 kernel at 45 TFLOP/s (0.72 ms for the 64->128 stage at 2560 frames, tools/conv_probe.py)."""

    @staticmethod
    def forward(ctx, x, weight):
        x = _cl(x)
        Cout, Cin = weight.shape[0], weight.shape[1]
        wb = torch.empty((Cout, Cin), device=x.device, dtype=torch.bfloat16)
        ops.cast_bf16(weight.detach().contiguous().view(Cout, Cin), wb)
        w4 = wb.view(Cout, Cin, 1, 1).contiguous(memory_format=torch.channels_last)
        y = torch.ops.aten.convolution(x, w4, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1)
        ctx.save_for_backward(x, wb)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wb = ctx.saved_tensors
        dy = _cl(dy)
        n, Cin, H, W = x.shape
        Cout = wb.shape[0]
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)   # channels_last storage [n][H][W][Cin]
            ops.conv1x1s2_dgrad(dy, wb, dx, n, H, W, Cin, Cout)
        if ctx.needs_input_grad[1]:
            w4 = wb.view(Cout, Cin, 1, 1).contiguous(memory_format=torch.channels_last)
            dw = torch.ops.aten.convolution_backward(dy, x, w4, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1,
                                                     [False, True, False])[1].float().reshape(Cout, Cin, 1, 1)
        return dx, dw


class Conv3x3C64(torch.autograd.Function):
    """A 3x3 / stride 1 / padding 1 convolution with 64 -> 64 channels (ResNet18 layer1, torchvision resnet.py BasicBlock).
    Forward and data gradient stay library calls (cuDNN's weight-stationary kernels run them at 1.2 PFLOP/s); the WEIGHT
    gradient is libsd_b200's implicit GEMM over the pixels (sd_conv3x3_wgrad_c64_bf16: 0.55 ms vs 0.87 ms for cuDNN's
    64x64x64 kernel at 2560 frames of 56x56, tools/wgrad_micro.py)."""

    @staticmethod
    def forward(ctx, x, weight):
        x = _cl(x)
        wb = weight.detach().to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        y = torch.ops.aten.convolution(x, wb, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1)
        ctx.save_for_backward(x, wb)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wb = ctx.saved_tensors
        dy = _cl(dy)
        n, _, H, W = x.shape
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.ops.aten.convolution_backward(dy, x, wb, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                     [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            dw = torch.empty((64, 64, 3, 3), device=x.device, dtype=torch.float32)
            ops.conv3x3_wgrad_c64(x, dy, dw, n, H, W)
        return dx, dw


class TrunkConv(torch.autograd.Function):
    """A convolution of the trunk's residual stages (3x3 / 1x1, stride 1 / 2, no bias; torchvision resnet.py) for the bf16 path.
    Forward: library call on the bf16 copy of the weight.  Backward: the data gradient on the current stream
    (sd_conv1x1s2_dgrad_bf16 for the downsample convolutions, the library otherwise), the WEIGHT gradient
    (sd_conv3x3_wgrad_c64_bf16 for layer1, the library otherwise) — when gradients are accumulated straight into the optimizer's
    flat buffer (runtime.direct_grads) — on the side stream of runtime.wgrad_stream(): it is a leaf of the backward pass, so the
    tensor-bound weight-gradient GEMMs overlap the HBM-bound BatchNorm backward kernels of the main chain; the optimizer /
    all-reduce side joins through runtime.join_wgrad_stream()."""

    @staticmethod
    def forward(ctx, x, weight, stride: int, padding: int):
        x = _cl(x)
        wb = weight.detach().to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        y = torch.ops.aten.convolution(x, wb, None, [stride, stride], [padding, padding], [1, 1], False, [0, 0], 1)
        ctx.save_for_backward(x, wb)
        ctx.conv = (stride, padding)
        ctx.param = weight
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wb = ctx.saved_tensors
        stride, padding = ctx.conv
        weight = ctx.param
        dy = _cl(dy)
        n, Cin, H, W = x.shape
        Cout, k = wb.shape[0], wb.shape[2]
        args = ([stride, stride], [padding, padding], [1, 1], False, [0, 0], 1)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if _OWN_DS_DGRAD and k == 1 and stride == 2 and padding == 0 and ops.conv1x1s2_dgrad_supported(H, W, Cin, Cout):
                dx = torch.empty_like(x)
                ops.conv1x1s2_dgrad(dy, wb.view(Cout, Cin), dx, n, H, W, Cin, Cout)
            else:
                dx = torch.ops.aten.convolution_backward(dy, x, wb, None, *args, [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            own = (_OWN_WGRAD_C64 and k == 3 and stride == 1 and padding == 1 and Cin == 64 and Cout == 64
                   and ops.conv3x3_wgrad_c64_supported(H, W))
            g = weight.grad
            direct = (runtime.direct_grads() and weight.is_leaf and g is not None and g.dtype == torch.float32 and g.is_contiguous()
                      and g.shape == weight.shape)

            def compute(into):
                if own:
                    out = into if into is not None else torch.empty((64, 64, 3, 3), device=x.device, dtype=torch.float32)
                    ops.conv3x3_wgrad_c64(x, dy, out, n, H, W, accumulate=into is not None)
                    return out
                dwb = torch.ops.aten.convolution_backward(dy, x, wb, None, *args, [False, True, False])[1]
                if into is not None:
                    into.add_(dwb.reshape(into.shape))
                    return into
                return dwb.float().reshape(weight.shape)

            if direct and runtime.wgrad_overlap():
                cur = torch.cuda.current_stream()
                side = runtime.wgrad_stream(x.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    compute(g)
                for t in (dy, x, wb):
                    t.record_stream(side)
                runtime.mark_grad_written(weight)
            elif direct:
                compute(g)
                runtime.mark_grad_written(weight)
            else:
                dw = compute(None)
        return dx, dw, None, None


def _trunk_conv_ok(conv, x) -> bool:
    return (_TRUNK_CONV and torch.is_grad_enabled() and x.dtype == torch.bfloat16 and x.dim() == 4 and conv.bias is None
            and conv.groups == 1 and conv.dilation == (1, 1) and conv.weight.dtype == torch.float32
            and conv.kernel_size[0] == conv.kernel_size[1] and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
            and (x.requires_grad or conv.weight.requires_grad))


def _own_wgrad_c64(conv, x) -> bool:
    return (_OWN_WGRAD_C64 and torch.is_grad_enabled() and conv.weight.requires_grad and x.dtype == torch.bfloat16 and x.dim() == 4
            and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1)
            and conv.groups == 1 and conv.bias is None and conv.in_channels == 64 and conv.out_channels == 64
            and conv.weight.dtype == torch.float32 and ops.conv3x3_wgrad_c64_supported(x.shape[2], x.shape[3]))


def _own_ds_dgrad(conv, x) -> bool:
    return (_OWN_DS_DGRAD and torch.is_grad_enabled() and x.requires_grad and x.dtype == torch.bfloat16 and x.dim() == 4
            and conv.kernel_size == (1, 1) and conv.stride == (2, 2) and conv.padding == (0, 0) and conv.dilation == (1, 1)
            and conv.groups == 1 and conv.bias is None and conv.weight.dtype == torch.float32
            and ops.conv1x1s2_dgrad_supported(x.shape[2], x.shape[3], conv.in_channels, conv.out_channels))


def _stem_conv_s2d(conv, images: torch.Tensor) -> torch.Tensor:
    """conv1 (7x7, stride 2, pad 3, Cin=3) evaluated as a 4x4 stride-1 convolution over the 2x2 space-to-depth
    image (Cin = 3*2*2 = 12, zero-padded to 16): the same products and sums (the extra taps/channels are exact
    zeros), but a tensor-core friendly shape instead of cuDNN's Cin=3 -> 8 padding path.  The weight transform is
    ordinary differentiable torch code, so the gradient lands on the original (64,3,7,7) parameter."""
    N, Cin, H, W = images.shape
    Cout = conv.weight.shape[0]
    Hp, Wp = (H + 6) // 2, (W + 6) // 2
    if images.dtype == torch.float32 and not images.requires_grad:
        # one pass: pad + space-to-depth + bf16 + NHWC (channel = (c, dy, dx), 12 -> 16)
        xp = torch.empty((N, Hp, Wp, 16), device=images.device, dtype=torch.bfloat16)
        ops.stem_pack(images.contiguous(), xp, N, H, W)
        x = xp.permute(0, 3, 1, 2)                                                       # NCHW view of NHWC storage
    else:
        x = F.pad(images.to(torch.bfloat16), (3, 3, 3, 3))                               # (N,3,H+6,W+6)
        x = x.view(N, Cin, Hp, 2, Wp, 2).permute(0, 1, 3, 5, 2, 4).reshape(N, Cin * 4, Hp, Wp)
        x = F.pad(x, (0, 0, 0, 0, 0, 16 - Cin * 4)).contiguous(memory_format=torch.channels_last)
    w = F.pad(conv.weight, (0, 1, 0, 1))                                                # (Cout,3,8,8): zero last tap
    w = w.view(Cout, Cin, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(Cout, Cin * 4, 4, 4)
    w = F.pad(w, (0, 0, 0, 0, 0, 16 - Cin * 4))
    return F.conv2d(x, w, conv.bias, 1, 0)


class StemConvS2D(torch.autograd.Function):
    """conv1 in the space-to-depth formulation: forward and weight gradient on libsd_b200's TMA + tcgen05 kernels
    (sd_stem_fprop_s2d_bf16 / sd_stem_wgrad_s2d_bf16; cuDNN when the shape is not supported).  The packed image is kept
    for the backward pass (1.1 GB at bs=256) instead of being re-packed."""

    @staticmethod
    def forward(ctx, images, weight, want_stats: bool = False):
        """Returns (y, sums): ``sums`` (float64[2*Cout], sum y / sum y^2 per channel from the convolution's epilogue) when
        ``want_stats`` and the tcgen05 kernel ran, else an empty tensor."""
        with torch.no_grad():
            sums = torch.empty(2 * weight.shape[0], device=images.device, dtype=torch.float64) if want_stats else None
            y, xp, have = _stem_conv_s2d_raw(images, weight, return_packed=True, sums=sums)
            if not have:
                sums = torch.empty(0, device=images.device, dtype=torch.float64)
        ctx.save_for_backward(images, weight, xp)
        ctx.mark_non_differentiable(sums)
        return y, sums

    @staticmethod
    def backward(ctx, dy, _dsums=None):
        images, weight, xp = ctx.saved_tensors
        N, Cin, H, W = images.shape
        Cout = weight.shape[0]
        if Cout == 64 and _USE_TC_STEM_WGRAD:
            dws = torch.empty((256, 64), device=images.device, dtype=torch.float32)
            ops.stem_wgrad(xp, _cl(dy), dws, N, H, W)
            g = dws.view(4, 4, 16, Cout)[:, :, : Cin * 4]                  # (kh, kw, ci=(c,dy,dx), cout)
            g = g.reshape(4, 4, Cin, 2, 2, Cout).permute(5, 2, 0, 3, 1, 4)  # (cout, c, kh, dy, kw, dx)
            gw = g.reshape(Cout, Cin, 8, 8)[:, :, :7, :7]
            return None, gw.to(weight.dtype).contiguous(), None
        x = normalize_u8(images).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        gw = torch.ops.aten.convolution_backward(_cl(dy), x, weight.to(torch.bfloat16), None, (2, 2), (3, 3), (1, 1), False,
                                                 (0, 0), 1, (False, True, False))[1]
        return None, gw.to(weight.dtype), None


def normalize_u8(images: torch.Tensor) -> torch.Tensor:
    """Device-side replica of the reference's host preprocessing for raw uint8 frames (v2.ToDtype(float32, scale=True)
    -> v2.Normalize(ImageNet mean/std); dataset/pytorch.py:198-204): same fp32 operations, identity for float input."""
    if images.dtype != torch.uint8:
        return images
    mean = torch.tensor(ops.IMAGENET_MEAN, device=images.device, dtype=torch.float32).view(-1, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD, device=images.device, dtype=torch.float32).view(-1, 1, 1)
    return images.to(torch.float32).mul_(1.0 / 255).sub_(mean).div_(std)


def _stem_conv_s2d_raw(images, weight, return_packed: bool = False, sums=None):
    """conv1 on the packed image.  With ``return_packed``: (y, packed image, whether ``sums`` was filled)."""
    N, Cin, H, W = images.shape
    Cout = weight.shape[0]
    Hp, Wp = (H + 6) // 2, (W + 6) // 2
    xp = torch.empty((N, Hp, Wp, 16), device=images.device, dtype=torch.bfloat16)
    if images.dtype == torch.uint8:
        ops.stem_pack_u8(images.contiguous(), xp, N, H, W)   # ToDtype(scale) + Normalize fused into the packing pass
    else:
        ops.stem_pack(images.contiguous(), xp, N, H, W)
    w = F.pad(weight.to(torch.bfloat16), (0, 1, 0, 1))
    w = w.view(Cout, Cin, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(Cout, Cin * 4, 4, 4)
    w = F.pad(w, (0, 0, 0, 0, 0, 16 - Cin * 4))                      # (Cout, 16, 4, 4)
    if Cout == 64 and _USE_TC_STEM_FPROP:
        # TMA + tcgen05 kernel: weights as [cout][kh][kw][ci] = [64][256]
        w2 = w.permute(0, 2, 3, 1).reshape(Cout, 256).contiguous()
        y = torch.empty((N, Cout, H // 2, W // 2), device=images.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
        if ops.stem_fprop(xp, w2, y, N, H, W, sums=sums):
            return (y, xp, sums is not None) if return_packed else y
    y = F.conv2d(xp.permute(0, 3, 1, 2), w.contiguous(memory_format=torch.channels_last), None, 1, 0)
    return (y, xp, False) if return_packed else y


def _stem_is_s2d_compatible(conv, images) -> bool:
    return (conv.kernel_size == (7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3) and conv.dilation == (1, 1)
            and conv.groups == 1 and conv.in_channels == 3 and images.shape[-1] % 2 == 0 and images.shape[-2] % 2 == 0)


def supported(encoder) -> bool:
    from torchvision.models.resnet import BasicBlock, Bottleneck

    ok = all(isinstance(b, (BasicBlock, Bottleneck)) for layer in (encoder.layer1, encoder.layer2, encoder.layer3,
                                                                   encoder.layer4) for b in layer)
    mp = encoder.maxpool
    return ok and mp.kernel_size == 3 and mp.stride == 2 and mp.padding == 1 and isinstance(encoder.bn1, torch.nn.BatchNorm2d)


_TRUNK_CONV = os.environ.get("SD_B200_TRUNK_CONV", "1") == "1"   # one autograd node per trunk convolution (split data / weight gradients)
_OWN_WGRAD_C64 = os.environ.get("SD_B200_OWN_WGRAD_C64", "1") == "1"   # layer1 weight gradients on libsd_b200's implicit GEMM
_OWN_DS_DGRAD = os.environ.get("SD_B200_OWN_DS_DGRAD", "1") == "1"   # downsample 1x1/s2 data gradient on libsd_b200's GEMM
_BN_FORK = os.environ.get("SD_B200_BN_FORK", "1") == "1"   # twin block outputs: gradients summed inside the BN backward kernels
# how many cuDNN algorithms the autotuner times per convolution shape (torch default 10; 0 = all)
_CUDNN_BENCHMARK_LIMIT = int(os.environ.get("SD_B200_CUDNN_BENCHMARK_LIMIT", "10"))


def resnet_trunk_bf16(encoder, images: torch.Tensor) -> torch.Tensor:
    """(n,3,R,R) -> (n,C,h,w) bf16 channels_last: conv1..layer4 of a torchvision ResNet."""
    from torchvision.models.resnet import BasicBlock

    # cuDNN autotuning for the (fixed) convolution shapes of the trunk; restored on exit
    with torch.backends.cudnn.flags(enabled=True, benchmark=True, benchmark_limit=_CUDNN_BENCHMARK_LIMIT), \
            torch.autocast("cuda", dtype=torch.bfloat16):
        bn1 = encoder.bn1
        presums = None
        if _stem_is_s2d_compatible(encoder.conv1, images):
            if images.dtype in (torch.float32, torch.uint8) and not images.requires_grad and encoder.conv1.bias is None:
                x, presums = StemConvS2D.apply(images, encoder.conv1.weight, bool(bn1.training))
            else:
                x = _stem_conv_s2d(encoder.conv1, normalize_u8(images))
        else:
            x = _conv(encoder.conv1, normalize_u8(images).to(dtype=torch.bfloat16, memory_format=torch.channels_last))
        if bn1.training and bn1.track_running_stats and bn1.num_batches_tracked is not None:
            bn1.num_batches_tracked.add_(1)
        x = StemBNReLUPool.apply(x, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var, bn1.training,
                                 float(0.1 if bn1.momentum is None else bn1.momentum), float(bn1.eps), presums)
        blocks = [blk for layer in (encoder.layer1, encoder.layer2, encoder.layer3, encoder.layer4) for blk in layer]
        # a block's output feeds the next block's first convolution (x_main) AND its identity path (x_skip): with
        # ``fork`` the two consumers get twin outputs, so their gradients reach the BatchNorm backward separately and
        # are summed inside its kernels (no elementwise-add pass over the activation gradient)
        x_main = x_skip = x
        n12 = len(encoder.layer1) + len(encoder.layer2)
        for i, blk in enumerate(blocks):
            if i == n12 and torch.is_grad_enabled() and runtime.grad_ready_enabled() and x_main.requires_grad:
                # backward-pass milestone: once the gradient reaches this point, layer3, layer4 and everything after the
                # trunk (89 % of the parameters) have their gradients — the data-parallel all-reduce of that bucket starts
                # here and overlaps the backward pass of layer2, layer1 and the stem
                from soccerdiffusion_b200.functional import GradReadyFn

                same = x_main is x_skip
                st = {"need": 1 if same else 2, "seen": 0}
                x_main = GradReadyFn.apply(x_main, "trunk.layer3_onward", st)
                x_skip = x_main if same else GradReadyFn.apply(x_skip, "trunk.layer3_onward", st)
            fork = _BN_FORK and torch.is_grad_enabled() and i + 1 < len(blocks)
            identity = x_skip
            if blk.downsample is not None:
                ds = blk.downsample[0]
                if _trunk_conv_ok(ds, x_skip):
                    ds_out = _conv(ds, x_skip)
                else:
                    ds_out = DownsampleConv1x1S2.apply(x_skip, ds.weight) if _own_ds_dgrad(ds, x_skip) else _conv(ds, x_skip)
                identity = _bn(blk.downsample[1], ds_out, None, False)
            out = _bn(blk.bn1, _conv(blk.conv1, x_main), None, True)
            if isinstance(blk, BasicBlock):
                res = _bn(blk.bn2, _conv(blk.conv2, out), identity, True, fork and blk.bn2.training)
            else:
                out = _bn(blk.bn2, _conv(blk.conv2, out), None, True)
                res = _bn(blk.bn3, _conv(blk.conv3, out), identity, True, fork and blk.bn3.training)
            x_main, x_skip = res if isinstance(res, tuple) else (res, res)
        x = x_main
    return x
