"""Autograd nodes of the hot path.  One ``torch.autograd.Function`` per transformer *stack* (an
encoder with its embedding, the whole denoiser), so that residual joins, LayerNorm, bias and
activation never appear as separate framework kernels: forward and backward are sequences of
libsd_b200 launches.

Reference semantics restated here (paths under /root/reference/soccer_diffusion/):
  BaseEncoder.forward            ml/model/encoder/base.py:41-53
  DiffusionActionGenerator       ml/model/decoder.py:38-54
  pre-LN encoder / decoder layer torch/nn/modules/transformer.py:944-950, 1131-1143
  packed in_proj of MHA          torch/nn/functional.py:5798-5856
  context concat + StepToken     ml/model/model.py:173-176, ml/model/misc.py:25-35
  mse_loss / add_noise           ml/training/train.py:218,229
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops
from .ops import ACT_GELU, KM, KN, MK, NK

ENC_PARAMS_PER_LAYER = 12   # in_w,in_b,out_w,out_b,l1_w,l1_b,l2_w,l2_b,n1_w,n1_b,n2_w,n2_b
DEC_PARAMS_PER_LAYER = 18   # sa(4) ca(4) l1_w,l1_b,l2_w,l2_b n1(2) n2(2) n3(2)


@dataclass
class RunCfg:
    precision: int = ops.PREC_FP32
    p: float = 0.0          # dropout probability (0 in eval)
    seed: int = 0
    stream_base: int = 0    # dropout stream namespace of this stack

    def drop(self, site: int):
        return (self.p, self.seed, self.stream_base + site) if self.p > 0.0 else None


def _empty(shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.float32)


def _p(t, off=0):
    return t.data_ptr() + 4 * off


# --------------------------------------------------------------------------------------------------
# blocks (forward returns (y, saved); backward returns dx and accumulates parameter grads)


def _sa_fwd(x, B, T, H, in_w, in_b, out_w, out_b, n_w, n_b, cfg: RunCfg, site: int, save: bool):
    d = x.shape[-1]
    M = B * T
    mean, rstd = ops.ln_stats(x, d)
    qkv = _empty((M, 3 * d), x)
    ops.gemm(x, d, MK, in_w, d, NK, qkv, 3 * d, M, 3 * d, d, precision=cfg.precision, ln=(mean, rstd, n_w, n_b), bias=in_b)
    attn = _empty((M, d), x)
    lse = _empty((B, H, T), x) if save else None
    ops.attention_fwd(_p(qkv), 3 * d, _p(qkv, d), 3 * d, _p(qkv, 2 * d), 3 * d, _p(attn), d,
                      None if lse is None else _p(lse), B, H, T, T, d // H, cfg.drop(site), cfg.precision)
    y = _empty((M, d), x)
    ops.gemm(attn, d, MK, out_w, d, NK, y, d, M, d, d, precision=cfg.precision, bias=out_b, dropout=cfg.drop(site + 1),
             residual=x, ldr=d)
    return y, ((x, mean, rstd, qkv, attn, lse) if save else None)


def _sa_bwd(dy, saved, B, T, H, in_w, out_w, n_w, n_b, g_in_w, g_in_b, g_out_w, g_out_b, g_n_w, g_n_b, cfg: RunCfg,
            site: int):
    x, mean, rstd, qkv, attn, lse = saved
    d = x.shape[-1]
    M = B * T
    g1 = dy if cfg.p == 0.0 else ops.dropout_apply(dy, cfg.p, cfg.seed, cfg.stream_base + site + 1)
    ops.colsum_accum(g1, d, M, d, g_out_b)
    ops.gemm(g1, d, KM, attn, d, KN, g_out_w, d, d, d, M, precision=cfg.precision, accumulate=True)
    dattn = _empty((M, d), x)
    ops.gemm(g1, d, MK, out_w, d, KN, dattn, d, M, d, d, precision=cfg.precision)
    dqkv = _empty((M, 3 * d), x)
    ops.attention_bwd(_p(qkv), 3 * d, _p(qkv, d), 3 * d, _p(qkv, 2 * d), 3 * d, _p(attn), d, _p(dattn), d, _p(lse),
                      _p(dqkv), 3 * d, _p(dqkv, d), 3 * d, _p(dqkv, 2 * d), 3 * d, B, H, T, T, d // H, cfg.drop(site), cfg.precision)
    ops.colsum_accum(dqkv, 3 * d, M, 3 * d, g_in_b)
    ops.gemm(dqkv, 3 * d, KM, x, d, KN, g_in_w, d, 3 * d, d, M, precision=cfg.precision, ln=(mean, rstd, n_w, n_b),
             accumulate=True)
    dxn = _empty((M, d), x)
    ops.gemm(dqkv, 3 * d, MK, in_w, d, KN, dxn, d, M, d, 3 * d, precision=cfg.precision)
    dx = _empty((M, d), x)
    ops.ln_bwd(dxn, x, mean, rstd, n_w, dy, dx, g_n_w, g_n_b, M, d)
    return dx


def _ffn_fwd(x, l1_w, l1_b, l2_w, l2_b, n_w, n_b, cfg: RunCfg, site: int, save: bool):
    d = x.shape[-1]
    ff = l1_w.shape[0]
    M = x.shape[0]
    mean, rstd = ops.ln_stats(x, d)
    hpre = _empty((M, ff), x) if save else None
    hact = _empty((M, ff), x)
    ops.gemm(x, d, MK, l1_w, d, NK, hact, ff, M, ff, d, precision=cfg.precision, ln=(mean, rstd, n_w, n_b), bias=l1_b,
             pre_out=hpre, ldp=ff, act=ACT_GELU, dropout=cfg.drop(site))
    y = _empty((M, d), x)
    ops.gemm(hact, ff, MK, l2_w, ff, NK, y, d, M, d, ff, precision=cfg.precision, bias=l2_b, dropout=cfg.drop(site + 1),
             residual=x, ldr=d)
    return y, ((x, mean, rstd, hpre, hact) if save else None)


def _ffn_bwd(dy, saved, l1_w, l2_w, n_w, n_b, g_l1_w, g_l1_b, g_l2_w, g_l2_b, g_n_w, g_n_b, cfg: RunCfg, site: int):
    x, mean, rstd, hpre, hact = saved
    d = x.shape[-1]
    ff = l1_w.shape[0]
    M = x.shape[0]
    g = dy if cfg.p == 0.0 else ops.dropout_apply(dy, cfg.p, cfg.seed, cfg.stream_base + site + 1)
    ops.colsum_accum(g, d, M, d, g_l2_b)
    ops.gemm(g, d, KM, hact, ff, KN, g_l2_w, ff, d, ff, M, precision=cfg.precision, accumulate=True)
    dhpre = _empty((M, ff), x)
    ops.gemm(g, d, MK, l2_w, ff, KN, dhpre, ff, M, ff, d, precision=cfg.precision, gelu_grad_src=hpre, ldg=ff,
             dropout=cfg.drop(site))
    ops.colsum_accum(dhpre, ff, M, ff, g_l1_b)
    ops.gemm(dhpre, ff, KM, x, d, KN, g_l1_w, d, ff, d, M, precision=cfg.precision, ln=(mean, rstd, n_w, n_b),
             accumulate=True)
    dxn = _empty((M, d), x)
    ops.gemm(dhpre, ff, MK, l1_w, d, KN, dxn, d, M, d, ff, precision=cfg.precision)
    dx = _empty((M, d), x)
    ops.ln_bwd(dxn, x, mean, rstd, n_w, dy, dx, g_n_w, g_n_b, M, d)
    return dx


def _ca_fwd(x, mem, B, T, Mm, H, in_w, in_b, out_w, out_b, n_w, n_b, cfg: RunCfg, site: int, save: bool):
    """x (B*T,d) queries (pre-LN), mem (B*Mm,d) keys/values (NOT layer-normed; transformer.py:1137)."""
    d = x.shape[-1]
    Mq = B * T
    mean, rstd = ops.ln_stats(x, d)
    q = _empty((Mq, d), x)
    ops.gemm(x, d, MK, in_w, d, NK, q, d, Mq, d, d, precision=cfg.precision, ln=(mean, rstd, n_w, n_b), bias=in_b)
    kv = _empty((B * Mm, 2 * d), x)
    ops.gemm(mem, d, MK, _p(in_w, d * d), d, NK, kv, 2 * d, B * Mm, 2 * d, d, precision=cfg.precision, bias=_p(in_b, d))
    attn = _empty((Mq, d), x)
    lse = _empty((B, H, T), x) if save else None
    ops.attention_fwd(_p(q), d, _p(kv), 2 * d, _p(kv, d), 2 * d, _p(attn), d, None if lse is None else _p(lse), B, H, T,
                      Mm, d // H, cfg.drop(site), cfg.precision)
    y = _empty((Mq, d), x)
    ops.gemm(attn, d, MK, out_w, d, NK, y, d, Mq, d, d, precision=cfg.precision, bias=out_b, dropout=cfg.drop(site + 1),
             residual=x, ldr=d)
    return y, ((x, mean, rstd, q, kv, attn, lse) if save else None)


def _ca_bwd(dy, saved, mem, dmem, B, T, Mm, H, in_w, out_w, n_w, n_b, g_in_w, g_in_b, g_out_w, g_out_b, g_n_w, g_n_b,
            cfg: RunCfg, site: int):
    x, mean, rstd, q, kv, attn, lse = saved
    d = x.shape[-1]
    Mq = B * T
    g1 = dy if cfg.p == 0.0 else ops.dropout_apply(dy, cfg.p, cfg.seed, cfg.stream_base + site + 1)
    ops.colsum_accum(g1, d, Mq, d, g_out_b)
    ops.gemm(g1, d, KM, attn, d, KN, g_out_w, d, d, d, Mq, precision=cfg.precision, accumulate=True)
    dattn = _empty((Mq, d), x)
    ops.gemm(g1, d, MK, out_w, d, KN, dattn, d, Mq, d, d, precision=cfg.precision)
    dq = _empty((Mq, d), x)
    dkv = _empty((B * Mm, 2 * d), x)
    ops.attention_bwd(_p(q), d, _p(kv), 2 * d, _p(kv, d), 2 * d, _p(attn), d, _p(dattn), d, _p(lse), _p(dq), d, _p(dkv),
                      2 * d, _p(dkv, d), 2 * d, B, H, T, Mm, d // H, cfg.drop(site), cfg.precision)
    # q projection (rows 0:d of in_proj) and k/v projection (rows d:3d)
    ops.colsum_accum(dq, d, Mq, d, g_in_b)
    ops.colsum_accum(dkv, 2 * d, B * Mm, 2 * d, _p(g_in_b, d))
    ops.gemm(dq, d, KM, x, d, KN, g_in_w, d, d, d, Mq, precision=cfg.precision, ln=(mean, rstd, n_w, n_b), accumulate=True)
    ops.gemm(dkv, 2 * d, KM, mem, d, KN, _p(g_in_w, d * d), d, 2 * d, d, B * Mm, precision=cfg.precision, accumulate=True)
    if dmem is not None:
        ops.gemm(dkv, 2 * d, MK, _p(in_w, d * d), d, KN, dmem, d, B * Mm, d, 2 * d, precision=cfg.precision,
                 accumulate=True)
    dxn = _empty((Mq, d), x)
    ops.gemm(dq, d, MK, in_w, d, KN, dxn, d, Mq, d, d, precision=cfg.precision)
    dx = _empty((Mq, d), x)
    ops.ln_bwd(dxn, x, mean, rstd, n_w, dy, dx, g_n_w, g_n_b, Mq, d)
    return dx


def _zero_grads(params):
    """Gradient buffers of ``params``: (buffers the kernels accumulate into, what backward() hands to autograd).

    Default: one flat zeroed buffer viewed per parameter, returned to autograd (which adds it into ``p.grad``).
    With ``runtime.set_direct_grads(True)`` a leaf parameter whose ``.grad`` already exists (``FusedAdamW`` keeps them as
    views of its flat, zeroed gradient buffer) is accumulated into IN PLACE and ``None`` is returned for it: the ~260
    per-parameter ``add_`` launches of autograd's accumulation disappear.  (Gradient hooks on such parameters do not
    fire in that mode — it is opt-in, used by the captured training step.)"""
    from . import runtime

    direct = runtime.direct_grads()
    in_place = [direct and p.is_leaf and p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                and p.grad.shape == p.shape for p in params]
    total = sum(p.numel() for p, ip in zip(params, in_place) if not ip)
    flat = torch.zeros(max(total, 1), device=params[0].device, dtype=torch.float32)
    bufs, rets, off = [], [], 0
    for p, ip in zip(params, in_place):
        if ip:
            bufs.append(p.grad)
            rets.append(None)
            runtime.mark_grad_written(p)   # autograd never sees this gradient: tell the optimizer it exists
        else:
            v = flat[off: off + p.numel()].view_as(p)
            off += p.numel()
            bufs.append(v)
            rets.append(v)
    return bufs, rets


def _fused_enc_ok(cfg: RunCfg, d: int, ff: int, S: int, H: int) -> bool:
    from . import runtime

    return (cfg.precision == ops.PREC_BF16 and runtime.fused_layers() and ops.enc_layer_supported(d, ff, S, H))


def _pack_enc_weights(layer_params, L: int, d: int):
    """bf16 copies of the L layers' GEMM weights as ONE [L*768][d] matrix (the TMA tensor map of the fused kernels)."""
    wp = torch.empty((L * ops.ENC_ROWS_PER_LAYER, d), device=layer_params[0].device, dtype=torch.bfloat16)
    mats = []
    for l in range(L):
        in_w, _, out_w, _, l1_w, _, l2_w = layer_params[l * ENC_PARAMS_PER_LAYER: l * ENC_PARAMS_PER_LAYER + 7]
        r0 = l * ops.ENC_ROWS_PER_LAYER
        mats += [(in_w, 3 * d, r0), (out_w, d, r0 + 3 * d), (l1_w, d, r0 + 4 * d), (l2_w, d, r0 + 5 * d)]
    ops.pack_weights_bf16(mats, wp, d)
    return wp


def _enc_stack_fused_bwd(ctx, dh, layer_params, g_layers):
    """Backward through the fused layers: per layer ONE data-gradient kernel + ONE launch for all six weight gradients."""
    cfg = ctx.cfg
    B, S, H, d, Kin, L = ctx.dims
    M = B * S
    wp = ctx.fused
    bf = lambda n: torch.empty((M, n), device=dh.device, dtype=torch.bfloat16)
    g2, dhpre, g1, dqkv = bf(d), bf(d), bf(d), bf(3 * d)   # reused by every layer (same stream: consumed before rewritten)
    dx = _empty((M, d), dh)                                 # never write into autograd's own gradient tensor
    for l in reversed(range(L)):
        in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b = layer_params[
            l * ENC_PARAMS_PER_LAYER: (l + 1) * ENC_PARAMS_PER_LAYER]
        (g_in_w, g_in_b, g_out_w, g_out_b, g_l1_w, g_l1_b, g_l2_w, g_l2_b, g_n1_w, g_n1_b, g_n2_w,
         g_n2_b) = g_layers[l * ENC_PARAMS_PER_LAYER: (l + 1) * ENC_PARAMS_PER_LAYER]
        h_in, x1, xn1, attn, xn2, hact = ctx.acts[l]
        ops.enc_layer_bwd(dh, dx, h_in, x1, xn1, xn2, g2, dhpre, g1, dqkv, g_n1_w, g_n1_b, g_n2_w, g_n2_b, B, S, H, wp,
                          l * ops.ENC_ROWS_PER_LAYER, in_b, l1_b, n1_w, n2_w, dropout=cfg.drop(8 * l))
        ops.wgrad_bf16([(dqkv, 0, xn1, 0, _p(g_in_w), d, _p(g_in_b)),
                        (dqkv, d, xn1, 0, _p(g_in_w, d * d), d, _p(g_in_b, d)),
                        (dqkv, 2 * d, xn1, 0, _p(g_in_w, 2 * d * d), d, _p(g_in_b, 2 * d)),
                        (g1, 0, attn, 0, _p(g_out_w), d, _p(g_out_b)),
                        (dhpre, 0, xn2, 0, _p(g_l1_w), d, _p(g_l1_b)),
                        (g2, 0, hact, 0, _p(g_l2_w), d, _p(g_l2_b))], M)
        dh = dx   # the next (earlier) layer reads and rewrites dx in place
    return dh


# --------------------------------------------------------------------------------------------------
class EncoderStackFn(torch.autograd.Function):
    """Conv1d patch embedding (as a GEMM) + PE + L pre-LN encoder layers (base.py:49-53)."""

    @staticmethod
    def forward(ctx, cfg: RunCfg, B: int, S: int, H: int, pe, x_in, emb_w, emb_b, *layer_params):
        d = emb_w.shape[0]
        Kin = emb_w.shape[1]
        M = B * S
        L = len(layer_params) // ENC_PARAMS_PER_LAYER
        save = any(ctx.needs_input_grad)  # grad mode is always off inside forward(); this reflects the caller's
        h = _empty((M, d), x_in)
        ops.gemm(x_in, Kin, MK, emb_w, Kin, NK, h, d, M, d, Kin, precision=cfg.precision, bias=emb_b, pe=pe, pe_period=S)
        if L > 0 and _fused_enc_ok(cfg, d, layer_params[4].shape[0], S, H):
            # layer-fused tensor-core path: ONE kernel per layer (weights by TMA, residual stream in registers).  Inference
            # updates the stream in place; training keeps each layer's input and the activations its backward needs.
            wp = _pack_enc_weights(layer_params, L, d)
            fsaves = []
            for l in range(L):
                (in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b) = layer_params[
                    l * ENC_PARAMS_PER_LAYER: (l + 1) * ENC_PARAMS_PER_LAYER]
                if save:
                    y = _empty((M, d), h)
                    x1 = _empty((M, d), h)
                    bf = [torch.empty((M, d), device=h.device, dtype=torch.bfloat16) for _ in range(4)]   # xn1, attn, xn2, hact
                    ops.enc_layer_fwd(h, y, B, S, H, wp, l * ops.ENC_ROWS_PER_LAYER, in_b, out_b, l1_b, l2_b, n1_w, n1_b,
                                      n2_w, n2_b, saves=(x1, *bf), dropout=cfg.drop(8 * l))
                    fsaves.append((h, x1, *bf))
                    h = y
                else:
                    ops.enc_layer_fwd(h, h, B, S, H, wp, l * ops.ENC_ROWS_PER_LAYER, in_b, out_b, l1_b, l2_b, n1_w, n1_b,
                                      n2_w, n2_b, saves=None, dropout=cfg.drop(8 * l))
            if save:
                ctx.cfg, ctx.dims, ctx.acts, ctx.fused = cfg, (B, S, H, d, Kin, L), fsaves, wp
                ctx.save_for_backward(x_in, emb_w, emb_b, *layer_params)
            return h.view(B, S, d)
        acts = []
        for l in range(L):
            in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b = layer_params[
                l * ENC_PARAMS_PER_LAYER: (l + 1) * ENC_PARAMS_PER_LAYER]
            h, s1 = _sa_fwd(h, B, S, H, in_w, in_b, out_w, out_b, n1_w, n1_b, cfg, 8 * l, save)
            h, s2 = _ffn_fwd(h, l1_w, l1_b, l2_w, l2_b, n2_w, n2_b, cfg, 8 * l + 2, save)
            acts.append((s1, s2))
        if save:
            ctx.cfg, ctx.dims, ctx.acts, ctx.fused = cfg, (B, S, H, d, Kin, L), acts, None
            ctx.save_for_backward(x_in, emb_w, emb_b, *layer_params)
        return h.view(B, S, d)

    @staticmethod
    def backward(ctx, dy):
        cfg = ctx.cfg
        B, S, H, d, Kin, L = ctx.dims
        x_in, emb_w, emb_b, *layer_params = ctx.saved_tensors
        M = B * S
        grads, rets = _zero_grads([emb_w, emb_b, *layer_params])
        g_emb_w, g_emb_b, g_layers = grads[0], grads[1], grads[2:]
        dh = dy.contiguous().view(M, d)
        if ctx.fused is not None:
            dh = _enc_stack_fused_bwd(ctx, dh, layer_params, g_layers)
        for l in reversed(range(L if ctx.fused is None else 0)):
            in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b = layer_params[
                l * ENC_PARAMS_PER_LAYER: (l + 1) * ENC_PARAMS_PER_LAYER]
            (g_in_w, g_in_b, g_out_w, g_out_b, g_l1_w, g_l1_b, g_l2_w, g_l2_b, g_n1_w, g_n1_b, g_n2_w,
             g_n2_b) = g_layers[l * ENC_PARAMS_PER_LAYER: (l + 1) * ENC_PARAMS_PER_LAYER]
            s1, s2 = ctx.acts[l]
            dh = _ffn_bwd(dh, s2, l1_w, l2_w, n2_w, n2_b, g_l1_w, g_l1_b, g_l2_w, g_l2_b, g_n2_w, g_n2_b, cfg, 8 * l + 2)
            dh = _sa_bwd(dh, s1, B, S, H, in_w, out_w, n1_w, n1_b, g_in_w, g_in_b, g_out_w, g_out_b, g_n1_w, g_n1_b,
                         cfg, 8 * l)
        ops.colsum_accum(dh, d, M, d, g_emb_b)
        ops.gemm(dh, d, KM, x_in, Kin, KN, g_emb_w, Kin, d, Kin, M, precision=cfg.precision, accumulate=True)
        dx_in = None
        if ctx.needs_input_grad[5]:
            dx_in = _empty((M, Kin), dh)
            ops.gemm(dh, d, MK, emb_w, Kin, KN, dx_in, Kin, M, Kin, d, precision=cfg.precision)
            dx_in = dx_in.view_as(x_in)
        ctx.acts = None
        return (None, None, None, None, None, dx_in, *rets)


def _pack_dec_weights(layer_params, L: int, d: int):
    """bf16 copies of the decoder layers' GEMM weights as ONE [L*1280][d] matrix (rows per layer: self_attn.in_proj |
    self_attn.out_proj | multihead_attn.in_proj | multihead_attn.out_proj | linear1 | linear2)."""
    wp = torch.empty((L * ops.DEC_ROWS_PER_LAYER, d), device=layer_params[0].device, dtype=torch.bfloat16)
    mats = []
    for l in range(L):
        P = layer_params[l * DEC_PARAMS_PER_LAYER: (l + 1) * DEC_PARAMS_PER_LAYER]
        r0 = l * ops.DEC_ROWS_PER_LAYER
        mats += [(P[0], 3 * d, r0), (P[2], d, r0 + 3 * d), (P[4], 3 * d, r0 + 4 * d), (P[6], d, r0 + 7 * d),
                 (P[8], d, r0 + 8 * d), (P[10], d, r0 + 9 * d)]
    ops.pack_weights_bf16(mats, wp, d)
    return wp


def _bf16(shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.bfloat16)


def _dec_layer_fused_fwd(h, mem2, B, T, Mm, H, P, wp, l, cfg: RunCfg, save: bool, kv_all=None):
    """One decoder layer as three fused kernels: self-attention block, cross-attention block (``kv_all``: the memory's K | V
    projections of all layers, bf16; None -> per-op kernels, shapes outside sd_ca_block_supported), feed-forward block."""
    (sa_in_w, sa_in_b, sa_out_w, sa_out_b, ca_in_w, ca_in_b, ca_out_w, ca_out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b,
     n3_w, n3_b) = P
    d = h.shape[-1]
    M = B * T
    r0 = l * ops.DEC_ROWS_PER_LAYER
    if save:
        x1 = _empty((M, d), h)
        xn1, attn = _bf16((M, d), h), _bf16((M, d), h)
        ops.enc_layer_fwd(h, x1, B, T, H, wp, r0, sa_in_b, sa_out_b, None, None, n1_w, n1_b, None, None,
                          saves=(None, xn1, attn, None, None), dropout=cfg.drop(8 * l), blocks=ops.LAYER_SA)
    else:
        x1, xn1, attn = h, None, None
        ops.enc_layer_fwd(h, h, B, T, H, wp, r0, sa_in_b, sa_out_b, None, None, n1_w, n1_b, None, None,
                          dropout=cfg.drop(8 * l), blocks=ops.LAYER_SA)
    if kv_all is None:
        x2, s2 = _ca_fwd(x1, mem2, B, T, Mm, H, ca_in_w, ca_in_b, ca_out_w, ca_out_b, n2_w, n2_b, cfg, 8 * l + 2, save)
    elif save:
        x2 = _empty((M, d), h)
        xn2, q2, attn2 = _bf16((M, d), h), _bf16((M, d), h), _bf16((M, d), h)
        stats, lse = _empty((M, 2), h), _empty((B, H, T), h)
        ops.ca_block_fwd(x1, x2, B, T, Mm, wp, r0 + 512, r0 + 896, kv_all, 256 * l, ca_in_b, ca_out_b, n2_w, n2_b,
                         saves=(xn2, q2, attn2, stats, lse), dropout=cfg.drop(8 * l + 2))
        s2 = (x1, xn2, q2, attn2, stats, lse)
    else:
        ops.ca_block_fwd(x1, x1, B, T, Mm, wp, r0 + 512, r0 + 896, kv_all, 256 * l, ca_in_b, ca_out_b, n2_w, n2_b)
        x2, s2 = x1, None
    drop = cfg.drop(8 * l + 4)
    kw = dict(dropout=drop, blocks=ops.LAYER_FFN, w_row_ffn=r0 + 1024, dropout_stream_ffn=cfg.stream_base + 8 * l + 4)
    if save:
        y = _empty((M, d), h)
        xn3, hact = _bf16((M, d), h), _bf16((M, d), h)
        ops.enc_layer_fwd(x2, y, B, T, H, wp, r0, None, None, l1_b, l2_b, None, None, n3_w, n3_b,
                          saves=(None, None, None, xn3, hact), **kw)
        return y, (h, xn1, attn, s2, x2, xn3, hact)
    ops.enc_layer_fwd(x2, x2, B, T, H, wp, r0, None, None, l1_b, l2_b, None, None, n3_w, n3_b, **kw)
    return x2, None


def _dec_layer_fused_bwd(dh, saved, mem2, dmem, B, T, Mm, H, P, G, wp, l, cfg: RunCfg, bufs, kvctx=None):
    (sa_in_w, sa_in_b, sa_out_w, sa_out_b, ca_in_w, ca_in_b, ca_out_w, ca_out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b,
     n3_w, n3_b) = P
    (g_sa_in_w, g_sa_in_b, g_sa_out_w, g_sa_out_b, g_ca_in_w, g_ca_in_b, g_ca_out_w, g_ca_out_b, g_l1_w, g_l1_b, g_l2_w, g_l2_b,
     g_n1_w, g_n1_b, g_n2_w, g_n2_b, g_n3_w, g_n3_b) = G
    h_in, xn1, attn, s2, x2, xn3, hact = saved
    g2, dhpre, g1, dqkv, dx = bufs
    d = h_in.shape[-1]
    M = B * T
    r0 = l * ops.DEC_ROWS_PER_LAYER
    # feed-forward block
    ops.enc_layer_bwd(dh, dx, None, x2, None, xn3, g2, dhpre, None, None, None, None, g_n3_w, g_n3_b, B, T, H, wp, r0, None, l1_b,
                      None, n3_w, dropout=cfg.drop(8 * l + 4), blocks=ops.LAYER_FFN, w_row_ffn=r0 + 1024,
                      dropout_stream_ffn=cfg.stream_base + 8 * l + 4)
    ops.wgrad_bf16([(dhpre, 0, xn3, 0, _p(g_l1_w), d, _p(g_l1_b)), (g2, 0, hact, 0, _p(g_l2_w), d, _p(g_l2_b))], M)
    # cross-attention block
    if kvctx is None:
        dh = _ca_bwd(dx, s2, mem2, dmem, B, T, Mm, H, ca_in_w, ca_out_w, n2_w, n2_b, g_ca_in_w, g_ca_in_b, g_ca_out_w, g_ca_out_b,
                     g_n2_w, g_n2_b, cfg, 8 * l + 2)
    else:
        kv_all, dkv_all = kvctx
        x1, xn2, q2, attn2, stats, lse = s2
        ops.ca_block_bwd(dx, dx, x1, q2, attn2, stats, lse, B, T, Mm, wp, r0 + 512, r0 + 896, kv_all, 256 * l, n2_w, g1, dhpre, dkv_all,
                         g_n2_w, g_n2_b, dropout=cfg.drop(8 * l + 2))   # g1 / dhpre: free bf16 buffers at this point
        ops.wgrad_bf16([(g1, 0, attn2, 0, _p(g_ca_out_w), d, _p(g_ca_out_b)), (dhpre, 0, xn2, 0, _p(g_ca_in_w), d, _p(g_ca_in_b))], M)
        dh = dx
    # self-attention block
    ops.enc_layer_bwd(dh, dx, h_in, None, xn1, None, None, None, g1, dqkv, g_n1_w, g_n1_b, None, None, B, T, H, wp, r0, sa_in_b,
                      None, n1_w, None, dropout=cfg.drop(8 * l), blocks=ops.LAYER_SA)
    ops.wgrad_bf16([(dqkv, 0, xn1, 0, _p(g_sa_in_w), d, _p(g_sa_in_b)),
                    (dqkv, d, xn1, 0, _p(g_sa_in_w, d * d), d, _p(g_sa_in_b, d)),
                    (dqkv, 2 * d, xn1, 0, _p(g_sa_in_w, 2 * d * d), d, _p(g_sa_in_b, 2 * d)),
                    (g1, 0, attn, 0, _p(g_sa_out_w), d, _p(g_sa_out_b))], M)
    return dx


def tc_sampler_supported(d: int, H: int, T: int, Mm: int, L: int, ff: int) -> bool:
    """The tensor-core batched sampler needs every block of the decoder layer on the fused kernels."""
    return (L > 0 and L <= ops.KV_MAX_LAYERS and ops.enc_layer_supported(d, ff, T, H) and ops.ca_block_fwd_supported(d, H, T, Mm))


def ddim_sample_tc(B: int, T: int, Mm: int, H: int, pe, x_T, mem, tok, emb_w, emb_b, fc_w, fc_b, layer_params, coefs, trace=None):
    """The complete eta=0 DDIM loop (ros.py:301-310, distill.py:179-189) for a BATCH of trajectories on the layer-fused
    tensor-core kernels, eval mode.  ``mem`` fp32 (B*Mm, d): the assembled context whose LAST row per trajectory is the step
    token (its content is irrelevant here); ``tok`` fp32 (S, d): the step token of every scheduled timestep; ``coefs``:
    per-step (sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)).  The context's K | V of all layers are projected ONCE
    (the reference re-projects them in each of the S*L cross-attention calls); per step only the step token's K | V row
    changes: it comes from an (S, L*256) table computed by the same GEMM.  Returns x_0 (B, T, J) fp32."""
    d, J = emb_w.shape
    L = len(layer_params) // DEC_PARAMS_PER_LAYER
    S = tok.shape[0]
    cfg = RunCfg(precision=ops.PREC_BF16)
    wp = _pack_dec_weights(layer_params, L, d)
    biases = [_p(layer_params[l * DEC_PARAMS_PER_LAYER + 5], d) for l in range(L)]
    mem_bf = _bf16((B * Mm, d), mem)
    ops.cast_bf16(mem, mem_bf)
    kv_all = _bf16((B * Mm, 256 * L), mem)
    ops.kv_proj_bf16(mem_bf, wp, 640, ops.DEC_ROWS_PER_LAYER, biases, kv_all)
    Sp = (S + 7) // 8 * 8   # sd_cast_bf16 works on multiples of 8 elements
    tok_bf = torch.zeros((Sp, d), device=mem.device, dtype=torch.bfloat16)
    tok_p = tok if Sp == S else torch.cat([tok, tok.new_zeros(Sp - S, d)])
    ops.cast_bf16(tok_p.contiguous(), tok_bf)
    kv_tok = _bf16((Sp, 256 * L), mem)
    ops.kv_proj_bf16(tok_bf, wp, 640, ops.DEC_ROWS_PER_LAYER, biases, kv_tok)
    Mq = B * T
    x = x_T.contiguous().view(Mq, J).clone()
    x_next = torch.empty_like(x)
    h = _empty((Mq, d), mem)
    eps = _empty((Mq, J), mem)
    from . import runtime

    # the loop is a chain of 16 short kernels per step: programmatic dependent launch lets each one run its prologue
    # (barriers, TMEM allocation, weight TMA loads) under its predecessor's tail
    pdl_before = ops.set_pdl(runtime.pdl_enabled())
    glue = J <= ops.DDIM_GLUE_MAX_J   # output projection + scheduler update + next embedding + K | V row broadcast in one launch
    try:
        for s in range(S):
            if s == 0 or not glue:
                ops.bcast_row_bf16(kv_all, Mm, Mm - 1, B, kv_tok.data_ptr() + 2 * s * 256 * L, 256 * L)
                ops.gemm(x, J, MK, emb_w, J, NK, h, d, Mq, d, J, precision=cfg.precision, bias=emb_b, pe=pe, pe_period=T)
            # per layer: self-attention | cross-attention | feed-forward, with the feed-forward half of layer l and the
            # self-attention half of layer l+1 as ONE launch (they are adjacent row-tile kernels): 2 L + 1 launches, not 3 L
            for l in range(L):
                (sa_in_w, sa_in_b, sa_out_w, sa_out_b, ca_in_w, ca_in_b, ca_out_w, ca_out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b,
                 n3_w, n3_b) = layer_params[l * DEC_PARAMS_PER_LAYER: (l + 1) * DEC_PARAMS_PER_LAYER]
                r0 = l * ops.DEC_ROWS_PER_LAYER
                if l == 0:
                    ops.enc_layer_fwd(h, h, B, T, H, wp, r0, sa_in_b, sa_out_b, None, None, n1_w, n1_b, None, None, blocks=ops.LAYER_SA)
                ops.ca_block_fwd(h, h, B, T, Mm, wp, r0 + 512, r0 + 896, kv_all, 256 * l, ca_in_b, ca_out_b, n2_w, n2_b)
                if l + 1 < L:
                    nx = layer_params[(l + 1) * DEC_PARAMS_PER_LAYER: (l + 2) * DEC_PARAMS_PER_LAYER]
                    ops.enc_layer_fwd(h, h, B, T, H, wp, r0 + ops.DEC_ROWS_PER_LAYER, nx[1], nx[3], l1_b, l2_b, nx[12], nx[13], n3_w, n3_b,
                                      blocks=ops.LAYER_SA | ops.LAYER_FFN | ops.LAYER_FFN_FIRST, w_row_ffn=r0 + 1024)
                else:
                    ops.enc_layer_fwd(h, h, B, T, H, wp, r0, None, None, l1_b, l2_b, None, None, n3_w, n3_b, blocks=ops.LAYER_FFN,
                                      w_row_ffn=r0 + 1024)
            if glue:
                last = s + 1 == S
                ops.ddim_glue(h, fc_w, fc_b, x, x_next, None if trace is None else trace[s], coefs[s],
                              emb=None if last else (emb_w, emb_b, pe, T),
                              kv_bcast=None if last else (kv_all, Mm, Mm - 1, B, kv_tok.data_ptr() + 2 * (s + 1) * 256 * L, 256 * L))
            else:
                ops.gemm(h, d, MK, fc_w, d, NK, eps, J, Mq, J, d, precision=cfg.precision, bias=fc_b)
                if trace is not None:
                    trace[s].view(Mq, J).copy_(eps)
                ops.ddim_step(x, eps, x_next, None, coefs[s])
            x, x_next = x_next, x
    finally:
        ops.set_pdl(pdl_before)
    return x.view(B, T, J)


class DenoiserFn(torch.autograd.Function):
    """Linear(J->d)+PE, L pre-LN decoder layers over memory, Linear(d->J) (decoder.py:47-54)."""

    @staticmethod
    def forward(ctx, cfg: RunCfg, B: int, T: int, Mm: int, H: int, pe, x, mem, emb_w, emb_b, fc_w, fc_b, *layer_params):
        d = emb_w.shape[0]
        J = emb_w.shape[1]
        Mq = B * T
        L = len(layer_params) // DEC_PARAMS_PER_LAYER
        save = any(ctx.needs_input_grad)
        mem2 = mem.contiguous().view(B * Mm, d)
        x2 = x.contiguous().view(Mq, J)
        h = _empty((Mq, d), mem2)
        ops.gemm(x2, J, MK, emb_w, J, NK, h, d, Mq, d, J, precision=cfg.precision, bias=emb_b, pe=pe, pe_period=T)
        acts = []
        fused = L > 0 and _fused_enc_ok(cfg, d, layer_params[8].shape[0], T, H)
        wp = _pack_dec_weights(layer_params, L, d) if fused else None
        mem_bf = kv_all = None
        # training (saves for the backward kernel): T <= 16; inference: T <= 64 through query-row groups
        if fused and L <= ops.KV_MAX_LAYERS and (ops.ca_block_supported(d, H, T, Mm) if save else ops.ca_block_fwd_supported(d, H, T, Mm)):
            # the memory is not layer-normed and every layer reads it: ONE bf16 copy, ONE GEMM for the K | V of all layers
            mem_bf = _bf16((B * Mm, d), mem2)
            ops.cast_bf16(mem2, mem_bf)
            kv_all = _bf16((B * Mm, 256 * L), mem2)
            ops.kv_proj_bf16(mem_bf, wp, 640, ops.DEC_ROWS_PER_LAYER, [_p(layer_params[l * DEC_PARAMS_PER_LAYER + 5], d) for l in range(L)],
                             kv_all)
        for l in range(L):
            P = layer_params[l * DEC_PARAMS_PER_LAYER: (l + 1) * DEC_PARAMS_PER_LAYER]
            if fused:
                h, sv = _dec_layer_fused_fwd(h, mem2, B, T, Mm, H, P, wp, l, cfg, save, kv_all)
                acts.append(sv)
                continue
            (sa_in_w, sa_in_b, sa_out_w, sa_out_b, ca_in_w, ca_in_b, ca_out_w, ca_out_b, l1_w, l1_b, l2_w, l2_b, n1_w,
             n1_b, n2_w, n2_b, n3_w, n3_b) = P
            h, s1 = _sa_fwd(h, B, T, H, sa_in_w, sa_in_b, sa_out_w, sa_out_b, n1_w, n1_b, cfg, 8 * l, save)
            h, s2 = _ca_fwd(h, mem2, B, T, Mm, H, ca_in_w, ca_in_b, ca_out_w, ca_out_b, n2_w, n2_b, cfg, 8 * l + 2, save)
            h, s3 = _ffn_fwd(h, l1_w, l1_b, l2_w, l2_b, n3_w, n3_b, cfg, 8 * l + 4, save)
            acts.append((s1, s2, s3))
        out = _empty((Mq, J), mem2)
        ops.gemm(h, d, MK, fc_w, d, NK, out, J, Mq, J, d, precision=cfg.precision, bias=fc_b)
        if save:
            ctx.cfg, ctx.dims, ctx.acts, ctx.h_last, ctx.fused = cfg, (B, T, Mm, H, d, J, L), acts, h, wp
            ctx.kv = (mem_bf, kv_all)
            ctx.save_for_backward(x2, mem2, emb_w, emb_b, fc_w, fc_b, *layer_params)
        return out.view(B, T, J)

    @staticmethod
    def backward(ctx, dout):
        cfg = ctx.cfg
        B, T, Mm, H, d, J, L = ctx.dims
        x2, mem2, emb_w, emb_b, fc_w, fc_b, *layer_params = ctx.saved_tensors
        Mq = B * T
        grads, rets = _zero_grads([emb_w, emb_b, fc_w, fc_b, *layer_params])
        g_emb_w, g_emb_b, g_fc_w, g_fc_b, g_layers = grads[0], grads[1], grads[2], grads[3], grads[4:]
        do = dout.contiguous().view(Mq, J)
        ops.colsum_accum(do, J, Mq, J, g_fc_b)
        ops.gemm(do, J, KM, ctx.h_last, d, KN, g_fc_w, d, J, d, Mq, precision=cfg.precision, accumulate=True)
        dh = _empty((Mq, d), mem2)
        ops.gemm(do, J, MK, fc_w, d, KN, dh, d, Mq, d, J, precision=cfg.precision)
        need_dmem = ctx.needs_input_grad[7]
        mem_bf, kv_all = ctx.kv
        kvctx = None
        if kv_all is not None:
            kvctx = (kv_all, _bf16((B * Mm, 256 * L), mem2))   # every element is written by the layers' backward kernels
            dmem = _empty((B * Mm, d), mem2) if need_dmem else None
        else:
            dmem = torch.zeros((B * Mm, d), device=mem2.device, dtype=torch.float32) if need_dmem else None
        bufs = None
        if ctx.fused is not None:
            bufs = (_bf16((Mq, d), mem2), _bf16((Mq, d), mem2), _bf16((Mq, d), mem2), _bf16((Mq, 3 * d), mem2), _empty((Mq, d), mem2))
        for l in reversed(range(L)):
            if ctx.fused is not None:
                dh = _dec_layer_fused_bwd(dh, ctx.acts[l], mem2, dmem, B, T, Mm, H,
                                          layer_params[l * DEC_PARAMS_PER_LAYER: (l + 1) * DEC_PARAMS_PER_LAYER],
                                          g_layers[l * DEC_PARAMS_PER_LAYER: (l + 1) * DEC_PARAMS_PER_LAYER], ctx.fused, l, cfg, bufs,
                                          kvctx)
                continue
            (sa_in_w, sa_in_b, sa_out_w, sa_out_b, ca_in_w, ca_in_b, ca_out_w, ca_out_b, l1_w, l1_b, l2_w, l2_b, n1_w,
             n1_b, n2_w, n2_b, n3_w, n3_b) = layer_params[l * DEC_PARAMS_PER_LAYER: (l + 1) * DEC_PARAMS_PER_LAYER]
            (g_sa_in_w, g_sa_in_b, g_sa_out_w, g_sa_out_b, g_ca_in_w, g_ca_in_b, g_ca_out_w, g_ca_out_b, g_l1_w, g_l1_b,
             g_l2_w, g_l2_b, g_n1_w, g_n1_b, g_n2_w, g_n2_b, g_n3_w,
             g_n3_b) = g_layers[l * DEC_PARAMS_PER_LAYER: (l + 1) * DEC_PARAMS_PER_LAYER]
            s1, s2, s3 = ctx.acts[l]
            dh = _ffn_bwd(dh, s3, l1_w, l2_w, n3_w, n3_b, g_l1_w, g_l1_b, g_l2_w, g_l2_b, g_n3_w, g_n3_b, cfg, 8 * l + 4)
            dh = _ca_bwd(dh, s2, mem2, dmem, B, T, Mm, H, ca_in_w, ca_out_w, n2_w, n2_b, g_ca_in_w, g_ca_in_b,
                         g_ca_out_w, g_ca_out_b, g_n2_w, g_n2_b, cfg, 8 * l + 2)
            dh = _sa_bwd(dh, s1, B, T, H, sa_in_w, sa_out_w, n1_w, n1_b, g_sa_in_w, g_sa_in_b, g_sa_out_w, g_sa_out_b,
                         g_n1_w, g_n1_b, cfg, 8 * l)
        if kvctx is not None:
            # k / v projection of the memory: weight / bias gradients of all layers (rows d:3d of each in_proj) and ONE data GEMM
            dkv_all = kvctx[1]
            jobs = []
            for l in range(L):
                g_w, g_b = g_layers[l * DEC_PARAMS_PER_LAYER + 4], g_layers[l * DEC_PARAMS_PER_LAYER + 5]
                jobs += [(dkv_all, 256 * l, mem_bf, 0, _p(g_w, d * d), d, _p(g_b, d)),
                         (dkv_all, 256 * l + 128, mem_bf, 0, _p(g_w, 2 * d * d), d, _p(g_b, 2 * d))]
            for i in range(0, len(jobs), ops.WGRAD_MAX_JOBS):
                ops.wgrad_bf16(jobs[i: i + ops.WGRAD_MAX_JOBS], B * Mm)
            if dmem is not None:
                ops.kv_dgrad_bf16(dkv_all, ctx.fused, 640, ops.DEC_ROWS_PER_LAYER, L, dmem, False)
        ctx.kv = None
        ops.colsum_accum(dh, d, Mq, d, g_emb_b)
        ops.gemm(dh, d, KM, x2, J, KN, g_emb_w, J, d, J, Mq, precision=cfg.precision, accumulate=True)
        dx = None
        if ctx.needs_input_grad[6]:
            dx = _empty((Mq, J), dh)
            ops.gemm(dh, d, MK, emb_w, J, KN, dx, J, Mq, J, d, precision=cfg.precision)
            dx = dx.view(B, T, J)
        ctx.acts = None
        ctx.h_last = None
        return (None, None, None, None, None, None, dx, None if dmem is None else dmem.view(B, Mm, d), *rets)


class LinearFn(torch.autograd.Function):
    """y = x W^T + b for 2-D x (image-token head: Conv2d 1x1 and fc; image.py:69-73)."""

    @staticmethod
    def forward(ctx, precision: int, x, w, b):
        M, K = x.shape
        N = w.shape[0]
        y = _empty((M, N), x)
        ops.gemm(x, K, MK, w, K, NK, y, N, M, N, K, precision=precision, bias=b)
        ctx.precision = precision
        ctx.save_for_backward(x, w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b = ctx.saved_tensors
        M, K = x.shape
        N = w.shape[0]
        dy = dy.contiguous()
        (gw, gb), (rw, rb) = _zero_grads([w, b])
        ops.colsum_accum(dy, N, M, N, gb)
        ops.gemm(dy, N, KM, x, K, KN, gw, K, N, K, M, precision=ctx.precision, accumulate=True)
        dx = None
        if ctx.needs_input_grad[1]:
            dx = _empty((M, K), x)
            ops.gemm(dy, N, MK, w, K, KN, dx, K, M, K, N, precision=ctx.precision)
        return None, dx, rw, rb


class AssembleContextFn(torch.autograd.Function):
    """torch.cat(context + [StepToken(t)], dim=1) (model.py:173-176) written by strided-copy kernels."""

    @staticmethod
    def forward(ctx, step, freqs, token, d: int, *parts):
        B = step.shape[0]
        lens = [p.shape[1] for p in parts]
        Mm = sum(lens) + 1
        mem = torch.empty((B, Mm, d), device=token.device, dtype=torch.float32)
        off = 0
        for p_, n in zip(parts, lens):
            pc = p_.contiguous()
            ops.copy_rows(pc.data_ptr(), n * d, d, mem.data_ptr() + 4 * off * d, Mm * d, d, B, n, d)
            off += n
        ops.step_token(step, freqs, token, mem.data_ptr() + 4 * off * d, Mm * d, B, d)
        ctx.lens, ctx.d, ctx.B = lens, d, B
        ctx.save_for_backward(token)
        return mem

    @staticmethod
    def backward(ctx, dmem):
        (token,) = ctx.saved_tensors
        d, B, lens = ctx.d, ctx.B, ctx.lens
        Mm = sum(lens) + 1
        dmem = dmem.contiguous()
        outs = []
        off = 0
        for i, n in enumerate(lens):
            if ctx.needs_input_grad[4 + i]:
                g = torch.empty((B, n, d), device=dmem.device, dtype=torch.float32)
                ops.copy_rows(dmem.data_ptr() + 4 * off * d, Mm * d, d, g.data_ptr(), n * d, d, B, n, d)
                outs.append(g)
            else:
                outs.append(None)
            off += n
        dtoken = None
        if ctx.needs_input_grad[2]:
            dtoken = torch.zeros_like(token)
            ops.step_token_bwd(dmem.data_ptr() + 4 * off * d, Mm * d, B, d, dtoken)
        return (None, None, dtoken, None, *outs)


class GatherRowsFn(torch.autograd.Function):
    """nn.Embedding lookup (game_state.py:27)."""

    @staticmethod
    def forward(ctx, table, idx):
        B, d = idx.numel(), table.shape[1]
        out = torch.empty((B, 1, d), device=table.device, dtype=torch.float32)
        err = torch.zeros(1, device=table.device, dtype=torch.int32)
        ops.gather_rows(table, idx, out, d, err)
        ctx.save_for_backward(idx)
        ctx.shape = table.shape
        ctx.err = err
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dtable = torch.zeros(ctx.shape, device=dout.device, dtype=torch.float32)
        dout = dout.contiguous()
        ops.scatter_add_rows(dout.data_ptr(), dout.shape[-1], idx, dtable)
        return dtable, None


class GradReadyFn(torch.autograd.Function):
    """Identity in the forward pass; its backward fires ``runtime.grad_ready(tag)`` once ALL ``state["need"]`` twins of the
    marked activation have received their gradient — every backward node downstream of the mark has then run."""

    @staticmethod
    def forward(ctx, x, tag: str, state: dict):
        ctx.tag, ctx.state = tag, state
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        from . import runtime

        ctx.state["seen"] += 1
        if ctx.state["seen"] == ctx.state["need"]:
            runtime.grad_ready(ctx.tag)
        return g, None, None


class MseLossFn(torch.autograd.Function):
    """F.mse_loss(pred, target) with mean reduction (train.py:229)."""

    @staticmethod
    def forward(ctx, pred, target):
        pred = pred.contiguous()
        target = target.contiguous()
        out = torch.empty((), device=pred.device, dtype=torch.float32)
        ops.mse_fwd(pred, target, out)
        ctx.save_for_backward(pred, target)
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        gp = torch.empty_like(pred)
        ops.mse_bwd(pred, target, g.contiguous(), gp)
        gt = None
        if ctx.needs_input_grad[1]:
            gt = -gp
        return gp, gt


def mse_loss(pred, target):
    return MseLossFn.apply(pred, target)
