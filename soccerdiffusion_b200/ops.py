"""Thin typed wrappers over the C ABI (one Python function per entry point of include/sd_b200.h).

Device memory comes from torch's caching allocator and kernels are launched on torch's current
stream — PyTorch is plumbing here; all arithmetic happens in libsd_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmDesc, check, stream_ptr

MK, KM, NK, KN = 0, 1, 0, 1
ACT_NONE, ACT_GELU = 0, 1
PREC_FP32, PREC_BF16 = 0, 1
LN_EPS = 1e-5

# number of kernels of this library launched since the last reset (bench.py's gpu_launches)
_launches = 0


def launches() -> int:
    return _launches


def reset_launches():
    global _launches
    _launches = 0


def _count(n=1):
    global _launches
    _launches += n


# optional per-entry-point device timing (CUDA events on the launching stream) for bench.py's roofline block;
# off by default: the events are only recorded while a profile is open
_profile = None


def profile_begin():
    global _profile
    _profile = []


def profile_end():
    """-> {kernel class: {"launches", "ms", "flops", "bytes"}} ; synchronises the device."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, flops, nbytes, detail in rec or []:
        ms = e0.elapsed_time(e1)
        for key in (name, name + " " + detail) if detail else (name,):
            d = out.setdefault(key, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += ms
            d["flops"] += flops
            d["bytes"] += nbytes
    return out


class _Timed:
    def __init__(self, name, flops=0.0, nbytes=0.0, detail=""):
        self.name, self.flops, self.nbytes, self.detail = name, flops, nbytes, detail

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None and exc[0] is None:
            self.e1.record()
            _profile.append((self.name, self.e0, self.e1, self.flops, self.nbytes, self.detail))
        return False


def _f32(t: torch.Tensor, name="tensor"):
    if t.dtype != torch.float32:
        raise _lib.SdError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise _lib.SdError(f"{name} must live on a CUDA device (no CPU fallback)")
    return t


def gemm(A, lda, a_layout, B, ldb, b_layout, C_, ldc, M, N, K, *, precision=PREC_FP32, ln=None, bias=None,
         pre_out=None, ldp=0, act=ACT_NONE, gelu_grad_src=None, ldg=0, dropout=None, pe=None, pe_period=0,
         residual=None, ldr=0, accumulate=False, alpha=1.0):
    """A, B, C_, ... are tensors or raw int pointers (for sliced views use ``t.data_ptr() + off*4``)."""
    d = GemmDesc()
    P = lambda t: (t if isinstance(t, int) or t is None else t.data_ptr())
    d.A, d.lda, d.a_layout = P(A), lda, a_layout
    d.B, d.ldb, d.b_layout = P(B), ldb, b_layout
    d.C, d.ldc = P(C_), ldc
    d.M, d.N, d.K = M, N, K
    d.precision = precision
    if ln is not None:
        d.ln_mean, d.ln_rstd, d.ln_gamma, d.ln_beta = (P(t) for t in ln)
    d.alpha = alpha
    d.bias = P(bias)
    d.pre_out, d.ldp = P(pre_out), ldp
    d.act = act
    d.gelu_grad_src, d.ldg = P(gelu_grad_src), ldg
    if dropout is not None and dropout[0] > 0.0:
        d.dropout_p, d.dropout_seed, d.dropout_stream = dropout
    d.pe, d.pe_period = P(pe), pe_period
    d.residual, d.ldr = P(residual), ldr
    d.accumulate = 1 if accumulate else 0
    kind = "gemm_wgrad" if a_layout == KM else ("gemm_dgrad" if b_layout == KN else "gemm_fwd")
    with _Timed(("tc_" if precision == PREC_BF16 else "f32_") + kind, 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N),
                f"[{M}x{N}x{K}]"):
        check(_lib.lib().sd_gemm(C.byref(d), stream_ptr()), "sd_gemm")
    _count()


def ln_stats(x2d: torch.Tensor, d: int):
    M = x2d.numel() // d
    mean = torch.empty(M, device=x2d.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x2d.device, dtype=torch.float32)
    check(_lib.lib().sd_ln_stats(x2d.data_ptr(), d, M, d, mean.data_ptr(), rstd.data_ptr(), LN_EPS, stream_ptr()),
          "sd_ln_stats")
    _count()
    return mean, rstd


def ln_bwd(g, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, M, d):
    check(_lib.lib().sd_ln_bwd(g.data_ptr(), d, x.data_ptr(), d, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                               None if dres is None else dres.data_ptr(), d, dx.data_ptr(), d, dgamma.data_ptr(),
                               dbeta.data_ptr(), M, d, stream_ptr()), "sd_ln_bwd")
    _count()


def _attn_tc_ok(ptrs, lds, T, M, dh):
    return (_lib.lib().sd_attention_tc_supported(T, M, dh) == 1 and all(p_ % 16 == 0 for p_ in ptrs)
            and all(l % 4 == 0 for l in lds))


def attention_fwd(Q, ldq, K, ldk, V, ldv, O, ldo, lse, B, H, T, M, dh, dropout=None, precision=PREC_FP32):
    p, seed, sid = dropout if dropout is not None else (0.0, 0, 0)
    tc = precision == PREC_BF16 and _attn_tc_ok((Q, K, V, O), (ldq, ldk, ldv, ldo), T, M, dh)
    fn = _lib.lib().sd_attention_tc_fwd if tc else _lib.lib().sd_attention_fwd
    with _Timed("tc_attention_fwd" if tc else "attention_fwd", 4.0 * B * H * T * M * dh, 4.0 * B * H * dh * (2 * T + 2 * M),
                f"[B{B} H{H} T{T} M{M} dh{dh}]"):
        check(fn(Q, ldq, K, ldk, V, ldv, O, ldo, lse, B, H, T, M, dh, p, seed, sid, stream_ptr()), "sd_attention_fwd")
    _count()


def attention_bwd(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, dQ, lddq, dK, lddk, dV, lddv, B, H, T, M, dh,
                  dropout=None, precision=PREC_FP32):
    p, seed, sid = dropout if dropout is not None else (0.0, 0, 0)
    tc = precision == PREC_BF16 and _attn_tc_ok((Q, K, V, O, dO, dQ, dK, dV), (ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv),
                                                T, M, dh)
    fn = _lib.lib().sd_attention_tc_bwd if tc else _lib.lib().sd_attention_bwd
    with _Timed("tc_attention_bwd" if tc else "attention_bwd", 10.0 * B * H * T * M * dh, 4.0 * B * H * dh * (4 * T + 4 * M),
                f"[B{B} H{H} T{T} M{M} dh{dh}]"):
        check(fn(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, dQ, lddq, dK, lddk, dV, lddv, B, H, T, M, dh, p, seed, sid,
                 stream_ptr()), "sd_attention_bwd")
    _count()


def step_token(t: torch.Tensor, freqs, token, out, ld_out, B, d):
    if t.dtype == torch.int64:
        is_float = 0
    elif t.dtype == torch.float32:
        is_float = 1
    else:
        raise _lib.SdError(f"step must be int64 or float32, got {t.dtype}")
    check(_lib.lib().sd_step_token(t.data_ptr(), is_float, freqs.data_ptr(), token.data_ptr(),
                                   out if isinstance(out, int) else out.data_ptr(), ld_out, B, d, stream_ptr()),
          "sd_step_token")
    _count()


def step_token_bwd(dout_ptr, ld, B, d, dtoken):
    check(_lib.lib().sd_step_token_bwd(dout_ptr, ld, B, d, dtoken.data_ptr(), stream_ptr()), "sd_step_token_bwd")
    _count()


def q_sample(joint_command, mean, std, noise, t, acp, x0_out, xt_out):
    B = joint_command.shape[0]
    inner = joint_command.numel() // max(B, 1)
    J = joint_command.shape[-1]
    check(_lib.lib().sd_q_sample(joint_command.data_ptr(), _lib.ptr(mean), _lib.ptr(std), noise.data_ptr(),
                                 t.data_ptr(), acp.data_ptr(), acp.numel(), _lib.ptr(x0_out), xt_out.data_ptr(), B,
                                 inner, J, stream_ptr()), "sd_q_sample")
    _count()


def ddim_step(x, eps, prev, x0_pred, coef):
    sb, sa, sap, sbp = coef
    check(_lib.lib().sd_ddim_step(x.data_ptr(), eps.data_ptr(), prev.data_ptr(), _lib.ptr(x0_pred), x.numel(), sb, sa,
                                  sap, sbp, stream_ptr()), "sd_ddim_step")
    _count()


def mse_fwd(pred, target, out):
    check(_lib.lib().sd_mse_fwd(pred.data_ptr(), target.data_ptr(), pred.numel(), out.data_ptr(), stream_ptr()),
          "sd_mse_fwd")
    _count()


def mse_bwd(pred, target, grad_loss, grad_pred):
    check(_lib.lib().sd_mse_bwd(pred.data_ptr(), target.data_ptr(), pred.numel(), _lib.ptr(grad_loss),
                                grad_pred.data_ptr(), stream_ptr()), "sd_mse_bwd")
    _count()


def affine_joints(x, mean, std, out, mode):
    check(_lib.lib().sd_affine_joints(x.data_ptr(), mean.data_ptr(), std.data_ptr(), out.data_ptr(), x.numel(),
                                      x.shape[-1], mode, stream_ptr()), "sd_affine_joints")
    _count()


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, wd, step, grad_scale=1.0):
    with _Timed("adamw", 0.0, 28.0 * p.numel()):
        check(_lib.lib().sd_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1,
                                       beta2, eps, wd, step, grad_scale, stream_ptr()), "sd_adamw_step")
    _count()


def adamw_step_dev(p, g, m, v, hyper_dev, step_dev=None):
    """``step_dev``: 1-element int32 device tensor holding t (bias corrections computed in the kernel), or None."""
    with _Timed("adamw", 0.0, 28.0 * p.numel()):
        check(_lib.lib().sd_adamw_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                           hyper_dev.data_ptr(), None if step_dev is None else step_dev.data_ptr(),
                                           stream_ptr()), "sd_adamw_step_dev")
    _count()


def set_dropout_seed_offset(counter):
    """counter: 1-element int64/uint64 CUDA tensor (kept alive by the caller) or None."""
    check(_lib.lib().sd_set_dropout_seed_offset(None if counter is None else counter.data_ptr()),
          "sd_set_dropout_seed_offset")


def gather_rows(table, idx, out, ld_out, err_flag=None):
    B, d = idx.numel(), table.shape[1]
    check(_lib.lib().sd_gather_rows(table.data_ptr(), idx.data_ptr(), table.shape[0],
                                    out if isinstance(out, int) else out.data_ptr(), ld_out, B, d,
                                    _lib.ptr(err_flag), stream_ptr()), "sd_gather_rows")
    _count()


def scatter_add_rows(dout_ptr, ld, idx, dtable):
    check(_lib.lib().sd_scatter_add_rows(dout_ptr, ld, idx.data_ptr(), dtable.shape[0], dtable.data_ptr(), idx.numel(),
                                         dtable.shape[1], stream_ptr()), "sd_scatter_add_rows")
    _count()


def colsum_accum(x, ld, M, N, out):
    check(_lib.lib().sd_colsum_accum(x if isinstance(x, int) else x.data_ptr(), ld, M, N,
                                     out if isinstance(out, int) else out.data_ptr(), stream_ptr()), "sd_colsum_accum")
    _count()


def copy_rows(src, sbs, sld, dst, dbs, dld, B, rows, cols, accumulate=False):
    check(_lib.lib().sd_copy_rows(src, sbs, sld, dst, dbs, dld, B, rows, cols, 1 if accumulate else 0, stream_ptr()),
          "sd_copy_rows")
    _count()


def add(a, b, y):
    check(_lib.lib().sd_add(a.data_ptr(), b.data_ptr(), y.data_ptr(), y.numel(), stream_ptr()), "sd_add")
    _count()


def dropout_mask(n, p, seed, stream_id, device):
    out = torch.empty(n, device=device, dtype=torch.float32)
    check(_lib.lib().sd_dropout_mask(out.data_ptr(), n, p, seed, stream_id, stream_ptr()), "sd_dropout_mask")
    _count()
    return out


def dropout_apply(x, p, seed, stream_id):
    y = torch.empty_like(x)
    check(_lib.lib().sd_dropout_apply(x.data_ptr(), y.data_ptr(), x.numel(), p, seed, stream_id, stream_ptr()),
          "sd_dropout_apply")
    _count()
    return y


# ---- image-trunk layers (bf16 NHWC) --------------------------------------------------------------------
def bn_stats(x, R, C, sums, eps, momentum, mean, invstd, running_mean, running_var):
    with _Timed("bn_stats", 0.0, 2.0 * R * C, f"[R{R} C{C}]"):
        check(_lib.lib().sd_bn_stats_nhwc_bf16(x.data_ptr(), R, C, sums.data_ptr(), eps, momentum, mean.data_ptr(),
                                               invstd.data_ptr(), _lib.ptr(running_mean), _lib.ptr(running_var),
                                               stream_ptr()), "sd_bn_stats_nhwc_bf16")
    _count(2)


def bn_apply(x, residual, mean, invstd, gamma, beta, relu, y, mask, R, C):
    with _Timed("bn_apply", 0.0, 2.0 * R * C * (3 if residual is not None else 2) + (R * C / 8 if mask is not None else 0),
                f"[R{R} C{C}]"):
        check(_lib.lib().sd_bn_apply_nhwc_bf16(x.data_ptr(), _lib.ptr(residual), mean.data_ptr(), invstd.data_ptr(),
                                               gamma.data_ptr(), beta.data_ptr(), 1 if relu else 0, y.data_ptr(),
                                               _lib.ptr(mask), R, C, stream_ptr()), "sd_bn_apply_nhwc_bf16")
    _count()


def bn_bwd(dy, y_relu, x, mean, invstd, gamma, sums, dx, dres, dgamma, dbeta, R, C, beta_recompute=None, dy2=None):
    """``dy2``: optional second incoming gradient (the output had two consumers); summed with ``dy`` inside the kernels."""
    nb = 2.0 * R * C * (2 * (2 + (1 if dy2 is not None else 0)) + 1 + (1 if dres is not None else 0)) \
        + (2.0 * R * C / 8 if y_relu is not None else 0)
    with _Timed("bn_bwd", 0.0, nb, f"[R{R} C{C}]"):
        check(_lib.lib().sd_bn_bwd2_nhwc_bf16(dy.data_ptr(), _lib.ptr(dy2), _lib.ptr(y_relu), x.data_ptr(), mean.data_ptr(),
                                              invstd.data_ptr(), gamma.data_ptr(), _lib.ptr(beta_recompute), sums.data_ptr(),
                                              dx.data_ptr(), _lib.ptr(dres), dgamma.data_ptr(), dbeta.data_ptr(), R, C,
                                              stream_ptr()),
              "sd_bn_bwd2_nhwc_bf16")
    _count(3)


def maxpool_fwd(x, y, idx, N, H, W, C):
    with _Timed("maxpool_fwd", 0.0, 2.0 * N * H * W * C * 1.25 + N * H * W * C / 4, f"[N{N} H{H} C{C}]"):
        check(_lib.lib().sd_maxpool3x3s2_nhwc_bf16_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), N, H, W, C, stream_ptr()),
              "sd_maxpool3x3s2_nhwc_bf16_fwd")
    _count()


def maxpool_bwd(dy, idx, dx, N, H, W, C):
    with _Timed("maxpool_bwd", 0.0, 2.0 * N * H * W * C * 1.25 + N * H * W * C / 4, f"[N{N} H{H} C{C}]"):
        check(_lib.lib().sd_maxpool3x3s2_nhwc_bf16_bwd(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), N, H, W, C, stream_ptr()),
              "sd_maxpool3x3s2_nhwc_bf16_bwd")
    _count()


def stem_fwd(x, mean, invstd, gamma, beta, y, idx, N, H, W, C):
    with _Timed("stem_bn_relu_pool_fwd", 0.0, 2.0 * N * H * W * C * 1.25 + N * H * W * C / 4, f"[N{N} H{H} C{C}]"):
        check(_lib.lib().sd_stem_bn_relu_pool_nhwc_bf16_fwd(x.data_ptr(), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                                            beta.data_ptr(), y.data_ptr(), idx.data_ptr(), N, H, W, C,
                                                            stream_ptr()), "sd_stem_bn_relu_pool_nhwc_bf16_fwd")
    _count()


def stem_band_supported(H, W, C) -> bool:
    return _lib.lib().sd_stem_band_supported(H, W, C) == 1


def stem_bwd(dpool, idx, x, mean, invstd, gamma, beta, sums, dx, dgamma, dbeta, N, H, W, C, y_pooled=None):
    """``y_pooled``: the forward's pooled output; with it the per-channel reductions run in the pooled domain (2 B + 2 B per
    pooled element) instead of a pass over x."""
    # dx pass: x (2 B/elem) + pooled gradient and taps (3 B per pooled elem) + dx; reductions: pooled pair or the same reads
    dx_pass = 2.0 * N * H * W * C * 2 + 3.0 * N * H * W * C / 4
    red_pass = 4.0 * N * H * W * C / 4 if y_pooled is not None else 2.0 * N * H * W * C + 3.0 * N * H * W * C / 4
    with _Timed("stem_bn_relu_pool_bwd", 0.0, dx_pass + red_pass, f"[N{N} H{H} C{C}]"):
        check(_lib.lib().sd_stem_bn_relu_pool_nhwc_bf16_bwd2(dpool.data_ptr(), idx.data_ptr(), x.data_ptr(), _lib.ptr(y_pooled),
                                                             mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                                             beta.data_ptr(), sums.data_ptr(), dx.data_ptr(), dgamma.data_ptr(),
                                                             dbeta.data_ptr(), N, H, W, C, stream_ptr()),
              "sd_stem_bn_relu_pool_nhwc_bf16_bwd2")
    _count(4 if y_pooled is not None else 3)


def stem_pack(images, out, N, H, W):
    with _Timed("stem_pack_s2d", 0.0, 12.0 * N * H * W + 32.0 * N * ((H + 6) // 2) * ((W + 6) // 2), f"[N{N} H{H}]"):
        check(_lib.lib().sd_stem_pack_s2d_bf16(images.data_ptr(), out.data_ptr(), N, H, W, stream_ptr()),
              "sd_stem_pack_s2d_bf16")
    _count()


IMAGENET_MEAN = (0.485, 0.456, 0.406)   # v2.Normalize constants of the reference (dataset/pytorch.py:202, ros.py:194)
IMAGENET_STD = (0.229, 0.224, 0.225)


def stem_pack_u8(images_u8, out, N, H, W, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """uint8 (N,3,H,W) -> normalised, space-to-depth packed bf16 image (the host preprocessing moved onto the device)."""
    with _Timed("stem_pack_s2d", 0.0, 3.0 * N * H * W + 32.0 * N * ((H + 6) // 2) * ((W + 6) // 2), f"[u8 N{N} H{H}]"):
        check(_lib.lib().sd_stem_pack_s2d_u8(images_u8.data_ptr(), out.data_ptr(), N, H, W, *[float(m) for m in mean],
                                             *[float(v) for v in std], stream_ptr()), "sd_stem_pack_s2d_u8")
    _count()


def stem_wgrad(xs2d, dy, dw, N, H, W):
    P = N * (H // 2) * (W // 2)
    with _Timed("stem_wgrad_s2d", 2.0 * 256 * 64 * P, 32.0 * N * ((H + 6) // 2) * ((W + 6) // 2) + 128.0 * P, f"[N{N} H{H}]"):
        check(_lib.lib().sd_stem_wgrad_s2d_bf16(xs2d.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, stream_ptr()),
              "sd_stem_wgrad_s2d_bf16")
    _count(2)


def stem_fprop(xs2d, w_s2d, y, N, H, W, sums=None) -> bool:
    """conv1 forward on the packed image; False when the TMA/tcgen05 kernel does not support this shape.  ``sums``
    (float64[128]): the epilogue also accumulates bn1's batch statistics (sum y, sum y^2 per channel)."""
    P = N * (H // 2) * (W // 2)
    with _Timed("stem_fprop_s2d", 2.0 * 256 * 64 * P, 32.0 * N * ((H + 6) // 2) * ((W + 6) // 2) + 128.0 * P, f"[N{N} H{H}]"):
        rc = _lib.lib().sd_stem_fprop_s2d_bf16_stats(xs2d.data_ptr(), w_s2d.data_ptr(), y.data_ptr(),
                                                     sums.data_ptr() if sums is not None else None, N, H, W, stream_ptr())
    if rc == -2:
        return False
    check(rc, "sd_stem_fprop_s2d_bf16_stats")
    _count()
    return True


def bn_finalize(sums, R, C, eps, momentum, mean, invstd, running_mean, running_var):
    check(_lib.lib().sd_bn_finalize(sums.data_ptr(), R, C, eps, momentum, mean.data_ptr(), invstd.data_ptr(),
                                    running_mean.data_ptr() if running_mean is not None else None,
                                    running_var.data_ptr() if running_var is not None else None, stream_ptr()),
          "sd_bn_finalize")
    _count()


# ---- layer-fused tensor-core path (bf16 mode, d_model = ff = 128) ------------------------------------------------------
ENC_ROWS_PER_LAYER = 768     # in_proj (384) | out_proj | linear1 | linear2
WGRAD_MAX_JOBS = _lib.WGRAD_MAX_JOBS
KV_MAX_LAYERS = _lib.KV_MAX_LAYERS
DEC_ROWS_PER_LAYER = 1280    # sa.in_proj (384) | sa.out_proj | ca.in_proj (384) | ca.out_proj | linear1 | linear2


def pack_weights_bf16(mats, dst: torch.Tensor, K: int):
    """``mats``: list of (fp32 weight tensor or raw pointer, rows, first destination row) -> rows of the packed bf16 matrix ``dst``
    (one launch per 64 matrices)."""
    lib = _lib.lib()
    for i0 in range(0, len(mats), _lib.PACK_MAX_SEGMENTS):
        chunk = mats[i0: i0 + _lib.PACK_MAX_SEGMENTS]
        a = _lib.PackArgs()
        for i, (w, rows, row0) in enumerate(chunk):
            a.src[i] = w if isinstance(w, int) else w.data_ptr()
            a.rows[i] = rows
            a.dst_row0[i] = row0
        a.K = K
        a.dst = dst.data_ptr()
        check(lib.sd_pack_weights_bf16(C.byref(a), len(chunk), stream_ptr()), "sd_pack_weights_bf16")
        _count()


def enc_layer_supported(d: int, ff: int, S: int, H: int) -> bool:
    return bool(_lib.lib().sd_enc_layer_supported(d, ff, S, H))


LAYER_SA, LAYER_FFN, LAYER_FFN_FIRST = 1, 2, 4   # sd_enc_layer_desc.blocks


def _dp(t):
    """tensor -> device pointer; raw integer pointers (views at an element offset) and None pass through"""
    return t if (t is None or isinstance(t, int)) else t.data_ptr()


def enc_layer_fwd(x, y, B, S, H, w_packed, w_row0, in_b, out_b, l1_b, l2_b, n1_w, n1_b, n2_w, n2_b, saves=None, dropout=None,
                  blocks=0, w_row_ffn=0, dropout_stream_ffn=0):
    """One fused layer forward (sd_enc_layer_fwd).  ``saves`` = (x1 fp32, xn1, attn, xn2, hact bf16) or None; ``blocks``:
    0 = whole encoder layer, LAYER_SA / LAYER_FFN = one half (decoder layers; unused parameters / saves may be None)."""
    d = _lib.EncLayerDesc()
    d.x, d.y, d.B, d.S, d.H = x.data_ptr(), y.data_ptr(), B, S, H
    d.w_packed, d.w_rows_total, d.w_row0 = w_packed.data_ptr(), w_packed.shape[0], w_row0
    d.in_b, d.out_b, d.l1_b, d.l2_b = _dp(in_b), _dp(out_b), _dp(l1_b), _dp(l2_b)
    d.n1_w, d.n1_b, d.n2_w, d.n2_b = _dp(n1_w), _dp(n1_b), _dp(n2_w), _dp(n2_b)
    if saves is not None:
        d.x1_save, d.xn1_save, d.attn_save, d.xn2_save, d.hact_save = (_dp(t) for t in saves)
    if dropout is not None and dropout[0] > 0.0:
        d.dropout_p, d.dropout_seed, d.dropout_stream = dropout
    d.blocks, d.w_row_ffn, d.dropout_stream_ffn = blocks, w_row_ffn, dropout_stream_ffn
    M = B * S
    sa, ffn = blocks == 0 or bool(blocks & LAYER_SA), blocks == 0 or bool(blocks & LAYER_FFN)
    # algorithmic work: attention block 8 S d^2 + 4 S^2 d, feed-forward block 4 S d^2 per sample
    flops = B * ((8.0 * S * 128 * 128 + 4.0 * S * S * 128) * sa + 4.0 * S * 128 * 128 * ffn)
    name = ("fused_enc_layer_fwd" if blocks == 0 else "fused_ffn_sa_blocks_fwd" if blocks & LAYER_FFN_FIRST else
            "fused_sa_block_fwd" if sa else "fused_ffn_block_fwd")
    with _Timed(name, flops, M * 128 * (8.0 + (4.0 + 8.0 if saves is not None else 0.0)), f"[B{B} S{S} H{H}]"):
        check(_lib.lib().sd_enc_layer_fwd(C.byref(d), stream_ptr()), "sd_enc_layer_fwd")
    _count()


def enc_layer_bwd(dy, dx, x, x1, xn1, xn2, g2, dhpre, g1, dqkv, g_n1_w, g_n1_b, g_n2_w, g_n2_b, B, S, H, w_packed, w_row0,
                  in_b, l1_b, n1_w, n2_w, dropout=None, blocks=0, w_row_ffn=0, dropout_stream_ffn=0):
    """Data-path backward of one fused layer / half layer (sd_enc_layer_bwd)."""
    d = _lib.EncLayerBwdDesc()
    d.dy, d.dx, d.x, d.x1, d.xn1, d.xn2 = (_dp(t) for t in (dy, dx, x, x1, xn1, xn2))
    d.g2, d.dhpre, d.g1, d.dqkv = (_dp(t) for t in (g2, dhpre, g1, dqkv))
    d.g_n1_w, d.g_n1_b, d.g_n2_w, d.g_n2_b = (_dp(t) for t in (g_n1_w, g_n1_b, g_n2_w, g_n2_b))
    d.B, d.S, d.H = B, S, H
    d.w_packed, d.w_rows_total, d.w_row0 = w_packed.data_ptr(), w_packed.shape[0], w_row0
    d.in_b, d.l1_b, d.n1_w, d.n2_w = _dp(in_b), _dp(l1_b), _dp(n1_w), _dp(n2_w)
    if dropout is not None and dropout[0] > 0.0:
        d.dropout_p, d.dropout_seed, d.dropout_stream = dropout
    d.blocks, d.w_row_ffn, d.dropout_stream_ffn = blocks, w_row_ffn, dropout_stream_ffn
    M = B * S
    sa, ffn = blocks in (0, LAYER_SA), blocks in (0, LAYER_FFN)
    # data gradients + the recomputed forward GEMMs (weight gradients are sd_wgrad_bf16's)
    flops = B * ((12.0 * S * 128 * 128 + 12.0 * S * S * 128) * sa + 6.0 * S * 128 * 128 * ffn)
    name = "fused_enc_layer_bwd" if blocks == 0 else ("fused_sa_block_bwd" if sa else "fused_ffn_block_bwd")
    with _Timed(name, flops, M * 128 * (4.0 * 4 + 2.0 * 2 + 2.0 * 6), f"[B{B} S{S} H{H}]"):
        check(_lib.lib().sd_enc_layer_bwd(C.byref(d), stream_ptr()), "sd_enc_layer_bwd")
    _count()


def wgrad_bf16(jobs, rows: int):
    """``jobs``: list of (G bf16 [rows][ldg], g_col0, X bf16 [rows][ldx], x_col0, dW fp32 view [128][ldw], ldw, db or None):
    dW += G[:, g_col0:g_col0+128]^T X[:, x_col0:x_col0+128], db += column sums of G (sd_wgrad_bf16, one launch)."""
    arr = (_lib.WgradJob * len(jobs))()
    for a, (G, g_col0, X, x_col0, dW, ldw, db) in zip(arr, jobs):
        a.G, a.ldg, a.g_col0 = G.data_ptr(), G.shape[1], g_col0
        a.X, a.ldx, a.x_col0 = X.data_ptr(), X.shape[1], x_col0
        a.dW = dW if isinstance(dW, int) else dW.data_ptr()
        a.ldw = ldw
        a.db = None if db is None else (db if isinstance(db, int) else db.data_ptr())
    with _Timed("tma_wgrad", 2.0 * rows * 128 * 128 * len(jobs), 2.0 * rows * 256 * len(jobs), f"[{len(jobs)}x{rows}]"):
        check(_lib.lib().sd_wgrad_bf16(arr, len(jobs), rows, stream_ptr()), "sd_wgrad_bf16")
    _count()


# ---- layer-fused cross-attention block (csrc/cross_attn.cu) --------------------------------------------------------
def set_pdl(on: bool) -> bool:
    """Programmatic dependent launch for the kernel chains (sd_set_pdl); returns the previous setting."""
    return bool(_lib.lib().sd_set_pdl(1 if on else 0))


def ca_block_supported(d: int, H: int, T: int, M: int) -> bool:
    return bool(_lib.lib().sd_ca_block_supported(d, H, T, M))


def ca_block_fwd_supported(d: int, H: int, T: int, M: int) -> bool:
    """forward without saves (inference): T <= 64"""
    return bool(_lib.lib().sd_ca_block_fwd_supported(d, H, T, M))


def cast_bf16(src: torch.Tensor, dst: torch.Tensor):
    """fp32 -> bf16 copy of a contiguous tensor (numel a multiple of 8)."""
    n = src.numel()
    with _Timed("cast_bf16", 0.0, 6.0 * n, f"[{n}]"):
        check(_lib.lib().sd_cast_bf16(_f32(src).data_ptr(), dst.data_ptr(), n, stream_ptr()), "sd_cast_bf16")
    _count()


def bcast_row_bf16(dst, block_rows: int, row: int, B: int, src_row_ptr: int, ncols: int):
    """dst[b*block_rows + row, :ncols] = the bf16 row at ``src_row_ptr`` for every b < B (sd_bcast_row_bf16)."""
    check(_lib.lib().sd_bcast_row_bf16(dst.data_ptr(), dst.shape[1], block_rows, row, B, src_row_ptr, ncols, stream_ptr()),
          "sd_bcast_row_bf16")
    _count()


DDIM_GLUE_MAX_J = 32


def ddim_glue(h, fc_w, fc_b, x, x_next, eps_out, coef, emb=None, kv_bcast=None):
    """sd_ddim_glue: eps = h fc_w^T + fc_b; x_next = DDIM(x, eps); optionally h <- embedding(x_next) + PE (``emb`` = (emb_w, emb_b,
    pe, T)) and the next step token's K | V row into every trajectory (``kv_bcast`` = (kv, block_rows, row, B, src_ptr, ncols))."""
    rows, J = x.shape
    sb, sa, sap, sbp = coef
    ew, eb, pe, T = emb if emb is not None else (None, None, None, 0)
    kv, block_rows, row, B, src, ncols = kv_bcast if kv_bcast is not None else (None, 0, 0, 0, None, 0)
    check(_lib.lib().sd_ddim_glue(h.data_ptr(), fc_w.data_ptr(), fc_b.data_ptr(), x.data_ptr(), x_next.data_ptr(), _dp(eps_out), rows, J,
                                  sb, sa, sap, sbp, _dp(ew), _dp(eb), _dp(pe), T, h.data_ptr() if emb is not None else None,
                                  _dp(kv), kv.shape[1] if kv is not None else 0, block_rows, row, B, src, ncols, stream_ptr()),
          "sd_ddim_glue")
    _count()


def kv_proj_bf16(mem_bf, w_packed, w_row0: int, w_stride: int, biases, kv_out):
    """kv_out[:, 256 l : 256 l + 256] = mem_bf @ w_packed[w_row0 + l*w_stride : +256].T + biases[l]  for every layer l
    (sd_kv_proj_bf16; ``biases``: list of raw pointers / tensors of 256 floats, one per layer)."""
    rows, L = mem_bf.shape[0], len(biases)
    arr = (_lib.c_f * L)(*[b if isinstance(b, int) else b.data_ptr() for b in biases])
    with _Timed("kv_proj_all_layers", 2.0 * rows * 128 * 256 * L, rows * (256.0 + 512.0 * L), f"[R{rows} L{L}]"):
        check(_lib.lib().sd_kv_proj_bf16(mem_bf.data_ptr(), rows, w_packed.data_ptr(), w_packed.shape[0], w_row0, w_stride, L, arr,
                                         kv_out.data_ptr(), kv_out.shape[1], stream_ptr()), "sd_kv_proj_bf16")
    _count()


def kv_dgrad_bf16(dkv, w_packed, w_row0: int, w_stride: int, L: int, dmem, accumulate: bool):
    """dmem (+)= dkv @ [Wkv_0; ...; Wkv_{L-1}]  (sd_kv_dgrad_bf16; dmem fp32 [rows][128])."""
    rows = dkv.shape[0]
    with _Timed("kv_dgrad_all_layers", 2.0 * rows * 128 * 256 * L, rows * (512.0 * L + 512.0), f"[R{rows} L{L}]"):
        check(_lib.lib().sd_kv_dgrad_bf16(dkv.data_ptr(), rows, dkv.shape[1], w_packed.data_ptr(), w_packed.shape[0], w_row0, w_stride,
                                          L, _f32(dmem).data_ptr(), dmem.shape[1], int(accumulate), stream_ptr()), "sd_kv_dgrad_bf16")
    _count()


def ca_block_fwd(x, y, B, T, M, w_packed, w_row_q, w_row_o, kv, kv_col0, q_b, out_b, n_w, n_b, saves=None, dropout=None):
    """Fused cross-attention block forward (sd_ca_block_fwd).  ``saves`` = (xn, q, attn bf16 [B*T][128], stats fp32 [B*T][2],
    lse fp32 [B][4][T]) or None; ``q_b`` / ``out_b`` may be raw pointers."""
    d = _lib.CaBlockDesc()
    d.x, d.y, d.B, d.T, d.M = x.data_ptr(), y.data_ptr(), B, T, M
    d.w_packed, d.w_rows_total, d.w_row_q, d.w_row_o = w_packed.data_ptr(), w_packed.shape[0], w_row_q, w_row_o
    d.kv, d.ldkv, d.kv_col0 = kv.data_ptr(), kv.shape[1], kv_col0
    d.q_b, d.out_b, d.n_w, d.n_b = _dp(q_b), _dp(out_b), _dp(n_w), _dp(n_b)
    if saves is not None:
        d.xn_save, d.q_save, d.attn_save, d.stats_save, d.lse_save = (_dp(t) for t in saves)
    if dropout is not None and dropout[0] > 0.0:
        d.dropout_p, d.dropout_seed, d.dropout_stream = dropout
    # LN + q-proj + out-proj (4 T d^2) + scores and P.V (4 T M d) per sample; reads x, K | V (bf16), writes y
    flops = B * (4.0 * T * 128 * 128 + 4.0 * T * M * 128)
    with _Timed("fused_ca_block_fwd", flops, B * (T * 128 * 8.0 + M * 512.0), f"[B{B} T{T} M{M}]"):
        check(_lib.lib().sd_ca_block_fwd(C.byref(d), stream_ptr()), "sd_ca_block_fwd")
    _count()


def ca_block_bwd(dy, dx, x, q, attn, stats, lse, B, T, M, w_packed, w_row_q, w_row_o, kv, kv_col0, n_w, g1, dq, dkv, g_n_w, g_n_b,
                 dropout=None):
    """Fused cross-attention block backward, data path (sd_ca_block_bwd)."""
    d = _lib.CaBlockBwdDesc()
    d.dy, d.dx, d.x, d.q, d.attn, d.stats, d.lse = (_dp(t) for t in (dy, dx, x, q, attn, stats, lse))
    d.B, d.T, d.M = B, T, M
    d.w_packed, d.w_rows_total, d.w_row_q, d.w_row_o = w_packed.data_ptr(), w_packed.shape[0], w_row_q, w_row_o
    d.kv, d.ldkv, d.kv_col0 = kv.data_ptr(), kv.shape[1], kv_col0
    d.n_w = _dp(n_w)
    d.g1, d.dq, d.dkv, d.lddkv = _dp(g1), _dp(dq), dkv.data_ptr(), dkv.shape[1]
    d.g_n_w, d.g_n_b = _dp(g_n_w), _dp(g_n_b)
    if dropout is not None and dropout[0] > 0.0:
        d.dropout_p, d.dropout_seed, d.dropout_stream = dropout
    flops = B * (4.0 * T * 128 * 128 + 10.0 * T * M * 128)
    with _Timed("fused_ca_block_bwd", flops, B * (T * 128 * 16.0 + M * 1024.0), f"[B{B} T{T} M{M}]"):
        check(_lib.lib().sd_ca_block_bwd(C.byref(d), stream_ptr()), "sd_ca_block_bwd")
    _count()


def conv1x1s2_dgrad_supported(Hin: int, Win: int, Cin: int, Cout: int) -> bool:
    return bool(_lib.lib().sd_conv1x1s2_dgrad_supported(Hin, Win, Cin, Cout))


def conv1x1s2_dgrad(dy, w_bf16, dx, frames, Hin, Win, Cin, Cout):
    """dx (bf16 NHWC storage [frames][Hin][Win][Cin]) = data gradient of a 1x1 stride-2 convolution (sd_conv1x1s2_dgrad_bf16)."""
    rows = frames * (Hin // 2) * (Win // 2)
    with _Timed("conv1x1s2_dgrad", 2.0 * rows * Cin * Cout, rows * 2.0 * Cout + frames * Hin * Win * Cin * 2.0,
                f"[N{frames} H{Hin} C{Cin}->{Cout}]"):
        check(_lib.lib().sd_conv1x1s2_dgrad_bf16(dy.data_ptr(), w_bf16.data_ptr(), dx.data_ptr(), frames, Hin, Win, Cin, Cout,
                                                 stream_ptr()), "sd_conv1x1s2_dgrad_bf16")
    _count()


def conv3x3_wgrad_c64_supported(H: int, W: int) -> bool:
    return bool(_lib.lib().sd_conv3x3_wgrad_c64_supported(H, W))


def conv3x3_wgrad_c64(x, dy, dW, frames, H, W, accumulate=False):
    """dW (fp32 [64][64][3][3]) (+)= weight gradient of a 3x3/s1/p1 64->64 convolution; x, dy bf16 NHWC storage
    (sd_conv3x3_wgrad_c64_bf16).  The per-CTA partial-sum scratch (21.8 MB) is allocated per call from the caching allocator
    (stream-ordered reuse; inside a CUDA-graph capture it comes from the graph's pool)."""
    sc = torch.empty(_lib.lib().sd_conv3x3_wgrad_c64_scratch_bytes() // 4, device=x.device, dtype=torch.float32)
    with _Timed("conv3x3_wgrad_c64", 2.0 * frames * H * W * 9 * 64 * 64, frames * H * W * 256.0, f"[N{frames} H{H}]"):
        check(_lib.lib().sd_conv3x3_wgrad_c64_bf16(x.data_ptr(), dy.data_ptr(), _f32(dW).data_ptr(), frames, H, W, sc.data_ptr(),
                                                   int(accumulate), stream_ptr()), "sd_conv3x3_wgrad_c64_bf16")
    _count(2)
