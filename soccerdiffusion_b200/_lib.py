"""ctypes binding of libsd_b200.so (the C ABI declared in include/sd_b200.h).

There is NO CPU or PyTorch fallback: if the shared library is missing or a kernel call fails, the
product path raises.  The library is built in-tree by ``soccerdiffusion_b200.build`` /
``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsd_b200.so")

c_f = C.c_void_p  # device pointers travel as integers
c_ll = C.c_longlong
c_i = C.c_int
c_fl = C.c_float
c_ull = C.c_ulonglong
c_u = C.c_uint


class SdError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", c_f), ("lda", c_ll), ("a_layout", c_i),
        ("B", c_f), ("ldb", c_ll), ("b_layout", c_i),
        ("C", c_f), ("ldc", c_ll),
        ("M", c_i), ("N", c_i), ("K", c_i),
        ("precision", c_i),
        ("ln_mean", c_f), ("ln_rstd", c_f), ("ln_gamma", c_f), ("ln_beta", c_f),
        ("alpha", c_fl),
        ("bias", c_f),
        ("pre_out", c_f), ("ldp", c_ll),
        ("act", c_i),
        ("gelu_grad_src", c_f), ("ldg", c_ll),
        ("dropout_p", c_fl), ("dropout_seed", c_ull), ("dropout_stream", c_u),
        ("pe", c_f), ("pe_period", c_i),
        ("residual", c_f), ("ldr", c_ll),
        ("accumulate", c_i),
    ]


class PlanConfig(C.Structure):
    _fields_ = [("d", c_i), ("heads", c_i), ("layers", c_i), ("T", c_i), ("J", c_i), ("ctx_tokens", c_i)]


class DecoderLayerWeights(C.Structure):
    _fields_ = [(n, c_f) for n in (
        "sa_in_w", "sa_in_b", "sa_out_w", "sa_out_b",
        "ca_in_w", "ca_in_b", "ca_out_w", "ca_out_b",
        "lin1_w", "lin1_b", "lin2_w", "lin2_b",
        "norm1_w", "norm1_b", "norm2_w", "norm2_b", "norm3_w", "norm3_b",
    )]


PACK_MAX_SEGMENTS = 64


class PackArgs(C.Structure):
    _fields_ = [("src", c_f * PACK_MAX_SEGMENTS), ("rows", c_i * PACK_MAX_SEGMENTS), ("dst_row0", c_i * PACK_MAX_SEGMENTS),
                ("K", c_i), ("dst", c_f)]


class EncLayerDesc(C.Structure):
    _fields_ = [("x", c_f), ("y", c_f), ("B", c_i), ("S", c_i), ("H", c_i),
                ("w_packed", c_f), ("w_rows_total", c_i), ("w_row0", c_i),
                ("in_b", c_f), ("out_b", c_f), ("l1_b", c_f), ("l2_b", c_f),
                ("n1_w", c_f), ("n1_b", c_f), ("n2_w", c_f), ("n2_b", c_f),
                ("x1_save", c_f), ("xn1_save", c_f), ("attn_save", c_f), ("xn2_save", c_f), ("hact_save", c_f),
                ("dropout_p", c_fl), ("dropout_seed", c_ull), ("dropout_stream", c_u),
                ("blocks", c_i), ("w_row_ffn", c_i), ("dropout_stream_ffn", c_u)]


class EncLayerBwdDesc(C.Structure):
    _fields_ = [("dy", c_f), ("dx", c_f), ("x", c_f), ("x1", c_f), ("xn1", c_f), ("xn2", c_f),
                ("g2", c_f), ("dhpre", c_f), ("g1", c_f), ("dqkv", c_f),
                ("g_n1_w", c_f), ("g_n1_b", c_f), ("g_n2_w", c_f), ("g_n2_b", c_f),
                ("B", c_i), ("S", c_i), ("H", c_i),
                ("w_packed", c_f), ("w_rows_total", c_i), ("w_row0", c_i),
                ("in_b", c_f), ("l1_b", c_f), ("n1_w", c_f), ("n2_w", c_f),
                ("dropout_p", c_fl), ("dropout_seed", c_ull), ("dropout_stream", c_u),
                ("blocks", c_i), ("w_row_ffn", c_i), ("dropout_stream_ffn", c_u)]


WGRAD_MAX_JOBS = 8


class WgradJob(C.Structure):
    _fields_ = [("G", c_f), ("ldg", c_ll), ("g_col0", c_i), ("X", c_f), ("ldx", c_ll), ("x_col0", c_i),
                ("dW", c_f), ("ldw", c_ll), ("db", c_f)]


KV_MAX_LAYERS = 16


class CaBlockDesc(C.Structure):
    _fields_ = [("x", c_f), ("y", c_f), ("B", c_i), ("T", c_i), ("M", c_i),
                ("w_packed", c_f), ("w_rows_total", c_i), ("w_row_q", c_i), ("w_row_o", c_i),
                ("kv", c_f), ("ldkv", c_ll), ("kv_col0", c_i),
                ("q_b", c_f), ("out_b", c_f), ("n_w", c_f), ("n_b", c_f),
                ("xn_save", c_f), ("q_save", c_f), ("attn_save", c_f), ("stats_save", c_f), ("lse_save", c_f),
                ("dropout_p", c_fl), ("dropout_seed", c_ull), ("dropout_stream", c_u)]


class CaBlockBwdDesc(C.Structure):
    _fields_ = [("dy", c_f), ("dx", c_f), ("x", c_f), ("q", c_f), ("attn", c_f), ("stats", c_f), ("lse", c_f),
                ("B", c_i), ("T", c_i), ("M", c_i),
                ("w_packed", c_f), ("w_rows_total", c_i), ("w_row_q", c_i), ("w_row_o", c_i),
                ("kv", c_f), ("ldkv", c_ll), ("kv_col0", c_i),
                ("n_w", c_f),
                ("g1", c_f), ("dq", c_f), ("dkv", c_f), ("lddkv", c_ll),
                ("g_n_w", c_f), ("g_n_b", c_f),
                ("dropout_p", c_fl), ("dropout_seed", c_ull), ("dropout_stream", c_u)]


# name -> argtypes (restype is always int unless noted); mirrors include/sd_b200.h one to one
SIGNATURES = {
    "sd_abi_version": [],
    "sd_gemm": [C.POINTER(GemmDesc), c_f],
    "sd_ln_stats": [c_f, c_ll, c_ll, c_i, c_f, c_f, c_fl, c_f],
    "sd_ln_bwd": [c_f, c_ll, c_f, c_ll, c_f, c_f, c_f, c_f, c_ll, c_f, c_ll, c_f, c_f, c_ll, c_i, c_f],
    "sd_attention_fwd": [c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_ull, c_u, c_f],
    "sd_attention_bwd": [c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_f, c_ll, c_f, c_ll, c_f, c_ll,
                         c_i, c_i, c_i, c_i, c_i, c_fl, c_ull, c_u, c_f],
    "sd_attention_tc_supported": [c_i, c_i, c_i],
    "sd_attention_tc_fwd": [c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_ull, c_u, c_f],
    "sd_attention_tc_bwd": [c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_ll, c_f, c_f, c_ll, c_f, c_ll, c_f, c_ll,
                            c_i, c_i, c_i, c_i, c_i, c_fl, c_ull, c_u, c_f],
    "sd_step_token": [c_f, c_i, c_f, c_f, c_f, c_ll, c_i, c_i, c_f],
    "sd_step_token_bwd": [c_f, c_ll, c_i, c_i, c_f, c_f],
    "sd_q_sample": [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_f],
    "sd_ddim_step": [c_f, c_f, c_f, c_f, c_ll, c_fl, c_fl, c_fl, c_fl, c_f],
    "sd_mse_fwd": [c_f, c_f, c_ll, c_f, c_f],
    "sd_mse_bwd": [c_f, c_f, c_ll, c_f, c_f, c_f],
    "sd_affine_joints": [c_f, c_f, c_f, c_f, c_ll, c_i, c_i, c_f],
    "sd_adamw_step": [c_f, c_f, c_f, c_f, c_ll, c_fl, c_fl, c_fl, c_fl, c_fl, c_i, c_fl, c_f],
    "sd_adamw_step_dev": [c_f, c_f, c_f, c_f, c_ll, c_f, c_f, c_f],
    "sd_set_dropout_seed_offset": [c_f],
    "sd_gather_rows": [c_f, c_f, c_i, c_f, c_ll, c_i, c_i, c_f, c_f],
    "sd_scatter_add_rows": [c_f, c_ll, c_f, c_i, c_f, c_i, c_i, c_f],
    "sd_colsum_accum": [c_f, c_ll, c_ll, c_i, c_f, c_f],
    "sd_copy_rows": [c_f, c_ll, c_ll, c_f, c_ll, c_ll, c_i, c_i, c_i, c_i, c_f],
    "sd_add": [c_f, c_f, c_f, c_ll, c_f],
    "sd_dropout_mask": [c_f, c_ll, c_fl, c_ull, c_u, c_f],
    "sd_dropout_apply": [c_f, c_f, c_ll, c_fl, c_ull, c_u, c_f],
    "sd_bn_stats_nhwc_bf16": [c_f, c_ll, c_i, c_f, c_fl, c_fl, c_f, c_f, c_f, c_f, c_f],
    "sd_bn_apply_nhwc_bf16": [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_ll, c_i, c_f],
    "sd_bn_bwd_nhwc_bf16": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_ll, c_i, c_f],
    "sd_bn_bwd2_nhwc_bf16": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_ll, c_i, c_f],
    "sd_stem_pack_s2d_bf16": [c_f, c_f, c_i, c_i, c_i, c_f],
    "sd_stem_pack_s2d_u8": [c_f, c_f, c_i, c_i, c_i, c_fl, c_fl, c_fl, c_fl, c_fl, c_fl, c_f],
    "sd_stem_fprop_s2d_bf16": [c_f, c_f, c_f, c_i, c_i, c_i, c_f],
    "sd_stem_fprop_s2d_bf16_stats": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_f],
    "sd_bn_finalize": [c_f, c_ll, c_i, c_fl, c_fl, c_f, c_f, c_f, c_f, c_f],
    "sd_stem_wgrad_s2d_bf16": [c_f, c_f, c_f, c_i, c_i, c_i, c_f],
    "sd_stem_band_supported": [c_i, c_i, c_i],
    "sd_stem_bn_relu_pool_nhwc_bf16_fwd": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f],
    "sd_stem_bn_relu_pool_nhwc_bf16_bwd": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f],
    "sd_stem_bn_relu_pool_nhwc_bf16_bwd2": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f],
    "sd_maxpool3x3s2_nhwc_bf16_fwd": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f],
    "sd_maxpool3x3s2_nhwc_bf16_bwd": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f],
    "sd_plan_create": [C.POINTER(PlanConfig), C.POINTER(C.c_void_p)],
    "sd_plan_destroy": [C.c_void_p],
    "sd_plan_set_layer": [C.c_void_p, c_i, C.POINTER(DecoderLayerWeights), c_f],
    "sd_plan_set_io": [C.c_void_p, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f],
    "sd_plan_set_schedule": [C.c_void_p, c_i, C.POINTER(c_ll), C.POINTER(c_fl), c_f],
    "sd_plan_set_context": [C.c_void_p, c_f, c_i, c_f],
    "sd_plan_sample": [C.c_void_p, c_f, c_f, c_f, c_i, c_i, c_f],
    "sd_plan_denoise": [C.c_void_p, c_f, c_f, c_i, c_f, c_i, c_f],
    "sd_plan_set_sampler": [C.c_void_p, c_i],
    "sd_plan_last_sampler": [C.c_void_p],
    "sd_plan_set_debug_stamps": [C.c_void_p, c_f],
    "sd_pack_weights_bf16": [C.POINTER(PackArgs), c_i, c_f],
    "sd_enc_layer_supported": [c_i, c_i, c_i, c_i],
    "sd_enc_layer_fwd": [C.POINTER(EncLayerDesc), c_f],
    "sd_enc_layer_bwd": [C.POINTER(EncLayerBwdDesc), c_f],
    "sd_wgrad_bf16": [C.POINTER(WgradJob), c_i, c_ll, c_f],
    "sd_set_pdl": [c_i],
    "sd_conv3x3_wgrad_c64_supported": [c_i, c_i],
    "sd_conv3x3_wgrad_c64_scratch_bytes": [],
    "sd_conv3x3_wgrad_c64_bf16": [c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_i, c_f],
    "sd_debug_shifted_mma": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f],
    "sd_debug_ldtm": [c_i, c_i, c_i, c_f, c_f, c_f],
    "sd_conv1x1s2_dgrad_supported": [c_i, c_i, c_i, c_i],
    "sd_conv1x1s2_dgrad_bf16": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f],
    "sd_ddim_glue": [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_fl, c_fl, c_fl, c_fl, c_f, c_f, c_f, c_i, c_f, c_f, c_ll, c_ll, c_ll,
                     c_i, c_f, c_i, c_f],
    "sd_cast_bf16": [c_f, c_f, c_ll, c_f],
    "sd_bcast_row_bf16": [c_f, c_ll, c_ll, c_ll, c_i, c_f, c_i, c_f],
    "sd_kv_proj_bf16": [c_f, c_ll, c_f, c_i, c_i, c_i, c_i, C.POINTER(c_f), c_f, c_ll, c_f],
    "sd_kv_dgrad_bf16": [c_f, c_ll, c_ll, c_f, c_i, c_i, c_i, c_i, c_f, c_ll, c_i, c_f],
    "sd_ca_block_supported": [c_i, c_i, c_i, c_i],
    "sd_ca_block_fwd_supported": [c_i, c_i, c_i, c_i],
    "sd_ca_block_fwd": [C.POINTER(CaBlockDesc), c_f],
    "sd_ca_block_bwd": [C.POINTER(CaBlockBwdDesc), c_f],
}

_lib = None


def load(build_if_missing: bool = False):
    """Loads the shared library and binds every symbol of the header. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build

            _build.build()
        else:
            raise SdError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU/PyTorch fallback). "
                "Run `python -m soccerdiffusion_b200.build` or `__graft_entry__.build()`."
            )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = c_i
    lib.sd_error_string.argtypes = [c_i]
    lib.sd_error_string.restype = C.c_char_p
    if lib.sd_abi_version() != 1:
        raise SdError("libsd_b200.so ABI version mismatch")
    _lib = lib
    return lib


def lib():
    return load()


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().sd_error_string(rc).decode()
        raise SdError(f"{what}: {msg} (code {rc})")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL). Requires CUDA + fp32/int64 contiguity checks by callers."""
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SdError(
                "soccerdiffusion_b200 runs on CUDA (sm_100a) only; got a CPU tensor. There is no CPU fallback."
            )
