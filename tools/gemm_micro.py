"""Micro-benchmark of the bf16 tcgen05 GEMM (sd_gemm, precision = bf16) on the shapes of the default.yaml training step at
bs=256: CUDA-event time per launch, algorithmic GB/s (fp32 operands in HBM) and TF/s.  Inputs are re-allocated per shape
and cycled through 8 buffer sets (> 126 MB L2 for the large shapes).  Usage: python tools/gemm_micro.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soccerdiffusion_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
NSET = 8


def bench(name, M, N, K, kind, reps=40, **kw):
    sets = []
    for _ in range(NSET):
        if kind == "fwd":      # C[M,N] = A[M,K] W[N,K]^T
            A, B, C = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.empty(M, N, device=dev)
            args = (A, K, ops.MK, B, K, ops.NK, C, N, M, N, K)
        elif kind == "dgrad":  # dX[M,N] = dY[M,K] W[K,N]
            A, B, C = torch.randn(M, K, device=dev), torch.randn(K, N, device=dev), torch.empty(M, N, device=dev)
            args = (A, K, ops.MK, B, N, ops.KN, C, N, M, N, K)
        else:                  # dW[M,N] = dY[K,M]^T X[K,N]
            A, B, C = torch.randn(K, M, device=dev), torch.randn(K, N, device=dev), torch.empty(M, N, device=dev)
            args = (A, M, ops.KM, B, N, ops.KN, C, N, M, N, K)
        extra = {}
        if kw.get("residual"):
            extra.update(residual=torch.randn(M, N, device=dev), ldr=N)
        if kw.get("bias"):
            extra.update(bias=torch.randn(N, device=dev))
        if kw.get("gelu"):
            extra.update(act=ops.ACT_GELU, pre_out=torch.empty(M, N, device=dev), ldp=N)
        if kw.get("dropout"):
            extra.update(dropout=(0.1, 1234, 7))
        if kw.get("ln"):
            extra.update(ln=(torch.zeros(M, device=dev), torch.ones(M, device=dev), torch.ones(K, device=dev), torch.zeros(K, device=dev)))
        sets.append((args, extra))
    run = lambda i: ops.gemm(*sets[i % NSET][0], precision=ops.PREC_BF16, **sets[i % NSET][1])
    for i in range(NSET):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    byts = 4.0 * (M * K + N * K + M * N) + (4.0 * M * N if kw.get("residual") else 0) + (4.0 * M * N if kw.get("gelu") else 0)
    print(f"{name:34s} {M:6d}x{N:4d}x{K:6d} {us:8.1f} us {byts / us / 1e3:7.0f} GB/s {2.0 * M * N * K / us / 1e6:7.1f} TF/s")


bench("qkv proj (LN on load, bias)", 25600, 384, 128, "fwd", ln=True, bias=True)
bench("out proj (bias, drop, residual)", 25600, 128, 128, "fwd", bias=True, dropout=True, residual=True)
bench("ffn1 (LN, bias, gelu, drop)", 25600, 128, 128, "fwd", ln=True, bias=True, gelu=True, dropout=True)
bench("cross K/V proj (bias)", 79872, 256, 128, "fwd", bias=True)
bench("decoder proj", 2560, 128, 128, "fwd", bias=True)
bench("dgrad", 25600, 128, 128, "dgrad")
bench("dgrad cross K/V", 79872, 128, 256, "dgrad")
bench("wgrad", 128, 128, 25600, "wgrad")
bench("wgrad cross K/V", 256, 128, 79872, "wgrad")
