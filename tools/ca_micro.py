"""Micro-benchmark of the fused cross-attention kernels (csrc/cross_attn.cu): each kernel graph-replayed alone between CUDA
events.  python tools/ca_micro.py [B] [T] [M] [L] [p]      (--once: one plain launch of each, for ncu)"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from soccerdiffusion_b200 import ops  # noqa: E402
from tools.fused_micro import timed_graph  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    once = "--once" in sys.argv
    B = int(args[0]) if len(args) > 0 else 256
    T = int(args[1]) if len(args) > 1 else 10
    M = int(args[2]) if len(args) > 2 else 312
    L = int(args[3]) if len(args) > 3 else 4
    p = float(args[4]) if len(args) > 4 else 0.1
    d = 128
    gen = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=gen).cuda()
    stride = ops.DEC_ROWS_PER_LAYER
    wp = (r(L * stride, d) / math.sqrt(d)).to(torch.bfloat16)
    mem = r(B * M, d)
    mem_bf = torch.empty(B * M, d, device="cuda", dtype=torch.bfloat16)
    kv = torch.empty(B * M, 256 * L, device="cuda", dtype=torch.bfloat16)
    dkv = torch.empty_like(kv)
    biases = [0.1 * r(256) for _ in range(L)]
    x, dy = r(B * T, d), r(B * T, d)
    y, dx = torch.empty_like(x), torch.empty_like(x)
    q_b, out_b, n_w, n_b = 0.1 * r(d), 0.1 * r(d), 1 + 0.1 * r(d), 0.1 * r(d)
    b16 = lambda: torch.empty(B * T, d, device="cuda", dtype=torch.bfloat16)
    saves = (b16(), b16(), b16(), torch.empty(B * T, 2, device="cuda"), torch.empty(B, 4, T, device="cuda"))
    g1, dq = b16(), b16()
    gw, gb = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    dmem = torch.empty(B * M, d, device="cuda")
    drop = (p, 1, 3) if p > 0 else None
    legs = {
        "cast_bf16": (lambda: ops.cast_bf16(mem, mem_bf), 0.0, B * M * d * 6.0),
        "kv_proj_all_layers": (lambda: ops.kv_proj_bf16(mem_bf, wp, 640, stride, biases, kv), 2.0 * B * M * d * 256 * L,
                               B * M * (256.0 + 512.0 * L)),
        "ca_block_fwd": (lambda: ops.ca_block_fwd(x, y, B, T, M, wp, 512, 896, kv, 256, q_b, out_b, n_w, n_b, saves=saves, dropout=drop),
                         B * (4.0 * T * d * d + 4.0 * T * M * d), B * (M * 512.0 + T * d * 14.0)),
        "ca_block_bwd": (lambda: ops.ca_block_bwd(dy, dx, x, saves[1], saves[2], saves[3], saves[4], B, T, M, wp, 512, 896, kv, 256, n_w,
                                                  g1, dq, dkv, gw, gb, dropout=drop),
                         B * (4.0 * T * d * d + 10.0 * T * M * d), B * (M * 1024.0 + T * d * 22.0)),
        "kv_dgrad_all_layers": (lambda: ops.kv_dgrad_bf16(dkv, wp, 640, stride, L, dmem, False), 2.0 * B * M * d * 256 * L,
                                B * M * (512.0 * L + 512.0)),
        "kv_wgrad (8 jobs)": (lambda: ops.wgrad_bf16([(dkv, 128 * j, mem_bf, 0, torch.zeros(d, d, device="cuda"), d, None) for j in range(min(8, 2 * L))],
                                                     B * M), 2.0 * B * M * d * d * min(8, 2 * L), B * M * 512.0 * min(8, 2 * L)),
    }
    for name, (fn, flops, nbytes) in legs.items():
        if once:
            fn()
            torch.cuda.synchronize()
            continue
        ms = timed_graph(fn)
        print(f"{name:24s} B={B} T={T} M={M} L={L} p={p}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TF/s  {nbytes/ms/1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
