"""Micro-benchmark of the layer-fused kernels against the kernel-per-op path (CUDA events, graph-replayed so that host
launch gaps are excluded).  python tools/fused_micro.py [B] [S] [H] [L]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from soccerdiffusion_b200 import ops, runtime  # noqa: E402
from soccerdiffusion_b200.functional import EncoderStackFn, RunCfg  # noqa: E402
from test_gpu_fused import make_layer  # noqa: E402


def timed_graph(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    L = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    d, kin = 128, 20
    gen = torch.Generator().manual_seed(0)
    layers = []
    for _ in range(L):
        P = make_layer(d, d, gen)
        layers += [P[k].cuda() for k in ("in_w", "in_b", "out_w", "out_b", "l1_w", "l1_b", "l2_w", "l2_b", "n1_w", "n1_b",
                                         "n2_w", "n2_b")]
    emb_w = (torch.randn(d, kin, generator=gen) / math.sqrt(kin)).cuda()
    emb_b = torch.zeros(d).cuda()
    pe = (0.1 * torch.randn(S, d, generator=gen)).cuda()
    x = torch.randn(B * S, kin, generator=gen).cuda()
    flops = B * L * (12.0 * S * d * d + 4.0 * S * S * d)
    for p in (0.0, 0.1):
        for fused in (True, False):
            runtime.set_fused_layers(fused)
            cfg = RunCfg(precision=ops.PREC_BF16, p=p, seed=3, stream_base=0)
            with torch.no_grad():
                ms = timed_graph(lambda: EncoderStackFn.apply(cfg, B, S, H, pe, x, emb_w, emb_b, *layers))
            print(f"enc stack fwd no_grad B={B} S={S} H={H} L={L} p={p} fused={fused}: {ms*1e3:.1f} us "
                  f"({flops/ms/1e9:.1f} TF/s algorithmic)")
    # training path: forward with saves + fused backward + TMA weight-gradient GEMM
    gout = torch.randn(B, S, d, device="cuda")
    params = [emb_w, emb_b, *layers]
    for t in params:
        t.requires_grad_(True)
    for p in (0.0, 0.1):
        for fused in (True, False):
            runtime.set_fused_layers(fused)
            cfg = RunCfg(precision=ops.PREC_BF16, p=p, seed=3, stream_base=0)

            def fb():
                for t in params:
                    t.grad = None
                EncoderStackFn.apply(cfg, B, S, H, pe, x, emb_w, emb_b, *layers).backward(gout)

            ms = timed_graph(fb)
            print(f"enc stack fwd+bwd B={B} S={S} H={H} L={L} p={p} fused={fused}: {ms*1e3:.1f} us ({3*flops/ms/1e9:.1f} TF/s algorithmic)")
    runtime.set_fused_layers(True)


if __name__ == "__main__":
    main()
