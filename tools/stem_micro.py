"""Micro-benchmark of the stem kernels after conv1 at the bs=256 shape (2560 frames, 112x112x64 bf16 NHWC): CUDA-event
times and algorithmic GB/s of maxpool3x3s2(relu(bn(x))) forward and backward.  Usage: python tools/stem_micro.py [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soccerdiffusion_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
C, H, W = 64, 112, 112
HO, WO = 56, 56
dev = torch.device("cuda", 0)
x = torch.randn(N, H, W, C, device=dev).to(torch.bfloat16)
gamma = torch.rand(C, device=dev) + 0.5
beta = torch.randn(C, device=dev) * 0.2
mean = torch.zeros(C, device=dev)
invstd = torch.ones(C, device=dev)
sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
y = torch.empty(N, HO, WO, C, device=dev, dtype=torch.bfloat16)
idx = torch.empty(N, HO, WO, C, device=dev, dtype=torch.uint8)
dp = torch.randn(N, HO, WO, C, device=dev).to(torch.bfloat16)
dx = torch.empty_like(x)
dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


elems = N * H * W * C
pooled = N * HO * WO * C
t = timed(lambda: ops.bn_stats(x, N * H * W, C, sums, 1e-5, 0.1, mean, invstd, rm, rv))
print(f"bn_stats       {t:7.3f} ms  {2 * elems / t / 1e6:7.0f} GB/s")
t = timed(lambda: ops.stem_fwd(x, mean, invstd, gamma, beta, y, idx, N, H, W, C))
print(f"stem_fwd       {t:7.3f} ms  {(2 * elems + 3 * pooled) / t / 1e6:7.0f} GB/s")
t = timed(lambda: ops.stem_bwd(dp, idx, x, mean, invstd, gamma, beta, sums, dx, dg, db, N, H, W, C))
print(f"stem_bwd (2p)  {t:7.3f} ms  {(6 * elems + 6 * pooled) / t / 1e6:7.0f} GB/s   (reductions over pixels)")
t = timed(lambda: ops.stem_bwd(dp, idx, x, mean, invstd, gamma, beta, sums, dx, dg, db, N, H, W, C, y_pooled=y))
print(f"stem_bwd pooled{t:7.3f} ms  {(4 * elems + 7 * pooled) / t / 1e6:7.0f} GB/s   (reductions over pooling windows)")

# conv1 itself (space-to-depth packed image): pack, tcgen05 forward, tcgen05 weight gradient
img = torch.randn(N, 3, 224, 224, device=dev)
xp = torch.empty(N, 115, 115, 16, device=dev, dtype=torch.bfloat16)
w2 = torch.randn(64, 256, device=dev).to(torch.bfloat16)
yc = torch.empty(N, 112, 112, 64, device=dev, dtype=torch.bfloat16)
dws = torch.empty(256, 64, device=dev, dtype=torch.float32)
P = N * 112 * 112
t = timed(lambda: ops.stem_pack(img, xp, N, 224, 224))
print(f"stem_pack      {t:7.3f} ms  {(12.0 * N * 224 * 224 + 32.0 * N * 115 * 115) / t / 1e6:7.0f} GB/s")
t = timed(lambda: ops.stem_fprop(xp, w2, yc, N, 224, 224))
print(f"stem_fprop     {t:7.3f} ms  {2.0 * 256 * 64 * P / t / 1e9:7.0f} TF/s  {(32.0 * N * 115 * 115 + 128.0 * P) / t / 1e6:7.0f} GB/s")
t = timed(lambda: ops.stem_wgrad(xp, dx.view(N, 112, 112, 64), dws, N, 224, 224))
print(f"stem_wgrad     {t:7.3f} ms  {2.0 * 256 * 64 * P / t / 1e9:7.0f} TF/s")
