"""Smallest program that launches the three layer-fused kernels on the default shapes (for ncu):
one encoder layer, B=256 samples of S=100 tokens, dropout 0.1, forward (with saves) + backward + weight gradients, 3 times."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from soccerdiffusion_b200 import ops
from soccerdiffusion_b200.functional import EncoderStackFn, RunCfg
from test_gpu_fused import _stack_inputs

B, S, H, L = 256, 100, 4, 1
x, emb_w, emb_b, pe, layers, gout = _stack_inputs(B, S, L, seed=1)
cfg = RunCfg(precision=ops.PREC_BF16, p=0.1, seed=3, stream_base=0)
for it in range(3):
    for t in (emb_w, emb_b, *layers):
        t.grad = None
    EncoderStackFn.apply(cfg, B, S, H, pe, x, emb_w, emb_b, *layers).backward(gout)
torch.cuda.synchronize()
print("ok")
