"""Small-shape calls of the round-2 kernels for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerdiffusion_b200 import ops  # noqa: E402

torch.manual_seed(0)
cl = torch.channels_last
d = 128
# layer1 weight gradient, downsample data gradient
x = torch.randn(3, 64, 8, 12, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
dy = torch.randn(3, 64, 8, 12, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
dW = torch.empty(64, 64, 3, 3, device="cuda")
ops.conv3x3_wgrad_c64(x, dy, dW, 3, 8, 12)
dyd = torch.randn(3, 128, 4, 6, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
wb = torch.randn(128, 64, device="cuda", dtype=torch.bfloat16)
ops.conv1x1s2_dgrad(dyd, wb, torch.empty_like(x), 3, 8, 12, 64, 128)
# K/V projection, its data gradient, cross-attention forward / backward, DDIM glue
B, T, M, L = 3, 10, 200, 2
stride = ops.DEC_ROWS_PER_LAYER
wp = (torch.randn(L * stride, d, device="cuda") / math.sqrt(d)).to(torch.bfloat16)
mem = torch.randn(B * M, d, device="cuda")
mem_bf = torch.empty(B * M, d, device="cuda", dtype=torch.bfloat16)
ops.cast_bf16(mem, mem_bf)
kv = torch.empty(B * M, 256 * L, device="cuda", dtype=torch.bfloat16)
biases = [0.1 * torch.randn(256, device="cuda") for _ in range(L)]
ops.kv_proj_bf16(mem_bf, wp, 640, stride, biases, kv)
xq, dyq = torch.randn(B * T, d, device="cuda"), torch.randn(B * T, d, device="cuda")
y, dx = torch.empty_like(xq), torch.empty_like(xq)
vec = lambda: 0.1 * torch.randn(d, device="cuda")
b16 = lambda: torch.empty(B * T, d, device="cuda", dtype=torch.bfloat16)
saves = (b16(), b16(), b16(), torch.empty(B * T, 2, device="cuda"), torch.empty(B, 4, T, device="cuda"))
q_b, out_b, n_w, n_b = vec(), vec(), 1 + vec(), vec()
for drop in (None, (0.1, 3, 5)):
    ops.ca_block_fwd(xq, y, B, T, M, wp, 512, 896, kv, 256, q_b, out_b, n_w, n_b, saves=saves, dropout=drop)
    dkv = torch.empty_like(kv)
    ops.ca_block_bwd(dyq, dx, xq, saves[1], saves[2], saves[3], saves[4], B, T, M, wp, 512, 896, kv, 256, n_w, b16(), b16(), dkv,
                     torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), dropout=drop)
dmem = torch.empty(B * M, d, device="cuda")
ops.kv_dgrad_bf16(dkv, wp, 640, stride, L, dmem, False)
J = 20
ops.ddim_glue(xq.clone(), torch.randn(J, d, device="cuda"), torch.randn(J, device="cuda"), torch.randn(B * T, J, device="cuda"),
              torch.empty(B * T, J, device="cuda"), None, (0.6, 0.8, 0.9, 0.43), emb=(torch.randn(d, J, device="cuda"), vec(), torch.randn(T, d, device="cuda"), T),
              kv_bcast=(kv, M, M - 1, B, kv.data_ptr(), 256 * L))
# fused layer kernel, feed-forward-first variant
wpe = (torch.randn(2 * stride, d, device="cuda") / math.sqrt(d)).to(torch.bfloat16)
h = torch.randn(B * T, d, device="cuda")
ops.enc_layer_fwd(h, h, B, T, 4, wpe, stride, torch.randn(3 * d, device="cuda"), vec(), vec(), vec(), 1 + vec(), vec(), 1 + vec(), vec(),
                  blocks=ops.LAYER_SA | ops.LAYER_FFN | ops.LAYER_FFN_FIRST, w_row_ffn=1024)
torch.cuda.synchronize()
print("sanitize run ok")
