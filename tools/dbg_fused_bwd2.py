import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_fused import make_layer
from util_gpu import rel
from soccerdiffusion_b200 import ops

B, S, H = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (2, 100, 4)
d = 128; dh = d // H
gen = torch.Generator().manual_seed(1)
P = {k: v.cuda() for k, v in make_layer(d, d, gen).items()}
x = torch.randn(B * S, d, generator=gen).cuda()
dy = torch.randn(B * S, d, generator=gen).cuda()
F = torch.nn.functional
xn1 = F.layer_norm(x, (d,), P["n1_w"], P["n1_b"], 1e-5)
qkv = (xn1 @ P["in_w"].T + P["in_b"]).detach().requires_grad_(True)
q, k, v = (t.view(B, S, H, dh).transpose(1, 2) for t in qkv.split(d, dim=-1))
a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1) @ v
attn = a.transpose(1, 2).reshape(B * S, d)
x1 = x + attn @ P["out_w"].T + P["out_b"]
xn2 = F.layer_norm(x1, (d,), P["n2_w"], P["n2_b"], 1e-5)
hact = F.gelu(xn2 @ P["l1_w"].T + P["l1_b"])
y = x1 + hact @ P["l2_w"].T + P["l2_b"]
y.backward(dy)
ref = qkv.grad
wp = torch.empty(768, d, device="cuda", dtype=torch.bfloat16)
ops.pack_weights_bf16([(P["in_w"], 384, 0), (P["out_w"], d, 384), (P["l1_w"], d, 512), (P["l2_w"], d, 640)], wp, d)
yk = torch.empty_like(x); x1k = torch.empty_like(x)
bf = [torch.empty(B * S, d, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
ops.enc_layer_fwd(x, yk, B, S, H, wp, 0, P["in_b"], P["out_b"], P["l1_b"], P["l2_b"], P["n1_w"], P["n1_b"], P["n2_w"], P["n2_b"], saves=(x1k, *bf))
g2, dhpre, g1 = (torch.empty(B * S, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
dqkv = torch.zeros(B * S, 3 * d, device="cuda", dtype=torch.bfloat16)
dx = torch.empty_like(x)
gn = [torch.zeros(d, device="cuda") for _ in range(4)]
ops.enc_layer_bwd(dy, dx, x, x1k, bf[0], bf[2], g2, dhpre, g1, dqkv, *gn, B, S, H, wp, 0, P["in_b"], P["l1_b"], P["n1_w"], P["n2_w"])
torch.cuda.synchronize()
got = dqkv.float()
for blk, nm in enumerate("qkv"):
    print(nm, "all", round(rel(got[:, blk*128:(blk+1)*128], ref[:, blk*128:(blk+1)*128]), 4))
g = got[:, 256:].view(B, S, d); r_ = ref[:, 256:].view(B, S, d)
for b in range(B):
    print(" dv sample", b, "by 16-row blocks:", " ".join(f"{rel(g[b, i:i+16], r_[b, i:i+16]):.2f}" for i in range(0, S, 16)))
    print(" dv sample", b, "by 16-col blocks:", " ".join(f"{rel(g[b, :, i:i+16], r_[b, :, i:i+16]):.2f}" for i in range(0, d, 16)))
print("ratio sample0 row0:", (g[0, 0, :8] / r_[0, 0, :8]).tolist())

