"""Per-phase cycle breakdown of the cluster sampler (clock64 stamps written by CTA 0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soccerdiffusion_b200 as sd  # noqa: E402
from soccerdiffusion_b200 import _lib, config  # noqa: E402
from soccerdiffusion_b200.schedulers import DDIMScheduler  # noqa: E402

dev = torch.device("cuda", 0)
hp = dict(config.DEFAULT)
torch.manual_seed(0)
model = config.build_model(hp).to(dev).eval()
sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
sch.set_timesteps(30)
ctx = [torch.randn(1, 311, 128, device=dev)]
x_T = torch.randn(1, 10, 20, device=dev)
with torch.no_grad():
    model.sample(ctx, x_T, sch)
    plan = next(iter(model._plans.values()))
    buf = torch.zeros(30 * 96, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().sd_plan_set_debug_stamps(plan.handle, buf.data_ptr()), "dbg")
    model.sample(ctx, x_T, sch)
    torch.cuda.synchronize()
    _lib.lib().sd_plan_set_debug_stamps(plan.handle, None)
st = buf.view(30, 96).cpu().numpy()
names = ["step start"]
for l in range(4):
    names += [f"L{l} start", "LN1", "qkv gemm+push", "sync", "self-attn core", "sa out+sync+res", "LN2+q gemm+sync",
              "scores", "softmax", "PV partial", "PV barrier", "PV reduce+push", "sync", "combine", "ca out gemm+push",
              "sync", "residual"]
names += ["all layers (LN3+ffn1+sync+ffn2+sync+res of last layer)", "fc_out+ddim (step end)"]
n = len(names)
d = np.diff(st[5:25, :n].astype(np.float64), axis=1).mean(axis=0)
tot = (st[5:25, n - 1] - st[5:25, 0]).mean()
print(f"cycles per step (mean of steps 5..24): {tot:.0f}")
agg = {}
for i in range(1, n):
    nm = names[i].split(" ", 1)[1] if names[i].startswith("L") and names[i][1].isdigit() else names[i]
    agg[nm] = agg.get(nm, 0.0) + d[i - 1]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.0f} cyc {100*v/tot:5.1f}%  {k}")
