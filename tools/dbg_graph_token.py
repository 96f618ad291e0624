import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import synth
from util_gpu import synth_model, to_dev
from soccerdiffusion_b200 import runtime
from soccerdiffusion_b200.dataset.pytorch import Normalizer
from soccerdiffusion_b200.ml.training import FusedAdamW, GraphedTrainStep, train_step
from soccerdiffusion_b200.schedulers import DDIMScheduler
hp = synth.PATCH_HP
B, seed = 4, 21
noise = synth.synth_noise("eps", hp, B, seed).cuda()
t = synth.synth_timesteps(B, seed).cuda()
batches = [to_dev(synth.synth_batch(hp, B, seed + i)) for i in range(4)]
sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
runtime.set_dropout(0.0)
m1, _ = synth_model(hp, seed); m1.train()
o1 = FusedAdamW(m1.parameters(), lr=1e-3)
l1 = torch.optim.lr_scheduler.OneCycleLR(o1, max_lr=1e-3, total_steps=20)
n1 = Normalizer(m1.mean, m1.std)
m2, _ = synth_model(hp, seed); m2.train()
o2 = FusedAdamW(m2.parameters(), lr=1e-3)
l2 = torch.optim.lr_scheduler.OneCycleLR(o2, max_lr=1e-3, total_steps=20)
g = GraphedTrainStep(m2, o2, sch, batches[0], lr_scheduler=l2, warmup_steps=3, noise=noise, timesteps=t)
names = {id(p): n for n, p in m2.named_parameters()}
print("graph: touched", len(o2._touched), "of", len(o2._flat[0]["params"]), "runs", o2._runs(o2._flat[0])[:5])
print("untouched (graph):", [names[id(p)] for p in o2._flat[0]["params"] if id(p) not in o2._touched][:10])
for i in range(6):
    print("  hp eager", o1.param_groups[0]["lr"], o1.param_groups[0]["betas"], o1._flat[0]["step"], " graph", o2.param_groups[0]["lr"], o2.param_groups[0]["betas"], o2._flat[0]["step"])
    l1v = train_step(m1, o1, sch, n1, batches[0], lr_scheduler=l1, noise=noise, timesteps=t).item()
    if i == 0:
        names1 = {id(p): n for n, p in m1.named_parameters()}
        print("eager: touched", len(o1._touched), "untouched:", [names1[id(p)] for p in o1._flat[0]["params"] if id(p) not in o1._touched][:10])
    l2 = g(batches[0]).item()
    tk1, tk2 = m1.step_encoding.token, m2.step_encoding.token
    print(i, l1v, l2, "token diff", (tk1 - tk2).abs().max().item(), "grad", tk1.grad.abs().max().item(), tk2.grad.abs().max().item(),
          "grad diff", (tk1.grad - tk2.grad).abs().max().item())
