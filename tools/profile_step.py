"""Kernel-level time breakdown of one training step (torch.profiler / CUPTI): which kernels — ours and the library
trunk's — the step spends its time in.  Usage: python tools/profile_step.py [--precision bf16] [--batch 256] [--workload full]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soccerdiffusion_b200 as sd  # noqa: E402
from soccerdiffusion_b200 import config  # noqa: E402
from soccerdiffusion_b200.dataset.pytorch import Normalizer  # noqa: E402
from soccerdiffusion_b200.ml.training import FusedAdamW, train_step  # noqa: E402
from soccerdiffusion_b200.schedulers import DDIMScheduler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--rows", type=int, default=45)
ap.add_argument("--cudnn-benchmark", action="store_true")
args = ap.parse_args()
torch.backends.cudnn.benchmark = args.cudnn_benchmark
sd.set_precision(args.precision)
dev = torch.device("cuda", 0)
hp = dict(config.DEFAULT)
torch.manual_seed(0)
model = config.build_model(hp).to(dev).train()
opt = FusedAdamW(model.parameters(), lr=1e-4)
sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
norm = Normalizer(model.mean, model.std)
batch = config.synthetic_batch(hp, args.batch, dev)
for _ in range(3):
    train_step(model, opt, sch, norm, batch)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    train_step(model, opt, sch, norm, batch)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"] if hasattr(prof.key_averages()[0], "device_type") else prof.key_averages()
rows = sorted(prof.key_averages(), key=lambda e: -getattr(e, "self_device_time_total", 0))
tot = sum(getattr(e, "self_device_time_total", 0) for e in rows)
print(f"total self device time {tot/1e3:.2f} ms")
for e in rows[: args.rows]:
    t = getattr(e, "self_device_time_total", 0)
    if t <= 0:
        break
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}% x{e.count:<5d} {e.key[:110]}")

print("\n--- aten::copy_ / add_ / contiguous by input shape ---")
rows2 = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ("aten::copy_", "aten::add_", "aten::contiguous", "aten::to", "aten::_to_copy", "aten::fill_", "aten::zero_", "aten::mul", "aten::clone")]
rows2.sort(key=lambda e: -getattr(e, "self_device_time_total", 0))
for e in rows2[:18]:
    print(f"{getattr(e, 'self_device_time_total', 0)/1e3:9.3f} ms x{e.count:<4d} {e.key:18s} {str(e.input_shapes)[:110]}")
