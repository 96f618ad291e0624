"""Latency / throughput of model.sample() per sampler kind and batch size (CUDA events around whole calls, p50 of `reps`).
python tools/sampler_micro.py [default|scaled] [B ...]"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import soccerdiffusion_b200 as sd  # noqa: E402
from soccerdiffusion_b200 import config  # noqa: E402
from soccerdiffusion_b200.schedulers import DDIMScheduler  # noqa: E402


def main():
    args = sys.argv[1:]
    name = args[0] if args and args[0] in ("default", "scaled") else "default"
    Bs = [int(a) for a in args if a.isdigit()] or [1, 8, 64, 256, 512]
    hp = dict(config.SCALED if name == "scaled" else config.DEFAULT)
    torch.manual_seed(0)
    model = config.build_model(hp).cuda().eval()
    d = hp["hidden_dim"]
    lens = [hp["action_context_length"], hp["imu_context_length"], hp["joint_state_context_length"], hp["image_context_length"], 1]
    sch = DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)
    sch.set_timesteps(30)
    for B in Bs:
        ctx = [torch.randn(B, n, d, device="cuda") for n in lens]
        x_T = torch.randn(B, hp["trajectory_prediction_length"], hp["num_joints"], device="cuda")
        for kind in ("tc", "cta", "cluster"):
            if kind == "cluster" and B > 8:
                continue
            sd.set_precision("bf16" if kind == "tc" else "fp32")
            if kind == "tc" and not model.tc_sampler_supported(sum(lens), x_T.shape[1]):
                print(f"{name} B={B} tc: unsupported shapes")
                continue
            for _ in range(3):
                model.sample(ctx, x_T, sch, sampler=kind)
            torch.cuda.synchronize()
            ts = []
            for _ in range(15 if B <= 64 else 5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                model.sample(ctx, x_T, sch, sampler=kind)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            p50 = statistics.median(ts)
            print(f"{name} B={B:4d} {kind:8s} ran={model.last_sampler:8s} p50 {p50:8.3f} ms  {B / p50 * 1e3:10.0f} trajectories/s", flush=True)


if __name__ == "__main__":
    main()
