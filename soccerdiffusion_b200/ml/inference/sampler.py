"""Inference entry points: the control-tick body of ml/inference/ros.py:259-318 without ROS.

    sample_loop         the reference's step-at-a-time loop (ros.py:301-310), kept working unchanged
    TrajectorySampler   encode once -> ONE persistent-kernel launch for all DDIM steps -> denormalise
"""
from __future__ import annotations

import torch


@torch.no_grad()
def sample_loop(model, scheduler, context, x_T, num_steps: int):
    """ros.py:301-310 verbatim against this package's model/scheduler (one launch per step + one per update)."""
    trajectory = x_T
    scheduler.set_timesteps(num_steps)
    B = x_T.shape[0]
    for t in scheduler.timesteps:
        noise_pred = model.forward_with_context(context, trajectory, torch.full((B,), int(t), device=x_T.device))
        trajectory = scheduler.step(noise_pred, t, trajectory).prev_sample
    return trajectory


class TrajectorySampler:
    def __init__(self, model, scheduler, num_inference_steps: int = 30, distilled: bool = False):
        self.model = model.eval()
        self.scheduler = scheduler
        self.num_inference_steps = num_inference_steps
        self.distilled = distilled
        scheduler.set_timesteps(num_inference_steps)

    @torch.no_grad()
    def __call__(self, batch: dict, x_T: torch.Tensor | None = None, denormalize: bool = True) -> torch.Tensor:
        m = self.model
        any_t = next(iter(batch.values()))
        B = any_t.shape[0]
        if x_T is None:
            x_T = torch.randn(B, m.diffusion_action_generator.max_seq_len, m.num_joints, device=any_t.device)
        ctx = m.encode_input_data(batch)
        return self.sample_with_context(ctx, x_T, denormalize)

    @torch.no_grad()
    def sample_with_context(self, ctx, x_T, denormalize: bool = True):
        m = self.model
        if self.distilled:  # ros.py:293-298
            x = m.forward_with_context(ctx, x_T, torch.zeros(x_T.shape[0], dtype=torch.int64, device=x_T.device))
            if denormalize:
                from soccerdiffusion_b200.dataset.pytorch import Normalizer

                x = Normalizer(m.mean, m.std).denormalize(x)
            return x
        return m.sample(ctx, x_T, self.scheduler, denormalize=denormalize)
