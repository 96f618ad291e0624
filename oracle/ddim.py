"""TEST INFRASTRUCTURE — CPU restatement of ``diffusers.DDIMScheduler`` (oracle; not shipped code).

PARITY UNPINNED against upstream: ``diffusers`` is an un-vendored third-party dependency of the
reference (pyproject.toml:15 ``^0.31.0``; poetry.lock:447-448 pins 0.31.0) and is not installed
here, and the reference holds no tests/golden vectors for it.  This file restates the published
algorithm of ``src/diffusers/schedulers/scheduling_ddim.py`` @ v0.31.0 for exactly the
configuration the reference uses at every call site
(``DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)``:
ml/training/train.py:185, ml/training/distill.py:151, ml/inference/ros.py:151, ml/inference/plot.py:71):

  * ``betas_for_alpha_bar`` (cosine):  beta_i = min(1 - abar((i+1)/T)/abar(i/T), 0.999),
    abar(s) = cos((s+0.008)/1.008 * pi/2)^2 evaluated in Python float64, stored float32.
  * ``alphas_cumprod = cumprod(1 - betas)`` in float32; ``final_alpha_cumprod = 1.0``
    (``set_alpha_to_one=True`` default).
  * ``set_timesteps`` "leading" spacing, ``steps_offset=0``.
  * ``step`` with ``prediction_type="epsilon"``, ``eta=0``, ``use_clipped_model_output=False``,
    ``thresholding=False``.
  * ``add_noise``.

The arithmetic is written with numpy float32 scalars/arrays in the same operation order as
upstream's torch float32 code.  Known-answer values are asserted by
tests/test_oracle_ddim.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F32 = np.float32


def betas_for_alpha_bar(num_diffusion_timesteps: int, max_beta: float = 0.999) -> np.ndarray:
    def alpha_bar(t: float) -> float:
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.asarray(betas, dtype=np.float64).astype(F32)


@dataclass
class StepOutput:
    prev_sample: np.ndarray
    pred_original_sample: np.ndarray


class DDIMOracle:
    def __init__(self, num_train_timesteps: int = 1000):
        self.num_train_timesteps = num_train_timesteps
        self.betas = betas_for_alpha_bar(num_train_timesteps)
        self.alphas = (F32(1.0) - self.betas).astype(F32)
        # torch.cumprod on CPU float32 accumulates in double and rounds each prefix once
        # (checked entry-by-entry against torch.cumprod in tests/test_oracle_ddim.py)
        self.alphas_cumprod = np.cumprod(self.alphas.astype(np.float64)).astype(F32)
        self.final_alpha_cumprod = F32(1.0)
        self.num_inference_steps = None
        self.timesteps = np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64)

    def set_timesteps(self, num_inference_steps: int):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        self.timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)

    def coefficients(self, timestep: int):
        """(sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)) as float32 — what ``step`` uses."""
        prev = int(timestep) - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[int(timestep)]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        beta_t = F32(F32(1.0) - a_t)
        return (
            F32(np.sqrt(beta_t)),
            F32(np.sqrt(a_t)),
            F32(np.sqrt(a_prev)),
            F32(np.sqrt(F32(F32(1.0) - a_prev) - F32(0.0))),
        )

    def step(self, model_output: np.ndarray, timestep: int, sample: np.ndarray) -> StepOutput:
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        sb, sa, sap, sbp = self.coefficients(timestep)
        eps = model_output.astype(F32)
        x = sample.astype(F32)
        pred_x0 = ((x - sb * eps) / sa).astype(F32)
        direction = (sbp * eps).astype(F32)
        prev = (sap * pred_x0 + direction).astype(F32)
        return StepOutput(prev, pred_x0)

    def add_noise(self, original: np.ndarray, noise: np.ndarray, timesteps: np.ndarray) -> np.ndarray:
        a = self.alphas_cumprod[np.asarray(timesteps, dtype=np.int64)]
        sa = np.sqrt(a).astype(F32)
        sb = np.sqrt((F32(1.0) - a).astype(F32)).astype(F32)
        shape = (-1,) + (1,) * (original.ndim - 1)
        return (sa.reshape(shape) * original.astype(F32) + sb.reshape(shape) * noise.astype(F32)).astype(F32)
