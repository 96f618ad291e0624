"""DDIMScheduler — the noise-scheduler step interface of the hot path.

Stands in for ``diffusers.schedulers.scheduling_ddim.DDIMScheduler`` (pinned 0.31.0 by the
reference, poetry.lock:447-448; not vendored) for exactly the surface the reference touches:

    DDIMScheduler(beta_schedule="squaredcos_cap_v2", clip_sample=False)   train.py:185, ros.py:151
    scheduler.config["num_train_timesteps"] (get / set)                   train.py:186,211
    scheduler.add_noise(x0, noise, timesteps)                             train.py:218
    scheduler.set_timesteps(n); scheduler.timesteps                       ros.py:301-302, distill.py:179
    scheduler.step(model_output, t, sample).prev_sample                   ros.py:310, distill.py:189

The tables (betas, alphas_cumprod) are built on the host with the same float64 -> float32 ->
``torch.cumprod`` sequence as upstream; the per-element arithmetic of ``add_noise`` and ``step``
runs in libsd_b200 (sd_q_sample / sd_ddim_step).  CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from soccerdiffusion_b200 import _lib, ops


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: torch.Tensor | None = None


class _Config(dict):
    """Item- and attribute-readable config (upstream: FrozenDict; the reference assigns items)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


def _betas_for_alpha_bar(n: int, max_beta: float = 0.999) -> torch.Tensor:
    def alpha_bar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = [min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)]
    return torch.tensor(betas, dtype=torch.float32)


class DDIMScheduler:
    order = 1

    def __init__(
        self,
        num_train_timesteps: int = 1000,
        beta_start: float = 0.0001,
        beta_end: float = 0.02,
        beta_schedule: str = "linear",
        clip_sample: bool = True,
        set_alpha_to_one: bool = True,
        steps_offset: int = 0,
        prediction_type: str = "epsilon",
        timestep_spacing: str = "leading",
    ):
        if beta_schedule == "squaredcos_cap_v2":
            self.betas = _betas_for_alpha_bar(num_train_timesteps)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if clip_sample:
            raise NotImplementedError("clip_sample=True is not on the reference's path (it always passes False)")
        if prediction_type != "epsilon":
            raise NotImplementedError("only prediction_type='epsilon' is on the reference's path")
        if timestep_spacing != "leading":
            raise NotImplementedError("only timestep_spacing='leading' is on the reference's path")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self.config = _Config(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, clip_sample=clip_sample, set_alpha_to_one=set_alpha_to_one,
            steps_offset=steps_offset, prediction_type=prediction_type, timestep_spacing=timestep_spacing,
        )
        # upstream reads the constructor value, not the (mutable) config item, for its arithmetic
        self._n_train = num_train_timesteps
        self._steps_offset = steps_offset
        self._acp_dev: dict = {}
        self._tables: dict = {}

    # ------------------------------------------------------------------------------------------
    def _acp_on(self, device) -> torch.Tensor:
        t = self._acp_dev.get(device)
        if t is None:
            t = self.alphas_cumprod.to(device)
            self._acp_dev[device] = t
        return t

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self._n_train:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self._n_train} as the unet model trained with this scheduler can only handle"
                f" maximal {self._n_train} timesteps."
            )
        self.num_inference_steps = num_inference_steps
        step_ratio = self._n_train // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self._steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def coefficients(self, timestep: int):
        """float32 (sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)) of one eta=0 step."""
        if self.num_inference_steps is None:
            raise ValueError(
                "Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler"
            )
        t = int(timestep)
        prev = t - self._n_train // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        # eta = 0  ->  variance term vanishes: sqrt(1 - a_prev - 0)
        return (float(beta_t ** 0.5), float(a_t ** 0.5), float(a_prev ** 0.5), float((1 - a_prev - 0.0) ** 0.5))

    def schedule_tables(self):
        """(timesteps list, (n,4) float32 coefficient table) for the persistent sampler (sd_plan_set_schedule)."""
        key = (self.num_inference_steps, self._steps_offset)
        hit = self._tables.get(key)
        if hit is None:
            ts = [int(t) for t in self.timesteps]
            coef = np.asarray([self.coefficients(t) for t in ts], dtype=np.float32)
            hit = self._tables[key] = (ts, coef)
        return hit

    # ------------------------------------------------------------------------------------------
    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(original_samples, noise)
        x0 = original_samples.float().contiguous()
        nz = noise.float().contiguous()
        t = timesteps.to(device=x0.device, dtype=torch.int64).contiguous()
        if t.ndim == 0:
            t = t.expand(x0.shape[0]).contiguous()
        out = torch.empty_like(x0)
        ops.q_sample(x0, None, None, nz, t, self._acp_on(x0.device), None, out)
        return out

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = True):
        if eta != 0.0:
            raise NotImplementedError("only eta=0 (deterministic DDIM) is on the reference's path")
        _lib.require_cuda(model_output, sample)
        coef = self.coefficients(int(timestep))
        x = sample.float().contiguous()
        eps = model_output.float().contiguous()
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x)
        ops.ddim_step(x, eps, prev, x0, coef)
        if not return_dict:
            return (prev, x0)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def __len__(self):
        return self._n_train
