"""Inference entry points: the control-tick body of ml/inference/ros.py:259-318 without ROS.

    sample_loop         the reference's step-at-a-time loop (ros.py:301-310), kept working unchanged
    TrajectorySampler   encode once -> ONE persistent-kernel launch for all DDIM steps -> denormalise
    FrameEmbeddingCache per-frame image embeddings computed once per camera frame (the TODO at ros.py:180-183)
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


class FrameEmbeddingCache:
    """Per-frame image embeddings across control ticks (the TODO at ml/inference/ros.py:180-183 and
    ml/model/model.py:134: "calculate the embedding first ... for now we calculate them every timestep for the whole
    sequence").  ``push(frame)`` runs trunk + token head on the NEW frame(s) only and keeps the last
    ``image_context_length`` tokens; a tick then passes ``batch["image_tokens"] = cache.tokens()`` instead of
    ``batch["image_data"]`` and only the frame-sequence encoder runs.  Eval mode only (BatchNorm with running
    statistics is per-frame independent, so the tokens equal those of the whole-sequence call)."""

    def __init__(self, model, context_length: int | None = None, use_cuda_graph: bool = False):
        if model.image_sequence_encoder is None:
            raise ValueError("the model has no image encoder")
        self.encoder = model.image_sequence_encoder.image_encoder
        self.use_cuda_graph = use_cuda_graph      # single-frame pushes: trunk + token head replayed from one captured graph
        self._graphs: dict = {}
        self.capacity = int(context_length or model.image_sequence_encoder.transformer_encoder.positional_encoding.pe.shape[1])
        self._tokens: list[torch.Tensor] = []

    def __len__(self) -> int:
        return len(self._tokens)

    @torch.no_grad()
    def push(self, frames: torch.Tensor) -> None:
        """``frames``: (3,R,R) or (n,3,R,R); normalised float32 like the reference's preprocessing output, or raw uint8."""
        if self.encoder.training:
            raise RuntimeError("FrameEmbeddingCache needs model.eval(): train-mode BatchNorm couples the frames of a batch")
        x = frames if frames.dim() == 4 else frames.unsqueeze(0)
        if self.use_cuda_graph and x.shape[0] == 1:
            tok = self._push_graphed(x)
        else:
            tok = self.encoder(x.unsqueeze(0))[0]                  # (n, d)
        self._tokens.extend(tok[i] for i in range(tok.shape[0]))
        self._tokens = self._tokens[-self.capacity:]

    def _push_graphed(self, x: torch.Tensor) -> torch.Tensor:
        sig = (tuple(x.shape), x.dtype)
        entry = self._graphs.get(sig)
        if entry is None:
            static_in = x.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                                  # allocations, cuDNN autotuning
                    self.encoder(static_in.unsqueeze(0))
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.encoder(static_in.unsqueeze(0))[0]
            entry = (graph, static_in, static_out)
            self._graphs[sig] = entry
        graph, static_in, static_out = entry
        static_in.copy_(x, non_blocking=True)
        graph.replay()
        return static_out.clone()

    def tokens(self) -> torch.Tensor:
        """(1, frames, d) — the cached embeddings, oldest first (ros.py:269 stacks the frame buffer the same way)."""
        if not self._tokens:
            raise RuntimeError("no frame has been pushed yet")
        return torch.stack(self._tokens, dim=0).unsqueeze(0)


class TrajectorySampler:
    """One control tick (ros.py:259-318 without ROS).  With ``use_cuda_graph=True`` the whole tick — context
    encoders (trunk + sequence encoders), K/V cache projection and the persistent DDIM sampler — is captured ONCE
    into a CUDA graph per input signature and replayed: per tick the host issues a few input copies and one graph
    launch instead of ~200 kernel launches."""

    def __init__(self, model, scheduler, num_inference_steps: int = 30, distilled: bool = False,
                 use_cuda_graph: bool = False):
        self.model = model.eval()
        self.scheduler = scheduler
        self.num_inference_steps = num_inference_steps
        self.distilled = distilled
        self.use_cuda_graph = use_cuda_graph
        self._graphs: dict = {}
        scheduler.set_timesteps(num_inference_steps)

    @torch.no_grad()
    def _tick(self, batch: dict, x_T: torch.Tensor, denormalize: bool) -> torch.Tensor:
        ctx = self.model.encode_input_data(batch)
        return self.sample_with_context(ctx, x_T, denormalize)

    @torch.no_grad()
    def __call__(self, batch: dict, x_T: torch.Tensor | None = None, denormalize: bool = True) -> torch.Tensor:
        m = self.model
        any_t = next(iter(batch.values()))
        B = any_t.shape[0]
        if x_T is None:
            x_T = torch.randn(B, m.diffusion_action_generator.max_seq_len, m.num_joints, device=any_t.device)
        if not self.use_cuda_graph:
            return self._tick(batch, x_T, denormalize)
        keys = [k for k in ("joint_command_history", "rotation", "joint_state", "image_data", "image_tokens", "game_state")
                if k in batch]
        sig = (tuple((k, tuple(batch[k].shape), batch[k].dtype) for k in keys), tuple(x_T.shape), denormalize)
        entry = self._graphs.get(sig)
        if entry is None:
            static_in = {k: batch[k].clone() for k in keys}
            static_x = x_T.clone()
            # warm-up on a side stream (allocations, cuDNN autotuning, plan caches), then capture
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._tick(static_in, static_x, denormalize)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._tick(static_in, static_x, denormalize)
            entry = (graph, static_in, static_x, static_out)
            self._graphs[sig] = entry
        graph, static_in, static_x, static_out = entry
        for k in keys:
            static_in[k].copy_(batch[k], non_blocking=True)
        static_x.copy_(x_T, non_blocking=True)
        graph.replay()
        return static_out.clone()

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
