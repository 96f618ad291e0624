"""The training-loop body (ml/training/train.py:193-240) captured ONCE as a CUDA graph and replayed.

Per iteration the host copies the batch into static device buffers, uploads eight optimizer scalars, launches ONE graph
(≈ 1000 kernels: context encoders, trunk, denoiser, loss, backward, gradient all-reduce, AdamW) and steps the
learning-rate schedule.  What changes between replays lives in device memory: the batch, torch's Philox state
(timesteps, noise), the dropout seed counter (``runtime.device_seed_counter``) and the optimizer hyper-parameters
(``FusedAdamW.prepare_captured_step``).
"""
from __future__ import annotations

import torch

from soccerdiffusion_b200 import ops, runtime
from soccerdiffusion_b200.functional import mse_loss
from soccerdiffusion_b200.ml.training.step import BucketedAllReduce, q_sample


class GraphedTrainStep:
    def __init__(self, model, optimizer, scheduler, example_batch: dict, *, lr_scheduler=None, data_parallel: bool = False,
                 group=None, decoder_pretraining: bool = False, warmup_steps: int = 3, noise=None, timesteps=None,
                 direct_grads: bool = True):
        self.model, self.opt, self.sch, self.lrs = model, optimizer, scheduler, lr_scheduler
        self.dp, self.group, self.pretrain = data_parallel, group, decoder_pretraining
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.fixed_noise, self.fixed_t = noise, timesteps   # parity tests pin the RNG inputs
        self.seed_counter = runtime.device_seed_counter(next(model.parameters()).device)
        self.direct_grads = direct_grads
        self.launches_per_replay = 0
        self.reducer = BucketedAllReduce(model, optimizer, group) if data_parallel else None
        # Warm-up must leave no trace (the reference's loop takes exactly epochs * len(dataloader) optimizer and
        # OneCycleLR steps, train.py:172,239-240): parameters, AdamW moments and step counters, BatchNorm running
        # statistics, the LR schedule and the dropout seed counter are snapshotted here and restored after the capture.
        snap = self._snapshot()
        # warm-up on a side stream: allocations, cuDNN autotuning, lazy kernel attributes, NCCL communicators
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup_steps):
                self.opt.prepare_captured_step(self._grad_scale())
                self._body()
                self.opt.commit_captured_step()
                self.opt._opt_called = True    # the optimizer step ran (step_captured): what lr_scheduler's order check tracks
                if self.lrs is not None:
                    self.lrs.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.prepare_captured_step(self._grad_scale())
        n0 = ops.launches()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.launches_per_replay = ops.launches() - n0
        self.opt.rollback_captured_step()      # the capture executed nothing: undo its step-counter increment
        self._restore(snap)
        self.warmup_steps = 0                  # optimizer / scheduler steps left behind by the constructor: none

    def _snapshot(self):
        import copy

        flats = [{k: f[k].clone() for k in ("p", "m", "v", "pstep_dev")} if f is not None else None for f in self.opt._flat]
        steps = [list(f["pstep"]) if f is not None else None for f in self.opt._flat]
        groups = [{k: copy.deepcopy(v) for k, v in g.items() if k != "params"} for g in self.opt.param_groups]
        buffers = [b.clone() for b in self.model.buffers()]
        lrs = copy.deepcopy(self.lrs.state_dict()) if self.lrs is not None else None
        return dict(flats=flats, steps=steps, groups=groups, buffers=buffers, lrs=lrs, seed=self.seed_counter.clone(),
                    opt_called=getattr(self.opt, "_opt_called", False))

    def _restore(self, snap):
        torch.cuda.synchronize()
        with torch.no_grad():
            for f, saved, step in zip(self.opt._flat, snap["flats"], snap["steps"]):
                if f is None:
                    continue
                for k in ("p", "m", "v", "pstep_dev"):
                    f[k].copy_(saved[k])
                f["g"].zero_()
                f["pstep"][:] = step
            for g, saved in zip(self.opt.param_groups, snap["groups"]):
                g.update(saved)
            for b, saved in zip(self.model.buffers(), snap["buffers"]):
                b.copy_(saved)
            self.seed_counter.copy_(snap["seed"])
        if self.lrs is not None:
            self.lrs.load_state_dict(snap["lrs"])
            # OneCycleLR also mirrors its state into the param groups (lr / betas): restored above
        runtime.bump_weights_generation()

    def _body(self):
        b = self.static
        jt = b["joint_command"]
        bs, dev = jt.size(0), jt.device
        self.seed_counter.add_(1)
        self.opt.zero_grad()
        if self.reducer is not None:
            self.reducer.begin()   # BEFORE the forward pass (it plants the backward milestone): the big gradient bucket is
                                   # all-reduced DURING the backward pass
        t = self.fixed_t if self.fixed_t is not None else torch.randint(0, self.sch.config["num_train_timesteps"], (bs,), device=dev)
        noise = self.fixed_noise if self.fixed_noise is not None else torch.randn(jt.shape, device=dev, dtype=torch.float32)
        noisy = q_sample(self.sch, self.model, jt, noise, t)
        if self.pretrain:
            ctx = torch.randn((bs, 10, self.model.hidden_dim), device=dev)
            pred = self.model.forward_with_context([ctx], noisy, t)
        else:
            pred = self.model(b, noisy, t)
        loss = mse_loss(pred, noise)
        prev = runtime.direct_grads()
        runtime.set_direct_grads(self.direct_grads)   # gradients land in FusedAdamW's flat buffer without autograd adds
        try:
            loss.backward()
        finally:
            runtime.set_direct_grads(prev)
            runtime.join_wgrad_stream()                # weight gradients launched on the side stream (encoder/trunk.py: TrunkConv)
            if self.reducer is not None:
                self.reducer.finish()                  # sum over ranks; the 1/world factor is the AdamW kernel's grad_scale
        self.opt.step_captured()
        return loss.detach()

    def _grad_scale(self) -> float:
        if not self.dp:
            return 1.0
        import torch.distributed as dist

        return 1.0 / dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1.0

    def __call__(self, batch: dict) -> torch.Tensor:
        """One training iteration on ``batch`` (device tensors); returns the (static) loss tensor."""
        if any(tuple(batch[k].shape) != tuple(v.shape) or batch[k].dtype != v.dtype for k, v in self.static.items()):
            # e.g. the smaller last batch of an epoch (train.py:193-199): the graph is shape-specialised, this step runs
            # kernel by kernel
            from soccerdiffusion_b200.ml.training.step import train_step

            return train_step(self.model, self.opt, self.sch, self.model, batch, lr_scheduler=self.lrs,
                              decoder_pretraining=self.pretrain, group=self.group, data_parallel=self.dp)
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        self.opt.prepare_captured_step(self._grad_scale())
        self.graph.replay()
        self.opt.commit_captured_step()
        ops._count(self.launches_per_replay)
        self.opt._opt_called = True            # the replayed graph contains the optimizer step
        if self.lrs is not None:
            self.lrs.step()
        runtime.bump_weights_generation()
        return self.loss
