"""Loop bodies of the reference's training entry points as functions.

    train_step    ml/training/train.py:193-240   (normalise -> t, eps -> add_noise -> model -> mse -> backward -> AdamW -> OneCycle)
    distill_step  ml/training/distill.py:160-205 (teacher 30-step DDIM under no_grad -> student one step at t=0 -> mse)

Data parallelism (new functionality: the reference is single-device): every rank
owns a full replica and a slice of the batch; ``allreduce_gradients`` averages the flat gradient
buffer of ``FusedAdamW`` with ONE NCCL all-reduce over NVLink (``gloo`` in the CPU tests).
"""
from __future__ import annotations

import torch

from soccerdiffusion_b200 import ops
from soccerdiffusion_b200.functional import mse_loss


def broadcast_parameters(model: torch.nn.Module, src: int = 0, group=None):
    """Identical replicas: rank ``src``'s parameters and buffers overwrite everyone's."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src, group=group)


def allreduce_gradients(optimizer_or_tensors, group=None, average: bool = True):
    """Mean of the gradients over ranks (``average=False``: the sum — the caller folds 1/world into the optimizer
    kernel's ``grad_scale``).  Accepts a FusedAdamW (flat buffers: one collective per parameter group) or an iterable of
    gradient tensors."""
    import torch.distributed as dist

    from soccerdiffusion_b200 import runtime

    runtime.join_wgrad_stream()
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    bufs = optimizer_or_tensors.flat_gradients() if hasattr(optimizer_or_tensors, "flat_gradients") else list(
        optimizer_or_tensors)
    for g in bufs:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        if average:
            g.mul_(1.0 / world)


class BucketedAllReduce:
    """Gradient SUM over the data-parallel ranks in two buckets, the first one overlapped with the backward pass.

    The flat gradient buffer of ``FusedAdamW`` follows ``model.parameters()``: [step token, the three sequence encoders,
    trunk conv1 .. layer2 | trunk layer3, layer4, image head, frame-sequence encoder, game state, denoiser].  The backward
    pass finishes the second part first (it holds 89 % of the parameters); when its gradient milestone fires
    (``runtime.grad_ready("trunk.layer3_onward")``, raised inside the trunk's backward) that bucket is all-reduced on a
    communication stream while layer2, layer1 and the stem are still being differentiated.  ``finish()`` reduces the rest
    (~5 MB) and joins the streams.  The mean (1/world) is applied by the optimizer kernel's ``grad_scale``.
    Works eagerly and inside CUDA-graph capture (the communication stream forks from / joins the capturing stream)."""

    SPLIT_PREFIX = "image_sequence_encoder.image_encoder.encoder.layer3."

    def __init__(self, model: torch.nn.Module, optimizer, group=None):
        self.opt, self.group = optimizer, group
        self.split = None          # (flat group index, element offset) of the first parameter of the early bucket
        flats = [f for f in optimizer._flat if f is not None]
        if len(flats) == 1:
            f = flats[0]
            off_of = {id(p): o for p, o in zip(f["params"], f["offs"])}
            for name, p in model.named_parameters():
                if name.startswith(self.SPLIT_PREFIX) and id(p) in off_of:
                    self.split = off_of[id(p)]
                    break
        self.comm = None
        self.early_done = False

    def _stream(self):
        if self.comm is None:
            self.comm = torch.cuda.Stream()
        return self.comm

    def begin(self):
        """Call before the FORWARD pass of a step (the forward pass plants the backward milestone)."""
        from soccerdiffusion_b200 import runtime

        self.early_done = False
        if self.split is not None:
            runtime.set_grad_ready_callback(self._on_ready)

    def _on_ready(self, tag: str):
        import torch.distributed as dist

        from soccerdiffusion_b200 import runtime

        if tag != "trunk.layer3_onward" or self.early_done or self.split is None:
            return
        g = self.opt.flat_gradients()[0]
        cur = torch.cuda.current_stream()
        comm = self._stream()
        comm.wait_stream(cur)
        runtime.join_wgrad_stream(comm, clear=False)   # layer3 / layer4 weight gradients run on the trunk's side stream
        with torch.cuda.stream(comm):
            dist.all_reduce(g[self.split:], op=dist.ReduceOp.SUM, group=self.group)
        self.early_done = True

    def finish(self):
        """Call after the backward pass: reduces what the milestone did not cover and joins the communication stream."""
        import torch.distributed as dist

        from soccerdiffusion_b200 import runtime

        runtime.set_grad_ready_callback(None)
        runtime.join_wgrad_stream()
        cur = torch.cuda.current_stream()
        if self.early_done:
            g = self.opt.flat_gradients()[0]
            dist.all_reduce(g[: self.split], op=dist.ReduceOp.SUM, group=self.group)
            cur.wait_stream(self.comm)
        else:
            for g in self.opt.flat_gradients():
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)


def q_sample(scheduler, normalizer_or_model, joint_command, noise, timesteps):
    """normalise (train.py:204) + add_noise (train.py:218) in one kernel. Returns x_t."""
    jc = joint_command.float().contiguous()
    mean = normalizer_or_model.mean.to(device=jc.device, dtype=torch.float32).contiguous()
    std = normalizer_or_model.std.to(device=jc.device, dtype=torch.float32).contiguous()
    xt = torch.empty_like(jc)
    ops.q_sample(jc, mean, std, noise.float().contiguous(), timesteps.to(jc.device, torch.int64).contiguous(),
                 scheduler._acp_on(jc.device), None, xt)
    return xt


def train_step(model, optimizer, scheduler, normalizer, batch, *, lr_scheduler=None, noise=None, timesteps=None,
               decoder_pretraining: bool = False, pretraining_context=None, hidden_dim: int | None = None,
               group=None, data_parallel: bool = False):
    """One iteration of train.py:193-240.  ``batch`` holds device tensors (the H2D copy is the caller's,
    train.py:193).  ``noise`` / ``timesteps`` default to the reference's RNG calls; pass them for parity tests.
    Returns the (detached, device) loss."""
    joint_targets = batch["joint_command"]
    bs = joint_targets.size(0)
    device = joint_targets.device
    optimizer.zero_grad()
    if timesteps is None:
        timesteps = torch.randint(0, scheduler.config["num_train_timesteps"], (bs,)).long().to(device)
    if noise is None:
        noise = torch.randn(joint_targets.shape, device=device, dtype=torch.float32)
    noisy = q_sample(scheduler, normalizer, joint_targets, noise, timesteps)
    if decoder_pretraining:
        if pretraining_context is None:
            pretraining_context = torch.randn((bs, 10, hidden_dim or model.hidden_dim), device=device)
        pred = model.forward_with_context([pretraining_context], noisy, timesteps)
    else:
        pred = model(batch, noisy, timesteps)
    loss = mse_loss(pred, noise)
    loss.backward()
    if data_parallel and hasattr(optimizer, "flat_gradients"):
        import torch.distributed as dist

        allreduce_gradients(optimizer, group, average=False)
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        optimizer.step(grad_scale=1.0 / world)      # the mean over ranks is taken inside the AdamW kernel
    else:
        if data_parallel:
            allreduce_gradients(optimizer, group)
        optimizer.step()
    if lr_scheduler is not None:
        lr_scheduler.step()
    return loss.detach()


def distill_step(teacher, student, optimizer, scheduler, batch, num_teacher_steps: int, *, lr_scheduler=None,
                 noise=None, group=None, data_parallel: bool = False):
    """One iteration of distill.py:160-205."""
    joint_targets = batch["joint_command"]
    device = joint_targets.device
    if noise is None:
        noise = torch.randn(joint_targets.shape, device=device, dtype=torch.float32)
    optimizer.zero_grad()
    with torch.no_grad():
        was_training = teacher.training
        teacher.eval()
        ctx = teacher.encode_input_data(batch)
        scheduler.set_timesteps(num_teacher_steps)
        trajectory = teacher.sample(ctx, noise, scheduler)
        teacher.train(was_training)
    pred = student.forward_with_context(ctx, noise, torch.zeros(joint_targets.size(0), device=device))
    loss = mse_loss(pred, trajectory)
    loss.backward()
    if data_parallel:
        allreduce_gradients(optimizer, group)
    optimizer.step()
    if lr_scheduler is not None:
        lr_scheduler.step()
    return loss.detach()
