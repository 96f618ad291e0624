"""AdamW over flat fp32 buffers: one libsd_b200 launch per parameter group and step.

Drop-in for ``torch.optim.AdamW(model.parameters(), lr=...)`` (reference: ml/training/train.py:162,
239; distill.py:139,203) including what ``OneCycleLR`` needs (train.py:172: it rewrites
``group["lr"]`` and cycles ``group["betas"][0]`` every step).  Parameters of a group are re-homed
into ONE contiguous buffer (``p.data`` become views), so are their gradients, ``exp_avg`` and
``exp_avg_sq``: the update is a single 28 B/param streaming kernel (sd_adamw_step) and the
data-parallel gradient exchange is a single NCCL all-reduce over the flat gradient.
"""
from __future__ import annotations

import torch

from soccerdiffusion_b200 import _lib, ops


def _round_up(n, m):
    return (n + m - 1) // m * m


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            _lib.require_cuda(*ps)
            for p in ps:
                if p.dtype != torch.float32:
                    raise _lib.SdError("FusedAdamW holds fp32 master parameters only")
            offs, total = [], 0
            for p in ps:
                offs.append(total)
                total += _round_up(p.numel(), 4)  # keep every view 16-byte aligned
            dev = ps[0].device
            flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
            flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
            flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
            flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
            # ONE step tensor per group, shared by the state of all its parameters and updated in place: state_dict()
            # (the reference checkpoints optimizer_state_dict, train.py:244-252) always carries the true step count,
            # whichever of step() / step_captured() ran, without 296 tensor constructions per iteration
            step_t = torch.tensor(0.0)
            for p, o in zip(ps, offs):
                n = p.numel()
                flat_p[o:o + n].copy_(p.data.reshape(-1))
                p.data = flat_p[o:o + n].view(p.shape)
                p.grad = flat_g[o:o + n].view(p.shape)
                self.state[p] = dict(step=step_t, exp_avg=flat_m[o:o + n].view(p.shape),
                                     exp_avg_sq=flat_v[o:o + n].view(p.shape))
                # torch.optim.AdamW skips parameters whose .grad is None (unused this step: e.g. every encoder under
                # --decoder-pretraining, train.py:221-224).  The flat gradient views always exist, so "received a gradient
                # this step" is tracked instead: by this hook for gradients accumulated by autograd, by
                # functional._zero_grads (runtime.mark_grad_written) for kernels that write p.grad in place.
                p.register_post_accumulate_grad_hook(self._mark)
            self._flat.append(dict(p=flat_p, g=flat_g, m=flat_m, v=flat_v, params=ps, offs=offs, step=0, step_t=step_t,
                                   index={id(p): i for i, p in enumerate(ps)}))
        self._touched: set[int] = set()
        self._frozen_runs = None
        from soccerdiffusion_b200 import runtime

        runtime.register_grad_listener(self._mark)

    def _mark(self, p):
        self._touched.add(id(p))

    def _runs(self, f):
        """Contiguous [begin, end) element ranges of the flat buffers covering exactly the parameters that received a
        gradient since the last zero_grad() (all of them in ordinary training: one range = one launch)."""
        runs, cur = [], None
        for p, o in zip(f["params"], f["offs"]):
            if id(p) in self._touched:
                end = o + _round_up(p.numel(), 4)
                if cur is not None and cur[1] == o:
                    cur[1] = end
                else:
                    cur = [o, end]
                    runs.append(cur)
        return [tuple(r) for r in runs]

    # gradients stay allocated so autograd accumulates into the flat buffer (set_to_none would drop the views)
    def zero_grad(self, set_to_none: bool = False):
        self._touched.clear()
        for f in self._flat:
            if f is None:
                continue
            f["g"].zero_()
            for p, o in zip(f["params"], f["offs"]):
                if p.grad is None or p.grad.data_ptr() != f["g"].data_ptr() + 4 * o:
                    p.grad = f["g"][o:o + p.numel()].view(p.shape)

    def flat_gradients(self):
        """The flat gradient buffers (one per parameter group) — what data parallelism all-reduces."""
        return [f["g"] for f in self._flat if f is not None]

    def flat_parameters(self):
        return [f["p"] for f in self._flat if f is not None]

    def _gather_stray_grads(self, f):
        # a caller may have replaced p.grad (e.g. zero_grad(set_to_none=True) by a framework): fold it back
        for p, o in zip(f["params"], f["offs"]):
            want = f["g"].data_ptr() + 4 * o
            if p.grad is None:   # no gradient this step: skipped like torch.optim.AdamW does (the view is restored)
                self._touched.discard(id(p))
                f["g"][o:o + p.numel()].zero_()
                p.grad = f["g"][o:o + p.numel()].view(p.shape)
            elif p.grad.data_ptr() != want:
                f["g"][o:o + p.numel()].copy_(p.grad.reshape(-1))
                p.grad = f["g"][o:o + p.numel()].view(p.shape)
                self._touched.add(id(p))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            self._gather_stray_grads(f)
            f["step"] += 1
            f["step_t"].fill_(float(f["step"]))
            b1, b2 = group["betas"]
            for a, b in self._runs(f):
                ops.adamw_step(f["p"][a:b], f["g"][a:b], f["m"][a:b], f["v"][a:b], float(group["lr"]), float(b1), float(b2),
                               float(group["eps"]), float(group["weight_decay"]), f["step"], grad_scale)
        from soccerdiffusion_b200 import runtime

        runtime.bump_weights_generation()
        return loss

    # ---- CUDA-graph friendly stepping: hyper-parameters travel through device memory ------------------------
    def prepare_captured_step(self):
        """Host side of one step (call BEFORE replaying a graph that contains ``step_captured``): advances the step
        counters and uploads {lr, betas, eps, wd, bias corrections} of every group (async, current stream)."""
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            if "hp_host" not in f:
                # PAGEABLE on purpose: cudaMemcpyAsync from pageable memory stages the bytes before it returns, so the
                # buffer can be rewritten for the next step while earlier graph replays are still queued (a pinned
                # buffer would be read by the DMA engine later — the host runs many steps ahead of the device)
                f["hp_host"] = torch.empty(8, dtype=torch.float32)
                f["hp_dev"] = torch.empty(8, dtype=torch.float32, device=f["p"].device)
            f["step"] += 1
            f["step_t"].fill_(float(f["step"]))
            b1, b2 = group["betas"]
            t = f["step"]
            bc1 = 1.0 - float(b1) ** t
            bc2 = 1.0 - float(b2) ** t
            hp = f["hp_host"]
            hp[0], hp[1], hp[2], hp[3], hp[4] = float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"])
            hp[5], hp[6], hp[7] = float(group["lr"]) / bc1, 1.0 / (bc2 ** 0.5), 1.0
            f["hp_dev"].copy_(hp, non_blocking=True)

    def rollback_captured_step(self):
        for f in self._flat:
            if f is not None:
                f["step"] -= 1
                f["step_t"].fill_(float(f["step"]))

    @torch.no_grad()
    def step_captured(self):
        """Device side: one sd_adamw_step_dev launch per group (this is what a CUDA graph captures)."""
        for f in self._flat:
            if f is None:
                continue
            if "hp_dev" not in f:
                raise RuntimeError("call prepare_captured_step() once before capturing step_captured()")
            self._gather_stray_grads(f)
            for a, b in self._runs(f):
                ops.adamw_step_dev(f["p"][a:b], f["g"][a:b], f["m"][a:b], f["v"][a:b], f["hp_dev"])
        from soccerdiffusion_b200 import runtime

        runtime.bump_weights_generation()

    # torch's loader would replace the state tensors by copies; keep the flat views instead
    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        for g_saved, group, f in zip(groups, self.param_groups, self._flat):
            if len(g_saved["params"]) != len(group["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            for k, v in g_saved.items():
                if k != "params":
                    group[k] = v
            step = 0
            for idx, p in zip(g_saved["params"], group["params"]):
                st = state_dict["state"].get(idx)
                if st is None or p not in self.state:
                    continue
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                step = max(step, int(st["step"]))
            if f is not None:
                f["step"] = step
                f["step_t"].fill_(float(step))
        from soccerdiffusion_b200 import runtime

        runtime.bump_weights_generation()
