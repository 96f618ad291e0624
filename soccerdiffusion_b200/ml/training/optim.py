"""AdamW over flat fp32 buffers: one libsd_b200 launch per parameter group and step.

Drop-in for ``torch.optim.AdamW(model.parameters(), lr=...)`` (reference: ml/training/train.py:162,
239; distill.py:139,203) including what ``OneCycleLR`` needs (train.py:172: it rewrites
``group["lr"]`` and cycles ``group["betas"][0]`` every step).  Parameters of a group are re-homed
into ONE contiguous buffer (``p.data`` become views), so are their gradients, ``exp_avg`` and
``exp_avg_sq``: the update is a single 28 B/param streaming kernel (sd_adamw_step) and the
data-parallel gradient exchange is a single NCCL all-reduce over the flat gradient.

torch semantics kept exactly:
  * a parameter that received NO gradient in a step (``.grad is None`` for torch: e.g. every encoder under
    ``--decoder-pretraining``, train.py:221-224) is skipped entirely — no weight decay, no moment decay, no step count;
  * the step count (bias correction) is PER PARAMETER.
The flat gradient views always exist, so "received a gradient" is tracked (autograd hook / functional._zero_grads), and
the update is launched over the contiguous runs of parameters that did and share a step count — one run, one launch, in
ordinary training.
"""
from __future__ import annotations

import torch

from soccerdiffusion_b200 import _lib, ops, runtime


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
        self._touched: set[int] = set()
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
            for p, o in zip(ps, offs):
                n = p.numel()
                flat_p[o:o + n].copy_(p.data.reshape(-1))
                p.data = flat_p[o:o + n].view(p.shape)
                p.grad = flat_g[o:o + n].view(p.shape)
                self.state[p] = dict(step=torch.tensor(0.0), exp_avg=flat_m[o:o + n].view(p.shape),
                                     exp_avg_sq=flat_v[o:o + n].view(p.shape))
                p.register_post_accumulate_grad_hook(self._mark)   # gradients accumulated by autograd
            self._flat.append(dict(p=flat_p, g=flat_g, m=flat_m, v=flat_v, params=ps, offs=offs,
                                   pstep=[0] * len(ps),                                           # host mirror, per parameter
                                   pstep_dev=torch.zeros(len(ps), device=dev, dtype=torch.int32),   # what captured steps read
                                   idx_cache={}, last_touched=None))
        runtime.register_grad_listener(self._mark)   # gradients written in place by the fused backward kernels

    # ---- bookkeeping ------------------------------------------------------------------------------------------------------
    def _mark(self, p):
        self._touched.add(id(p))

    def _touched_indices(self, f):
        return tuple(i for i, p in enumerate(f["params"]) if id(p) in self._touched)

    def _runs(self, f, touched):
        """Contiguous [begin, end) ranges of the flat buffers covering exactly the touched parameters, split where the
        per-parameter step count changes -> [(begin, end, index of the run's first parameter)]."""
        runs, cur = [], None
        tset = set(touched)
        for i, (p, o) in enumerate(zip(f["params"], f["offs"])):
            if i not in tset:
                cur = None
                continue
            end = o + _round_up(p.numel(), 4)
            if cur is not None and cur[1] == o and f["pstep"][cur[2]] == f["pstep"][i]:
                cur[1] = end
            else:
                cur = [o, end, i]
                runs.append(cur)
        return [tuple(r) for r in runs]

    def _idx(self, f, touched):
        t = f["idx_cache"].get(touched)
        if t is None:
            t = torch.tensor(touched, device=f["p"].device, dtype=torch.int64)
            f["idx_cache"][touched] = t
        return t

    @property
    def _steps(self):
        """Largest per-parameter step count of each group (what a single global counter would show)."""
        return [max(f["pstep"]) if f is not None else None for f in self._flat]

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

    # ---- eager step ---------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        runtime.join_wgrad_stream()   # weight gradients the trunk launched on its side stream (no-op when none are pending)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            self._gather_stray_grads(f)
            touched = self._touched_indices(f)
            if not touched:
                continue
            b1, b2 = group["betas"]
            for a, b, first in self._runs(f, touched):
                ops.adamw_step(f["p"][a:b], f["g"][a:b], f["m"][a:b], f["v"][a:b], float(group["lr"]), float(b1), float(b2),
                               float(group["eps"]), float(group["weight_decay"]), f["pstep"][first] + 1, grad_scale)
            for i in touched:
                f["pstep"][i] += 1
            f["pstep_dev"].index_add_(0, self._idx(f, touched), torch.ones((), device=f["p"].device, dtype=torch.int32).expand(len(touched)))
        runtime.bump_weights_generation()
        return loss

    # ---- CUDA-graph friendly stepping: hyper-parameters AND step counts travel through device memory -------------------------
    def prepare_captured_step(self, grad_scale: float = 1.0):
        """Host side of one step (call BEFORE replaying a graph that contains ``step_captured``): uploads {lr, betas, eps,
        wd, grad_scale} of every group (async, current stream).  The step counts are NOT host state of the replay: the
        captured graph advances a device-resident counter per parameter and the kernel derives the bias corrections."""
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            if "hp_host" not in f:
                # PAGEABLE on purpose: cudaMemcpyAsync from pageable memory stages the bytes before it returns, so the
                # buffer can be rewritten for the next step while earlier graph replays are still queued (a pinned
                # buffer would be read by the DMA engine later — the host runs many steps ahead of the device)
                f["hp_host"] = torch.zeros(8, dtype=torch.float32)
                f["hp_dev"] = torch.zeros(8, dtype=torch.float32, device=f["p"].device)
            b1, b2 = group["betas"]
            hp = f["hp_host"]
            hp[0], hp[1], hp[2], hp[3], hp[4] = float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"])
            hp[7] = float(grad_scale)
            f["hp_dev"].copy_(hp, non_blocking=True)

    @torch.no_grad()
    def step_captured(self):
        """Device side (this is what a CUDA graph captures): advance the touched parameters' device step counters, then one
        sd_adamw_step_dev launch per run."""
        runtime.join_wgrad_stream()
        for f in self._flat:
            if f is None:
                continue
            if "hp_dev" not in f:
                raise RuntimeError("call prepare_captured_step() once before capturing step_captured()")
            self._gather_stray_grads(f)
            touched = self._touched_indices(f)
            f["last_touched"] = touched
            if not touched:
                continue
            f["pstep_dev"].index_add_(0, self._idx(f, touched), torch.ones((), device=f["p"].device, dtype=torch.int32).expand(len(touched)))
            for a, b, first in self._runs(f, touched):
                ops.adamw_step_dev(f["p"][a:b], f["g"][a:b], f["m"][a:b], f["v"][a:b], f["hp_dev"], f["pstep_dev"][first:first + 1])
        runtime.bump_weights_generation()

    def commit_captured_step(self):
        """Host mirror of one executed ``step_captured`` (eager call or graph replay): the per-parameter step counts."""
        for f in self._flat:
            if f is not None and f["last_touched"]:
                for i in f["last_touched"]:
                    f["pstep"][i] += 1
        runtime.bump_weights_generation()

    def rollback_captured_step(self):
        """Kept for callers of the previous protocol: capture executes nothing and advances no host state any more."""

    # ---- (de)serialisation: torch's format, per-parameter ``step`` -------------------------------------------------------------
    def state_dict(self):
        for f in self._flat:
            if f is None:
                continue
            for p, t in zip(f["params"], f["pstep"]):
                self.state[p]["step"] = torch.tensor(float(t))
        return super().state_dict()

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
            index = {id(p): i for i, p in enumerate(f["params"])} if f is not None else {}
            for idx, p in zip(g_saved["params"], group["params"]):
                st = state_dict["state"].get(idx)
                if st is None or p not in self.state:
                    continue
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                self.state[p]["step"] = torch.tensor(float(int(st["step"])))
                if id(p) in index:
                    f["pstep"][index[id(p)]] = int(st["step"])
            if f is not None:
                f["pstep_dev"].copy_(torch.tensor(f["pstep"], dtype=torch.int32))
        runtime.bump_weights_generation()
