"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the REAL reference modules.

Build-container only (needs /root/reference).  Weights and inputs are NOT stored: they are
re-synthesised bit-identically from oracle/synth.py at test time; only the reference's outputs are
committed.  Each case records which torch code path produced it (encoder fast path in
eval+no_grad vs slow path with grad enabled).

    python -m oracle.gen_golden            # writes tests/golden/{tiny,patch,default,decoder_only}.npz + manifest.json
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from oracle.ddim import DDIMOracle  # noqa: E402
from oracle.load_reference import build_reference_model  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def zero_dropout(model):
    """p=0 everywhere (nn.Dropout modules and nn.MultiheadAttention.dropout) so train-mode parity is defined."""
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0


def ref_sample(model, ctx, x_T, steps):
    """ros.py:301-310 with the restated scheduler (diffusers is not installable offline)."""
    sch = DDIMOracle(1000)
    sch.set_timesteps(steps)
    x = x_T.clone()
    eps_all = []
    for t in sch.timesteps:
        with torch.no_grad():
            eps = model.forward_with_context(ctx, x, torch.full((x.shape[0],), int(t), dtype=torch.int64))
        eps_all.append(eps.numpy().copy())
        x = torch.from_numpy(sch.step(eps.numpy(), int(t), x.numpy()).prev_sample)
    return x.numpy(), np.stack(eps_all)


def run_case(name, hp, batch_size, ddim_steps, seed, with_grads=True, full_grads=None):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model = build_reference_model(hp)
    sd = synth.synth_state_dict(model.state_dict(), seed)
    model.load_state_dict(sd)
    batch = synth.synth_batch(hp, batch_size, seed)
    x_T = synth.synth_noise("x_T", hp, batch_size, seed)
    noise = synth.synth_noise("eps", hp, batch_size, seed)
    t = synth.synth_timesteps(batch_size, seed)
    out = {}

    # --- inference semantics: eval + no_grad (encoder fast path) ---------------------------------
    model.eval()
    with torch.no_grad():
        ctx = model.encode_input_data(batch)
        for i, c in enumerate(ctx):
            out[f"ctx{i}"] = c.numpy().copy()
        out["eps_eval"] = model.forward_with_context(ctx, x_T, t).numpy().copy()
        out["eps_eval_float_t0"] = model.forward_with_context(ctx, x_T, torch.zeros(batch_size)).numpy().copy()
    x0, eps_trace = ref_sample(model, ctx, x_T, ddim_steps)
    out["ddim_x0"] = x0
    out["ddim_eps_trace"] = eps_trace

    # --- training semantics: train mode (BN batch statistics, slow path), dropout forced to p=0 --
    if with_grads:
        model.train()
        zero_dropout(model)
        model.load_state_dict(sd)  # fresh BN running statistics
        sch = DDIMOracle(1000)
        x0n = (batch["joint_command"] - sd["mean"]) / sd["std"]
        x_t = torch.from_numpy(sch.add_noise(x0n.numpy(), noise.numpy(), t.numpy()))
        model.zero_grad()
        pred = model(batch, x_t, t)
        loss = torch.nn.functional.mse_loss(pred, noise)
        loss.backward()
        out["train_x_t"] = x_t.numpy().copy()
        out["train_pred"] = pred.detach().numpy().copy()
        out["train_loss"] = np.asarray(loss.item(), dtype=np.float64)
        names, norms = [], []
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            names.append(n)
            norms.append(float(p.grad.double().norm()))
            # full gradients only for the in-scope stack (small); trunk gradients by norm
            if ".image_encoder.encoder." not in n or n.endswith(("fc.weight", "fc.bias", "avgpool.weight", "avgpool.bias")):
                if p.grad.numel() <= (70000 if hp["hidden_dim"] <= 128 else 20000) and (full_grads is None or any(k in n for k in full_grads)):
                    out["grad/" + n] = p.grad.numpy().copy()
        out["grad_names"] = np.asarray(names)
        out["grad_norms"] = np.asarray(norms, dtype=np.float64)

    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **out)
    return dict(hp=hp, batch_size=batch_size, ddim_steps=ddim_steps, seed=seed,
                state_dict_names=list(model.state_dict().keys()),
                state_dict_shapes=[list(v.shape) for v in model.state_dict().values()],
                n_params=int(sum(p.numel() for p in model.parameters())),
                paths="eps_eval/ctx*/ddim_*: eval()+no_grad (encoder fast path); train_*/grad/*: train(), dropout p=0 (slow path)")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {"torch": torch.__version__, "generator": "oracle/gen_golden.py", "cases": {}}
    manifest["cases"]["tiny"] = run_case("tiny", synth.TINY_HP, 1, 10, 0)
    manifest["cases"]["patch"] = run_case("patch", synth.PATCH_HP, 3, 10, 1)
    manifest["cases"]["default"] = run_case(
        "default", synth.DEFAULT_HP, 2, 30, 2,
        full_grads=("layers.0.", "layers.3.", "embedding", "fc_out", "step_encoding", "avgpool.bias", "fc.bias"))
    # the reference's other shipped configurations (frame resolution reduced, see oracle/synth.py) and the scaled-up one
    small = ("layers.0.", "embedding", "fc_out", "step_encoding", "fc.bias", "norm")
    manifest["cases"]["larger"] = run_case("larger", synth.LARGER_HP, 1, 10, 5, full_grads=small)
    manifest["cases"]["sim_scratch"] = run_case("sim_scratch", synth.SIM_SCRATCH_HP, 2, 10, 6, full_grads=small)
    manifest["cases"]["scaled_img"] = run_case("scaled_img", synth.SCALED_IMG_HP, 2, 10, 7, full_grads=small)
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
