"""Pins oracle/model_ref.py (the CPU restatement) against
  (1) golden outputs of the REAL reference modules (tests/golden, made by oracle/gen_golden.py), always;
  (2) the live reference imported from /root/reference, when that tree exists (build container only).
Weights/inputs are re-synthesised from oracle/synth.py (bit-identical everywhere)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import model_ref, synth
from oracle.load_reference import build_reference_model, reference_available

CASES = {"tiny": synth.TINY_HP, "patch": synth.PATCH_HP, "default": synth.DEFAULT_HP}


def _template(manifest, case):
    c = manifest["cases"][case]
    return {n: torch.empty(s) for n, s in zip(c["state_dict_names"], c["state_dict_shapes"])}, c


@pytest.mark.parametrize("case", ["tiny", "patch", "default"])
def test_restatement_matches_golden_inference(manifest, case):
    tmpl, c = _template(manifest, case)
    hp, B, seed = CASES[case], c["batch_size"], c["seed"]
    sd = synth.synth_state_dict(tmpl, seed)
    g = load_golden(case)
    batch = synth.synth_batch(hp, B, seed)
    x_T = synth.synth_noise("x_T", hp, B, seed)
    t = synth.synth_timesteps(B, seed)
    torch.set_num_threads(8)
    with torch.no_grad():
        ctx = model_ref.encode_input_data(batch, sd, hp)
        for i, cx in enumerate(ctx):
            assert rel_l2(cx.numpy(), g[f"ctx{i}"]) < 2e-6, f"ctx{i}"
        eps = model_ref.forward_with_context(ctx, x_T, t, sd, hp)
        assert rel_l2(eps.numpy(), g["eps_eval"]) < 2e-6
        eps0 = model_ref.forward_with_context(ctx, x_T, torch.zeros(B), sd, hp)
        assert rel_l2(eps0.numpy(), g["eps_eval_float_t0"]) < 2e-6
        if case != "default":  # 30 steps of the default model on CPU are covered once, in the sampler test below
            x0, trace = model_ref.sample_ddim(ctx, x_T, sd, hp, c["ddim_steps"])
            assert rel_l2(torch.stack(trace).numpy(), g["ddim_eps_trace"]) < 1e-5
            assert rel_l2(x0.numpy(), g["ddim_x0"]) < 1e-5


@pytest.mark.parametrize("case", ["tiny", "patch"])
def test_restatement_matches_golden_training(manifest, case):
    tmpl, c = _template(manifest, case)
    hp, B, seed = CASES[case], c["batch_size"], c["seed"]
    sd = synth.synth_state_dict(tmpl, seed)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k and k not in ("mean", "std") else v)
          for k, v in sd.items()}
    g = load_golden(case)
    batch = synth.synth_batch(hp, B, seed)
    noise = synth.synth_noise("eps", hp, B, seed)
    t = synth.synth_timesteps(B, seed)
    loss, pred = model_ref.training_loss(batch, noise, t, sd, hp, masks=None, train_bn=True)
    assert rel_l2(pred.detach().numpy(), g["train_pred"]) < 5e-6
    assert abs(loss.item() - float(g["train_loss"])) < 1e-5 * max(1.0, abs(float(g["train_loss"])))
    loss.backward()
    checked = 0
    for key in g.files:
        if not key.startswith("grad/"):
            continue
        name = key[5:]
        got = sd[name].grad
        assert got is not None, name
        ref = g[key]
        denom = max(np.linalg.norm(ref), 1e-8)
        assert np.linalg.norm(got.numpy() - ref) / denom < 5e-4, name
        checked += 1
    assert checked > 20


def test_default_sampler_matches_golden(manifest):
    tmpl, c = _template(manifest, "default")
    hp, B, seed = synth.DEFAULT_HP, c["batch_size"], c["seed"]
    sd = synth.synth_state_dict(tmpl, seed)
    g = load_golden("default")
    x_T = synth.synth_noise("x_T", hp, B, seed)
    ctx = [torch.from_numpy(g[f"ctx{i}"]) for i in range(5)]
    torch.set_num_threads(8)
    with torch.no_grad():
        x0, trace = model_ref.sample_ddim(ctx, x_T, sd, hp, c["ddim_steps"])
    assert rel_l2(torch.stack(trace).numpy(), g["ddim_eps_trace"]) < 1e-5
    assert rel_l2(x0.numpy(), g["ddim_x0"]) < 1e-5


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("case", ["tiny", "patch"])
def test_restatement_matches_live_reference(case):
    hp = CASES[case]
    B, seed = 2, 7
    model = build_reference_model(hp)
    sd = synth.synth_state_dict(model.state_dict(), seed)
    model.load_state_dict(sd)
    model.eval()
    batch = synth.synth_batch(hp, B, seed)
    x = synth.synth_noise("x", hp, B, seed)
    t = synth.synth_timesteps(B, seed)
    with torch.no_grad():
        ref_ctx = model.encode_input_data(batch)
        ref = model.forward_with_context(ref_ctx, x, t)
        ctx = model_ref.encode_input_data(batch, sd, hp)
        got = model_ref.forward_with_context(ctx, x, t, sd, hp)
    for a, b in zip(ctx, ref_ctx):
        assert rel_l2(a.numpy(), b.numpy()) < 2e-6
    assert rel_l2(got.numpy(), ref.numpy()) < 2e-6
    # foreign context of another length (train.py:221-224)
    with torch.no_grad():
        fc = [torch.from_numpy(synth.normal("fc", (B, 10, hp["hidden_dim"]), seed))]
        assert rel_l2(model_ref.forward_with_context(fc, x, t, sd, hp).numpy(),
                      model.forward_with_context(fc, x, t).numpy()) < 2e-6
