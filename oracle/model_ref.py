"""TEST INFRASTRUCTURE — CPU restatement of the reference model arithmetic (oracle; not shipped code).

Explicit-formula restatement (plain torch ops on CPU, float32 or float64) of
``soccer_diffusion/ml/model`` for a reference-format ``state_dict``:

  * StepToken            ml/model/misc.py:25-35
  * PositionalEncoding   ml/model/misc.py:38-65
  * BaseEncoder          ml/model/encoder/base.py:41-53  (Conv1d k=stride=patch -> +PE -> pre-LN encoder layers)
  * GameStateEncoder     ml/model/encoder/game_state.py:19-27
  * image token head     ml/model/encoder/image.py:38-52, 69-73 (trunk itself = torchvision)
  * DiffusionActionGenerator  ml/model/decoder.py:38-54
  * End2EndDiffusionTransformer.{encode_input_data, forward_with_context, forward}  ml/model/model.py:123-179
  * layer semantics of torch.nn.TransformerEncoderLayer / TransformerDecoderLayer with
    norm_first=True, activation="gelu" (erf), dim_feedforward=d, LayerNorm eps 1e-5, and
    nn.MultiheadAttention's packed in_proj (torch/nn/modules/transformer.py:944-950,1131-1143;
    torch/nn/functional.py:5798-5856).

PINNING: the reference ships no tests or golden vectors for ml/ (SURVEY.md §4), so this
restatement is pinned against the *live* reference modules imported from /root/reference
(tests/test_oracle_vs_reference.py, build container only) and against golden outputs of those
modules committed under tests/golden/ (made by oracle/gen_golden.py).

Dropout: the reference trains with p=0.1 (torch default).  ``masks`` (name -> 0/1 tensor already
divided by keep-prob) lets a test inject the masks produced by the CUDA path; with ``masks=None``
dropout is off (eval semantics).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------
# small pieces


def positional_encoding_table(d_model: int, max_len: int) -> torch.Tensor:
    """misc.py:51-56 — float32 table (max_len, d_model)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def step_token_freqs(dim: int) -> torch.Tensor:
    """misc.py:31-32 — int64 arange * numpy float64 scalar -> float32 tensor, / (half-1), exp."""
    half = dim // 4
    return torch.exp(torch.arange(half) * -np.log(10000) / (half - 1))


def step_token(steps: torch.Tensor, token: torch.Tensor, dim: int) -> torch.Tensor:
    """misc.py:25-35.  sin/cos are evaluated in float32 exactly like the reference, then cast."""
    freqs = step_token_freqs(dim)
    arg = steps[:, None] * freqs[None, :]  # int64*f32 -> f32 ; f32*f32 -> f32
    arg = arg.to(torch.float32)
    emb = torch.cat((arg.sin().to(token.dtype), arg.cos().to(token.dtype), token.expand(steps.size(0), dim // 2)), dim=-1)
    return emb.unsqueeze(1)


def _drop(x, masks, key):
    if masks is None or key not in masks:
        return x
    return x * masks[key].to(x.dtype)


def mha(q_in, kv_in, w_in, b_in, w_out, b_out, heads: int, masks=None, key=""):
    """nn.MultiheadAttention forward, batch_first, no masks (functional.py:5798-5856, 6623-6690)."""
    B, T, d = q_in.shape
    M = kv_in.shape[1]
    dh = d // heads
    q = F.linear(q_in, w_in[:d], b_in[:d])
    k = F.linear(kv_in, w_in[d : 2 * d], b_in[d : 2 * d])
    v = F.linear(kv_in, w_in[2 * d :], b_in[2 * d :])
    q = q.view(B, T, heads, dh).transpose(1, 2)
    k = k.view(B, M, heads, dh).transpose(1, 2)
    v = v.view(B, M, heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
    p = torch.softmax(s, dim=-1)
    p = _drop(p, masks, key + ".attn")
    o = (p @ v).transpose(1, 2).reshape(B, T, d)
    return F.linear(o, w_out, b_out)


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def _ffn(x, sd, prefix, masks, key):
    h = F.gelu(F.linear(x, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"]))
    h = _drop(h, masks, key + ".ffn_inner")
    return F.linear(h, sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])


def encoder_layer(x, sd, prefix, heads, masks=None):
    """x = x + SA(LN1 x); x = x + FFN(LN2 x)   (transformer.py:944-950)"""
    a = prefix + ".self_attn"
    h = _ln(x, sd, prefix + ".norm1")
    x = x + _drop(
        mha(h, h, sd[a + ".in_proj_weight"], sd[a + ".in_proj_bias"], sd[a + ".out_proj.weight"], sd[a + ".out_proj.bias"], heads, masks, prefix + ".sa"),
        masks,
        prefix + ".sa.out",
    )
    x = x + _drop(_ffn(_ln(x, sd, prefix + ".norm2"), sd, prefix, masks, prefix), masks, prefix + ".ffn.out")
    return x


def decoder_layer(x, mem, sd, prefix, heads, masks=None):
    """x += SA(LN1 x); x += CA(LN2 x, mem, mem); x += FFN(LN3 x)   (transformer.py:1131-1143)"""
    a = prefix + ".self_attn"
    c = prefix + ".multihead_attn"
    h = _ln(x, sd, prefix + ".norm1")
    x = x + _drop(
        mha(h, h, sd[a + ".in_proj_weight"], sd[a + ".in_proj_bias"], sd[a + ".out_proj.weight"], sd[a + ".out_proj.bias"], heads, masks, prefix + ".sa"),
        masks,
        prefix + ".sa.out",
    )
    h = _ln(x, sd, prefix + ".norm2")
    x = x + _drop(
        mha(h, mem, sd[c + ".in_proj_weight"], sd[c + ".in_proj_bias"], sd[c + ".out_proj.weight"], sd[c + ".out_proj.bias"], heads, masks, prefix + ".ca"),
        masks,
        prefix + ".ca.out",
    )
    x = x + _drop(_ffn(_ln(x, sd, prefix + ".norm3"), sd, prefix, masks, prefix), masks, prefix + ".ffn.out")
    return x


def _num_layers(sd, prefix):
    n = 0
    while f"{prefix}.{n}.norm1.weight" in sd:
        n += 1
    return n


def base_encoder(x, sd, prefix, heads, masks=None):
    """base.py:49-53.  x (B,S,in) -> (B,S/p,d)."""
    w = sd[prefix + ".embedding.weight"]  # (d, in, p)
    b = sd[prefix + ".embedding.bias"]
    d, cin, p = w.shape
    B, S, _ = x.shape
    S2 = (S - p) // p + 1
    # non-overlapping patches: token j = sum_{c,k} w[:,c,k] * x[b, j*p+k, c]
    xp = x[:, : S2 * p].reshape(B, S2, p, cin).permute(0, 1, 3, 2).reshape(B, S2, cin * p)
    h = F.linear(xp, w.reshape(d, cin * p), b)
    pe = positional_encoding_table(d, max(S2, 1)).to(h.dtype)
    h = h + pe[None, :S2]
    lp = prefix + ".transformer_encoder.layers"
    for i in range(_num_layers(sd, lp)):
        h = encoder_layer(h, sd, f"{lp}.{i}", heads, masks)
    return h


def denoiser(x, mem, sd, heads=4, masks=None, prefix="diffusion_action_generator"):
    """decoder.py:47-54."""
    h = F.linear(x, sd[prefix + ".embedding.weight"], sd[prefix + ".embedding.bias"])
    d = h.shape[-1]
    pe = positional_encoding_table(d, x.shape[1]).to(h.dtype)
    h = h + pe[None]
    lp = prefix + ".transformer_decoder.layers"
    for i in range(_num_layers(sd, lp)):
        h = decoder_layer(h, mem, sd, f"{lp}.{i}", heads, masks)
    return F.linear(h, sd[prefix + ".fc_out.weight"], sd[prefix + ".fc_out.bias"])


# ------------------------------------------------------------------------------------------------
# image path


def build_trunk(hp: dict):
    """torchvision trunk with the reference's head surgery (image.py:55-73, 86-100)."""
    import torchvision.models as tvm
    from torch import nn

    kind = hp["image_encoder_type"]
    d = hp["hidden_dim"]
    if kind in ("resnet18", "resnet50"):
        net = getattr(tvm, kind)(weights=None)
        if hp.get("image_use_final_avgpool", True):
            net.fc = nn.Linear(net.fc.in_features, d)
        else:
            r = hp.get("image_resolution", 480)
            r = (r - 7 + 6) // 2 + 1
            r = (r - 3 + 2) // 2 + 1
            r = r // 2 // 2 // 2
            net.avgpool = nn.Conv2d(net.fc.in_features, 32, 1)
            net.fc = nn.Linear(r * r * 32, d)
    elif kind in ("swin_transformer_tiny", "swin_transformer_small"):
        net = tvm.swin_t() if kind.endswith("tiny") else tvm.swin_s()
        net.head = nn.Linear(net.head.in_features, d)
    else:
        raise ValueError(f"Invalid image encoder type: {kind}")
    return net


def image_tokens(images, sd, hp, train_bn: bool = False):
    """image.py:38-52: (B,F,3,R,R) -> (B,F,d) through the torchvision trunk + head."""
    has_seq = hp["image_sequence_encoder_type"] == "transformer"
    prefix = "image_sequence_encoder.image_encoder.encoder." if has_seq else "image_sequence_encoder.encoder."
    net = build_trunk(hp)
    sub = {k[len(prefix) :]: v for k, v in sd.items() if k.startswith(prefix)}
    net = net.to(images.dtype)
    net.train(train_bn)
    # running statistics are updated in place in train mode: hand functional_call private copies
    sub = {k: (v.clone() if ("running_" in k or "num_batches" in k) else v) for k, v in sub.items()}
    B, Fr = images.shape[:2]
    tok = torch.func.functional_call(net, sub, (images.reshape(B * Fr, *images.shape[2:]),), strict=True)
    return tok.view(B, Fr, -1)


# ------------------------------------------------------------------------------------------------
# top level


def encode_input_data(batch: dict, sd: dict, hp: dict, masks=None, train_bn=False):
    """model.py:123-148 — list of context tensors in the reference's fixed order."""
    ctx = []
    if hp["use_action_history"]:
        ctx.append(base_encoder(batch["joint_command_history"], sd, "action_history_encoder", 4, masks))
    if hp["use_imu"]:
        ctx.append(base_encoder(batch["rotation"], sd, "imu_encoder", 4, masks))
    if hp["use_joint_states"]:
        ctx.append(base_encoder(batch["joint_state"], sd, "joint_states_encoder", 4, masks))
    if hp["use_images"]:
        tok = image_tokens(batch["image_data"], sd, hp, train_bn)
        if hp["image_sequence_encoder_type"] == "transformer":
            tok = base_encoder(tok, sd, "image_sequence_encoder.transformer_encoder", 8, masks)
        ctx.append(tok)
    if hp["use_gamestate"]:
        ctx.append(sd["game_state_encoder.embedding.weight"][batch["game_state"]].unsqueeze(1))
    return ctx


def forward_with_context(ctx: list, x, step, sd: dict, hp: dict, masks=None):
    """model.py:159-179."""
    tok = step_token(step, sd["step_encoding.token"], hp["hidden_dim"])
    mem = torch.cat(list(ctx) + [tok.to(x.dtype)], dim=1)
    return denoiser(x, mem, sd, 4, masks)


def forward(batch, x, step, sd, hp, masks=None, train_bn=False):
    """model.py:150-157."""
    return forward_with_context(encode_input_data(batch, sd, hp, masks, train_bn), x, step, sd, hp, masks)


def cast_state_dict(sd: dict, dtype) -> dict:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def sample_ddim(ctx, x_T, sd, hp, num_steps: int, num_train_timesteps: int = 1000):
    """ros.py:301-310 / distill.py:179-189 — DDIM eta=0 loop; scheduler arithmetic in float32 (oracle/ddim.py).

    Returns (x_0, list of eps_hat per step)."""
    from oracle.ddim import DDIMOracle

    sch = DDIMOracle(num_train_timesteps)
    sch.set_timesteps(num_steps)
    x = x_T.clone()
    eps_all = []
    B = x.shape[0]
    for t in sch.timesteps:
        eps = forward_with_context(ctx, x, torch.full((B,), int(t), dtype=torch.int64), sd, hp)
        eps_all.append(eps)
        if x.dtype == torch.float32:
            x = torch.from_numpy(sch.step(eps.numpy(), int(t), x.numpy()).prev_sample)
        else:
            sb, sa, sap, sbp = (float(c) for c in sch.coefficients(int(t)))
            x = sap * ((x - sb * eps) / sa) + sbp * eps
    return x, eps_all


def training_loss(batch, noise, t, sd, hp, masks=None, train_bn=True, context_override=None):
    """train.py:204-229: normalise -> add_noise -> model -> mse."""
    from oracle.ddim import DDIMOracle

    x0 = (batch["joint_command"] - sd["mean"]) / sd["std"]
    sch = DDIMOracle(hp.get("train_denoising_timesteps", 1000))
    a = torch.from_numpy(sch.alphas_cumprod)[t].to(x0.dtype)
    x_t = (a**0.5)[:, None, None] * x0 + ((1 - a) ** 0.5)[:, None, None] * noise
    if context_override is not None:
        pred = forward_with_context(context_override, x_t, t, sd, hp, masks)
    else:
        pred = forward(batch, x_t, t, sd, hp, masks, train_bn)
    return F.mse_loss(pred, noise), pred
