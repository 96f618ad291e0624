"""TEST / BENCH INFRASTRUCTURE — the reference's model assembled from the STOCK PyTorch modules it is made of.

The reference (soccer_diffusion/ml/model) is `torch.nn.TransformerEncoder/Decoder` layers (norm_first, GELU,
dim_feedforward = d), `nn.Conv1d` patch embeddings, `nn.Embedding`, a torchvision ResNet18 with the head surgery of
ml/model/encoder/image.py:55-73, and the step token of ml/model/misc.py:25-35.  The live reference cannot travel to the GPU
box, so this file rebuilds the same module graph from the same stock classes; bench.py's ``torch_eager_b200`` leg runs it
eagerly ON THE B200 (fp32, and bf16 autocast + channels_last) — the "before" a drop-in user has today.
Nothing in the product path imports this file.

    encoders       ml/model/encoder/base.py:7-53, joint.py, imu.py (quaternion embedding), game_state.py:19-27
    image path     ml/model/encoder/image.py:38-52, 55-73, 103-121
    denoiser       ml/model/decoder.py:6-54
    model          ml/model/model.py:16-179
    training step  ml/training/train.py:193-240 (AdamW, mse, add_noise restated inline: sqrt(a) x0 + sqrt(1-a) eps)
"""
from __future__ import annotations

import math

import torch
from torch import nn


class _PE(nn.Module):
    def __init__(self, d, max_len):
        super().__init__()
        pos = torch.arange(max_len).unsqueeze(1)
        div = torch.exp(torch.arange(0, d, 2) * (-math.log(10000.0) / d))
        pe = torch.zeros(1, max_len, d)
        pe[0, :, 0::2] = torch.sin(pos * div)
        pe[0, :, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe, persistent=False)

    def forward(self, x):
        return x + self.pe[:, : x.size(1)]


class _Encoder(nn.Module):
    def __init__(self, cin, patch, d, layers, heads, max_len):
        super().__init__()
        self.embedding = nn.Conv1d(cin, d, kernel_size=patch, stride=patch)
        self.pe = _PE(d, max_len)
        self.enc = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=d, nhead=heads, dim_feedforward=d, batch_first=True, norm_first=True,
                                       activation="gelu"), num_layers=layers, enable_nested_tensor=False)

    def forward(self, x):
        return self.enc(self.pe(self.embedding(x.permute(0, 2, 1)).permute(0, 2, 1)))


class _StepToken(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.d = d
        self.token = nn.Parameter(torch.randn(1, d // 2))

    def forward(self, steps):
        half = self.d // 4
        freqs = torch.exp(torch.arange(half, device=steps.device) * -(math.log(10000.0) / (half - 1)))
        a = steps[:, None] * freqs[None, :]
        return torch.cat((a.sin(), a.cos(), self.token.expand(steps.size(0), self.d // 2)), dim=-1).unsqueeze(1)


class StockModel(nn.Module):
    def __init__(self, hp):
        super().__init__()
        import torchvision.models as tvm

        d, J, p = hp["hidden_dim"], hp["num_joints"], hp["encoder_patch_size"]
        self.action = _Encoder(J, p, d, hp["num_action_history_encoder_layers"], 4, hp["action_context_length"])
        self.imu = _Encoder(4, p, d, hp["num_imu_encoder_layers"], 4, hp["imu_context_length"])
        self.joints = _Encoder(J, p, d, hp["joint_state_encoder_layers"], 4, hp["joint_state_context_length"])
        net = tvm.resnet18(weights=None)
        r = hp.get("image_resolution", 480)
        r = ((r - 7 + 6) // 2 + 1 - 3 + 2) // 2 + 1
        r = r // 2 // 2 // 2
        net.avgpool = nn.Conv2d(512, 32, 1)
        net.fc = nn.Linear(r * r * 32, d)
        self.trunk = net
        self.frames = _Encoder(d, 1, d, hp["num_image_sequence_encoder_layers"], 8, hp["image_context_length"])
        self.game = nn.Embedding(4, d)
        self.step = _StepToken(d)
        self.emb = nn.Linear(J, d)
        self.pe = _PE(d, hp["trajectory_prediction_length"])
        self.dec = nn.TransformerDecoder(
            nn.TransformerDecoderLayer(d_model=d, nhead=4, dim_feedforward=d, batch_first=True, norm_first=True,
                                       activation="gelu"), num_layers=hp["num_decoder_layers"])
        self.out = nn.Linear(d, J)

    def encode(self, b):
        img = b["image_data"]
        B, F = img.shape[:2]
        tok = self.trunk(img.reshape(B * F, *img.shape[2:])).view(B, F, -1)
        return [self.action(b["joint_command_history"]), self.imu(b["rotation"]), self.joints(b["joint_state"]),
                self.frames(tok), self.game(b["game_state"]).unsqueeze(1)]

    def forward(self, b, x, t):
        mem = torch.cat(self.encode(b) + [self.step(t.float())], dim=1)
        return self.out(self.dec(self.pe(self.emb(x)), mem))


def alphas_cumprod(T=1000):
    """squaredcos_cap_v2 (diffusers 0.31.0 betas_for_alpha_bar): float64 betas -> float32 cumprod."""
    f = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = torch.tensor([min(1 - f((i + 1) / T) / f(i / T), 0.999) for i in range(T)], dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


def train_step(model, opt, batch, acp, mean, std, autocast_bf16: bool):
    """ml/training/train.py:193-240 on one device batch; returns the loss tensor."""
    jt = (batch["joint_command"] - mean) / std
    opt.zero_grad()
    t = torch.randint(0, acp.numel(), (jt.size(0),), device=jt.device)
    noise = torch.randn_like(jt)
    a = acp[t][:, None, None]
    noisy = a.sqrt() * jt + (1 - a).sqrt() * noise
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast_bf16):
        pred = model(batch, noisy, t)
    loss = torch.nn.functional.mse_loss(pred.float(), noise)
    loss.backward()
    opt.step()
    return loss.detach()
