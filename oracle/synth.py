"""TEST INFRASTRUCTURE — deterministic synthetic weights and inputs.

Everything is generated from a counter-based integer hash (splitmix64) in numpy
uint64 arithmetic, so the same (name, shape, seed) gives bit-identical float32
arrays on every machine and library version.  That is what lets the golden
vectors under tests/golden/ hold only OUTPUTS of the real reference: weights and
inputs are re-synthesised at test time on the GPU box (where /root/reference
does not exist).

Shapes and distributions follow SURVEY.md §8(d).
"""
from __future__ import annotations

import zlib

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def _key(name: str, seed: int) -> np.uint64:
    return np.uint64(((zlib.crc32(name.encode()) & 0xFFFFFFFF) << 32) ^ (seed & 0xFFFFFFFF))


def uniform01(name: str, shape, seed: int = 0) -> np.ndarray:
    """float64 uniforms in [0,1) with 53 random bits."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + _splitmix64(np.array([_key(name, seed)], dtype=np.uint64))[0]
    bits = _splitmix64(ctr)
    u = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return u.reshape(shape)


def uniform(name, shape, lo, hi, seed=0) -> np.ndarray:
    return (lo + (hi - lo) * uniform01(name, shape, seed)).astype(np.float32)


def normal(name, shape, seed=0, std=1.0) -> np.ndarray:
    """Box-Muller in float64, rounded once to float32."""
    u1 = uniform01(name + "/u1", shape, seed)
    u2 = uniform01(name + "/u2", shape, seed)
    z = np.sqrt(-2.0 * np.log1p(-u1)) * np.cos(2.0 * np.pi * u2)
    return (std * z).astype(np.float32)


def randint(name, shape, lo, hi, seed=0) -> np.ndarray:
    return (lo + np.floor(uniform01(name, shape, seed) * (hi - lo))).astype(np.int64)


# --------------------------------------------------------------------------------------
# weights


def synth_state_dict(template: dict, seed: int = 0) -> dict:
    """Fills every tensor of ``template`` (name -> torch tensor / shape-like) deterministically.

    Rules: LayerNorm / BatchNorm weights ~ 1 + U(-.2,.2); biases ~ U(-.1,.1) (so that every
    parameter is exercised by parity tests); running_var ~ U(.5,1.5); running_mean ~ U(-.1,.1);
    matrices/conv kernels ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default bound);
    ``mean``/``std`` buffers: mean ~ U(2.5,3.5), std ~ U(.8,1.8).
    """
    import torch

    out = {}
    for name, t in template.items():
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=torch.int64)
            continue
        if name == "mean":
            a = uniform(name, shape, 2.5, 3.5, seed)
        elif name == "std":
            a = uniform(name, shape, 0.8, 1.8, seed)
        elif name.endswith("running_mean"):
            a = uniform(name, shape, -0.1, 0.1, seed)
        elif name.endswith("running_var"):
            a = uniform(name, shape, 0.5, 1.5, seed)
        elif len(shape) == 1 and (".norm" in name or ".bn" in name or "downsample.1" in name) and name.endswith("weight"):
            a = uniform(name, shape, 0.8, 1.2, seed)
        elif len(shape) == 1:
            a = uniform(name, shape, -0.1, 0.1, seed)
        elif name.endswith("step_encoding.token"):
            a = normal(name, shape, seed)
        elif "game_state_encoder.embedding" in name:
            a = normal(name, shape, seed)
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            a = uniform(name, shape, -bound, bound, seed)
        out[name] = torch.from_numpy(np.ascontiguousarray(a))
    return out


# --------------------------------------------------------------------------------------
# inputs


def synth_batch(hp: dict, batch: int, seed: int = 0, with_images: bool = True) -> dict:
    import torch

    J = hp["num_joints"]
    two_pi = 2.0 * np.pi
    b = {}
    b["joint_command_history"] = uniform("jch", (batch, hp["action_context_length"], J), 0.0, two_pi, seed)
    b["joint_state"] = uniform("js", (batch, hp["joint_state_context_length"], J), 0.0, two_pi, seed)
    imu_dim = 4 if hp["imu_orientation_embedding_method"] == "quaternion" else 5
    q = normal("rot", (batch, hp["imu_context_length"], imu_dim), seed).astype(np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    b["rotation"] = q.astype(np.float32)
    if with_images and hp.get("use_images", True):
        R = hp.get("image_resolution", 480)
        b["image_data"] = normal("img", (batch, hp["image_context_length"], 3, R, R), seed)
    b["game_state"] = randint("gs", (batch,), 0, 4, seed)
    T = hp["trajectory_prediction_length"]
    b["joint_command"] = uniform("jc", (batch, T, J), 0.0, two_pi, seed)
    out = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in b.items()}
    return out


def synth_noise(name: str, hp: dict, batch: int, seed: int = 0):
    import torch

    return torch.from_numpy(normal(name, (batch, hp["trajectory_prediction_length"], hp["num_joints"]), seed))


def synth_timesteps(batch: int, seed: int = 0, T: int = 1000):
    import torch

    return torch.from_numpy(randint("t", (batch,), 0, T, seed))


# --------------------------------------------------------------------------------------
# named configurations

DEFAULT_HP = dict(
    hidden_dim=128,
    action_context_length=100,
    trajectory_prediction_length=10,
    train_denoising_timesteps=1000,
    image_context_length=10,
    imu_context_length=100,
    joint_state_context_length=100,
    num_joints=20,
    use_action_history=True,
    num_action_history_encoder_layers=2,
    use_imu=True,
    imu_orientation_embedding_method="quaternion",
    num_imu_encoder_layers=2,
    use_joint_states=True,
    joint_state_encoder_layers=2,
    use_images=True,
    image_sequence_encoder_type="transformer",
    image_encoder_type="resnet18",
    image_resolution=224,
    image_use_final_avgpool=False,
    num_image_sequence_encoder_layers=1,
    num_decoder_layers=4,
    distill_teacher_inference_steps=30,
    use_gamestate=True,
    encoder_patch_size=1,
    lr=1e-4,
)

TINY_HP = dict(
    DEFAULT_HP,
    hidden_dim=32,
    action_context_length=20,
    imu_context_length=20,
    joint_state_context_length=20,
    image_context_length=2,
    image_resolution=64,
    num_action_history_encoder_layers=1,
    num_imu_encoder_layers=1,
    joint_state_encoder_layers=1,
    num_image_sequence_encoder_layers=1,
    num_decoder_layers=2,
)

# decoder_only.yaml (train.py:221-224): denoiser-only pretraining, d=256, foreign context
DECODER_ONLY_HP = dict(
    DEFAULT_HP,
    hidden_dim=256,
    use_action_history=False,
    use_imu=False,
    use_joint_states=False,
    use_images=False,
    use_gamestate=False,
    encoder_patch_size=10,
)

SCALED_HP = dict(
    DEFAULT_HP,
    num_action_history_encoder_layers=4,
    num_imu_encoder_layers=4,
    joint_state_encoder_layers=4,
    num_image_sequence_encoder_layers=2,
    num_decoder_layers=8,
    image_context_length=20,
    trajectory_prediction_length=20,
)

# ml/training/config/larger_model.yaml:1-27 verbatim (d=512, dh=128, encoders 4/4/4/1, 8 decoder layers; the YAML sets
# neither image_resolution nor image_use_final_avgpool -> train.py's .get defaults 480 / True), except the resolution of the
# synthetic frames (the avg-pooled head makes every other shape independent of it)
LARGER_HP = dict(
    DEFAULT_HP,
    hidden_dim=512,
    num_action_history_encoder_layers=4,
    num_imu_encoder_layers=4,
    joint_state_encoder_layers=4,
    num_image_sequence_encoder_layers=1,
    num_decoder_layers=8,
    image_use_final_avgpool=True,
    image_resolution=96,
)

# ml/training/config/sim_scratch.yaml verbatim (d=256, patch 5, five_dim IMU, no joint states, no game state, 6 decoder
# layers, un-pooled image head), except the frame resolution (224 -> 64: the fc layer of the head follows it)
SIM_SCRATCH_HP = dict(
    DEFAULT_HP,
    hidden_dim=256,
    num_action_history_encoder_layers=4,
    imu_orientation_embedding_method="five_dim",
    num_imu_encoder_layers=2,
    use_joint_states=False,
    joint_state_encoder_layers=4,
    image_use_final_avgpool=False,
    image_resolution=64,
    num_decoder_layers=6,
    use_gamestate=False,
    encoder_patch_size=5,
)

# the scaled-up configuration WITH its image branch, at a frame size a fixture can afford
SCALED_IMG_HP = dict(SCALED_HP, image_resolution=64)

# a patchified, five_dim, J=22 variant exercising the non-default branches (sim_scratch-like)
PATCH_HP = dict(
    TINY_HP,
    hidden_dim=64,
    num_joints=22,
    encoder_patch_size=5,
    imu_orientation_embedding_method="five_dim",
    use_gamestate=False,
    use_images=False,
    num_decoder_layers=1,
)
