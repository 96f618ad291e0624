"""Hyper-parameter dictionaries (the flat YAML keys of ml/training/config/*.yaml) -> model.

``build_model`` mirrors the constructor call of ml/training/train.py:113-139 (same keys, same ``.get``
defaults: image_resolution 480, image_use_final_avgpool True)."""
from __future__ import annotations

# ml/training/config/default.yaml
DEFAULT = dict(
    hidden_dim=128, action_context_length=100, trajectory_prediction_length=10, epochs=10, batch_size=64, lr=1e-4,
    train_denoising_timesteps=1000, image_context_length=10, imu_context_length=100, joint_state_context_length=100,
    num_normalization_samples=1000, num_joints=20, use_action_history=True, num_action_history_encoder_layers=2,
    use_imu=True, imu_orientation_embedding_method="quaternion", num_imu_encoder_layers=2, use_joint_states=True,
    joint_state_encoder_layers=2, use_images=True, image_sequence_encoder_type="transformer",
    image_encoder_type="resnet18", image_resolution=224, image_use_final_avgpool=False,
    num_image_sequence_encoder_layers=1, num_decoder_layers=4, distill_teacher_inference_steps=30, use_gamestate=True,
    encoder_patch_size=1,
)

# BASELINE.json configs[4]: 2x depth, longer image history and action horizon
SCALED = dict(DEFAULT, num_action_history_encoder_layers=4, num_imu_encoder_layers=4, joint_state_encoder_layers=4,
              num_image_sequence_encoder_layers=2, num_decoder_layers=8, image_context_length=20,
              trajectory_prediction_length=20)


def build_model(params: dict):
    from soccerdiffusion_b200.ml.model.encoder.image import ImageEncoderType, SequenceEncoderType
    from soccerdiffusion_b200.ml.model.encoder.imu import IMUEncoder
    from soccerdiffusion_b200.ml.model.model import End2EndDiffusionTransformer

    return End2EndDiffusionTransformer(
        num_joints=params["num_joints"],
        hidden_dim=params["hidden_dim"],
        use_action_history=params["use_action_history"],
        num_action_history_encoder_layers=params["num_action_history_encoder_layers"],
        max_action_context_length=params["action_context_length"],
        encoder_patch_size=params["encoder_patch_size"],
        use_imu=params["use_imu"],
        imu_orientation_embedding_method=IMUEncoder.OrientationEmbeddingMethod(params["imu_orientation_embedding_method"]),
        num_imu_encoder_layers=params["num_imu_encoder_layers"],
        imu_context_length=params["imu_context_length"],
        use_joint_states=params["use_joint_states"],
        joint_state_encoder_layers=params["joint_state_encoder_layers"],
        joint_state_context_length=params["joint_state_context_length"],
        use_images=params["use_images"],
        image_encoder_type=ImageEncoderType(params["image_encoder_type"]),
        image_sequence_encoder_type=SequenceEncoderType(params["image_sequence_encoder_type"]),
        num_image_sequence_encoder_layers=params["num_image_sequence_encoder_layers"],
        image_context_length=params["image_context_length"],
        image_use_final_avgpool=params.get("image_use_final_avgpool", True),
        image_resolution=params.get("image_resolution", 480),
        use_gamestate=params["use_gamestate"],
        num_decoder_layers=params["num_decoder_layers"],
        trajectory_prediction_length=params["trajectory_prediction_length"],
    )


def synthetic_batch(params: dict, batch: int, device, seed: int = 0, pin: bool = False, uint8_images: bool = False) -> dict:
    """Synthetic inputs of the named shapes: joints U[0,2pi), unit quaternions, N(0,1) images
    (``uint8_images``: raw U{0..255} frames instead — the model then runs the reference's preprocessing on the device)."""
    import math

    import torch

    g = torch.Generator().manual_seed(seed)
    J = params["num_joints"]
    b = {}
    b["joint_command_history"] = torch.rand(batch, params["action_context_length"], J, generator=g) * (2 * math.pi)
    b["joint_state"] = torch.rand(batch, params["joint_state_context_length"], J, generator=g) * (2 * math.pi)
    imu_dim = 4 if params["imu_orientation_embedding_method"] == "quaternion" else 5
    q = torch.randn(batch, params["imu_context_length"], imu_dim, generator=g)
    b["rotation"] = q / q.norm(dim=-1, keepdim=True)
    if params.get("use_images", True):
        R = params.get("image_resolution", 480)
        if uint8_images:
            b["image_data"] = torch.randint(0, 256, (batch, params["image_context_length"], 3, R, R), generator=g, dtype=torch.uint8)
        else:
            b["image_data"] = torch.randn(batch, params["image_context_length"], 3, R, R, generator=g)
    b["game_state"] = torch.randint(0, 4, (batch,), generator=g)
    b["joint_command"] = torch.rand(batch, params["trajectory_prediction_length"], J, generator=g) * (2 * math.pi)
    if pin:
        b = {k: v.pin_memory() for k, v in b.items()}
    if device is not None and str(device) != "cpu":
        b = {k: v.to(device, non_blocking=True) for k, v in b.items()}
    return b
