"""Helpers of the GPU parity tests (everything goes through the package -> ctypes -> C ABI)."""
import numpy as np
import torch

from oracle import synth


def build_model(hp, device="cuda"):
    from soccerdiffusion_b200.ml.model.encoder.image import ImageEncoderType, SequenceEncoderType
    from soccerdiffusion_b200.ml.model.encoder.imu import IMUEncoder
    from soccerdiffusion_b200.ml.model.model import End2EndDiffusionTransformer

    m = End2EndDiffusionTransformer(
        num_joints=hp["num_joints"], hidden_dim=hp["hidden_dim"], use_action_history=hp["use_action_history"],
        num_action_history_encoder_layers=hp["num_action_history_encoder_layers"],
        max_action_context_length=hp["action_context_length"], encoder_patch_size=hp["encoder_patch_size"],
        use_imu=hp["use_imu"],
        imu_orientation_embedding_method=IMUEncoder.OrientationEmbeddingMethod(hp["imu_orientation_embedding_method"]),
        num_imu_encoder_layers=hp["num_imu_encoder_layers"], imu_context_length=hp["imu_context_length"],
        use_joint_states=hp["use_joint_states"], joint_state_encoder_layers=hp["joint_state_encoder_layers"],
        joint_state_context_length=hp["joint_state_context_length"], use_images=hp["use_images"],
        image_encoder_type=ImageEncoderType(hp["image_encoder_type"]),
        image_sequence_encoder_type=SequenceEncoderType(hp["image_sequence_encoder_type"]),
        num_image_sequence_encoder_layers=hp["num_image_sequence_encoder_layers"],
        image_context_length=hp["image_context_length"], image_use_final_avgpool=hp.get("image_use_final_avgpool", True),
        image_resolution=hp.get("image_resolution", 480), use_gamestate=hp["use_gamestate"],
        num_decoder_layers=hp["num_decoder_layers"], trajectory_prediction_length=hp["trajectory_prediction_length"])
    return m.to(device)


def synth_model(hp, seed, device="cuda"):
    m = build_model(hp, "cpu")
    sd = synth.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    return m.to(device), sd


def to_dev(batch, device="cuda"):
    return {k: v.to(device) for k, v in batch.items()}


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
