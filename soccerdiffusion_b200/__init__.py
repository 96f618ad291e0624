"""soccerdiffusion_b200 — B200-native (sm_100a) denoiser + DDIM sampler, drop-in for ``soccer_diffusion.ml``.

Public surface (mirrors the reference import paths):
    soccerdiffusion_b200.ml.model.End2EndDiffusionTransformer
    soccerdiffusion_b200.ml.model.encoder.imu.IMUEncoder
    soccerdiffusion_b200.ml.model.encoder.image.{ImageEncoderType, SequenceEncoderType}
    soccerdiffusion_b200.schedulers.scheduling_ddim.DDIMScheduler   (for diffusers.schedulers.scheduling_ddim)
    soccerdiffusion_b200.dataset.pytorch.Normalizer
    soccerdiffusion_b200.ml.training / soccerdiffusion_b200.ml.inference   (loop bodies of train.py / distill.py / ros.py)
"""
from .runtime import manual_seed, precision_name, set_precision  # noqa: F401

__version__ = "0.1.0"
