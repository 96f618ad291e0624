from soccerdiffusion_b200.ml.model.model import End2EndDiffusionTransformer  # noqa: F401
