from soccerdiffusion_b200.ml.inference.sampler import TrajectorySampler, sample_loop  # noqa: F401
