from soccerdiffusion_b200.ml.inference.sampler import FrameEmbeddingCache, TrajectorySampler, sample_loop  # noqa: F401
