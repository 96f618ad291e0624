"""Only the piece of ``soccer_diffusion.dataset`` that sits on the hot path: ``Normalizer``."""
