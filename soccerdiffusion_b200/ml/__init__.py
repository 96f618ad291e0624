"""Mirror of ``soccer_diffusion.ml`` for the hot path (model, noise scheduler, train/inference loop bodies)."""
import logging

logger = logging.getLogger("soccerdiffusion_b200.ml")
