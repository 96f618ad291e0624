"""JointEncoder (reference: soccer_diffusion/ml/model/encoder/joint.py:4-29)."""
from soccerdiffusion_b200.ml.model.encoder.base import BaseEncoder


class JointEncoder(BaseEncoder):
    def __init__(
        self, num_joints: int, patch_size: int, hidden_dim: int, num_layers: int, num_heads: int, max_seq_len: int
    ):
        super().__init__(
            input_dim=num_joints,
            patch_size=patch_size,
            hidden_dim=hidden_dim,
            num_layers=num_layers,
            num_heads=num_heads,
            max_seq_len=max_seq_len,
        )
