"""IMUEncoder (reference: soccer_diffusion/ml/model/encoder/imu.py:6-53)."""
from enum import Enum

from soccerdiffusion_b200.ml.model.encoder.base import BaseEncoder


class IMUEncoder(BaseEncoder):
    class OrientationEmbeddingMethod(Enum):
        QUATERNION = "quaternion"
        FIVE_DIM = "five_dim"  # axis-angle with a 2-D vector for the angle

    _INPUT_FEATURES = {"quaternion": 4, "five_dim": 5}

    def __init__(
        self,
        orientation_embedding_method: "IMUEncoder.OrientationEmbeddingMethod",
        patch_size: int,
        hidden_dim: int,
        num_layers: int,
        num_heads: int,
        max_seq_len: int,
    ):
        method = IMUEncoder.OrientationEmbeddingMethod(getattr(orientation_embedding_method, "value",
                                                               orientation_embedding_method))
        super().__init__(
            input_dim=self._INPUT_FEATURES[method.value],
            patch_size=patch_size,
            hidden_dim=hidden_dim,
            num_layers=num_layers,
            num_heads=num_heads,
            max_seq_len=max_seq_len,
        )
