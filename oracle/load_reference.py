"""TEST INFRASTRUCTURE — imports the *real* reference modules from /root/reference.

Only usable in the build container (``/root/reference`` does not exist on the GPU
box).  Used by ``oracle/gen_golden.py`` and by the ``not gpu`` tests to pin the
restated oracle (``oracle/model_ref.py``) against the live reference code.

Three offline stubs are needed:
  1. ``importlib.metadata.version("soccer_diffusion")`` raises PackageNotFoundError
     (soccer_diffusion/__init__.py:8) and the package wants a writable log dir
     (soccer_diffusion/__init__.py:12-37).
  2. ml/model/encoder/game_state.py:4 imports dataset.models (SQLAlchemy missing);
     only ``len(RobotState) == 4`` (dataset/models.py:13-25) is used.
  3. ml/model/encoder/image.py:64,66 ask for ImageNet weights (network download).
"""
from __future__ import annotations

import enum
import importlib
import importlib.metadata
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("SD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "soccer_diffusion", "ml", "model"))


_loaded = None


def load_reference():
    """Returns the imported ``soccer_diffusion.ml.model`` package namespace as a dict."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")

    # stub 1: package metadata + log dir
    os.environ.setdefault("SOCCER_DIFFUSION_LOG_DIR", tempfile.mkdtemp(prefix="sd_ref_logs_"))
    real_version = importlib.metadata.version

    def _version(name):
        if name == "soccer_diffusion":
            return "1.0.0"
        return real_version(name)

    importlib.metadata.version = _version

    # stub 2: dataset.models.RobotState (4 members; dataset/models.py:13-25)
    class RobotState(str, enum.Enum):
        PLAYING = "PLAYING"
        POSITIONING = "POSITIONING"
        STOPPED = "STOPPED"
        UNKNOWN = "UNKNOWN"

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        import soccer_diffusion  # noqa: F401
    finally:
        importlib.metadata.version = real_version

    ds_pkg = types.ModuleType("soccer_diffusion.dataset")
    ds_pkg.__path__ = []  # mark as package, nothing importable from disk
    ds_models = types.ModuleType("soccer_diffusion.dataset.models")
    ds_models.RobotState = RobotState
    sys.modules.setdefault("soccer_diffusion.dataset", ds_pkg)
    sys.modules["soccer_diffusion.dataset.models"] = ds_models

    # the ml package __init__ sets up rich logging; import it as-is
    image = importlib.import_module("soccer_diffusion.ml.model.encoder.image")

    # stub 3: random-init trunks instead of downloading ImageNet weights
    import torchvision.models as tvm

    image.resnet18 = lambda weights=None: tvm.resnet18(weights=None)
    image.resnet50 = lambda weights=None: tvm.resnet50(weights=None)

    model = importlib.import_module("soccer_diffusion.ml.model.model")
    imu = importlib.import_module("soccer_diffusion.ml.model.encoder.imu")
    misc = importlib.import_module("soccer_diffusion.ml.model.misc")
    decoder = importlib.import_module("soccer_diffusion.ml.model.decoder")
    base = importlib.import_module("soccer_diffusion.ml.model.encoder.base")
    _loaded = dict(
        End2EndDiffusionTransformer=model.End2EndDiffusionTransformer,
        IMUEncoder=imu.IMUEncoder,
        ImageEncoderType=image.ImageEncoderType,
        SequenceEncoderType=image.SequenceEncoderType,
        StepToken=misc.StepToken,
        PositionalEncoding=misc.PositionalEncoding,
        DiffusionActionGenerator=decoder.DiffusionActionGenerator,
        BaseEncoder=base.BaseEncoder,
    )
    return _loaded


def build_reference_model(hp: dict):
    """Builds the real reference model from a flat hyper-parameter dict (train.py:113-139)."""
    ref = load_reference()
    return ref["End2EndDiffusionTransformer"](
        num_joints=hp["num_joints"],
        hidden_dim=hp["hidden_dim"],
        use_action_history=hp["use_action_history"],
        num_action_history_encoder_layers=hp["num_action_history_encoder_layers"],
        max_action_context_length=hp["action_context_length"],
        encoder_patch_size=hp["encoder_patch_size"],
        use_imu=hp["use_imu"],
        imu_orientation_embedding_method=ref["IMUEncoder"].OrientationEmbeddingMethod(
            hp["imu_orientation_embedding_method"]
        ),
        num_imu_encoder_layers=hp["num_imu_encoder_layers"],
        imu_context_length=hp["imu_context_length"],
        use_joint_states=hp["use_joint_states"],
        joint_state_encoder_layers=hp["joint_state_encoder_layers"],
        joint_state_context_length=hp["joint_state_context_length"],
        use_images=hp["use_images"],
        image_encoder_type=ref["ImageEncoderType"](hp["image_encoder_type"]),
        image_sequence_encoder_type=ref["SequenceEncoderType"](hp["image_sequence_encoder_type"]),
        num_image_sequence_encoder_layers=hp["num_image_sequence_encoder_layers"],
        image_context_length=hp["image_context_length"],
        image_use_final_avgpool=hp.get("image_use_final_avgpool", True),
        image_resolution=hp.get("image_resolution", 480),
        use_gamestate=hp["use_gamestate"],
        num_decoder_layers=hp["num_decoder_layers"],
        trajectory_prediction_length=hp["trajectory_prediction_length"],
    )
