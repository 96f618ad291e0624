"""conv1 forward (with bn1 statistics) and weight gradient at the bs=256 shape (2560 frames of 224x224): CUDA-event times.
python tools/stem_conv_micro.py [frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerdiffusion_b200.ml.model.encoder import trunk as T  # noqa: E402
from tools.conv_probe import timed  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
    img = torch.randint(0, 255, (frames, 3, 224, 224), device="cuda", dtype=torch.uint8)
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
    sums = torch.zeros(128, device="cuda", dtype=torch.float64)
    packed = T._stem_conv_s2d_raw(img, w, return_packed=True, sums=sums)
    assert packed[2], "TMA stem kernel not used"
    gf = 2.0 * frames * 112 * 112 * 64 * 256 / 1e9
    from soccerdiffusion_b200 import ops

    xp = packed[1]
    w2 = torch.randn(64, 256, device="cuda").to(torch.bfloat16)
    y = torch.empty((frames, 64, 112, 112), device="cuda", dtype=torch.bfloat16, memory_format=torch.channels_last)
    t = timed(lambda: ops.stem_fprop(xp, w2, y, frames, 224, 224, sums=sums))
    tp = timed(lambda: ops.stem_pack_u8(img, xp, frames, 224, 224))
    dy = torch.randn(frames, 112, 112, 64, device="cuda").to(torch.bfloat16)
    dw = torch.empty(256, 64, device="cuda")
    tw = timed(lambda: ops.stem_wgrad(xp, dy, dw, frames, 224, 224))
    print(f"{frames} frames: conv1 weight gradient {tw:.3f} ms ({gf / tw:.0f} TF/s)  [SD_B200_STEM_WGRAD_PAIR={os.environ.get('SD_B200_STEM_WGRAD_PAIR', '1')}]")
    print(f"{frames} frames: conv1 forward + statistics {t:.3f} ms ({gf / t:.0f} TF/s, {(xp.numel() * 2 + y.numel() * 2) / t / 1e6:.0f} GB/s);  uint8 packing {tp:.3f} ms")


if __name__ == "__main__":
    main()
