"""layer1 weight gradient: sd_conv3x3_wgrad_c64_bf16 vs cuDNN (CUDA events, back-to-back launches).  python tools/wgrad_micro.py [frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerdiffusion_b200 import ops  # noqa: E402
from tools.conv_probe import timed  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
    H = W = 56
    cl = torch.channels_last
    x = torch.randn(frames, 64, H, W, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    dy = torch.randn(frames, 64, H, W, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    w = torch.randn(64, 64, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
    dW = torch.empty(64, 64, 3, 3, device="cuda")
    torch.backends.cudnn.benchmark = True
    gf = 2.0 * frames * H * W * 9 * 64 * 64 / 1e9
    a = timed(lambda: ops.conv3x3_wgrad_c64(x, dy, dW, frames, H, W))
    b = timed(lambda: torch.ops.aten.convolution_backward(dy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False]))
    print(f"layer1 wgrad, {frames} frames: libsd_b200 {a:.3f} ms ({gf / a:.0f} TF/s, {frames * H * W * 256 / a / 1e6:.0f} GB/s algorithmic)   "
          f"cuDNN {b:.3f} ms ({gf / b:.0f} TF/s)")


if __name__ == "__main__":
    main()
