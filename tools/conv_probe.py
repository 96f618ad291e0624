"""Per-shape device time of the trunk's LIBRARY convolutions (cuDNN through aten, bf16 channels_last): forward, data gradient and
weight gradient separately, CUDA events, back-to-back launches.  python tools/conv_probe.py [frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import RESNET18_CONVS  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
    torch.backends.cudnn.benchmark = True
    seen = {}
    for shape in RESNET18_CONVS:
        seen[shape] = seen.get(shape, 0) + 1
    print("| cin | cout | k | stride | H_in | count | GF | fprop ms (TF/s) | dgrad ms (TF/s) | wgrad ms (TF/s) |\n|---|---|---|---|---|---|---|---|---|---|")
    tot = [0.0, 0.0, 0.0]
    for (cin, cout, k, st, hin), mult in seen.items():
        cl = torch.channels_last
        x = torch.randn(frames, cin, hin, hin, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
        w = torch.randn(cout, cin, k, k, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=cl)
        args = ([st, st], [k // 2, k // 2], [1, 1], False, [0, 0], 1)
        y = torch.ops.aten.convolution(x, w, None, *args)
        gy = torch.randn_like(y)
        gf = 2.0 * frames * y.shape[2] * y.shape[3] * cout * cin * k * k / 1e9
        f = timed(lambda: torch.ops.aten.convolution(x, w, None, *args))
        d = timed(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, *args, [True, False, False]))
        g = timed(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, *args, [False, True, False]))
        for i, t in enumerate((f, d, g)):
            tot[i] += mult * t
        print(f"| {cin} | {cout} | {k} | {st} | {hin} | {mult} | {gf:.0f} | {f:.3f} ({gf / f:.0f}) | {d:.3f} ({gf / d:.0f}) | {g:.3f} ({gf / g:.0f}) |", flush=True)
        del x, w, y, gy
        torch.cuda.empty_cache()
    print(f"\ntotal per step: fprop {tot[0]:.2f} ms, dgrad {tot[1]:.2f} ms, wgrad {tot[2]:.2f} ms = {sum(tot):.2f} ms")


if __name__ == "__main__":
    main()
