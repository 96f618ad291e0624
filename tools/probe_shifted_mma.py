"""Does tcgen05.mma accept MN-major SWIZZLE_128B operands that are row-shifted, overlapping views of one TMA tile?
(csrc/debug_mma.cu)   python tools/probe_shifted_mma.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerdiffusion_b200 import _lib  # noqa: E402


def run(X, Y, shift_a, lbo_a, shift_b, lbo_b, nblk_b, ksteps, base, reps=1):
    lib = _lib.lib()
    N = 64 * nblk_b
    D = torch.zeros(128, N, device="cuda")
    cyc = torch.zeros(1, device="cuda", dtype=torch.int64)
    _lib.check(lib.sd_debug_shifted_mma(X.data_ptr(), Y.data_ptr(), X.shape[0], shift_a, lbo_a, shift_b, lbo_b, nblk_b, ksteps, base, reps,
                                        D.data_ptr(), cyc.data_ptr(), _lib.stream_ptr()), "probe")
    torch.cuda.synchronize()
    return D, int(cyc.item())


def ref(X, Y, shift_a, lbo_a, shift_b, lbo_b, nblk_b, ksteps):
    K = 16 * ksteps
    Xf, Yf = X.float(), Y.float()
    A = torch.cat([Xf[shift_a + blk * (lbo_a // 128): shift_a + blk * (lbo_a // 128) + K].T for blk in range(2)])           # [128][K]
    B = torch.cat([Yf[shift_b + blk * (lbo_b // 128): shift_b + blk * (lbo_b // 128) + K].T for blk in range(nblk_b)])     # [N][K]
    return A @ B.T


def main():
    torch.manual_seed(0)
    rows = 128
    X = torch.randn(rows, 64, device="cuda").to(torch.bfloat16)
    Y = torch.randn(rows, 64, device="cuda").to(torch.bfloat16)
    cases = [("aligned views, blocks 8 rows apart", 0, 1024, 0, 1024, 3, 2),
             ("A shifted 1 row", 1, 1024, 0, 1024, 3, 2),
             ("A shifted 3 rows, B shifted 5 rows", 3, 1024, 5, 1024, 3, 2),
             ("A blocks 1 row apart (lbo 128)", 0, 128, 0, 1024, 3, 2),
             ("A shifted 1, blocks 1 row apart; B blocks 58 rows apart... (rows limit: 20)", 1, 128, 2, 20 * 128, 3, 2),
             ("everything odd", 7, 128, 3, 13 * 128, 4, 3)]
    for name, sa, la, sb, lb, nb, ks in cases:
        want = ref(X, Y, sa, la, sb, lb, nb, ks)
        for base in (0, 1):
            got, _ = run(X, Y, sa, la, sb, lb, nb, ks, base)
            err = float((got - want).norm() / want.norm())
            print(f"{name:75s} base_offset={base}: rel err {err:.3e} {'OK' if err < 1e-2 else 'WRONG'}")
    # issue rate: 128 x N x 16 instructions back to back
    for nb in (1, 2, 3, 4):
        for sa, la in ((0, 1024), (1, 128)):
            _, c = run(X, Y, sa, la, 0, 1024, nb, 4, 1, reps=256)
            print(f"N={64*nb:3d} A shift {sa} lbo {la:4d}: {c / (256 * 4):7.1f} clk per 128x{64*nb}x16 MMA  (floor {64*nb//2})")


def kmajor_rate():
    X = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    Y = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    for nb in (1, 2, 3, 4):
        _, c = run(X, Y, 0, 1024, 0, 1024, nb, 4, 2, reps=256)
        print(f"K-major operands N={64*nb:3d}: {c / (256 * 4):7.1f} clk per 128x{64*nb}x16 MMA  (floor {64*nb//2})")


if __name__ == "__main__":
    kmajor_rate()
    main()
