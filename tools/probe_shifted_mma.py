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
    cyc = torch.zeros(4, device="cuda", dtype=torch.int64)
    _lib.check(lib.sd_debug_shifted_mma(X.data_ptr(), Y.data_ptr(), X.shape[0], shift_a, lbo_a, shift_b, lbo_b, nblk_b, ksteps, base, reps,
                                        D.data_ptr(), cyc.data_ptr(), _lib.stream_ptr()), "probe")
    torch.cuda.synchronize()
    return D, (cyc.tolist() if base & 4 else int(cyc[0].item()))


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


def ld_concurrency():
    """Do tcgen05.ld (epilogue warps) and tcgen05.mma (other TMEM columns) overlap?  cycles: [MMA chain, ld loop of warps 1..3]"""
    X = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    Y = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    nld = 256
    for nb in (2, 4):
        _, c_mma = run(X, Y, 0, 1024, 0, 1024, nb, 4, 2, reps=128)
        _, c_ld = run(X, Y, nld, 1024, 0, 1024, nb, 1, 2 | 4, reps=1)
        _, c_both = run(X, Y, nld, 1024, 0, 1024, nb, 4, 2 | 4, reps=128)
        print(f"N={64*nb}: 512 MMAs alone {c_mma} clk | {nld} x 3 warps tcgen05.ld.32x32b.x32 alone {c_ld[1:]} clk "
              f"({3 * nld * 4096 / max(c_ld[1:]):.1f} B/clk) | together: MMAs {c_both[0]} clk, loads {c_both[1:]} clk")


def ldtm_rate():
    """TMEM read rate per lane quadrant (csrc/debug_mma.cu: dbg_ldtm_kernel)."""
    lib = _lib.lib()
    sink = torch.zeros(1, device="cuda")
    count = 512
    for mode, name, nbytes in ((0, ".x32, wait after each", 4096), (1, ".x32, two in flight", 4096), (2, ".x16, wait after each", 2048)):
        for nw in (1, 4, 8, 16):
            cyc = torch.zeros(16, device="cuda", dtype=torch.int64)
            _lib.check(lib.sd_debug_ldtm(nw, count, mode, cyc.data_ptr(), sink.data_ptr(), _lib.stream_ptr()), "ldtm")
            torch.cuda.synchronize()
            c = max(cyc.tolist())
            per_quadrant = (nw + 3) // 4
            print(f"tcgen05.ld.32x32b{name:24s} {nw:2d} warps ({per_quadrant} per lane quadrant): {c / count:7.1f} clk per load and warp, "
                  f"{per_quadrant * nbytes * count / c:5.1f} B/clk per quadrant, {nw * nbytes * count / c:6.1f} B/clk per SM")


def round_trip():
    """issue n MMAs -> tcgen05.commit -> mbarrier wait by the issuing thread: the fixed part is the hand-off latency"""
    X = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    Y = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    for n in (1, 2, 4, 8, 16):
        cs = [run(X, Y, 0, 1024, 0, 1024, 2, n if n <= 4 else 4, 2, reps=max(1, n // 4))[1] for _ in range(3)]
        print(f"{n:2d} x (128x128x16 MMA) + commit + wait: {min(cs)} clk")


def commit_cost():
    """512 K-major 128x128x16 MMAs with a tcgen05.commit after every n-th"""
    X = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    Y = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    for every in (0, 8, 4, 2, 1):
        _, c = run(X, Y, 0, 1024, every, 1024, 2, 4, 2, reps=128)
        print(f"commit after every {every or 'none':>4} MMAs: {c / 512:6.1f} clk per MMA")


if __name__ == "__main__":
    commit_cost()
    round_trip()
    ldtm_rate()
    ld_concurrency()
    kmajor_rate()
    main()
